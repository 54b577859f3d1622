// Launchers and tensor-map encoding of the TMA-fed kernels (kernels_tma.cuh); separate translation unit of libdasm.so.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>

#include "tma_launch.h"

namespace dasm
{
  namespace
  {
    typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

    EncodeTiled
    encode_fn()
    {
      static EncodeTiled fn = []() -> EncodeTiled {
        void *                           p = nullptr;
        cudaDriverEntryPointQueryResult  q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
          return nullptr;
        return (EncodeTiled)p;
      }();
      return fn;
    }

    template <typename T, int n>
    void
    eo_fill(EOMat<T, n> &E, const double *P, const double *Q)
    {
      constexpr int m = (n + 1) / 2, h = n / 2;
      for (int i = 0; i < m * m; ++i)
        E.P[i] = (T)P[i];
      for (int i = 0; i < h * h; ++i)
        E.Q[i] = (T)Q[i];
    }

    // DASM_VERBOSE: resident blocks per SM of a kernel (once per kernel)
    void
    report_occupancy(const char *name, const void *kern, const int threads, const size_t smem)
    {
      static const bool verbose = getenv("DASM_VERBOSE") != nullptr;
      if (!verbose)
        return;
      static std::map<const void *, int> seen;
      if (seen.count(kern))
        return;
      int occ = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
      seen[kern] = occ;
      fprintf(stderr, "[dasm] %s: %d threads, %zu bytes of shared memory, %d resident blocks per SM\n", name, threads, smem, occ);
    }

    void
    check(cudaError_t e, const char *what)
    {
      if (e != cudaSuccess)
        throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
    }
  } // namespace

  bool
  tma_encode_maps(TmaMaps &out, const void *vec, int k, int esize, long long n_lex, std::string &err)
  {
    EncodeTiled fn = encode_fn();
    if (fn == nullptr)
      {
        err = "cuTensorMapEncodeTiled is not available";
        return false;
      }
    if ((reinterpret_cast<uintptr_t>(vec) & 15) != 0 || n_lex <= 0)
      {
        err = "vector is not 16-byte aligned";
        return false;
      }
    const cuuint64_t R = 4 * k, XW = 16 / esize;
    const cuuint32_t RZ = (cuuint32_t)(tma_item_layers(k, esize) * k); // z extent of a work item (kernels_tma.cuh TmaItem)
    const cuuint64_t dims[4]    = {R, R, R, (cuuint64_t)n_lex};
    const cuuint64_t strides[3] = {R * esize, R * R * esize, R * R * R * esize};
    const cuuint32_t estr[4]    = {1, 1, 1, 1};
    const CUtensorMapDataType dt = esize == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    struct
    {
      CUtensorMap *m;
      cuuint32_t   box[4];
    } specs[8] = {{&out.main, {(cuuint32_t)R, (cuuint32_t)R, RZ, 1}},
                  {&out.fx, {(cuuint32_t)XW, (cuuint32_t)R, RZ, 1}},
                  {&out.fy, {(cuuint32_t)R, 1, RZ, 1}},
                  {&out.fz, {(cuuint32_t)R, (cuuint32_t)R, 1, 1}},
                  {&out.exy, {(cuuint32_t)XW, 1, RZ, 1}},
                  {&out.exz, {(cuuint32_t)XW, (cuuint32_t)R, 1, 1}},
                  {&out.eyz, {(cuuint32_t)R, 1, 1, 1}},
                  {&out.cxyz, {(cuuint32_t)XW, 1, 1, 1}}};
    if (k == 4)
      {
        // TmaLayout PERM: the box enumerated as (x, (Y>>2)&1, Y&3, (Y>>3) + 2 Z + 32 brick); double: 128-byte swizzle
        const cuuint64_t row        = R * esize;
        const cuuint64_t pdims[4]   = {R, 2, 4, (cuuint64_t)(32 * n_lex)};
        const cuuint64_t pstr[3]    = {4 * row, row, 8 * row};
        const cuuint32_t pbox[4]    = {(cuuint32_t)R, 2, 4, 2 * RZ};
        const CUresult   r = fn(&out.main, dt, 4, const_cast<void *>(vec), pdims, pstr, pbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              esize == 8 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS)
          {
            err = "cuTensorMapEncodeTiled (permuted box) failed with code " + std::to_string((int)r);
            return false;
          }
      }
    for (auto &s : specs)
      {
        if (k == 4 && s.m == &out.main)
          continue;
        const CUresult r = fn(s.m, dt, 4, const_cast<void *>(vec), dims, strides, s.box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS)
          {
            err = "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r);
            return false;
          }
      }
    return true;
  }

  size_t
  tma_laplace_smem(int k, int esize)
  {
#define DASM_TMA_SMEM(K) (esize == 8 ? TmaSmem<K, double, TmaItem<K, double>::BZ>::bytes(2, 1) : TmaSmem<K, float, TmaItem<K, float>::BZ>::bytes(2, 1))
    switch (k)
      {
        case 2:
          return DASM_TMA_SMEM(2);
        case 3:
          return DASM_TMA_SMEM(3);
        case 4:
          return DASM_TMA_SMEM(4);
        case 5:
          return DASM_TMA_SMEM(5);
        case 6:
          return DASM_TMA_SMEM(6);
      }
#undef DASM_TMA_SMEM
    return (size_t)-1;
  }

  size_t
  tma_fdm_smem(int k, int esize)
  {
#define DASM_TMA_SMEM(K) (esize == 8 ? TmaSmem<K, double, TmaItem<K, double>::BZ>::bytes(1, 2) + 8 * (K + 1) * (K + 1) * (K + 1) : TmaSmem<K, float, TmaItem<K, float>::BZ>::bytes(1, 2) + 4 * (K + 1) * (K + 1) * (K + 1))
    switch (k)
      {
        case 2:
          return DASM_TMA_SMEM(2);
        case 3:
          return DASM_TMA_SMEM(3);
        case 4:
          return DASM_TMA_SMEM(4);
        case 5:
          return DASM_TMA_SMEM(5);
        case 6:
          return DASM_TMA_SMEM(6);
      }
#undef DASM_TMA_SMEM
    return (size_t)-1;
  }

  template <int K, typename T>
  static void
  launch_laplace_tma_k(cudaStream_t stream, int grid, const T *src, T *dst, T *acc, const Epilogue<T> &epi, const double (*P)[25],
                       const double (*Q)[25], const TmaMaps &maps, const CUtensorMap &omap0, int shared_mode, const NextInit<T> &ni, const TmaList &list, int dbg)
  {
    using G = TmaGeom<K, T, TmaItem<K, T>::BZ>;
    FastLaplaceMats<T, K + 1> mats;
    eo_fill(mats.M, P[0], Q[0]);
    eo_fill(mats.K0, P[1], Q[1]);
    eo_fill(mats.K1, P[2], Q[2]);
    eo_fill(mats.K2, P[3], Q[3]);
    constexpr size_t smem  = TmaSmem<K, T, TmaItem<K, T>::BZ>::bytes(2, 1);
    const bool       need0 = (epi.kind == EPI_RESIDUAL || epi.kind == EPI_CHEB);
    auto             kern  = list.any_mode1 ? (need0 ? laplace_tma_kernel<K, T, 1, true> : laplace_tma_kernel<K, T, 0, true>) :
                                              (need0 ? laplace_tma_kernel<K, T, 1, false> : laplace_tma_kernel<K, T, 0, false>);
    check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "laplace_tma_kernel attribute");
    check(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), "carveout");
    report_occupancy("laplace_tma_kernel", (const void *)kern, G::NT, smem);
    const FastMaps fm = {nullptr, nullptr, nullptr, nullptr, 0, nullptr, dbg};
    kern<<<grid, G::NT, smem, stream>>>(src, dst, acc, epi, mats, maps, omap0, shared_mode, ni, list, fm);
    check(cudaGetLastError(), "laplace_tma_kernel launch");
  }

  template <int K, typename T>
  static void
  launch_fdm_tma_k(cudaStream_t stream, int grid, const T *src, T *dst, T *acc, const Epilogue<T> &epi, const double (*P)[25],
                   const double (*Q)[25], const double *inv, const TmaMaps &maps, const CUtensorMap &omap0, const CUtensorMap &omap1, int shared_mode, const NextInit<T> &ni,
                   const TmaList &list, int dbg)
  {
    using G         = TmaGeom<K, T, TmaItem<K, T>::BZ>;
    constexpr int n = K + 1;
    FastFdmMats<T, n> mats;
    eo_fill(mats.Ax, P[0], Q[0]);
    eo_fill(mats.Ay, P[1], Q[1]);
    eo_fill(mats.Az, P[2], Q[2]);
    eo_fill(mats.Bx, P[3], Q[3]);
    eo_fill(mats.By, P[4], Q[4]);
    eo_fill(mats.Bz, P[5], Q[5]);
    for (int i = 0; i < n * n * n; ++i)
      mats.inv[i] = (T)inv[i];
    constexpr size_t smem  = TmaSmem<K, T, TmaItem<K, T>::BZ>::bytes(1, 2);
    const bool       need0 = (epi.kind == EPI_RESIDUAL || epi.kind == EPI_CHEB);
    const bool       need1 = (epi.kind == EPI_CHEB && epi.f1 != T(0) && epi.v1 != nullptr);
    auto             kern  = list.any_mode1 ? (need1 ? fdm_tma_kernel<K, T, 2, true> : (need0 ? fdm_tma_kernel<K, T, 1, true> : fdm_tma_kernel<K, T, 0, true>)) :
                                              (need1 ? fdm_tma_kernel<K, T, 2, false> : (need0 ? fdm_tma_kernel<K, T, 1, false> : fdm_tma_kernel<K, T, 0, false>));
    check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "fdm_tma_kernel attribute");
    check(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), "carveout");
    report_occupancy("fdm_tma_kernel", (const void *)kern, G::NT, smem);
    const FastMaps fm = {nullptr, nullptr, nullptr, nullptr, 0, nullptr, dbg};
    kern<<<grid, G::NT, smem, stream>>>(src, dst, acc, epi, mats, maps, omap0, omap1, shared_mode, ni, list, fm);
    check(cudaGetLastError(), "fdm_tma_kernel launch");
  }

  template <typename T>
  void
  launch_laplace_tma(int k, cudaStream_t stream, int grid, const T *src, T *dst, T *acc, const Epilogue<T> &epi, const double (*P)[25],
                     const double (*Q)[25], const TmaMaps &maps, const CUtensorMap &omap0, int shared_mode, const NextInit<T> &ni, const TmaList &list, int dbg)
  {
    switch (k)
      {
        case 2:
          return launch_laplace_tma_k<2, T>(stream, grid, src, dst, acc, epi, P, Q, maps, omap0, shared_mode, ni, list, dbg);
        case 3:
          return launch_laplace_tma_k<3, T>(stream, grid, src, dst, acc, epi, P, Q, maps, omap0, shared_mode, ni, list, dbg);
        case 4:
          return launch_laplace_tma_k<4, T>(stream, grid, src, dst, acc, epi, P, Q, maps, omap0, shared_mode, ni, list, dbg);
        case 5:
          return launch_laplace_tma_k<5, T>(stream, grid, src, dst, acc, epi, P, Q, maps, omap0, shared_mode, ni, list, dbg);
        case 6:
          return launch_laplace_tma_k<6, T>(stream, grid, src, dst, acc, epi, P, Q, maps, omap0, shared_mode, ni, list, dbg);
      }
    throw std::runtime_error("laplace_tma_kernel: degree not instantiated");
  }

  template <typename T>
  void
  launch_fdm_tma(int k, cudaStream_t stream, int grid, const T *src, T *dst, T *acc, const Epilogue<T> &epi, const double (*P)[25],
                 const double (*Q)[25], const double *inv, const TmaMaps &maps, const CUtensorMap &omap0, const CUtensorMap &omap1, int shared_mode, const NextInit<T> &ni, const TmaList &list,
                 int dbg)
  {
    switch (k)
      {
        case 2:
          return launch_fdm_tma_k<2, T>(stream, grid, src, dst, acc, epi, P, Q, inv, maps, omap0, omap1, shared_mode, ni, list, dbg);
        case 3:
          return launch_fdm_tma_k<3, T>(stream, grid, src, dst, acc, epi, P, Q, inv, maps, omap0, omap1, shared_mode, ni, list, dbg);
        case 4:
          return launch_fdm_tma_k<4, T>(stream, grid, src, dst, acc, epi, P, Q, inv, maps, omap0, omap1, shared_mode, ni, list, dbg);
        case 5:
          return launch_fdm_tma_k<5, T>(stream, grid, src, dst, acc, epi, P, Q, inv, maps, omap0, omap1, shared_mode, ni, list, dbg);
        case 6:
          return launch_fdm_tma_k<6, T>(stream, grid, src, dst, acc, epi, P, Q, inv, maps, omap0, omap1, shared_mode, ni, list, dbg);
      }
    throw std::runtime_error("fdm_tma_kernel: degree not instantiated");
  }

  template void launch_laplace_tma<double>(int, cudaStream_t, int, const double *, double *, double *, const Epilogue<double> &,
                                           const double (*)[25], const double (*)[25], const TmaMaps &, const CUtensorMap &, int,
                                           const NextInit<double> &, const TmaList &, int);
  template void launch_laplace_tma<float>(int, cudaStream_t, int, const float *, float *, float *, const Epilogue<float> &,
                                          const double (*)[25], const double (*)[25], const TmaMaps &, const CUtensorMap &, int,
                                          const NextInit<float> &, const TmaList &, int);
  template void launch_fdm_tma<double>(int, cudaStream_t, int, const double *, double *, double *, const Epilogue<double> &,
                                       const double (*)[25], const double (*)[25], const double *, const TmaMaps &, const CUtensorMap &,
                                       const CUtensorMap &, int, const NextInit<double> &, const TmaList &, int);
  template void launch_fdm_tma<float>(int, cudaStream_t, int, const float *, float *, float *, const Epilogue<float> &, const double (*)[25],
                                      const double (*)[25], const double *, const TmaMaps &, const CUtensorMap &, const CUtensorMap &, int,
                                      const NextInit<float> &, const TmaList &, int);
} // namespace dasm
