// Exact-block additive Schwarz on the device (reference: RestrictedPreconditioner over Restrictors::ElementCenteredRestrictor with
// RestrictedMatrixView blocks inverted by gauss_jordan - include/preconditioners.h:528-605, 744-813; include/restrictors.h:17-378).
// It is the accuracy reference of the fast-diagonalisation preconditioner: same patches, same weights, exact block inverse.
//
// Set-up:  B_c = R_c A R_c^T is obtained matrix-free: the operator is applied to unit vectors placed at patch entry j of all cells of
//          one colour at once (cells of a colour are >= 3 cells apart in some direction, so neither the supports of the columns nor
//          the patches that read them overlap): 27 colours x m^3 operator applications.  Entries outside the domain / constrained
//          get identity rows.  The blocks are inverted in place by a batched Gauss-Jordan kernel (one thread block per cell; the
//          blocks are symmetric positive definite: no pivoting) and stored transposed in the number type of the operator.
// Apply:   one thread block per cell: gather (x pre-weight) -> dense m^3 x m^3 mat-vec, block read coalesced (the only HBM traffic
//          that matters: m^6 S bytes per cell) -> scatter with atomic adds (x post-weight).
#include <cuda_runtime.h>

#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

#include "dasm.h"

namespace
{
  constexpr uint32_t INVALID = 0xFFFFFFFFu;

#define BA_CUDA_CHECK(x)                                                                                                                 \
  do                                                                                                                                     \
    {                                                                                                                                    \
      cudaError_t e_ = (x);                                                                                                              \
      if (e_ != cudaSuccess)                                                                                                             \
        throw std::runtime_error(std::string("CUDA error ") + cudaGetErrorString(e_) + " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
    }                                                                                                                                    \
  while (0)
#define BA_CALL(x)                                   \
  do                                                 \
    {                                                \
      if ((x) != 0)                                  \
        throw std::runtime_error(dasm_last_error()); \
    }                                                \
  while (0)
#define BA_API_BEGIN try {
#define BA_API_END                 \
  return 0;                        \
  }                                \
  catch (const std::exception &e)  \
  {                                \
    dasm_set_last_error(e.what()); \
    return 1;                      \
  }

  // x[idx[c][j]] = 1 for the cells c of the list
  template <typename T>
  __global__ void
  set_unit_kernel(T *x, const uint32_t *__restrict__ idx, const uint32_t *__restrict__ cells, const int n_list, const int m3, const int j)
  {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_list)
      {
        const uint32_t g = idx[(size_t)cells[i] * m3 + j];
        if (g != INVALID)
          x[g] = T(1);
      }
  }

  // column j of the blocks of the listed cells from y = A x
  template <typename T>
  __global__ void
  read_column_kernel(double *__restrict__ blocks, const T *__restrict__ y, const uint32_t *__restrict__ idx, const uint32_t *__restrict__ cells,
                     const int n_list, const int m3, const int j)
  {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n_list * m3)
      return;
    const uint32_t c = cells[t / m3];
    const int      i = (int)(t % m3);
    const uint32_t gj = idx[(size_t)c * m3 + j], gi = idx[(size_t)c * m3 + i];
    double         v;
    if (gj == INVALID || gi == INVALID)
      v = (i == j) ? 1. : 0.;
    else
      v = (double)y[gi];
    blocks[((size_t)c * m3 + i) * m3 + j] = v;
  }

  // in-place Gauss-Jordan inversion, one thread block per matrix (row-major, no pivoting: symmetric positive definite blocks)
  __global__ void __launch_bounds__(256)
  gauss_jordan_kernel(double *__restrict__ blocks, const int m3)
  {
    double *       A = blocks + (size_t)blockIdx.x * m3 * m3;
    extern __shared__ double sh[]; // pivot row | pivot column
    double *       prow = sh, *pcol = sh + m3;
    for (int p = 0; p < m3; ++p)
      {
        const double piv = 1. / A[(size_t)p * m3 + p];
        for (int i = threadIdx.x; i < m3; i += blockDim.x)
          {
            prow[i] = A[(size_t)p * m3 + i] * piv;
            pcol[i] = A[(size_t)i * m3 + p];
          }
        __syncthreads();
        for (int e = threadIdx.x; e < m3 * m3; e += blockDim.x)
          {
            const int i = e / m3, j = e % m3;
            double    v;
            if (i == p)
              v = (j == p) ? piv : prow[j];
            else if (j == p)
              v = -pcol[i] * piv;
            else
              v = A[e] - pcol[i] * prow[j];
            A[e] = v;
          }
        __syncthreads();
      }
  }

  // out[c][j][i] = (T) in[c][i][j]
  template <typename T>
  __global__ void
  transpose_convert_kernel(T *__restrict__ out, const double *__restrict__ in, const int m3, const long long n_cells)
  {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_cells * m3 * m3)
      return;
    const long long c = t / ((long long)m3 * m3);
    const int       e = (int)(t % ((long long)m3 * m3));
    const int       j = e / m3, i = e % m3;
    out[t]            = (T)in[((size_t)c * m3 + i) * m3 + j];
  }

  // dst += sum_c W_post R_c^T B_c^-1 R_c W_pre src
  template <typename T>
  __global__ void __launch_bounds__(128)
  asm_apply_kernel(T *__restrict__ dst, const T *__restrict__ src, const T *__restrict__ binv_t, const uint32_t *__restrict__ idx,
                   const T *__restrict__ w, const int m3, const int w_pre, const int w_post)
  {
    extern __shared__ unsigned char sh_raw[];
    T *             v = reinterpret_cast<T *>(sh_raw);
    const long long c = blockIdx.x;
    for (int i = threadIdx.x; i < m3; i += blockDim.x)
      {
        const uint32_t g = idx[(size_t)c * m3 + i];
        T              x = (g == INVALID) ? T(0) : src[g];
        if (w_pre && g != INVALID)
          x *= w[(size_t)c * m3 + i];
        v[i] = x;
      }
    __syncthreads();
    const T *B = binv_t + (size_t)c * m3 * m3;
    for (int i = threadIdx.x; i < m3; i += blockDim.x)
      {
        const uint32_t g = idx[(size_t)c * m3 + i];
        if (g == INVALID)
          continue; // (uniform work per row otherwise; the row of an entry outside the patch is not needed)
        T s = 0;
        for (int j = 0; j < m3; ++j)
          s += B[(size_t)j * m3 + i] * v[j];
        if (w_post)
          s *= w[(size_t)c * m3 + i];
        atomicAdd(dst + g, s);
      }
  }
} // namespace

struct dasm_asm
{
  dasm_fdm *   layout = nullptr;
  dasm_op *    op     = nullptr;
  cudaStream_t stream = nullptr;
  int          ntype  = DASM_F64;
  int          m3     = 0;
  long long    n_cells = 0;
  uint32_t *   d_idx  = nullptr;
  void *       d_w    = nullptr;
  void *       d_binv = nullptr; // [cell][j][i] transposed inverse blocks
  int          w_pre = 0, w_post = 0;
  size_t       bytes = 0;
};

template <typename T>
static void
asm_setup(dasm_asm *a)
{
  dasm_op *       op = a->op;
  const int       m3 = a->m3;
  const long long nc = a->n_cells;
  cudaStream_t    s  = a->stream;
  int             ncd[3], per[3];
  BA_CALL(dasm_mesh_global_size(dasm_op_mesh(op), ncd, per));
  std::vector<int> coord((size_t)nc * 3);
  BA_CALL(dasm_mesh_cell_coordinates(dasm_op_mesh(op), coord.data()));
  // colour stride per direction: cells of one colour are >= 3 apart (also across a periodic boundary)
  int stride[3];
  for (int d = 0; d < 3; ++d)
    {
      stride[d] = 3;
      if (per[d])
        {
          stride[d] = ncd[d];
          for (int q = 3; q <= ncd[d]; ++q)
            if (ncd[d] % q == 0)
              {
                stride[d] = q;
                break;
              }
        }
      stride[d] = std::max(1, std::min(stride[d], ncd[d]));
    }
  const int                          n_colours = stride[0] * stride[1] * stride[2];
  std::vector<std::vector<uint32_t>> lists(n_colours);
  for (long long c = 0; c < nc; ++c)
    lists[(coord[3 * c] % stride[0]) + stride[0] * ((coord[3 * c + 1] % stride[1]) + stride[1] * (coord[3 * c + 2] % stride[2]))].push_back(
      (uint32_t)c);
  double *d_blocks = nullptr;
  BA_CUDA_CHECK(cudaMalloc(&d_blocks, std::max<size_t>(1, (size_t)nc * m3 * m3) * sizeof(double)));
  T *x = nullptr, *y = nullptr;
  BA_CALL(dasm_op_vec_alloc(op, (void **)&x));
  BA_CALL(dasm_op_vec_alloc(op, (void **)&y));
  const size_t vbytes = (size_t)dasm_op_vec_size(op) * sizeof(T);
  uint32_t *   d_list = nullptr;
  size_t       max_list = 1;
  for (const auto &l : lists)
    max_list = std::max(max_list, l.size());
  BA_CUDA_CHECK(cudaMalloc(&d_list, max_list * sizeof(uint32_t)));
  for (const auto &l : lists)
    {
      if (l.empty())
        continue;
      const int nl = (int)l.size();
      BA_CUDA_CHECK(cudaMemcpyAsync(d_list, l.data(), l.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
      for (int j = 0; j < m3; ++j)
        {
          BA_CUDA_CHECK(cudaMemsetAsync(x, 0, vbytes, s));
          set_unit_kernel<T><<<(nl + 127) / 128, 128, 0, s>>>(x, a->d_idx, d_list, nl, m3, j);
          BA_CALL(dasm_op_vmult(op, y, x));
          const long long nt = (long long)nl * m3;
          read_column_kernel<T><<<(unsigned)((nt + 255) / 256), 256, 0, s>>>(d_blocks, y, a->d_idx, d_list, nl, m3, j);
        }
      BA_CUDA_CHECK(cudaStreamSynchronize(s)); // the host list buffer is reused
    }
  if (nc > 0)
    {
      gauss_jordan_kernel<<<(unsigned)nc, 256, 2 * m3 * sizeof(double), s>>>(d_blocks, m3);
      const long long ne = nc * m3 * m3;
      BA_CUDA_CHECK(cudaMalloc(&a->d_binv, (size_t)ne * sizeof(T)));
      transpose_convert_kernel<T><<<(unsigned)((ne + 255) / 256), 256, 0, s>>>((T *)a->d_binv, d_blocks, m3, nc);
    }
  BA_CUDA_CHECK(cudaGetLastError());
  BA_CUDA_CHECK(cudaStreamSynchronize(s));
  cudaFree(d_blocks);
  cudaFree(d_list);
  dasm_op_vec_free(op, x);
  dasm_op_vec_free(op, y);
}

extern "C" int
dasm_asm_create(dasm_fdm *layout, dasm_asm **out)
{
  BA_API_BEGIN
  auto a     = new dasm_asm;
  a->layout  = layout;
  a->op      = dasm_fdm_op(layout);
  a->stream  = (cudaStream_t)dasm_ctx_stream(dasm_op_ctx(a->op));
  a->ntype   = dasm_op_number_type(a->op);
  const int m = dasm_fdm_patch_size_1d(layout);
  a->m3      = m * m * m;
  a->n_cells = dasm_op_n_cells(a->op);
  if (dasm_op_n_ghost(a->op) > 0)
    throw std::runtime_error("exact-block ASM is built for one rank (the coloured probing writes owned DoFs only)");
  const size_t es = a->ntype == DASM_F64 ? 8 : 4;
  a->bytes        = (size_t)a->n_cells * a->m3 * a->m3 * es;
  if ((double)a->n_cells * a->m3 * a->m3 * (8. + es) > 64e9)
    throw std::runtime_error("exact-block ASM: the dense blocks of this mesh need more than 64 GB");
  BA_CUDA_CHECK(cudaMalloc(&a->d_idx, std::max<size_t>(1, (size_t)a->n_cells * a->m3) * sizeof(uint32_t)));
  BA_CUDA_CHECK(cudaMalloc(&a->d_w, std::max<size_t>(1, (size_t)a->n_cells * a->m3) * es));
  BA_CALL(dasm_fdm_export_patches(layout, a->d_idx, a->d_w, &a->w_pre, &a->w_post));
  if (a->ntype == DASM_F64)
    asm_setup<double>(a);
  else
    asm_setup<float>(a);
  *out = a;
  BA_API_END
}

extern "C" int
dasm_asm_destroy(dasm_asm *a)
{
  if (a)
    {
      cudaFree(a->d_idx);
      cudaFree(a->d_w);
      cudaFree(a->d_binv);
      delete a;
    }
  return 0;
}

extern "C" int
dasm_asm_vmult(dasm_asm *a, void *dst, const void *src)
{
  BA_API_BEGIN
  const size_t es = a->ntype == DASM_F64 ? 8 : 4;
  BA_CUDA_CHECK(cudaMemsetAsync(dst, 0, (size_t)dasm_op_vec_size(a->op) * es, a->stream));
  if (a->n_cells > 0)
    {
      if (a->ntype == DASM_F64)
        asm_apply_kernel<double><<<(unsigned)a->n_cells, 128, a->m3 * sizeof(double), a->stream>>>((double *)dst, (const double *)src,
                                                                                                  (const double *)a->d_binv, a->d_idx,
                                                                                                  (const double *)a->d_w, a->m3, a->w_pre, a->w_post);
      else
        asm_apply_kernel<float><<<(unsigned)a->n_cells, 128, a->m3 * sizeof(float), a->stream>>>((float *)dst, (const float *)src,
                                                                                                (const float *)a->d_binv, a->d_idx,
                                                                                                (const float *)a->d_w, a->m3, a->w_pre, a->w_post);
    }
  BA_CUDA_CHECK(cudaGetLastError());
  BA_API_END
}

extern "C" long long
dasm_asm_memory_consumption(const dasm_asm *a)
{
  return (long long)a->bytes;
}

extern "C" int
dasm_asm_block(const dasm_asm *a, long long cell, double *out)
{
  BA_API_BEGIN
  if (cell < 0 || cell >= a->n_cells)
    throw std::runtime_error("cell index out of range");
  const size_t n = (size_t)a->m3 * a->m3;
  BA_CUDA_CHECK(cudaStreamSynchronize(a->stream));
  if (a->ntype == DASM_F64)
    {
      std::vector<double> h(n);
      BA_CUDA_CHECK(cudaMemcpy(h.data(), (const double *)a->d_binv + (size_t)cell * n, n * sizeof(double), cudaMemcpyDeviceToHost));
      for (int j = 0; j < a->m3; ++j)
        for (int i = 0; i < a->m3; ++i)
          out[(size_t)i * a->m3 + j] = h[(size_t)j * a->m3 + i];
    }
  else
    {
      std::vector<float> h(n);
      BA_CUDA_CHECK(cudaMemcpy(h.data(), (const float *)a->d_binv + (size_t)cell * n, n * sizeof(float), cudaMemcpyDeviceToHost));
      for (int j = 0; j < a->m3; ++j)
        for (int i = 0; i < a->m3; ++i)
          out[(size_t)i * a->m3 + j] = h[(size_t)j * a->m3 + i];
    }
  BA_API_END
}

// host copy of the patch layout (Restrictors::ElementCenteredRestrictor: indices and weights, include/restrictors.h:48-338)
extern "C" int
dasm_fdm_patches_host(dasm_fdm *fdm, uint32_t *idx, double *w, int *w_pre, int *w_post)
{
  BA_API_BEGIN
  dasm_op *       op = dasm_fdm_op(fdm);
  const int       m  = dasm_fdm_patch_size_1d(fdm);
  const size_t    n  = (size_t)dasm_op_n_cells(op) * m * m * m;
  const bool      f64 = dasm_op_number_type(op) == DASM_F64;
  cudaStream_t    s   = (cudaStream_t)dasm_ctx_stream(dasm_op_ctx(op));
  uint32_t *      d_idx = nullptr;
  void *          d_w   = nullptr;
  BA_CUDA_CHECK(cudaMalloc(&d_idx, std::max<size_t>(1, n) * sizeof(uint32_t)));
  BA_CUDA_CHECK(cudaMalloc(&d_w, std::max<size_t>(1, n) * 8));
  BA_CALL(dasm_fdm_export_patches(fdm, d_idx, d_w, w_pre, w_post));
  BA_CUDA_CHECK(cudaStreamSynchronize(s));
  BA_CUDA_CHECK(cudaMemcpy(idx, d_idx, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  if (f64)
    BA_CUDA_CHECK(cudaMemcpy(w, d_w, n * sizeof(double), cudaMemcpyDeviceToHost));
  else
    {
      std::vector<float> h(n);
      BA_CUDA_CHECK(cudaMemcpy(h.data(), d_w, n * sizeof(float), cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < n; ++i)
        w[i] = h[i];
    }
  cudaFree(d_idx);
  cudaFree(d_w);
  BA_API_END
}
