// Host-side launchers of the TMA-fed kernels (kernels_tma.cuh), compiled in their own translation unit (tma_launch.cu).
#pragma once
#include <cuda_runtime.h>

#include "kernels_tma.cuh"

namespace dasm
{
  // tensor maps (x, y, z, brick) of a vector whose first n_lex * 64 k^3 entries are the boxes of the lex bricks;
  // false: the pointer is not 16-byte aligned or the driver entry point is missing
  bool tma_encode_maps(TmaMaps &out, const void *vec, int k, int esize, long long n_lex, std::string &err);

  size_t tma_laplace_smem(int k, int esize);
  size_t tma_fdm_smem(int k, int esize);

  // P / Q: even-odd blocks (EOMat) of M, g0 K, g1 K, g2 K   (Laplace) and of Ax Ay Az Bx By Bz (FDM)
  template <typename T>
  void launch_laplace_tma(int k, cudaStream_t stream, int grid, const T *src, T *dst, T *acc, const Epilogue<T> &epi, const double (*P)[25],
                          const double (*Q)[25], const TmaMaps &maps, const CUtensorMap &omap0, int shared_mode, const NextInit<T> &ni,
                          const TmaList &list, int dbg);

  template <typename T>
  void launch_fdm_tma(int k, cudaStream_t stream, int grid, const T *src, T *dst, T *acc, const Epilogue<T> &epi, const double (*P)[25],
                      const double (*Q)[25], const double *inv, const TmaMaps &maps, const CUtensorMap &omap0, const CUtensorMap &omap1,
                      int shared_mode, const NextInit<T> &ni, const TmaList &list, int dbg);
} // namespace dasm
