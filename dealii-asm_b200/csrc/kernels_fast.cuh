// Building blocks of the warp-specialised kernels of kernels_tma.cuh (sm_100a): even-odd form of the 1-D matrices, the
// register-plane contraction mat_vec, named-barrier / mbarrier / bulk-copy primitives and the merge of the x / y neighbour
// contributions by warp shuffles.
//
//   compute warps   thread (cell c, plane t), lanes = cells.  The operators are applied in their
//                   Kronecker form with ALL 1-D matrices as kernel parameters (constant-bank operands of the
//                   FMAs, no matrix loads):
//                     Laplace (operator.h:866-875 on a Cartesian cell; N^T D^T W D N = K, N^T W N = M exactly):
//                       A_cell = g0 K(x)M(x)M + g1 M(x)K(x)M + g2 M(x)M(x)K          7 sweeps instead of 12
//                       phase A  planes y = t:  q = Mx Mz v,  p = (g0 Kx Mz + g2 Mx Kz) v
//                       phase B  planes z = t:  r = My p + g1 Ky q
//                     FDM (matrix_free.h:1046-1052 apply_inverse, weights 1366-1488 folded into the first / last
//                     1-D matrices when they are a tensor product):
//                       phase A  planes z = t:  Sx^T Wx, Sy^T Wy      phase B  planes y = t:  Sz^T Wz, 1/(lx+ly+lz), Sz, Sx
//                       phase C  planes z = t:  Wy' Sy
//                   Every 1-D contraction runs in even-odd form (EOMat / mat_vec below): M and K are centrosymmetric,
//                   the eigenvectors of the symmetric 1-D problem are even or odd and ordered even-first by the host.
//                   The last contraction is linear and uses the same matrix in every cell, so the plane z = k of the
//                   cell below is added to the plane z = 0 BEFORE it; the contributions of the x / y neighbours
//                   are merged with warp shuffles (lane - 1, lane - 4).  Every DoF of the brick closure is then
//                   final in exactly one thread.
#pragma once
#include "kernels_brick.cuh"

namespace dasm
{
  // debugging switches of the kernels (DASM_FAST_DBG, results invalid): 1 no compute phases, 2 no epilogue
  struct FastMaps
  {
    const uint16_t *ltab;
    const uint16_t *ftab;
    const uint32_t *foreign_gidx;
    const uint32_t *brick_ids;
    int             n;
    long long *     prof;
    int             dbg;
  };

  // Even-odd form of an n x n 1-D matrix whose rows / columns are symmetric or antisymmetric under i -> n-1-i
  // (mass and stiffness matrices on symmetric nodes; eigenvector matrices of a symmetric 1-D problem with the even
  // eigenvectors ordered first): an m x m block acting on the even parts and an h x h block acting on the odd parts,
  // m = ceil(n/2), h = floor(n/2).  13 multiplications instead of 25 for n = 5.
  template <typename T, int n>
  struct EOMat
  {
    static constexpr int m = (n + 1) / 2, h = n / 2;
    T                    P[m * m], Q[h * h > 0 ? h * h : 1];
  };

  template <typename T, int n>
  struct FastLaplaceMats
  {
    EOMat<T, n> M, K0, K1, K2; // Kd = g_d K
  };

  template <typename T, int n>
  struct FastFdmMats
  {
    EOMat<T, n> Ax, Ay, Az; // first stage S^T diag(w_pre): nodal -> eigen space (even eigenvectors first)
    EOMat<T, n> Bx, By, Bz; // second stage diag(w_post) S: eigen -> nodal space
    T           inv[n * n * n]; // 1 / (lx[x] + ly[y] + lz[z]) at (z n + y) n + x, eigenvalues in the even-first order
  };

  enum
  {
    FB_OUT_FULL  = 1, // compute -> movers: the result of a brick is in the output tile
    FB_OUT_EMPTY = 2, // movers -> compute: the output tile has been stored
    FB_COMPUTE   = 3, // compute warps only
    FB_MOVERS    = 4  // mover warps only
  };

  __device__ __forceinline__ void
  bar_sync(const int id, const int count)
  {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
  }
  __device__ __forceinline__ void
  bar_arrive(const int id, const int count)
  {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
  }
  __device__ __forceinline__ void
  cp_async_wait_all()
  {
    asm volatile("cp.async.wait_all;\n" ::: "memory");
  }

  // ---- 1-D bulk copy (TMA engine, bypasses the LSU) with mbarrier completion -------------------------------------------
  __device__ __forceinline__ void
  mbar_init(const unsigned mbar, const unsigned count)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  __device__ __forceinline__ void
  mbar_expect_tx(const unsigned mbar, const unsigned bytes)
  {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
  }
  __device__ __forceinline__ void
  bulk_load(const unsigned dst, const void *src, const unsigned bytes, const unsigned mbar)
  {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(mbar)
                 : "memory");
  }
  __device__ __forceinline__ void
  mbar_wait(const unsigned mbar, const unsigned parity)
  {
    unsigned ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok)
                   : "r"(mbar), "r"(parity)
                   : "memory");
  }

  // load issued exactly here (volatile): descriptors / indices of later bricks are requested one phase before their use
  __device__ __forceinline__ uint32_t
  ldg_early(const uint32_t *p)
  {
    uint32_t v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
  }

  // r (+)= A v in even-odd form.  PRE: the input is nodal (split into even / odd parts), else it is already in the
  // even-first eigen order; POST: the output is nodal (recombined), else even-first.
  template <int n, typename T, bool PRE, bool POST, bool ADD>
  __device__ __forceinline__ void
  mat_vec(T (&r)[n], const EOMat<T, n> &A, const T (&v)[n])
  {
    constexpr int m = (n + 1) / 2, h = n / 2;
    T             e[m], o[h > 0 ? h : 1];
    if (PRE)
      {
#pragma unroll
        for (int i = 0; i < h; ++i)
          {
            e[i] = v[i] + v[n - 1 - i];
            o[i] = v[i] - v[n - 1 - i];
          }
        if (m > h)
          e[h] = v[h];
      }
    else
      {
#pragma unroll
        for (int i = 0; i < m; ++i)
          e[i] = v[i];
#pragma unroll
        for (int i = 0; i < h; ++i)
          o[i] = v[m + i];
      }
    T p[m], q[h > 0 ? h : 1];
#pragma unroll
    for (int a = 0; a < m; ++a)
      p[a] = A.P[a * m] * e[0];
#pragma unroll
    for (int i = 1; i < m; ++i)
#pragma unroll
      for (int a = 0; a < m; ++a)
        p[a] += A.P[a * m + i] * e[i];
    if (h > 0)
      {
#pragma unroll
        for (int a = 0; a < h; ++a)
          q[a] = A.Q[a * h] * o[0];
#pragma unroll
        for (int i = 1; i < h; ++i)
#pragma unroll
          for (int a = 0; a < h; ++a)
            q[a] += A.Q[a * h + i] * o[i];
      }
    if (POST)
      {
#pragma unroll
        for (int a = 0; a < h; ++a)
          {
            if (ADD)
              {
                r[a] += p[a] + q[a];
                r[n - 1 - a] += p[a] - q[a];
              }
            else
              {
                r[a]         = p[a] + q[a];
                r[n - 1 - a] = p[a] - q[a];
              }
          }
        if (m > h)
          {
            if (ADD)
              r[h] += p[h];
            else
              r[h] = p[h];
          }
      }
    else
      {
#pragma unroll
        for (int a = 0; a < m; ++a)
          r[a] = ADD ? r[a] + p[a] : p[a];
#pragma unroll
        for (int a = 0; a < h; ++a)
          r[m + a] = ADD ? r[m + a] + q[a] : q[a];
      }
  }

  // ---- compute side: merge of the x / y neighbour contributions and exclusive store into the output tile -----
  template <int k, typename T>
  __device__ __forceinline__ void
  fast_merge(T (&r)[k + 1][k + 1], const int cx, const int cy)
  {
#pragma unroll
    for (int y = 0; y <= k; ++y)
      {
        const T from = __shfl_up_sync(0xffffffffu, r[y][k], 1);
        if (cx > 0)
          r[y][0] += from;
      }
#pragma unroll
    for (int x = 0; x <= k; ++x)
      {
        const T from = __shfl_up_sync(0xffffffffu, r[k][x], 4);
        if (cy > 0)
          r[0][x] += from;
      }
  }

} // namespace dasm
