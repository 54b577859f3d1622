// Warp-specialised brick kernels (sm_100a) for the regular part of a mesh: full 4 x 4 x 4 bricks whose six
// faces are shared with other bricks, without constrained DoFs, uniform Cartesian geometry (Laplace) or one
// 1-D eigen-decomposition triple for all cells of the brick (FDM).  Everything else is processed by the kernels
// of kernels_brick.cuh in a second launch (the shared-face protocol is order independent).
//
//   block = 64 n compute threads + 64 mover threads (n = k + 1)
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
//                   written exactly once to the output tile: no slot reduction, no shared-memory atomics.
//                   The tile of the NEXT brick is gathered by the compute threads themselves with cp.async (fire and
//                   forget, coalesced: the own DoFs of a brick are one contiguous range) as soon as the current tile
//                   has been read (after phase A), and awaited at the top of the next brick.
//   mover warps     stage the epilogue operands (b, or x and x_old) of a brick in shared memory one brick ahead (1-D bulk
//                   copies of the TMA engine with mbarrier completion; cp.async for unaligned user vectors), and run the fused vector epilogue + coalesced stores / red.add of the PREVIOUS
//                   result from the output tile while the compute warps work on the next brick; also the
//                   pre-initialisation of the next kernel's destination on the brick's shared DoFs.
//                   Hand-over through named barriers (bar.arrive / bar.sync producer-consumer pairs).
//
// Tile layout: point (X, Y, Z) at Z * SZ + Y * TP + SKEW * (Y / k) + X; SKEW / SZ make the plane accesses of a
// half-warp (16 cells (cx, cy)) hit 16 different 8-byte banks (32 lanes / 32 banks for float).
#pragma once
#include "kernels_brick.cuh"

namespace dasm
{
  template <int k, int ES>
  struct FastSkew
  {
    static constexpr int skew = 0, padz = 0;
  };
  template <> struct FastSkew<3, 8> { static constexpr int skew = 5, padz = 0; };
  template <> struct FastSkew<3, 4> { static constexpr int skew = 1, padz = 7; };
  template <> struct FastSkew<4, 8> { static constexpr int skew = 1, padz = 0; };
  template <> struct FastSkew<4, 4> { static constexpr int skew = 1, padz = 7; };
  template <> struct FastSkew<5, 8> { static constexpr int skew = 3, padz = 0; };
  template <> struct FastSkew<5, 4> { static constexpr int skew = 3, padz = 11; };

  template <int k, typename T>
  struct FastGeom
  {
    static constexpr int n      = k + 1;
    static constexpr int TP     = 4 * k + 1;
    static constexpr int SKEW   = FastSkew<k, (int)sizeof(T)>::skew;
    static constexpr int SZ     = TP * TP + 4 * SKEW + FastSkew<k, (int)sizeof(T)>::padz;
    static constexpr int TILE   = (TP * SZ + 3) / 4 * 4;
    static constexpr int NCELLS = 64;
    static constexpr int NCT    = NCELLS * n; // compute threads
    static constexpr int NMT    = 64;         // mover threads
    static constexpr int NT     = NCT + NMT;
    static constexpr int CS     = (n * n * n) | 1;
    static constexpr int NOWN   = 64 * k * k * k;
    static constexpr int NPRIV  = (4 * k - 1) * (4 * k - 1) * (4 * k - 1);
    static constexpr int NFOR   = TP * TP * TP - NOWN;
    static constexpr int NFP    = (NFOR + 3) / 4 * 4;
    static constexpr int NOWNP  = (NOWN + 7) / 8 * 8;
    static constexpr int NPRIVP = (NPRIV + 3) / 4 * 4;
    // mbarrier (16 B) | tile | [out] | X (n_x slots of 64 CS) | operands (n_ops x NPRIVP) | ltab u16[NOWNP] | ftab u16[NFP]
    // (Laplace: n_x = 2 and the output tile aliases the first X slot; FDM: n_x = 1 and a separate output tile)
    static constexpr size_t
    smem_bytes(int n_tiles, int n_x, int n_ops)
    {
      return 16 + (size_t)(n_tiles * TILE + n_x * NCELLS * CS + 4 + n_ops * NPRIVP) * sizeof(T) + 16 +
             (size_t)(NOWNP + NFP) * sizeof(uint16_t);
    }
    __host__ __device__ static constexpr int
    addr(int X, int Y, int Z)
    {
      return Z * SZ + Y * TP + SKEW * (Y / k) + X;
    }
  };

  // tables of the regular brick variant and the list of bricks the kernel processes
  struct FastMaps
  {
    const uint16_t *ltab;         // [NOWN]  tile address of own DoF base + i
    const uint16_t *ftab;         // [NFP]   tile address of the j-th foreign point
    const uint32_t *foreign_gidx; // [brick][NFP] global index of the j-th foreign point
    const uint32_t *brick_ids;    // bricks to process
    int             n;
    long long *     prof;         // optional timing instrumentation (DASM_FAST_PROF): [block][16 iterations][16 events]
    int             dbg;          // timing experiments (DASM_FAST_DBG, results invalid): 1 no compute phases, 2 no private stores,
                                  // 4 no red.add, 8 no gather, 16 no operand staging
  };

  __device__ __forceinline__ void
  fast_prof(const FastMaps &maps, const int iter, const int event, const bool who)
  {
#ifdef DASM_FAST_PROF_BUILD
    if (maps.prof != nullptr && who && iter < 16)
      maps.prof[((size_t)blockIdx.x * 16 + iter) * 16 + event] = clock64();
#else
    (void)maps;
    (void)iter;
    (void)event;
    (void)who;
#endif
  }

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

  // ---- gather of a brick closure into the tile by the compute threads (cp.async, awaited one brick later) ----
  template <int k, typename T>
  struct FastCounts
  {
    static constexpr int NFT = (FastGeom<k, T>::NFOR + FastGeom<k, T>::NCT - 1) / FastGeom<k, T>::NCT; // foreign points per compute thread
    static constexpr int NFM = (FastGeom<k, T>::NFOR + FastGeom<k, T>::NMT - 1) / FastGeom<k, T>::NMT; // ... per mover thread
    static constexpr int NSI = (FastGeom<k, T>::NOWN - FastGeom<k, T>::NPRIV + FastGeom<k, T>::NMT - 1) / FastGeom<k, T>::NMT;
  };

  template <int k, typename T, int NTH, int NF>
  __device__ __forceinline__ void
  fast_load_foreign_idx(uint32_t (&gf)[NF], const FastMaps &maps, const uint32_t brick, const int tid)
  {
    using G             = FastGeom<k, T>;
    const uint32_t *src = maps.foreign_gidx + (size_t)brick * G::NFP;
#pragma unroll
    for (int jj = 0; jj < NF; ++jj)
      {
        const int j = tid + jj * NTH;
        gf[jj]      = (j < G::NFOR) ? ldg_early(src + j) : 0u;
      }
  }

  template <int k, typename T>
  __device__ __forceinline__ void
  fast_gather(T *tile, const uint16_t *ltab, const uint16_t *ftab, const uint32_t (&gf)[FastCounts<k, T>::NFT], const T *__restrict__ src,
              const uint32_t base, const int tid)
  {
    using G    = FastGeom<k, T>;
    const T *s = src + base;
#pragma unroll 4
    for (int i = tid; i < G::NOWN; i += G::NCT)
      cp_async_value(tile + ltab[i], s + i);
#pragma unroll
    for (int jj = 0; jj < FastCounts<k, T>::NFT; ++jj)
      {
        const int j = tid + jj * G::NCT;
        if (j < G::NFOR)
          cp_async_value(tile + ftab[j], src + gf[jj]);
      }
  }

  // ---- mover side ---------------------------------------------------------------------------------------
  // contiguous copy of the epilogue operands on the private DoFs of a brick into shared memory: one bulk copy per
  // operand (TMA engine: no LSU work, completion on the mbarrier) when the global range is 16-byte aligned, else cp.async
  template <int k, typename T>
  __device__ __forceinline__ void
  fast_stage_ops(T *ops0, T *ops1, const Epilogue<T> &epi, const bool need0, const bool need1, const bool bulk, const uint32_t base,
                 const int m, const unsigned mbar)
  {
    using G = FastGeom<k, T>;
    if (bulk)
      {
        if (m == 0 && need0)
          {
            constexpr unsigned bytes = (unsigned)(G::NPRIVP * sizeof(T)); // up to 3 elements beyond the private range: still own DoFs
            mbar_expect_tx(mbar, need1 ? 2 * bytes : bytes);
            bulk_load((unsigned)__cvta_generic_to_shared(ops0), epi.v0 + base, bytes, mbar);
            if (need1)
              bulk_load((unsigned)__cvta_generic_to_shared(ops1), epi.v1 + base, bytes, mbar);
          }
        return;
      }
    constexpr int V = 16 / (int)sizeof(T);
    auto copy       = [&](T *s, const T *g) {
      if ((reinterpret_cast<uintptr_t>(g) & 15) == 0)
        {
          for (int i = m; i < G::NPRIV / V; i += G::NMT)
            {
              const unsigned sa = (unsigned)__cvta_generic_to_shared(s + V * i);
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(g + V * i));
            }
          for (int i = (G::NPRIV / V) * V + m; i < G::NPRIV; i += G::NMT)
            cp_async_value(s + i, g + i);
        }
      else
        {
#pragma unroll 4
          for (int i = m; i < G::NPRIV; i += G::NMT)
            cp_async_value(s + i, g + i);
        }
    };
    if (need0)
      copy(ops0, epi.v0 + base);
    if (need1)
      copy(ops1, epi.v1 + base);
  }

  // pre-initialisation of the next kernel's destination on the brick's own shared DoFs: all loads are issued at the
  // top of a mover iteration, the stores follow after the epilogue stores of the brick (their latency is hidden)
  template <int k, typename T>
  struct FastInitRegs
  {
    T a[FastCounts<k, T>::NSI], b[FastCounts<k, T>::NSI];
  };

  template <int k, typename T>
  __device__ __forceinline__ void
  fast_next_init_load(FastInitRegs<k, T> &r, const NextInit<T> &ni, const uint32_t sh_base, const int m)
  {
    using G           = FastGeom<k, T>;
    constexpr int NSI = FastCounts<k, T>::NSI;
    const bool    h0 = (ni.out != nullptr) && ni.v0 != nullptr, h1 = (ni.out != nullptr) && (ni.v1 != nullptr && ni.f1 != T(0));
#pragma unroll
    for (int it = 0; it < NSI; ++it)
      {
        const int  i  = m + it * G::NMT;
        const bool in = i < G::NOWN - G::NPRIV;
        r.a[it]       = (h0 && in) ? __ldg(ni.v0 + sh_base + i) : T(0);
        r.b[it]       = (h1 && in) ? __ldg(ni.v1 + sh_base + i) : T(0);
      }
  }

  template <int k, typename T>
  __device__ __forceinline__ void
  fast_next_init_store(const FastInitRegs<k, T> &r, const NextInit<T> &ni, const uint32_t sh_base, const int m)
  {
    using G           = FastGeom<k, T>;
    constexpr int NSI = FastCounts<k, T>::NSI;
    if (ni.out == nullptr)
      return;
#pragma unroll
    for (int it = 0; it < NSI; ++it)
      {
        const int i = m + it * G::NMT;
        if (i < G::NOWN - G::NPRIV)
          ni.out[sh_base + i] = r.a[it] + ni.f1 * (r.a[it] - r.b[it]);
      }
  }

  template <int k, typename T>
  __device__ __forceinline__ void
  fast_mover_loop(const T *out, T *ops0, T *ops1, const uint16_t *ltab, const uint16_t *ftab, T *__restrict__ dst, T *__restrict__ acc,
                  const Epilogue<T> &epi, const BrickDesc *__restrict__ bricks, const FastMaps &maps, const int shared_mode,
                  const NextInit<T> &ni, const int m, const unsigned mbar)
  {
    using G           = FastGeom<k, T>;
    constexpr int NFM = FastCounts<k, T>::NFM;
    const bool need0  = (epi.kind == EPI_RESIDUAL || epi.kind == EPI_CHEB);
    const bool need1  = (epi.kind == EPI_CHEB && epi.f1 != T(0) && epi.v1 != nullptr);
    const T    alpha  = (epi.kind == EPI_RESIDUAL) ? T(-1) : ((epi.kind == EPI_CHEB || epi.kind == EPI_SCALE) ? epi.f2 : T(1));
    T *        sh_dst = (shared_mode == SHARED_DIRECT) ? dst : acc;
    const T    sh_a   = (shared_mode == SHARED_DIRECT) ? alpha : T(1);
    const int  G1 = (int)gridDim.x;
    int        it = blockIdx.x;
    // descriptors: current brick, next brick (loaded one iteration ahead), id of the brick after it
    uint32_t bid = ldg_early(maps.brick_ids + it);
    uint32_t bid_n = (it + G1 < maps.n) ? ldg_early(maps.brick_ids + it + G1) : 0u;
    uint32_t base = ldg_early(&bricks[bid].base), sh_base = ldg_early(&bricks[bid].sh_base);
    uint32_t base_n = ldg_early(&bricks[bid_n].base), sh_base_n = ldg_early(&bricks[bid_n].sh_base);
    // bulk staging needs 16-byte aligned global ranges (brick bases are aligned by the host-side eligibility test)
    // (measured: the bulk copies shorten the P-sweep (two operands) by 3.5 %; neutral for the A-sweep)
    const bool bulk  = ((reinterpret_cast<uintptr_t>(epi.v0) | reinterpret_cast<uintptr_t>(epi.v1)) & 15) == 0;
    unsigned   phase = 0;
    fast_stage_ops<k, T>(ops0, ops1, epi, need0, need1, bulk, base, m, mbar);
    for (; it < maps.n; it += G1)
      {
        const bool     has_next = it + G1 < maps.n;
        const uint32_t bid_nn   = (it + 2 * G1 < maps.n) ? ldg_early(maps.brick_ids + it + 2 * G1) : 0u;
        uint32_t       gf[NFM];
        const int      li = (it - (int)blockIdx.x) / G1;
        fast_prof(maps, li, 8, m == 0);
        fast_load_foreign_idx<k, T, G::NMT, NFM>(gf, maps, bid, m);
        FastInitRegs<k, T> nir;
        fast_next_init_load<k, T>(nir, ni, sh_base, m);
        fast_prof(maps, li, 9, m == 0);
        bar_sync(FB_OUT_FULL, G::NT); // the result of this brick is in the output tile
        fast_prof(maps, li, 10, m == 0);
        const uint32_t base_nn = ldg_early(&bricks[bid_nn].base), sh_base_nn = ldg_early(&bricks[bid_nn].sh_base);
        if (bulk)
          {
            if (need0)
              mbar_wait(mbar, phase); // the bulk copies of the operands have landed
            phase ^= 1u;
          }
        else
          {
            cp_async_wait_all();
            bar_sync(FB_MOVERS, G::NMT); // operands staged by all movers are visible
          }
        {
          T *d = dst + base;
          // private DoFs: fused epilogue, coalesced plain stores (one straight-line loop per epilogue kind so that the
          // shared-memory loads of an unrolled batch are in flight together)
          if (maps.dbg & 2)
            {
            }
          else if (epi.kind == EPI_CHEB && need1)
            {
              const T f1 = epi.f1, f2 = epi.f2;
#pragma unroll 16
              for (int i = m; i < G::NPRIV; i += G::NMT)
                {
                  const T a = ops0[i];
                  d[i]      = a + f2 * out[ltab[i]] + f1 * (a - ops1[i]);
                }
            }
          else if (epi.kind == EPI_CHEB)
            {
              const T f1 = epi.f1, f2 = epi.f2;
#pragma unroll 16
              for (int i = m; i < G::NPRIV; i += G::NMT)
                {
                  const T a = ops0[i];
                  d[i]      = a + f2 * out[ltab[i]] + f1 * (a - T(0));
                }
            }
          else if (epi.kind == EPI_RESIDUAL)
            {
#pragma unroll 16
              for (int i = m; i < G::NPRIV; i += G::NMT)
                d[i] = ops0[i] - out[ltab[i]];
            }
          else if (epi.kind == EPI_SCALE)
            {
              const T f2 = epi.f2;
#pragma unroll 16
              for (int i = m; i < G::NPRIV; i += G::NMT)
                d[i] = f2 * out[ltab[i]];
            }
          else
            {
#pragma unroll 16
              for (int i = m; i < G::NPRIV; i += G::NMT)
                d[i] = out[ltab[i]];
            }
          fast_prof(maps, li, 15, m == 0);
          // own DoFs on the lower (shared) faces
          T *sd = sh_dst + base;
          if (!(maps.dbg & 4))
#pragma unroll 4
          for (int i = G::NPRIV + m; i < G::NOWN; i += G::NMT)
            atomic_add(sd + i, sh_a * out[ltab[i]]);
            // points owned by other bricks
#pragma unroll
          for (int jj = 0; jj < NFM; ++jj)
            {
              const int j = m + jj * G::NMT;
              if (j < G::NFOR && !(maps.dbg & 4))
                atomic_add(sh_dst + gf[jj], sh_a * out[ftab[j]]);
            }
        }
        fast_prof(maps, li, 11, m == 0);
        if (has_next)
          bar_arrive(FB_OUT_EMPTY, G::NT);
        fast_next_init_store<k, T>(nir, ni, sh_base, m);
        bar_sync(FB_MOVERS, G::NMT); // all movers have read the operands: stage those of the next brick
        if (has_next && !(maps.dbg & 16))
          fast_stage_ops<k, T>(ops0, ops1, epi, need0, need1, bulk, base_n, m, mbar);
        fast_prof(maps, li, 12, m == 0);
        bid       = bid_n;
        base      = base_n;
        sh_base   = sh_base_n;
        bid_n     = bid_nn;
        base_n    = base_nn;
        sh_base_n = sh_base_nn;
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

  template <int k, typename T>
  __device__ __forceinline__ void
  fast_out_store(const T (&r)[k + 1][k + 1], T *op, const int cx, const int cy)
  {
    using G = FastGeom<k, T>;
#pragma unroll
    for (int y = 0; y <= k; ++y)
#pragma unroll
      for (int x = 0; x <= k; ++x)
        {
          const bool w = (x < k || cx == 3) && (y < k || cy == 3);
          if (w)
            op[y * G::TP + (y == k ? G::SKEW : 0) + x] = r[y][x];
        }
  }

  // ---- Laplace, uniform Cartesian geometry ---------------------------------------------------------------------
  template <int k, typename T>
  __global__ void __launch_bounds__(FastGeom<k, T>::NT, 1)
  laplace_fast_kernel(const T *__restrict__ src,
                      T *__restrict__ dst,
                      T *__restrict__ acc,
                      const Epilogue<T> epi,
                      const BrickDesc *__restrict__ bricks,
                      const __grid_constant__ FastLaplaceMats<T, k + 1> mats,
                      const int         shared_mode,
                      const NextInit<T> ni,
                      const FastMaps    maps)
  {
    using G           = FastGeom<k, T>;
    constexpr int n   = k + 1;
    constexpr int NFT = FastCounts<k, T>::NFT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *       tile = reinterpret_cast<T *>(smem_raw + 16); // the first 16 bytes hold the mbarrier of the operand staging
    T *       Xq   = tile + G::TILE;
    T *       Xp   = Xq + G::NCELLS * G::CS;
    T *       out  = Xq; // the output tile aliases the first exchange slot (written after all reads of it)
    T *       ops0 = Xp + G::NCELLS * G::CS; // 64 CS elements per slot: 16-byte aligned
    uint16_t *ltab = reinterpret_cast<uint16_t *>(ops0 + G::NPRIVP);
    uint16_t *ftab = ltab + G::NOWNP;
    if ((int)blockIdx.x >= maps.n)
      return;
    for (int i = threadIdx.x; i < G::NOWN; i += G::NT)
      ltab[i] = maps.ltab[i];
    for (int i = threadIdx.x; i < G::NFP; i += G::NT)
      ftab[i] = maps.ftab[i];
    if (threadIdx.x == 0)
      mbar_init((unsigned)__cvta_generic_to_shared(smem_raw), 1);
    __syncthreads();

    if (threadIdx.x >= G::NCT)
      {
        fast_mover_loop<k, T>(out, ops0, ops0, ltab, ftab, dst, acc, epi, bricks, maps, shared_mode, ni, threadIdx.x - G::NCT,
                              (unsigned)__cvta_generic_to_shared(smem_raw));
        return;
      }
    const int  tid = threadIdx.x;
    const int  c = tid % G::NCELLS, t = tid / G::NCELLS;
    const int  cx = c & 3, cy = (c >> 2) & 3, cz = c >> 4;
    const bool skip_last = ((t == k) && (cz < 2)) || (maps.dbg & 1); // whole warp: its plane z = k belongs to the cell above
    const T *  tp = tile + G::addr(k * cx, k * cy, k * cz) + t * G::TP + (t == k ? G::SKEW : 0); // plane y = t of the cell
    T *        xq = Xq + c * G::CS, *xp = Xp + c * G::CS;
    T *        op = out + G::addr(k * cx, k * cy, k * cz + t);
    const int  G1 = (int)gridDim.x;
    int        it = blockIdx.x;
    // descriptors: next brick (loaded one iteration ahead), id of the brick after it
    uint32_t bid_next  = (it + G1 < maps.n) ? ldg_early(maps.brick_ids + it + G1) : 0u;
    uint32_t base_next = 0;
    {
      const uint32_t bid = ldg_early(maps.brick_ids + it);
      uint32_t       gf[NFT];
      fast_load_foreign_idx<k, T, G::NCT, NFT>(gf, maps, bid, tid);
      base_next = ldg_early(&bricks[bid_next].base);
      fast_gather<k, T>(tile, ltab, ftab, gf, src, ldg_early(&bricks[bid].base), tid);
    }
    bool first = true;
    for (; it < maps.n; it += G1)
      {
        const bool     has_next = it + G1 < maps.n;
        const uint32_t bid_nn   = (it + 2 * G1 < maps.n) ? ldg_early(maps.brick_ids + it + 2 * G1) : 0u;
        uint32_t       gfn[NFT];
        fast_load_foreign_idx<k, T, G::NCT, NFT>(gfn, maps, bid_next, tid);
        const int li = (it - (int)blockIdx.x) / (int)gridDim.x;
        fast_prof(maps, li, 0, tid == 0);
        cp_async_wait_all();
        bar_sync(FB_COMPUTE, G::NCT); // the tile of this brick has landed
        fast_prof(maps, li, 1, tid == 0);
        // phase A: plane y = t, [z][x]: q = Mx Mz v, p = (g0 Kx Mz + g2 Mx Kz) v
        if (!(maps.dbg & 1))
        {
          T a[n][n], b[n][n];
#pragma unroll
          for (int z = 0; z < n; ++z)
            {
              T v[n];
#pragma unroll
              for (int x = 0; x < n; ++x)
                v[x] = tp[z * G::SZ + x];
              mat_vec<n, T, true, true, false>(a[z], mats.M, v);
              mat_vec<n, T, true, true, false>(b[z], mats.K0, v);
            }
          fast_prof(maps, li, 2, tid == 0);
          if (!first)
            bar_sync(FB_OUT_EMPTY, G::NT); // the previous result (aliased with Xq) has been stored
          fast_prof(maps, li, 3, tid == 0);
#pragma unroll
          for (int x = 0; x < n; ++x)
            {
              T ca[n], cb[n], q[n], p[n];
#pragma unroll
              for (int z = 0; z < n; ++z)
                {
                  ca[z] = a[z][x];
                  cb[z] = b[z][x];
                }
              mat_vec<n, T, true, true, false>(q, mats.M, ca);
              mat_vec<n, T, true, true, false>(p, mats.M, cb);
              mat_vec<n, T, true, true, true>(p, mats.K2, ca);
#pragma unroll
              for (int z = 0; z < n; ++z)
                {
                  xq[(z * n + t) * n + x] = q[z];
                  xp[(z * n + t) * n + x] = p[z];
                }
            }
        }
        fast_prof(maps, li, 4, tid == 0);
        if ((maps.dbg & 1) && !first)
          bar_sync(FB_OUT_EMPTY, G::NT);
        bar_sync(FB_COMPUTE, G::NCT);
        fast_prof(maps, li, 5, tid == 0);
        // the tile is dead: gather the next brick into it
        if (has_next && !(maps.dbg & 8))
          fast_gather<k, T>(tile, ltab, ftab, gfn, src, base_next, tid);
        bid_next  = bid_nn;
        base_next = ldg_early(&bricks[bid_nn].base);
        fast_prof(maps, li, 6, tid == 0);
        // phase B: plane z = t, [y][x]: r = My p + g1 Ky q  (+ plane z = k of the cell below for t = 0)
        T r[n][n];
        if (!skip_last)
          {
            const bool below = (t == 0) && (cz > 0);
#pragma unroll
            for (int x = 0; x < n; ++x)
              {
                T qi[n], pi[n], rc[n];
#pragma unroll
                for (int i = 0; i < n; ++i)
                  {
                    qi[i] = xq[(t * n + i) * n + x];
                    pi[i] = xp[(t * n + i) * n + x];
                  }
                if (t == 0)
                  {
#pragma unroll
                    for (int i = 0; i < n; ++i)
                      {
                        qi[i] += below ? xq[-16 * G::CS + (k * n + i) * n + x] : T(0);
                        pi[i] += below ? xp[-16 * G::CS + (k * n + i) * n + x] : T(0);
                      }
                  }
                mat_vec<n, T, true, true, false>(rc, mats.M, pi);
                mat_vec<n, T, true, true, true>(rc, mats.K1, qi);
#pragma unroll
                for (int y = 0; y < n; ++y)
                  r[y][x] = rc[y];
              }
            fast_merge<k, T>(r, cx, cy);
          }
        fast_prof(maps, li, 7, tid == 0);
        bar_sync(FB_COMPUTE, G::NCT); // all reads of the exchange slots are done: the output tile may overwrite Xq
        fast_prof(maps, li, 13, tid == 0);
        if (!skip_last && (t < k || cz == 3))
          fast_out_store<k, T>(r, op, cx, cy);
        bar_arrive(FB_OUT_FULL, G::NT);
        fast_prof(maps, li, 14, tid == 0);
        first = false;
      }
  }

  // ---- FDM, one eigen-decomposition triple, tensor-product weights folded into the matrices ------------------
  template <int k, typename T>
  __global__ void __launch_bounds__(FastGeom<k, T>::NT, 1)
  fdm_fast_kernel(const T *__restrict__ src,
                  T *__restrict__ dst,
                  T *__restrict__ acc,
                  const Epilogue<T> epi,
                  const BrickDesc *__restrict__ bricks,
                  const __grid_constant__ FastFdmMats<T, k + 1> mats,
                  const int         shared_mode,
                  const NextInit<T> ni,
                  const FastMaps    maps)
  {
    using G           = FastGeom<k, T>;
    constexpr int n   = k + 1;
    constexpr int NFT = FastCounts<k, T>::NFT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *       tile = reinterpret_cast<T *>(smem_raw + 16); // the first 16 bytes hold the mbarrier of the operand staging
    T *       out  = tile + G::TILE;
    T *       X    = out + G::TILE;
    T *       ops0 = X + G::NCELLS * G::CS; // 64 CS elements per slot: 16-byte aligned
    T *       ops1 = ops0 + G::NPRIVP;
    uint16_t *ltab = reinterpret_cast<uint16_t *>(ops1 + G::NPRIVP);
    uint16_t *ftab = ltab + G::NOWNP;
    __shared__ T s_inv[n * n * n];
    if ((int)blockIdx.x >= maps.n)
      return;
    for (int i = threadIdx.x; i < G::NOWN; i += G::NT)
      ltab[i] = maps.ltab[i];
    for (int i = threadIdx.x; i < G::NFP; i += G::NT)
      ftab[i] = maps.ftab[i];
    for (int i = threadIdx.x; i < n * n * n; i += G::NT)
      s_inv[i] = mats.inv[i];
    if (threadIdx.x == 0)
      mbar_init((unsigned)__cvta_generic_to_shared(smem_raw), 1);
    __syncthreads();

    if (threadIdx.x >= G::NCT)
      {
        fast_mover_loop<k, T>(out, ops0, ops1, ltab, ftab, dst, acc, epi, bricks, maps, shared_mode, ni, threadIdx.x - G::NCT,
                              (unsigned)__cvta_generic_to_shared(smem_raw));
        return;
      }
    const int  tid = threadIdx.x;
    const int  c = tid % G::NCELLS, t = tid / G::NCELLS;
    const int  cx = c & 3, cy = (c >> 2) & 3, cz = c >> 4;
    const bool skip_last = ((t == k) && (cz < 2)) || (maps.dbg & 1);
    const T *  tp = tile + G::addr(k * cx, k * cy, k * cz + t); // plane z = t of the cell
    T *        xs = X + c * G::CS;
    T *        op = out + G::addr(k * cx, k * cy, k * cz + t);
    // inverse eigenvalue sums, row (z, y = t) of this thread's plane in phase B: broadcast reads from shared memory
    const T *inv = s_inv + t * n;
    const int  G1 = (int)gridDim.x;
    int        it = blockIdx.x;
    // descriptors: next brick (loaded one iteration ahead), id of the brick after it
    uint32_t bid_next  = (it + G1 < maps.n) ? ldg_early(maps.brick_ids + it + G1) : 0u;
    uint32_t base_next = 0;
    {
      const uint32_t bid = ldg_early(maps.brick_ids + it);
      uint32_t       gf[NFT];
      fast_load_foreign_idx<k, T, G::NCT, NFT>(gf, maps, bid, tid);
      base_next = ldg_early(&bricks[bid_next].base);
      fast_gather<k, T>(tile, ltab, ftab, gf, src, ldg_early(&bricks[bid].base), tid);
    }
    bool first = true;
    for (; it < maps.n; it += G1)
      {
        const bool     has_next = it + G1 < maps.n;
        const uint32_t bid_nn   = (it + 2 * G1 < maps.n) ? ldg_early(maps.brick_ids + it + 2 * G1) : 0u;
        uint32_t       gfn[NFT];
        fast_load_foreign_idx<k, T, G::NCT, NFT>(gfn, maps, bid_next, tid);
        const int li = (it - (int)blockIdx.x) / (int)gridDim.x;
        fast_prof(maps, li, 0, tid == 0);
        cp_async_wait_all();
        bar_sync(FB_COMPUTE, G::NCT); // the tile of this brick has landed
        fast_prof(maps, li, 1, tid == 0);
        // phase A: plane z = t, [y][x]: Ax in x, Ay in y
        if (!(maps.dbg & 1))
        {
          T a[n][n];
#pragma unroll
          for (int y = 0; y < n; ++y)
            {
              T v[n];
#pragma unroll
              for (int x = 0; x < n; ++x)
                v[x] = tp[y * G::TP + (y == k ? G::SKEW : 0) + x];
              mat_vec<n, T, true, false, false>(a[y], mats.Ax, v);
            }
#pragma unroll
          for (int x = 0; x < n; ++x)
            {
              T ca[n], q[n];
#pragma unroll
              for (int y = 0; y < n; ++y)
                ca[y] = a[y][x];
              mat_vec<n, T, true, false, false>(q, mats.Ay, ca);
#pragma unroll
              for (int y = 0; y < n; ++y)
                xs[(t * n + y) * n + x] = q[y];
            }
        }
        fast_prof(maps, li, 2, tid == 0);
        bar_sync(FB_COMPUTE, G::NCT);
        fast_prof(maps, li, 3, tid == 0);
        // the tile is dead: gather the next brick into it
        if (has_next && !(maps.dbg & 8))
          fast_gather<k, T>(tile, ltab, ftab, gfn, src, base_next, tid);
        bid_next  = bid_nn;
        base_next = ldg_early(&bricks[bid_nn].base);
        fast_prof(maps, li, 4, tid == 0);
        // phase B: plane y = t, [z][x]: Az, scale, Bz in z; Bx in x
        if (!(maps.dbg & 1))
        {
          T w[n][n];
#pragma unroll
          for (int x = 0; x < n; ++x)
            {
              T col[n], u[n];
#pragma unroll
              for (int z = 0; z < n; ++z)
                col[z] = xs[(z * n + t) * n + x];
              mat_vec<n, T, true, false, false>(u, mats.Az, col);
#pragma unroll
              for (int z = 0; z < n; ++z)
                u[z] *= inv[z * n * n + x];
              mat_vec<n, T, false, true, false>(col, mats.Bz, u);
#pragma unroll
              for (int z = 0; z < n; ++z)
                w[z][x] = col[z];
            }
#pragma unroll
          for (int z = 0; z < n; ++z)
            {
              T u[n];
              mat_vec<n, T, false, true, false>(u, mats.Bx, w[z]);
#pragma unroll
              for (int x = 0; x < n; ++x)
                xs[(z * n + t) * n + x] = u[x];
            }
        }
        fast_prof(maps, li, 5, tid == 0);
        bar_sync(FB_COMPUTE, G::NCT);
        fast_prof(maps, li, 6, tid == 0);
        // phase C: plane z = t, [y][x]: By in y (+ plane z = k of the cell below for t = 0)
        T r[n][n];
        if (!skip_last)
          {
            const bool below = (t == 0) && (cz > 0);
#pragma unroll
            for (int x = 0; x < n; ++x)
              {
                T vi[n], rc[n];
#pragma unroll
                for (int i = 0; i < n; ++i)
                  vi[i] = xs[(t * n + i) * n + x];
                if (t == 0)
                  {
#pragma unroll
                    for (int i = 0; i < n; ++i)
                      vi[i] += below ? xs[-16 * G::CS + (k * n + i) * n + x] : T(0);
                  }
                mat_vec<n, T, false, true, false>(rc, mats.By, vi);
#pragma unroll
                for (int y = 0; y < n; ++y)
                  r[y][x] = rc[y];
              }
            fast_merge<k, T>(r, cx, cy);
          }
        fast_prof(maps, li, 7, tid == 0);
        if (!first)
          bar_sync(FB_OUT_EMPTY, G::NT); // the previous result has been stored
        fast_prof(maps, li, 13, tid == 0);
        if (!skip_last && (t < k || cz == 3))
          fast_out_store<k, T>(r, op, cx, cy);
        bar_arrive(FB_OUT_FULL, G::NT);
        fast_prof(maps, li, 14, tid == 0);
        first = false;
      }
  }
} // namespace dasm
