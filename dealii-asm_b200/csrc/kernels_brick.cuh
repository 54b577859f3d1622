// Tuned "brick" kernels (sm_100a): one thread block per brick of 4 x 4 x BZ cells, persistent over bricks.
//
// Data flow per brick:
//   1. tile load    the (4k+1)^2 (BZ k+1) closure of the brick is gathered ONCE from HBM into shared
//                   memory (global index of every tile point decoded from the 27 compressed indices of its
//                   canonical cell and kept in a shared u32 array for the store phase)
//   2. cell phases  n = k+1 threads per cell; every thread owns one n x n PLANE of the cell in registers
//                   and performs the 1-D contractions of two directions per phase with the n x n
//                   matrices as constant-bank operands of the FMAs; planes are exchanged between
//                   phases through a single n^3 shared-memory slot per cell (in place)
//   3. reduction    every tile point sums the <= 8 cell contributions from the slots in a fixed order
//                   (deterministic, no shared-memory atomics); points touched only by cells of this brick
//                   ("private") run the fused vector epilogue (residual / Chebyshev update) and are
//                   written with plain stores; points on faces shared with other bricks are added to a
//                   zero-invariant accumulator with red.global.add and finished by finalize_shared_kernel
//
// This replaces the reference's "first/last touched" DoF-range scheduling of the pre/post hooks
// (include/matrix_free.h:420-532, include/matrix_free_internal.h:309-359), which assumes a sequential
// cell order.
#pragma once
#include "kernels.cuh"

namespace dasm
{
  struct BrickDesc
  {
    uint32_t first_cell; // processing index of the first cell (cells of a brick are consecutive, x fastest)
    uint8_t  b[3];       // cells per direction
    uint8_t  shared;     // bit (2 d + side): the tile face is shared with cells outside the brick; bit 6: BRICK_LEX
    uint32_t sh_base;    // first DoF of the contiguous range of shared-face DoFs this brick owns
    uint32_t sh_count;   // (only meaningful when the kernel brick is a whole mesh brick)
    uint32_t base;       // first DoF owned by the brick: [base, base + npriv) private, then the shared range
    uint16_t npriv;      // number of private DoFs
    uint16_t variant;    // index of the brick's tile-map variant (BrickMaps), 0xFFFF: none
  };

  // Per-variant maps between the brick's contiguous range of owned DoFs and its tile (bricks of equal shape,
  // boundary flags and constraints share a variant) + per-brick global indices of the tile points owned by
  // other bricks.
  //   load_tab[i]   tile point of own DoF base + i (0xFFFF: constrained / unused)
  //   store_tab[i]  for own DoF base + i: bits 13-25 slot offset of the primary cell contribution, 26-28 mask of
  //                 the directions with a second contribution (lower neighbour cell inside the brick);
  //                 0xFFFFFFFF unused (store_off: reserved)
  //   for_tab[j]    tile points owned by other bricks sorted by mask: bits 0-12 tile point, 13-25 slot offset,
  //                 26-28 mask; for_off[m] class offsets (8 + end); foreign_gidx[brick][j] their global index
  struct BrickMaps
  {
    const uint16_t *load_tab;     // [variant][stride]
    const uint32_t *store_tab;    // [variant][stride]
    const uint32_t *store_off;    // [variant][17]
    const uint32_t *for_tab;      // [variant][nfp]
    const uint32_t *for_off;      // [variant][9]
    const uint32_t *flags;        // [variant] bit 0: the tile has unreferenced (constrained) points -> zero it first
    const uint32_t *foreign_gidx; // [brick][nfp]
    int             stride;
    int             nfp;
    const uint16_t *sh_tab;       // [variant][sh_stride] offsets of the own DoFs on shared lower faces
    const uint32_t *sh_cnt;       // [variant]
    int             sh_stride;
  };

  constexpr uint8_t  BRICK_LEX    = 0x40u;       // BrickDesc::shared: the brick's own DoFs are the lexicographic box [0, 4k)^3 (mesh.h)
  constexpr uint32_t STORE_SHARED = 0x20000000u; // store_tab entry: own DoF on a shared lower face (red.add instead of a store)

  // compressed weights as codes: code = patch valence of the entity (0: weight 0), value from this table
  template <typename T>
  struct WeightTable
  {
    T v[16];
  };

  // how contributions to DoFs on shared brick faces are combined
  enum
  {
    SHARED_ACC    = 0, // red.add y into the zero-invariant accumulator, finish_shared_kernel runs the epilogue
    SHARED_DIRECT = 1  // dst is ZERO on the shared DoFs before the kernel (zeroed by the previous kernel of the sequence); the brick
                       // that owns a shared DoF adds the full epilogue value (base + alpha y), every other brick alpha y
  };

  // the NEXT kernel's destination is zeroed on the shared DoFs owned by a brick (fused into the current kernel): `out`.
  // (v0 / v1 / f1 are only used by the stand-alone init_shared_kernel.)
  template <typename T>
  struct NextInit
  {
    T *      out;
    const T *v0;
    const T *v1;
    T        f1;
  };

  enum
  {
    EPI_STORE    = 0, // dst = y
    EPI_RESIDUAL = 1, // dst = v0 - y
    EPI_CHEB     = 2, // dst = v0 + f1 (v0 - v1) + f2 y      (v1 == nullptr: v1 = 0)
    EPI_SCALE    = 3  // dst = f2 y
  };

  template <typename T>
  struct Epilogue
  {
    int      kind;
    T        f1, f2;
    const T *v0;
    const T *v1;
  };

  // operands of the epilogue at DoF g (loaded separately so that a batch of loads can be in flight)
  template <typename T>
  __device__ __forceinline__ void
  epilogue_load(const Epilogue<T> &e, const uint32_t g, T &a, T &b)
  {
    a = T(0);
    b = T(0);
    if (e.kind == EPI_RESIDUAL || e.kind == EPI_CHEB)
      a = e.v0[g];
    if (e.kind == EPI_CHEB && e.f1 != T(0) && e.v1 != nullptr)
      b = e.v1[g];
  }

  template <typename T>
  __device__ __forceinline__ T
  epilogue_compute(const Epilogue<T> &e, const T y, const T a, const T b)
  {
    if (e.kind == EPI_RESIDUAL)
      return a - y;
    if (e.kind == EPI_CHEB)
      return a + e.f2 * y + e.f1 * (a - b);
    if (e.kind == EPI_SCALE)
      return e.f2 * y;
    return y;
  }

  template <typename T>
  __device__ __forceinline__ T
  epilogue_apply(const Epilogue<T> &e, const T y, const uint32_t g)
  {
    T a, b;
    epilogue_load(e, g, a, b);
    return epilogue_compute(e, y, a, b);
  }

  template <int k, int BZ>
  struct BrickGeom
  {
    static constexpr int n      = k + 1;
    static constexpr int BX     = 4;
    static constexpr int BY     = 4;
    static constexpr int TX     = BX * k + 1;
    static constexpr int TY     = BY * k + 1;
    static constexpr int TZ     = BZ * k + 1;
    static constexpr int NPTS   = TX * TY * TZ;
    static constexpr int NCELLS = BX * BY * BZ;
    static constexpr int NT     = NCELLS * n;   // threads per block
    static constexpr int MINB   = (NT <= 160) ? 2 : 1; // resident blocks per SM the register budget is set for
    static constexpr int CS     = (n * n * n) | 1; // slot stride per cell, odd: with cell-major lanes (lane = cell) every
                                                   // plane access of a half-warp hits 16 different banks
    // registers per thread with one resident block per SM (64 K registers, allocation granularity 8)
    // (registers are allocated per warp in units of 512)
    static constexpr int NWARPS = (NT + 31) / 32;
    static constexpr int MAXREG = ((65536 / NWARPS / 512) * 512 / 32) >= 255 ? 255 : ((65536 / NWARPS / 512) * 512 / 32);
    static constexpr int NFOREIGN = NPTS - (BX * k) * (BY * k) * (BZ * k); // tile points a brick can not own
    static constexpr int NFP      = (NFOREIGN + 3) / 4 * 4;                 // padded to 16 bytes
    static constexpr int NLT      = (NPTS + 1) / 2 * 2;                     // u16 load table padded to 4 bytes
    // non-lin layout: tile | operand tiles (n_ops) | slots | gidx[NPTS] | cidx[2]
    // lin layout    : tile | operand tiles (n_ops) | slots | gidx_f[3][NFP] | store_tab[NPTS] | for_tab[NFP] | offsets[32] |
    //                 load_tab u16[NLT]
    template <typename T>
    static constexpr size_t
    smem_bytes(int n_ops, bool lin)
    {
      return (size_t)(1 + n_ops) * NPTS * sizeof(T) + (size_t)NCELLS * CS * sizeof(T) +
             16 + (lin ? (size_t)(3 * NFP + NPTS + NFP + 32) * sizeof(uint32_t) + (size_t)NLT * sizeof(uint16_t) :
                    (size_t)NPTS * sizeof(uint32_t) + (size_t)2 * NCELLS * 27 * sizeof(uint32_t));
    }
  };

  // ---- register-plane contractions --------------------------------------------------------------
  // All loops are written input-major (for i: for o: r[o] += M[o][i] v[i]) so that n independent FMA
  // chains are interleaved in program order (the FP64 pipe needs >= 4 independent FMAs in flight per
  // warp at 10 warps/SM, see tools/fp64_peak.cu).
  // v[a][b]: apply M along b (fast index): v[a][:] = M v[a][:]   (TRANS: M^T)
  template <int n, typename T, bool TRANS>
  __device__ __forceinline__ void
  apply_fast(T (&v)[n][n], const T *M)
  {
#pragma unroll
    for (int a = 0; a < n; ++a)
      {
        T r[n];
#pragma unroll
        for (int o = 0; o < n; ++o)
          r[o] = (TRANS ? M[o] : M[o * n]) * v[a][0];
#pragma unroll
        for (int i = 1; i < n; ++i)
#pragma unroll
          for (int o = 0; o < n; ++o)
            r[o] += (TRANS ? M[i * n + o] : M[o * n + i]) * v[a][i];
#pragma unroll
        for (int o = 0; o < n; ++o)
          v[a][o] = r[o];
      }
  }

  // apply M along a (slow index): v[:][b] = M v[:][b]
  template <int n, typename T, bool TRANS>
  __device__ __forceinline__ void
  apply_slow(T (&v)[n][n], const T *M)
  {
#pragma unroll
    for (int b = 0; b < n; ++b)
      {
        T r[n];
#pragma unroll
        for (int o = 0; o < n; ++o)
          r[o] = (TRANS ? M[o] : M[o * n]) * v[0][b];
#pragma unroll
        for (int i = 1; i < n; ++i)
#pragma unroll
          for (int o = 0; o < n; ++o)
            r[o] += (TRANS ? M[i * n + o] : M[o * n + i]) * v[i][b];
#pragma unroll
        for (int o = 0; o < n; ++o)
          v[o][b] = r[o];
      }
  }

  // r[a][:] += D^T diag(c * w[:]) D v[a][:]   for every a (derivative along the fast index)
  template <int n, typename T>
  __device__ __forceinline__ void
  laplace_1d_fast(T (&r)[n][n], const T (&v)[n][n], const T *D, const T *wfast, const T (&wslow)[n])
  {
#pragma unroll
    for (int a = 0; a < n; ++a)
      {
        T f[n];
#pragma unroll
        for (int q = 0; q < n; ++q)
          f[q] = D[q * n] * v[a][0];
#pragma unroll
        for (int i = 1; i < n; ++i)
#pragma unroll
          for (int q = 0; q < n; ++q)
            f[q] += D[q * n + i] * v[a][i];
#pragma unroll
        for (int q = 0; q < n; ++q)
          f[q] *= (wfast[q] * wslow[a]);
#pragma unroll
        for (int q = 0; q < n; ++q)
#pragma unroll
          for (int o = 0; o < n; ++o)
            r[a][o] += D[q * n + o] * f[q];
      }
  }

  // r[:][b] += D^T diag(c * w) D v[:][b]   for every b (derivative along the slow index)
  template <int n, typename T>
  __device__ __forceinline__ void
  laplace_1d_slow(T (&r)[n][n], const T (&v)[n][n], const T *D, const T *wfast, const T (&wslow)[n])
  {
#pragma unroll
    for (int b = 0; b < n; ++b)
      {
        T f[n];
#pragma unroll
        for (int q = 0; q < n; ++q)
          f[q] = D[q * n] * v[0][b];
#pragma unroll
        for (int i = 1; i < n; ++i)
#pragma unroll
          for (int q = 0; q < n; ++q)
            f[q] += D[q * n + i] * v[i][b];
#pragma unroll
        for (int q = 0; q < n; ++q)
          f[q] *= (wfast[b] * wslow[q]);
#pragma unroll
        for (int q = 0; q < n; ++q)
#pragma unroll
          for (int o = 0; o < n; ++o)
            r[o][b] += D[q * n + o] * f[q];
      }
  }

  // ---- tile helpers ----------------------------------------------------------------------------------
  // Thread -> tile point mapping: every thread owns "pencils" (px, py) of the tile and walks along pz, so
  // that everything depending on (px, py) (canonical cell, entity codes, offsets, slot addresses) is
  // computed once and the per-point work is a handful of integer instructions.  The 27 compressed indices
  // of the brick's cells are staged in shared memory by one contiguous copy (the cells of a brick are
  // consecutive) that is issued one brick ahead.  All global->shared traffic uses cp.async (LDGSTS): the
  // source values are awaited before the first cell phase, the epilogue operands only before the store
  // phase, so their latency hides behind the sum factorisation.
  __device__ __forceinline__ void
  cp_async_4(void *smem, const void *gmem)
  {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
  }
  __device__ __forceinline__ void
  cp_async_8(void *smem, const void *gmem)
  {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
  }
  template <typename T>
  __device__ __forceinline__ void
  cp_async_value(T *smem, const T *gmem)
  {
    if (sizeof(T) == 8)
      cp_async_8(smem, gmem);
    else
      cp_async_4(smem, gmem);
  }
  __device__ __forceinline__ void
  cp_async_commit()
  {
    asm volatile("cp.async.commit_group;\n" ::);
  }
  template <int N>
  __device__ __forceinline__ void
  cp_async_wait()
  {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
  }

  template <int k, int BZ>
  __device__ __forceinline__ void
  brick_stage_cidx_async(const BrickDesc &bd, const uint32_t *__restrict__ cidx, uint32_t *s_cidx)
  {
    using G             = BrickGeom<k, BZ>;
    const int       n   = bd.b[0] * bd.b[1] * bd.b[2] * 27;
    const uint32_t *src = cidx + (size_t)bd.first_cell * 27;
    for (int i = threadIdx.x; i < n; i += G::NT)
      cp_async_4(s_cidx + i, src + i);
  }

  // gather of the brick closure: global index of every tile point (kept in gidx), source values -> tile
  // (commit group 1), epilogue operands at private points -> operand tiles (commit group 2)
  template <int k, int BZ, typename T>
  __device__ __forceinline__ void
  brick_issue_loads(const BrickDesc &bd, const uint32_t *s_cidx, T *tile, T *ops0, T *ops1, uint32_t *gidx,
                    const T *__restrict__ src, const T *__restrict__ v0, const T *__restrict__ v1)
  {
    using G             = BrickGeom<k, BZ>;
    constexpr int NPENC = G::TX * G::TY;
    constexpr int PITER = (NPENC + G::NT - 1) / G::NT;
    const int     ex = bd.b[0] * k + 1, ey = bd.b[1] * k + 1, ez = bd.b[2] * k + 1;
#pragma unroll 1
    for (int pi = 0; pi < PITER; ++pi)
      {
        const int q  = threadIdx.x + pi * G::NT;
        const int px = q % G::TX, py = q / G::TX;
        if (q >= NPENC || px >= ex || py >= ey)
          continue;
        const int cx = min(px / k, bd.b[0] - 1), cy = min(py / k, bd.b[1] - 1);
        const int lx = px - cx * k, ly = py - cy * k;
        const int ecx = (lx == 0) ? 0 : ((lx == k) ? 2 : 1), ecy = (ly == 0) ? 0 : ((ly == k) ? 2 : 1);
        const int ox = (ecx == 1) ? lx - 1 : 0, oy = (ecy == 1) ? ly - 1 : 0;
        const int sx = (ecx == 1) ? (k - 1) : 1, sy = (ecy == 1) ? (k - 1) : 1;
        const int cxy = cy * bd.b[0] + cx, exy = ecx + 3 * ecy, oxy = ox + sx * oy, sxy = sx * sy;
        const int czs = bd.b[0] * bd.b[1];
        const int base = py * G::TX + px;
#pragma unroll
        for (int pz = 0; pz < G::TZ; ++pz)
          {
            const int cz = min(pz / k, bd.b[2] - 1), lz = pz - cz * k;
            const int ecz = (lz == 0) ? 0 : ((lz == k) ? 2 : 1), oz = (ecz == 1) ? lz - 1 : 0;
            if (pz < ez)
              {
                const uint32_t st = s_cidx[(cz * czs + cxy) * 27 + exy + 9 * ecz];
                const uint32_t g  = (st == DEV_INVALID) ? DEV_INVALID :
                                    ((st & DEV_LEX_FLAG) ? (st & ~DEV_LEX_FLAG) + ox + 4 * k * (oy + 4 * k * oz) : st + oxy + sxy * oz);
                const int      p  = pz * NPENC + base;
                gidx[p]           = g;
                if (g == DEV_INVALID)
                  tile[p] = T(0);
                else
                  cp_async_value(tile + p, src + g);
              }
          }
      }
  }

  // epilogue operands (v0, v1 of the epilogue) at the private points of the tile -> operand tiles
  template <int k, int BZ, typename T>
  __device__ __forceinline__ void
  brick_issue_loads_ops(const BrickDesc &bd, const uint32_t *gidx, T *ops0, T *ops1, const Epilogue<T> &epi)
  {
    using G             = BrickGeom<k, BZ>;
    constexpr int NPENC = G::TX * G::TY;
    constexpr int PITER = (NPENC + G::NT - 1) / G::NT;
    const bool    need0 = (epi.kind == EPI_RESIDUAL || epi.kind == EPI_CHEB);
    const bool    need1 = (epi.kind == EPI_CHEB && epi.f1 != T(0) && epi.v1 != nullptr);
    if (!need0)
      return;
    const int ex = bd.b[0] * k + 1, ey = bd.b[1] * k + 1, ez = bd.b[2] * k + 1;
#pragma unroll 1
    for (int pi = 0; pi < PITER; ++pi)
      {
        const int q  = threadIdx.x + pi * G::NT;
        const int px = q % G::TX, py = q / G::TX;
        if (q >= NPENC || px >= ex || py >= ey)
          continue;
        const unsigned shxy = ((px == 0) ? (bd.shared & 1u) : 0u) | ((px == ex - 1) ? (bd.shared & 2u) : 0u) |
                              ((py == 0) ? (bd.shared & 4u) : 0u) | ((py == ey - 1) ? (bd.shared & 8u) : 0u);
        const int base = py * G::TX + px;
#pragma unroll
        for (int pz = 0; pz < G::TZ; ++pz)
          if (pz < ez)
            {
              const int      p  = pz * NPENC + base;
              const uint32_t g  = gidx[p]; // written by this thread in brick_issue_loads
              // every point the brick owns (not on an upper face shared with the neighbour brick)
              const bool     up = ((shxy & 10u) | ((pz == ez - 1) ? (bd.shared & 32u) : 0u)) != 0;
              if (g != DEV_INVALID && !up)
                {
                  cp_async_value(ops0 + p, epi.v0 + g);
                  if (need1)
                    cp_async_value(ops1 + p, epi.v1 + g);
                }
            }
      }
  }

  // ---- linear (coalesced) gather / store through the per-variant maps ------------------------------------------
  // The brick's own DoFs are one contiguous range, so consecutive lanes read / write consecutive global
  // addresses (2 lines per warp instruction instead of ~12 with the tile-ordered access); only the tile points
  // owned by other bricks (upper faces) are addressed through the compressed indices.
  // staged tables of the current variant (shared memory)
  struct LinTables
  {
    const uint16_t *load_tab;
    const uint32_t *store_tab;
    const uint32_t *for_tab;
    const uint32_t *off; // [0..16] store classes, [17..25] foreign classes
    unsigned        flags;  // bit 0: zero the tile first; bit 1: private / shared own DoFs are interleaved (lex brick)
    const uint16_t *sh_tab; // global memory: offsets of the own shared DoFs
    unsigned        sh_cnt;
  };

  template <int k, int BZ, typename T>
  __device__ __forceinline__ void
  brick_issue_loads_lin(const BrickDesc &bd, const LinTables &tb, const uint32_t *gidx_f, T *tile, const T *__restrict__ src)
  {
    using G         = BrickGeom<k, BZ>;
    const int n_own = bd.npriv + bd.sh_count, n_for = tb.off[25];
    if (tb.flags & 1u)
      {
        for (int p = threadIdx.x; p < G::NPTS; p += G::NT)
          tile[p] = T(0);
        __syncthreads();
      }
#pragma unroll 4
    for (int i = threadIdx.x; i < n_own; i += G::NT)
      {
        const unsigned p = tb.load_tab[i];
        if (p != 0xFFFFu)
          cp_async_value(tile + p, src + bd.base + i);
      }
    for (int j = threadIdx.x; j < n_for; j += G::NT)
      {
        const int      p = tb.for_tab[j] & 0x1FFFu;
        const uint32_t g = gidx_f[j];
        if (g != DEV_INVALID)
          cp_async_value(tile + p, src + g);
        else
          tile[p] = T(0);
      }
  }

  template <int k, int BZ, typename T>
  __device__ __forceinline__ void
  brick_issue_ops_lin(const BrickDesc &bd, T *ops0, T *ops1, const Epilogue<T> &epi)
  {
    using G          = BrickGeom<k, BZ>;
    const bool need0 = (epi.kind == EPI_RESIDUAL || epi.kind == EPI_CHEB);
    const bool need1 = (epi.kind == EPI_CHEB && epi.f1 != T(0) && epi.v1 != nullptr);
    if (!need0)
      return;
    // the whole own range: the operands on the brick's own shared DoFs enter the value it adds there
    const int n_stage = (int)(bd.npriv + bd.sh_count);
#pragma unroll 4
    for (int i = threadIdx.x; i < n_stage; i += G::NT)
      {
        cp_async_value(ops0 + i, epi.v0 + bd.base + i);
        if (need1)
          cp_async_value(ops1 + i, epi.v1 + bd.base + i);
      }
  }

  // sum of the cell contributions of one DoF in a fixed order (deterministic); m = mask of the directions with
  // a second contribution.  Branch-free: absent contributions are read from the primary slot with weight 0
  // (the branchy form costs more in divergence than the extra shared-memory reads).
  template <typename T>
  __device__ __forceinline__ T
  slot_sum(const T *slots, const int o, const unsigned m, const int dx, const int dy, const int dz)
  {
    const int ex = (m & 1u) ? dx : 0, ey = (m & 2u) ? dy : 0, ez = (m & 4u) ? dz : 0;
    const T   wx = (m & 1u) ? T(1) : T(0), wy = (m & 2u) ? T(1) : T(0), wz = (m & 4u) ? T(1) : T(0);
    const T   a  = slots[o] + wx * slots[o + ex];
    const T   b  = slots[o + ey] + wx * slots[o + ey + ex];
    const T   c  = slots[o + ez] + wx * slots[o + ez + ex];
    const T   d  = slots[o + ez + ey] + wx * slots[o + ez + ey + ex];
    return (a + wy * b) + wz * (c + wy * d);
  }

  // store of a brick in the order of its own DoF range (coalesced global access; a class-sorted order with
  // compile-time masks was measured slower: the stores lose their coalescing)
  template <int k, int BZ, typename T>
  __device__ __forceinline__ void
  brick_store_lin(const BrickDesc &bd, const LinTables &tb, const T *slots, const T *ops0, const T *ops1, const uint32_t *gidx_f,
                  T *__restrict__ dst, T *__restrict__ acc, const Epilogue<T> &epi, const int shared_mode)
  {
    using G          = BrickGeom<k, BZ>;
    constexpr int n  = k + 1;
    const int     dx = -G::CS + k, dy = -bd.b[0] * G::CS + k * n, dz = -bd.b[0] * bd.b[1] * G::CS + k * n * n;
    const bool    need0 = (epi.kind == EPI_RESIDUAL || epi.kind == EPI_CHEB);
    const bool    need1 = (epi.kind == EPI_CHEB && epi.f1 != T(0) && epi.v1 != nullptr);
    const T       alpha = (epi.kind == EPI_RESIDUAL) ? T(-1) : ((epi.kind == EPI_CHEB || epi.kind == EPI_SCALE) ? epi.f2 : T(1));
    T *           sh_dst = (shared_mode == SHARED_DIRECT) ? dst : acc;
    const T       sh_a   = (shared_mode == SHARED_DIRECT) ? alpha : T(1);
    const int     n_own = bd.npriv + bd.sh_count, n_for = tb.off[25];
    // own DoFs in the order of the own range (coalesced): private ones get the fused epilogue and a plain store, the
    // ones on shared lower faces (STORE_SHARED; the tail of the range, or interleaved in a lex brick) a red.add
#pragma unroll 4
    for (int i = threadIdx.x; i < n_own; i += G::NT)
      {
        const uint32_t e = tb.store_tab[i];
        if (e == 0xFFFFFFFFu)
          continue;
        const T y = slot_sum(slots, (e >> 13) & 0x1FFFu, (e >> 26) & 7u, dx, dy, dz);
        const T v = epilogue_compute(epi, y, need0 ? ops0[i] : T(0), need1 ? ops1[i] : T(0));
        if (e & STORE_SHARED)
          atomic_add(sh_dst + bd.base + i, shared_mode == SHARED_DIRECT ? v : y);
        else
          dst[bd.base + i] = v;
      }
    // tile points owned by other bricks
    for (int j = threadIdx.x; j < n_for; j += G::NT)
      {
        const uint32_t g = gidx_f[j];
        if (g != DEV_INVALID)
          {
            const uint32_t e = tb.for_tab[j];
            atomic_add(sh_dst + g, sh_a * slot_sum(slots, (e >> 13) & 0x1FFFu, (e >> 26) & 7u, dx, dy, dz));
          }
      }
  }

  // stage the foreign index list of a brick (contiguous, 16-byte cp.async)
  template <int k, int BZ>
  __device__ __forceinline__ void
  brick_stage_foreign_async(const BrickMaps &maps, const int brick, uint32_t *gidx_f)
  {
    using G             = BrickGeom<k, BZ>;
    const uint32_t *src = maps.foreign_gidx + (size_t)brick * maps.nfp;
    for (int i = threadIdx.x; i < maps.nfp / 4; i += G::NT)
      {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(gidx_f + 4 * i);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(src + 4 * i));
      }
  }

  // copy the tables of a variant into shared memory
  template <int k, int BZ>
  __device__ __forceinline__ void
  brick_stage_tables(const BrickMaps &maps, const int variant, uint16_t *s_load, uint32_t *s_store, uint32_t *s_for, uint32_t *s_off,
                     LinTables &tb)
  {
    using G = BrickGeom<k, BZ>;
    for (int i = threadIdx.x; i < G::NPTS; i += G::NT)
      {
        s_load[i]  = maps.load_tab[(size_t)variant * maps.stride + i];
        s_store[i] = maps.store_tab[(size_t)variant * maps.stride + i];
      }
    for (int i = threadIdx.x; i < maps.nfp; i += G::NT)
      s_for[i] = maps.for_tab[(size_t)variant * maps.nfp + i];
    if (threadIdx.x < 17)
      s_off[threadIdx.x] = maps.store_off[variant * 17 + threadIdx.x];
    if (threadIdx.x >= 32 && threadIdx.x < 41)
      s_off[17 + threadIdx.x - 32] = maps.for_off[variant * 9 + threadIdx.x - 32];
    tb.load_tab  = s_load;
    tb.store_tab = s_store;
    tb.for_tab   = s_for;
    tb.off       = s_off;
    tb.flags     = maps.flags[variant];
    tb.sh_tab    = maps.sh_tab + (size_t)variant * maps.sh_stride;
    tb.sh_cnt    = maps.sh_cnt[variant];
    __syncthreads();
  }

  // fused pre-initialisation of the next kernel's destination on the brick's own shared DoFs: the operand loads
  // are issued first thing for a brick, the stores follow after the wait for the tile data, so that the two
  // latencies overlap
  template <int k, int BZ>
  struct NextInitRegs
  {
    static constexpr int IT = (BrickGeom<k, BZ>::NPTS - (BrickGeom<k, BZ>::BX * k - 1) * (BrickGeom<k, BZ>::BY * k - 1) * (BZ * k - 1) +
                               BrickGeom<k, BZ>::NT - 1) / BrickGeom<k, BZ>::NT; // upper bound of shared DoFs per thread
  };

  // the next kernel's destination is zeroed on the brick's own shared DoFs
  // lex == nullptr: they are the contiguous range [sh_base, sh_base + sh_count); else base + lex[i]
  template <int k, int BZ, typename T>
  __device__ __forceinline__ void
  brick_next_init_store(const BrickDesc &bd, const NextInit<T> &ni, const uint16_t *__restrict__ lex = nullptr)
  {
    using G = BrickGeom<k, BZ>;
    if (ni.out == nullptr)
      return;
    for (uint32_t i = threadIdx.x; i < bd.sh_count; i += G::NT)
      ni.out[lex ? bd.base + lex[i] : bd.sh_base + i] = T(0);
  }

  // reduction of the cell results (slots) per tile point in a fixed order (deterministic, no shared-memory
  // atomics), fused epilogue at private points (plain stores), red.global.add into the zero-invariant
  // accumulator at points on shared brick faces (completed by finish_shared_kernel)
  template <int k, int BZ, typename T>
  __device__ __forceinline__ void
  brick_reduce_store(const BrickDesc &bd, const T *slots, const T *ops0, const T *ops1, const uint32_t *gidx,
                     T *__restrict__ dst, T *__restrict__ acc, const Epilogue<T> &epi, const int shared_mode)
  {
    // affine form of the epilogue: dst = base + alpha * y
    const T alpha = (epi.kind == EPI_RESIDUAL) ? T(-1) : ((epi.kind == EPI_CHEB || epi.kind == EPI_SCALE) ? epi.f2 : T(1));
    using G             = BrickGeom<k, BZ>;
    constexpr int n     = k + 1;
    constexpr int NPENC = G::TX * G::TY;
    constexpr int PITER = (NPENC + G::NT - 1) / G::NT;
    const int     ex = bd.b[0] * k + 1, ey = bd.b[1] * k + 1, ez = bd.b[2] * k + 1;
    const bool    need0 = (epi.kind == EPI_RESIDUAL || epi.kind == EPI_CHEB);
    const bool    need1 = (epi.kind == EPI_CHEB && epi.f1 != T(0) && epi.v1 != nullptr);
#pragma unroll 1
    for (int pi = 0; pi < PITER; ++pi)
      {
        const int q  = threadIdx.x + pi * G::NT;
        const int px = q % G::TX, py = q / G::TX;
        if (q >= NPENC || px >= ex || py >= ey)
          continue;
        // the <= 2 x 2 cells containing the pencil in x and y: primary (c = p / k, l = p % k) and, on a cell
        // boundary, the lower neighbour with l = k.  Missing ones get mask 0 and a valid dummy offset.
        int o0, o1, o2, o3;
        T   m0, m1, m2, m3;
        {
          const int  chx = px / k, llx = px - chx * k, chy = py / k, lly = py - chy * k;
          const bool xa = chx < bd.b[0], xb = (llx == 0 && chx > 0), ya = chy < bd.b[1], yb = (lly == 0 && chy > 0);
          const int  cxa = xa ? chx : chx - 1, lxa = xa ? llx : k, cxb = xb ? chx - 1 : cxa, lxb = xb ? k : lxa;
          const int  cya = ya ? chy : chy - 1, lya = ya ? lly : k, cyb = yb ? chy - 1 : cya, lyb = yb ? k : lya;
          // (xa || xb) and (ya || yb) always hold for points of the tile
          o0 = (cya * bd.b[0] + cxa) * G::CS + lya * n + lxa;
          o1 = (cya * bd.b[0] + cxb) * G::CS + lya * n + lxb;
          o2 = (cyb * bd.b[0] + cxa) * G::CS + lyb * n + lxa;
          o3 = (cyb * bd.b[0] + cxb) * G::CS + lyb * n + lxb;
          m0 = T(1);
          m1 = (xa && xb) ? T(1) : T(0);
          m2 = (ya && yb) ? T(1) : T(0);
          m3 = m1 * m2;
        }
        const unsigned shxy = ((px == 0) ? (bd.shared & 1u) : 0u) | ((px == ex - 1) ? (bd.shared & 2u) : 0u) |
                              ((py == 0) ? (bd.shared & 4u) : 0u) | ((py == ey - 1) ? (bd.shared & 8u) : 0u);
        const int base = py * G::TX + px;
        const int czs  = bd.b[0] * bd.b[1] * G::CS;
#pragma unroll
        for (int pz = 0; pz < G::TZ; ++pz)
          {
            if (pz < ez)
              {
                const int      p = pz * NPENC + base;
                const uint32_t g = gidx[p];
                if (g != DEV_INVALID)
                  {
                    constexpr int dummy = 0;
                    (void)dummy;
                    const int  chz = pz / k, llz = pz - chz * k;
                    const bool za  = chz < bd.b[2];
                    const int  oza = (za ? chz : chz - 1) * czs + (za ? llz : k) * n * n;
                    T          y   = slots[oza + o0] + m1 * slots[oza + o1] + m2 * slots[oza + o2] + m3 * slots[oza + o3];
                    if (llz == 0 && chz > 0)
                      {
                        const T   mz  = za ? T(1) : T(0); // the lower cell was already taken as primary if !za
                        const int ozb = (chz - 1) * czs + k * n * n;
                        y += mz * (slots[ozb + o0] + m1 * slots[ozb + o1] + m2 * slots[ozb + o2] + m3 * slots[ozb + o3]);
                      }
                    const bool sh = (shxy | ((pz == 0) ? (bd.shared & 16u) : 0u) | ((pz == ez - 1) ? (bd.shared & 32u) : 0u)) != 0;
                    const bool up = ((shxy & 10u) | ((pz == ez - 1) ? (bd.shared & 32u) : 0u)) != 0; // owned by a neighbour brick
                    if (sh)
                      {
                        if (shared_mode == SHARED_DIRECT)
                          atomic_add(dst + g, up ? alpha * y : epilogue_compute(epi, y, need0 ? ops0[p] : T(0), need1 ? ops1[p] : T(0)));
                        else
                          atomic_add(acc + g, y);
                      }
                    else
                      dst[g] = epilogue_compute(epi, y, need0 ? ops0[p] : T(0), need1 ? ops1[p] : T(0));
                  }
              }
          }
      }
  }

  // ---- K1 (tuned): Laplace brick kernel -----------------------------------------------------------------
  // GEOM 0: uniform Cartesian;  GEOM 1: merged coefficients geom[cell][6][n^3];  GEOM 2: Jacobian rebuilt per
  // quadrature point from the 27 x 3 monomial coefficients of the triquadratic cell map geom[cell][27][3]
  // ("quadratic geometry", operator.h:1035-1159; "linear geometry" 916-1033 uses the same code with the
  // trilinear coefficients of the 8 vertices)
  template <int k, typename T, int BZ, int GEOM>
  __global__ void __launch_bounds__(BrickGeom<k, BZ>::NT, BrickGeom<k, BZ>::MINB)
  laplace_brick_kernel(const T *__restrict__ src,
                       T *__restrict__ dst,
                       T *__restrict__ acc,
                       const Epilogue<T> epi,
                       const uint32_t *__restrict__ cidx,
                       const BrickDesc *__restrict__ bricks,
                       const int n_bricks,
                       const T *__restrict__ geom,
                       const CartesianCoef cart,
                       const int n_ops,
                       const int shared_mode,
                       const NextInit<T> ni,
                       const BrickMaps maps,
                       const uint32_t *__restrict__ order) // optional: bricks to process (nullptr: all, in order)
  {
    auto BID = [&](const int i) { return order != nullptr ? (int)order[i] : i; };
    using G         = BrickGeom<k, BZ>;
    constexpr int n = k + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *       tile   = reinterpret_cast<T *>(smem_raw);
    T *       ops0   = (n_ops > 0) ? tile + G::NPTS : tile;
    T *       ops1   = (n_ops > 1) ? tile + 2 * G::NPTS : ops0;
    T *       slots  = tile + (1 + n_ops) * G::NPTS;
    uint32_t *gidx   = reinterpret_cast<uint32_t *>((reinterpret_cast<uintptr_t>(slots + G::NCELLS * G::CS) + 15) & ~uintptr_t(15));
    // non-lin: gidx[NPTS] | cidx[2];   lin: gidx_f[3][NFP] | store_tab[NPTS] | for_tab[NFP] | offsets[32] | load_tab u16
    uint32_t * s_cidx  = gidx + G::NPTS; // non-lin only: two buffers of NCELLS * 27
    uint32_t * s_store = gidx + 3 * G::NFP;
    uint32_t * s_for   = s_store + G::NPTS;
    uint32_t * s_off   = s_for + G::NFP;
    uint16_t * s_load  = reinterpret_cast<uint16_t *>(s_off + 32);
    int        cur_variant = -1;
    LinTables  tb;

    const auto &B = BasisOf<T>::template get<k>();
    const int   c = threadIdx.x % G::NCELLS; // cell in brick (lane-major: conflict-free slot access)
    const int   t = threadIdx.x / G::NCELLS; // plane index (warp-uniform)

    // Software pipeline over the bricks of this block (LIN: kernel brick == mesh brick, coalesced access through
    // the tile maps).  The tile is dead after phase A, so the gather of the NEXT brick is issued right after
    // phase A and overlaps with phases B-E and the store of the current brick.  cp.async groups per iteration:
    //   [ops(i)] at the top, [tile(i+1)] and [cidx(i+2)] after phase A.
    constexpr bool LIN = (BZ == 4);
    if (blockIdx.x >= n_bricks)
      return;
    int       buf = 0, fb = 0; // buf: parity of the brick; fb: ring position (mod 3) of its foreign-index buffer
    BrickDesc bd_next   = bricks[BID(blockIdx.x)];
    bool      late_tile = false;
    if (LIN)
      {
        brick_stage_foreign_async<k, BZ>(maps, BID(blockIdx.x), gidx);
        cp_async_commit();
        cp_async_wait<0>();
        cur_variant = bd_next.variant;
        brick_stage_tables<k, BZ>(maps, cur_variant, s_load, s_store, s_for, s_off, tb); // ends with __syncthreads
        brick_issue_loads_lin<k, BZ, T>(bd_next, tb, gidx, tile, src);
        cp_async_commit(); // [tile(b0)]
        if (blockIdx.x + gridDim.x < (unsigned)n_bricks)
          brick_stage_foreign_async<k, BZ>(maps, BID(blockIdx.x + gridDim.x), gidx + G::NFP);
        cp_async_commit(); // [foreign(b1)]
      }
    else
      {
        brick_stage_cidx_async<k, BZ>(bd_next, cidx, s_cidx);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
      }
    for (int bi = blockIdx.x; bi < n_bricks; bi += gridDim.x, buf ^= 1, fb = (fb + 1) % 3)
      {
        const BrickDesc bd       = bd_next;
        const bool      has_next = bi + (int)gridDim.x < n_bricks;
        if (has_next)
          bd_next = bricks[BID(bi + gridDim.x)]; // descriptor of the next brick: latency hidden behind this brick
        const int       ncells   = bd.b[0] * bd.b[1] * bd.b[2];
        const uint32_t *cur_cidx = s_cidx + buf * (G::NCELLS * 27);
        uint32_t *      cur_gidx = LIN ? gidx + fb * G::NFP : gidx;
        const uint16_t *ni_lex = (LIN && (bd.shared & BRICK_LEX)) ? tb.sh_tab : nullptr;
        if (LIN)
          {
            brick_issue_ops_lin<k, BZ, T>(bd, ops0, ops1, epi);
            cp_async_commit(); // [ops(i)]
            if (late_tile)
              cp_async_wait<1>(); // pending: ops(i), tile(i) issued late
            else
              cp_async_wait<2>(); // pending: ops(i), cidx(i+1), tile(i)
          }
        else
          {
            brick_issue_loads<k, BZ, T>(bd, cur_cidx, tile, ops0, ops1, gidx, src, (const T *)nullptr, (const T *)nullptr);
            cp_async_commit();
            brick_issue_loads_ops<k, BZ, T>(bd, gidx, ops0, ops1, epi);
            cp_async_commit();
            if (has_next)
              brick_stage_cidx_async<k, BZ>(bd_next, cidx, s_cidx + (buf ^ 1) * (G::NCELLS * 27));
            cp_async_commit();
            cp_async_wait<2>();
          }
        brick_next_init_store<k, BZ, T>(bd, ni, ni_lex);
        __syncthreads();

        const bool act = (c < ncells);
        const int  cx = c % bd.b[0], cy = (c / bd.b[0]) % bd.b[1], cz = c / (bd.b[0] * bd.b[1]);
        T *        S  = slots + c * G::CS;
        T          r[n][n]; // partial result of the x/z directions, plane y = t, [z][x]

        // phase A: plane z = t, [y][x]: interpolate in x and y
        if (act)
          {
            T          v[n][n];
            const T *  tp = tile + ((cz * k + t) * G::TY + cy * k) * G::TX + cx * k;
#pragma unroll
            for (int y = 0; y < n; ++y)
#pragma unroll
              for (int x = 0; x < n; ++x)
                v[y][x] = tp[y * G::TX + x];
            apply_fast<n, T, false>(v, B.N);
            apply_slow<n, T, false>(v, B.N);
#pragma unroll
            for (int y = 0; y < n; ++y)
#pragma unroll
              for (int x = 0; x < n; ++x)
                S[(t * n + y) * n + x] = v[y][x];
          }
        if (LIN)
          cp_async_wait<1>(); // indices of the next brick (staged one iteration ago); pending: ops(i)
        __syncthreads();
        if (LIN)
          {
            // the tile is dead from here on: gather the next brick into it while this brick is computed
            late_tile = has_next && ((int)bd_next.variant != cur_variant);
            if (has_next && !late_tile)
              brick_issue_loads_lin<k, BZ, T>(bd_next, tb, gidx + ((fb + 1) % 3) * G::NFP, tile, src);
            cp_async_commit(); // [tile(i+1)]
            if (bi + 2 * (int)gridDim.x < n_bricks)
              brick_stage_foreign_async<k, BZ>(maps, BID(bi + 2 * gridDim.x), gidx + ((fb + 2) % 3) * G::NFP);
            cp_async_commit(); // [foreign(i+2)]
          }
        if (GEOM == 0)
          {
            // phase B: plane y = t, [z][x]: interpolate in z; x and z parts of the Laplacian
            if (act)
              {
                T w[n][n];
#pragma unroll
                for (int z = 0; z < n; ++z)
#pragma unroll
                  for (int x = 0; x < n; ++x)
                    w[z][x] = S[(z * n + t) * n + x];
                apply_slow<n, T, false>(w, B.N);
#pragma unroll
                for (int z = 0; z < n; ++z)
#pragma unroll
                  for (int x = 0; x < n; ++x)
                    {
                      S[(z * n + t) * n + x] = w[z][x];
                      r[z][x]                = 0;
                    }
                const T wt = B.qw[t];
                T       wz_x[n], wz_z[n];
#pragma unroll
                for (int z = 0; z < n; ++z)
                  {
                    wz_x[z] = T(cart.g[0]) * wt * B.qw[z]; // weight of the slow index for the x derivative
                    wz_z[z] = T(cart.g[2]) * wt * B.qw[z]; // weight along z (slow) for the z derivative
                  }
                laplace_1d_fast<n, T>(r, w, B.Dq, B.qw, wz_x); // f = Dx w * (qw[x] * g0 wt qw[z])
                laplace_1d_slow<n, T>(r, w, B.Dq, B.qw, wz_z); // f = Dz w * (qw[x] * g2 wt qw[zq])
              }
            __syncthreads();
            // phase C: plane z = t, [y][x]: y part
            if (act)
              {
                T w[n][n], ry[n][n];
#pragma unroll
                for (int y = 0; y < n; ++y)
#pragma unroll
                  for (int x = 0; x < n; ++x)
                    {
                      w[y][x]  = S[(t * n + y) * n + x];
                      ry[y][x] = 0;
                    }
                const T wt = B.qw[t];
                T       wy[n];
#pragma unroll
                for (int y = 0; y < n; ++y)
                  wy[y] = T(cart.g[1]) * wt * B.qw[y];
                laplace_1d_slow<n, T>(ry, w, B.Dq, B.qw, wy);
#pragma unroll
                for (int y = 0; y < n; ++y)
#pragma unroll
                  for (int x = 0; x < n; ++x)
                    S[(t * n + y) * n + x] = ry[y][x];
              }
            __syncthreads();
          }
        else
          {
            // general geometry: gradients, quadrature-point operation with the 6 merged coefficients
            T gx[n][n], gz[n][n];
            if (act)
              {
                T w[n][n];
#pragma unroll
                for (int z = 0; z < n; ++z)
#pragma unroll
                  for (int x = 0; x < n; ++x)
                    w[z][x] = S[(z * n + t) * n + x];
                apply_slow<n, T, false>(w, B.N);
#pragma unroll
                for (int z = 0; z < n; ++z)
#pragma unroll
                  for (int x = 0; x < n; ++x)
                    {
                      S[(z * n + t) * n + x] = w[z][x];
                      gx[z][x]               = w[z][x];
                      gz[z][x]               = w[z][x];
                    }
                apply_fast<n, T, false>(gx, B.Dq);
                apply_slow<n, T, false>(gz, B.Dq);
              }
            __syncthreads();
            if (act) // plane z = t: gy
              {
                T w[n][n];
#pragma unroll
                for (int y = 0; y < n; ++y)
#pragma unroll
                  for (int x = 0; x < n; ++x)
                    w[y][x] = S[(t * n + y) * n + x];
                apply_slow<n, T, false>(w, B.Dq);
#pragma unroll
                for (int y = 0; y < n; ++y)
#pragma unroll
                  for (int x = 0; x < n; ++x)
                    S[(t * n + y) * n + x] = w[y][x];
              }
            __syncthreads();
            if (act) // plane y = t: quadrature-point operation
              {
                constexpr int n3 = n * n * n;
                const T *     Gc = geom + (size_t)(bd.first_cell + c) * (GEOM == 1 ? 6 * n3 : 81);
                const T       eta = B.qp[t];
#pragma unroll
                for (int z = 0; z < n; ++z)
                  {
                    // GEOM 2: coefficients reduced in zeta and eta for this (z, t): position derivative pieces
                    T Bx[3][3], By[3][3], Bz[3][3]; // [power of xi][component]
                    if (GEOM == 2)
                      {
                        const T zeta = B.qp[z];
#pragma unroll
                        for (int i = 0; i < 3; ++i)
#pragma unroll
                          for (int d = 0; d < 3; ++d)
                            {
                              T A[3], dA[3];
#pragma unroll
                              for (int j = 0; j < 3; ++j)
                                {
                                  const T v0 = Gc[(3 * j + i) * 3 + d], v1 = Gc[(9 + 3 * j + i) * 3 + d], v2 = Gc[(18 + 3 * j + i) * 3 + d];
                                  A[j]  = v0 + zeta * (v1 + zeta * v2);
                                  dA[j] = v1 + (zeta + zeta) * v2;
                                }
                              Bx[i][d] = A[0] + eta * (A[1] + eta * A[2]);
                              By[i][d] = A[1] + (eta + eta) * A[2];
                              Bz[i][d] = dA[0] + eta * (dA[1] + eta * dA[2]);
                            }
                      }
#pragma unroll
                    for (int x = 0; x < n; ++x)
                      {
                        const int q   = (z * n + t) * n + x;
                        const T   gyv = S[q];
                        const T   a = gx[z][x], cc = gz[z][x];
                        T         gxx, gxy, gxz, gyy, gyz, gzz;
                        if (GEOM == 1)
                          {
                            gxx = Gc[q];
                            gxy = Gc[n3 + q];
                            gxz = Gc[2 * n3 + q];
                            gyy = Gc[3 * n3 + q];
                            gyz = Gc[4 * n3 + q];
                            gzz = Gc[5 * n3 + q];
                          }
                        else
                          {
                            const T xi = B.qp[x];
                            T       J[3][3]; // J[d][e] = d x_d / d xi_e
#pragma unroll
                            for (int d = 0; d < 3; ++d)
                              {
                                J[d][0] = Bx[1][d] + (xi + xi) * Bx[2][d];
                                J[d][1] = By[0][d] + xi * (By[1][d] + xi * By[2][d]);
                                J[d][2] = Bz[0][d] + xi * (Bz[1][d] + xi * Bz[2][d]);
                              }
                            const T c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2],
                                    c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
                            const T det  = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
                            const T idet = T(1) / det;
                            // inverse I[e][d] = d xi_e / d x_d
                            const T I00 = c00 * idet, I01 = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * idet,
                                    I02 = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * idet;
                            const T I10 = c01 * idet, I11 = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * idet,
                                    I12 = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * idet;
                            const T I20 = c02 * idet, I21 = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * idet,
                                    I22 = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * idet;
                            const T jxw = det * (B.qw[x] * B.qw[t] * B.qw[z]);
                            gxx         = jxw * (I00 * I00 + I01 * I01 + I02 * I02);
                            gxy         = jxw * (I00 * I10 + I01 * I11 + I02 * I12);
                            gxz         = jxw * (I00 * I20 + I01 * I21 + I02 * I22);
                            gyy         = jxw * (I10 * I10 + I11 * I11 + I12 * I12);
                            gyz         = jxw * (I10 * I20 + I11 * I21 + I12 * I22);
                            gzz         = jxw * (I20 * I20 + I21 * I21 + I22 * I22);
                          }
                        gx[z][x] = gxx * a + gxy * gyv + gxz * cc;
                        S[q]     = gxy * a + gyy * gyv + gyz * cc;
                        gz[z][x] = gxz * a + gyz * gyv + gzz * cc;
                      }
                  }
                apply_fast<n, T, true>(gx, B.Dq);
                apply_slow<n, T, true>(gz, B.Dq);
#pragma unroll
                for (int z = 0; z < n; ++z)
#pragma unroll
                  for (int x = 0; x < n; ++x)
                    r[z][x] = gx[z][x] + gz[z][x];
              }
            __syncthreads();
            if (act) // plane z = t: Dy^T fy
              {
                T w[n][n];
#pragma unroll
                for (int y = 0; y < n; ++y)
#pragma unroll
                  for (int x = 0; x < n; ++x)
                    w[y][x] = S[(t * n + y) * n + x];
                apply_slow<n, T, true>(w, B.Dq);
#pragma unroll
                for (int y = 0; y < n; ++y)
#pragma unroll
                  for (int x = 0; x < n; ++x)
                    S[(t * n + y) * n + x] = w[y][x];
              }
            __syncthreads();
          }
        // phase D: plane y = t: add the y part, N^T in z
        if (act)
          {
#pragma unroll
            for (int z = 0; z < n; ++z)
#pragma unroll
              for (int x = 0; x < n; ++x)
                r[z][x] += S[(z * n + t) * n + x];
            apply_slow<n, T, true>(r, B.N);
#pragma unroll
            for (int z = 0; z < n; ++z)
#pragma unroll
              for (int x = 0; x < n; ++x)
                S[(z * n + t) * n + x] = r[z][x];
          }
        __syncthreads();
        // phase E: plane z = t: N^T in y and x -> final cell result in the slot
        if (act)
          {
            T v[n][n];
#pragma unroll
            for (int y = 0; y < n; ++y)
#pragma unroll
              for (int x = 0; x < n; ++x)
                v[y][x] = S[(t * n + y) * n + x];
            apply_slow<n, T, true>(v, B.N);
            apply_fast<n, T, true>(v, B.N);
#pragma unroll
            for (int y = 0; y < n; ++y)
#pragma unroll
              for (int x = 0; x < n; ++x)
                S[(t * n + y) * n + x] = v[y][x];
          }
        if (LIN)
          cp_async_wait<2>(); // epilogue operands have landed; pending: cidx(i+2), tile(i+1)
        else
          cp_async_wait<1>();
        __syncthreads();
        if (LIN)
          brick_store_lin<k, BZ, T>(bd, tb, slots, ops0, ops1, cur_gidx, dst, acc, epi, shared_mode);
        else
          brick_reduce_store<k, BZ, T>(bd, slots, ops0, ops1, gidx, dst, acc, epi, shared_mode);
        if (LIN)
          {
            if (late_tile)
              {
                // the next brick has another tile-map variant: its maps can only be staged once this brick is stored
                __syncthreads();
                cur_variant = bd_next.variant;
                brick_stage_tables<k, BZ>(maps, cur_variant, s_load, s_store, s_for, s_off, tb);
                brick_issue_loads_lin<k, BZ, T>(bd_next, tb, gidx + ((fb + 1) % 3) * G::NFP, tile, src);
                cp_async_commit();
              }
          }
        else
          {
            cp_async_wait<0>(); // indices of the next brick
            __syncthreads();
          }
      }
  }

  // ---- K4 (tuned): FDM brick kernel, n_overlap = 1 (patch = cell closure), compressed weights -----------------
  template <int k, typename T, int BZ>
  __global__ void __launch_bounds__(BrickGeom<k, BZ>::NT, BrickGeom<k, BZ>::MINB)
  fdm_brick_kernel(const T *__restrict__ src,
                   T *__restrict__ dst,
                   T *__restrict__ acc,
                   const Epilogue<T> epi,
                   const uint32_t *__restrict__ cidx,
                   const BrickDesc *__restrict__ bricks,
                   const int n_bricks,
                   const uint32_t *__restrict__ inst,
                   const T *__restrict__ Smat,
                   const T *__restrict__ lam,
                   const uint8_t *__restrict__ cw, // weight codes [cell][32] (27 used) or nullptr
                   const WeightTable<T> wtab,      // weight value of every code
                   const uint4 *__restrict__ brick_tri, // per brick: instance triple of its first cell, w = all cells equal
                   const int w_pre,
                   const int w_post,
                   const int n_ops,
                   const int shared_mode,
                   const NextInit<T> ni,
                   const BrickMaps maps,
                   const int dbg,
                   const uint32_t *__restrict__ order) // optional: bricks to process (nullptr: all, in order)
  {
    auto BID = [&](const int i) { return order != nullptr ? (int)order[i] : i; };
    using G          = BrickGeom<k, BZ>;
    constexpr int n  = k + 1;
    constexpr int n2 = n * n;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *       tile   = reinterpret_cast<T *>(smem_raw);
    T *       ops0   = (n_ops > 0) ? tile + G::NPTS : tile;
    T *       ops1   = (n_ops > 1) ? tile + 2 * G::NPTS : ops0;
    T *       slots  = tile + (1 + n_ops) * G::NPTS;
    uint32_t *gidx   = reinterpret_cast<uint32_t *>((reinterpret_cast<uintptr_t>(slots + G::NCELLS * G::CS) + 15) & ~uintptr_t(15));
    // non-lin: gidx[NPTS] | cidx[2];   lin: gidx_f[3][NFP] | store_tab[NPTS] | for_tab[NFP] | offsets[32] | load_tab u16
    uint32_t * s_cidx  = gidx + G::NPTS; // non-lin only: two buffers of NCELLS * 27
    uint32_t * s_store = gidx + 3 * G::NFP;
    uint32_t * s_for   = s_store + G::NPTS;
    uint32_t * s_off   = s_for + G::NFP;
    uint16_t * s_load  = reinterpret_cast<uint16_t *>(s_off + 32);
    int        cur_variant = -1;
    LinTables  tb;

    const int c = threadIdx.x % G::NCELLS; // cell in brick (lane-major: conflict-free slot access)
    const int t = threadIdx.x / G::NCELLS; // plane index (warp-uniform)

    // cache of the 1-D eigen-decompositions when all cells of a brick share one instance triple (Cartesian
    // meshes): matrices, eigenvalues and the table of inverse eigenvalue sums live in shared memory and are
    // re-staged only when the triple changes; otherwise (deformed meshes) every thread reads its cell's
    // matrices from global memory and divides on the fly
    __shared__ T        s_M[3][n * n];
    __shared__ T        s_lam[3][n];
    __shared__ T        s_inv[n * n * n];
    __shared__ uint32_t s_tri[3];
    __shared__ __align__(16) uint8_t s_wc[2][G::NCELLS * 32];
    if (threadIdx.x < 3)
      s_tri[threadIdx.x] = 0xFFFFFFFFu;

    // Software pipeline over the bricks of this block (LIN: kernel brick == mesh brick, coalesced access through
    // the tile maps).  The tile is dead after phase A, so the gather of the NEXT brick is issued right after
    // phase A and overlaps with phases B-E and the store of the current brick.  cp.async groups per iteration:
    //   [ops(i)] at the top, [tile(i+1)] and [cidx(i+2)] after phase A.
    constexpr bool LIN = (BZ == 4);
    if (blockIdx.x >= n_bricks)
      return;
    int       buf = 0, fb = 0; // buf: parity of the brick; fb: ring position (mod 3) of its foreign-index buffer
    BrickDesc bd_next   = bricks[BID(blockIdx.x)];
    bool      late_tile = false;
    auto      stage_codes = [&](const BrickDesc &b2, const int which) {
      if (cw != nullptr)
        {
          const int nb16 = b2.b[0] * b2.b[1] * b2.b[2] * 2; // 32 bytes per cell
          for (int i = threadIdx.x; i < nb16; i += G::NT)
            {
              const unsigned sa = (unsigned)__cvta_generic_to_shared(&s_wc[which][i * 16]);
              asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(cw + (size_t)b2.first_cell * 32 + i * 16));
            }
        }
    };
    if (LIN)
      {
        brick_stage_foreign_async<k, BZ>(maps, BID(blockIdx.x), gidx);
        cp_async_commit();
        cp_async_wait<0>();
        cur_variant = bd_next.variant;
        brick_stage_tables<k, BZ>(maps, cur_variant, s_load, s_store, s_for, s_off, tb); // ends with __syncthreads
        stage_codes(bd_next, 0);
        brick_issue_loads_lin<k, BZ, T>(bd_next, tb, gidx, tile, src);
        cp_async_commit(); // [tile(b0)]
        if (blockIdx.x + gridDim.x < (unsigned)n_bricks)
          brick_stage_foreign_async<k, BZ>(maps, BID(blockIdx.x + gridDim.x), gidx + G::NFP);
        cp_async_commit(); // [foreign(b1)]
      }
    else
      {
        brick_stage_cidx_async<k, BZ>(bd_next, cidx, s_cidx);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
      }
    for (int bi = blockIdx.x; bi < n_bricks; bi += gridDim.x, buf ^= 1, fb = (fb + 1) % 3)
      {
        const BrickDesc bd       = bd_next;
        const bool      has_next = bi + (int)gridDim.x < n_bricks;
        if (has_next)
          bd_next = bricks[BID(bi + gridDim.x)]; // descriptor of the next brick: latency hidden behind this brick
        const int       ncells   = bd.b[0] * bd.b[1] * bd.b[2];
        const uint32_t *cur_cidx = s_cidx + buf * (G::NCELLS * 27);
        uint32_t *      cur_gidx = LIN ? gidx + fb * G::NFP : gidx;
        const uint16_t *ni_lex = (LIN && (bd.shared & BRICK_LEX)) ? tb.sh_tab : nullptr;
        if (LIN)
          {
            brick_issue_ops_lin<k, BZ, T>(bd, ops0, ops1, epi);
            cp_async_commit(); // [ops(i)]
            if (late_tile)
              cp_async_wait<1>(); // pending: ops(i), tile(i) issued late
            else
              cp_async_wait<2>(); // pending: ops(i), cidx(i+1), tile(i)
          }
        else
          {
            stage_codes(bd, buf);
            brick_issue_loads<k, BZ, T>(bd, cur_cidx, tile, ops0, ops1, gidx, src, (const T *)nullptr, (const T *)nullptr);
            cp_async_commit();
            brick_issue_loads_ops<k, BZ, T>(bd, gidx, ops0, ops1, epi);
            cp_async_commit();
            if (has_next)
              brick_stage_cidx_async<k, BZ>(bd_next, cidx, s_cidx + (buf ^ 1) * (G::NCELLS * 27));
            cp_async_commit();
            cp_async_wait<2>();
          }
        brick_next_init_store<k, BZ, T>(bd, ni, ni_lex);
        __syncthreads();

        const bool     act  = (c < ncells) && !(dbg & 2);
        const int      cx = c % bd.b[0], cy = (c / bd.b[0]) % bd.b[1], cz = c / (bd.b[0] * bd.b[1]);
        const uint32_t cell = bd.first_cell + c;
        T *            S    = slots + c * G::CS;
        const uint4    tri     = brick_tri[BID(bi)];
        const uint32_t r0 = tri.x, r1 = tri.y, r2 = tri.z;
        const bool     uniform = tri.w != 0;
        uint32_t       i0 = r0, i1 = r1, i2 = r2;
        if (act && !uniform)
          {
            i0 = inst[(size_t)cell * 3 + 0];
            i1 = inst[(size_t)cell * 3 + 1];
            i2 = inst[(size_t)cell * 3 + 2];
          }
        if (uniform && (s_tri[0] != r0 || s_tri[1] != r1 || s_tri[2] != r2))
          {
            __syncthreads();
            for (int i = threadIdx.x; i < 3 * n2; i += G::NT)
              s_M[i / n2][i % n2] = Smat[(size_t)(i / n2 == 0 ? r0 : (i / n2 == 1 ? r1 : r2)) * n2 + i % n2];
            for (int i = threadIdx.x; i < 3 * n; i += G::NT)
              s_lam[i / n][i % n] = lam[(size_t)(i / n == 0 ? r0 : (i / n == 1 ? r1 : r2)) * n + i % n];
            __syncthreads();
            for (int i = threadIdx.x; i < n * n2; i += G::NT)
              s_inv[i] = T(1) / (s_lam[0][i % n] + s_lam[1][(i / n) % n] + s_lam[2][i / n2]);
            if (threadIdx.x < 3)
              s_tri[threadIdx.x] = (threadIdx.x == 0 ? r0 : (threadIdx.x == 1 ? r1 : r2));
            __syncthreads();
          }
        // matrix d of this thread's cell into registers: LDS (broadcast) for uniform bricks, else global
        auto load_matrix = [&](T(&M)[n2], const int d, const uint32_t id) {
          if (uniform)
            {
#pragma unroll
              for (int i = 0; i < n2; ++i)
                M[i] = s_M[d][i];
            }
          else
            {
#pragma unroll
              for (int i = 0; i < n2; ++i)
                M[i] = Smat[(size_t)id * n2 + i];
            }
        };
        const int et = (t == 0) ? 0 : ((t == k) ? 2 : 1); // entity code of the plane index

        // phase A: plane z = t, [y][x]: (pre-weights) S0^T in x, S1^T in y
        if (act)
          {
            T        v[n][n];
            const T *tp = tile + ((cz * k + t) * G::TY + cy * k) * G::TX + cx * k;
#pragma unroll
            for (int y = 0; y < n; ++y)
#pragma unroll
              for (int x = 0; x < n; ++x)
                v[y][x] = tp[y * G::TX + x];
            if (cw != nullptr && w_pre && !(dbg & 4))
              {
                T wloc[9];
#pragma unroll
                for (int e = 0; e < 9; ++e)
                  wloc[e] = wtab.v[s_wc[buf][c * 32 + e + 9 * et]];
#pragma unroll
                for (int y = 0; y < n; ++y)
#pragma unroll
                  for (int x = 0; x < n; ++x)
                    v[y][x] *= wloc[((x == 0) ? 0 : ((x == k) ? 2 : 1)) + 3 * ((y == 0) ? 0 : ((y == k) ? 2 : 1))];
              }
            {
              T M[n2];
              load_matrix(M, 0, i0);
              apply_fast<n, T, true>(v, M);
            }
            {
              T M[n2];
              load_matrix(M, 1, i1);
              apply_slow<n, T, true>(v, M);
            }
#pragma unroll
            for (int y = 0; y < n; ++y)
#pragma unroll
              for (int x = 0; x < n; ++x)
                S[(t * n + y) * n + x] = v[y][x];
          }
        if (LIN)
          cp_async_wait<1>(); // indices of the next brick (staged one iteration ago); pending: ops(i)
        __syncthreads();
        if (LIN)
          {
            // the tile is dead from here on: gather the next brick into it while this brick is computed
            late_tile = has_next && ((int)bd_next.variant != cur_variant);
            if (has_next)
              stage_codes(bd_next, buf ^ 1);
            if (has_next && !late_tile)
              brick_issue_loads_lin<k, BZ, T>(bd_next, tb, gidx + ((fb + 1) % 3) * G::NFP, tile, src);
            cp_async_commit(); // [tile(i+1)]
            if (bi + 2 * (int)gridDim.x < n_bricks)
              brick_stage_foreign_async<k, BZ>(maps, BID(bi + 2 * gridDim.x), gidx + ((fb + 2) % 3) * G::NFP);
            cp_async_commit(); // [foreign(i+2)]
          }
        // phase B: plane y = t, [z][x]: S2^T in z, scale by 1/(l0[x] + l1[t] + l2[z]), S2 in z, S0 in x
        if (act)
          {
            T w[n][n];
#pragma unroll
            for (int z = 0; z < n; ++z)
#pragma unroll
              for (int x = 0; x < n; ++x)
                w[z][x] = S[(z * n + t) * n + x];
            T M[n2];
            load_matrix(M, 2, i2);
            apply_slow<n, T, true>(w, M);
            if (uniform)
              {
#pragma unroll
                for (int z = 0; z < n; ++z)
#pragma unroll
                  for (int x = 0; x < n; ++x)
                    w[z][x] *= s_inv[(z * n + t) * n + x];
              }
            else
              {
                T       l0[n], l2[n];
                const T l1 = lam[(size_t)i1 * n + t];
#pragma unroll
                for (int i = 0; i < n; ++i)
                  {
                    l0[i] = lam[(size_t)i0 * n + i];
                    l2[i] = lam[(size_t)i2 * n + i];
                  }
#pragma unroll
                for (int z = 0; z < n; ++z)
#pragma unroll
                  for (int x = 0; x < n; ++x)
                    w[z][x] = w[z][x] / (l0[x] + l1 + l2[z]);
              }
            apply_slow<n, T, false>(w, M);
            load_matrix(M, 0, i0);
            apply_fast<n, T, false>(w, M);
#pragma unroll
            for (int z = 0; z < n; ++z)
#pragma unroll
              for (int x = 0; x < n; ++x)
                S[(z * n + t) * n + x] = w[z][x];
          }
        __syncthreads();
        // phase C: plane z = t, [y][x]: S1 in y, (post-weights) -> final cell result in the slot
        if (act)
          {
            T v[n][n];
#pragma unroll
            for (int y = 0; y < n; ++y)
#pragma unroll
              for (int x = 0; x < n; ++x)
                v[y][x] = S[(t * n + y) * n + x];
            T M[n2];
            load_matrix(M, 1, i1);
            apply_slow<n, T, false>(v, M);
            if (cw != nullptr && w_post && !(dbg & 4))
              {
                T wloc[9];
#pragma unroll
                for (int e = 0; e < 9; ++e)
                  wloc[e] = wtab.v[s_wc[buf][c * 32 + e + 9 * et]];
#pragma unroll
                for (int y = 0; y < n; ++y)
#pragma unroll
                  for (int x = 0; x < n; ++x)
                    v[y][x] *= wloc[((x == 0) ? 0 : ((x == k) ? 2 : 1)) + 3 * ((y == 0) ? 0 : ((y == k) ? 2 : 1))];
              }
#pragma unroll
            for (int y = 0; y < n; ++y)
#pragma unroll
              for (int x = 0; x < n; ++x)
                S[(t * n + y) * n + x] = v[y][x];
          }
        if (LIN)
          cp_async_wait<2>(); // epilogue operands have landed; pending: cidx(i+2), tile(i+1)
        else
          cp_async_wait<1>();
        __syncthreads();
        if (LIN)
          brick_store_lin<k, BZ, T>(bd, tb, slots, ops0, ops1, cur_gidx, dst, acc, epi, shared_mode);
        else
          brick_reduce_store<k, BZ, T>(bd, slots, ops0, ops1, gidx, dst, acc, epi, shared_mode);
        if (LIN)
          {
            if (late_tile)
              {
                // the next brick has another tile-map variant: its maps can only be staged once this brick is stored
                __syncthreads();
                cur_variant = bd_next.variant;
                brick_stage_tables<k, BZ>(maps, cur_variant, s_load, s_store, s_for, s_off, tb);
                brick_issue_loads_lin<k, BZ, T>(bd_next, tb, gidx + ((fb + 1) % 3) * G::NFP, tile, src);
                cp_async_commit();
              }
          }
        else
          {
            cp_async_wait<0>(); // indices of the next brick
            __syncthreads();
          }
      }
  }

  // ---- shared-face DoFs: finish (dst = epilogue(acc), acc = 0) over the precomputed list of shared DoFs ---------
  template <typename T>
  __global__ void
  finish_shared_kernel(T *__restrict__ dst, T *__restrict__ acc, const Epilogue<T> epi, const uint32_t *__restrict__ list,
                       const long long n)
  {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
      {
        const uint32_t g = list[i];
        const T        y = acc[g];
        T              a, b;
        epilogue_load(epi, g, a, b);
        acc[g] = T(0);
        dst[g] = epilogue_compute(epi, y, a, b);
      }
  }

  // stand-alone pre-initialisation of a destination on all shared DoFs (first kernel of a fused sequence)
  template <typename T>
  __global__ void
  init_shared_kernel(const NextInit<T> ni, const uint32_t *__restrict__ list, const long long n)
  {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
      {
        const uint32_t g = list[i];
        const T        a = (ni.v0 != nullptr) ? ni.v0[g] : T(0);
        const T        b = (ni.v1 != nullptr && ni.f1 != T(0)) ? ni.v1[g] : T(0);
        ni.out[g]        = a + ni.f1 * (a - b);
      }
  }

  // epilogue on an index list (constrained DoFs): y = src[i] (unit-matrix operation) or 0
  template <typename T>
  __global__ void
  epilogue_indexed_kernel(T *__restrict__ dst, const T *__restrict__ src, const Epilogue<T> epi, const uint32_t *__restrict__ idx,
                          const long long n)
  {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
      {
        const uint32_t g = idx[i];
        dst[g]           = epilogue_apply(epi, src != nullptr ? src[g] : T(0), g);
      }
  }
} // namespace dasm
