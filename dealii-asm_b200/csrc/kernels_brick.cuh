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
    uint8_t  shared;     // bit (2 d + side): the tile face is shared with cells outside the brick
    uint16_t ib[3];      // position in the lattice of kernel bricks
    uint16_t pad;
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
    static constexpr int CS     = n * n * n;    // slot stride per cell
    // registers per thread with one resident block per SM (64 K registers, allocation granularity 8)
    // (registers are allocated per warp in units of 512)
    static constexpr int NWARPS = (NT + 31) / 32;
    static constexpr int MAXREG = ((65536 / NWARPS / 512) * 512 / 32) >= 255 ? 255 : ((65536 / NWARPS / 512) * 512 / 32);
    template <typename T>
    static constexpr size_t
    smem_bytes()
    {
      return (size_t)NPTS * sizeof(T) + (size_t)NCELLS * CS * sizeof(T) + (size_t)NPTS * sizeof(uint32_t) +
             (size_t)NCELLS * 27 * sizeof(uint32_t);
    }
  };

  // ---- register-plane contractions --------------------------------------------------------------
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
          {
            T s = (TRANS ? M[o] : M[o * n]) * v[a][0];
#pragma unroll
            for (int i = 1; i < n; ++i)
              s += (TRANS ? M[i * n + o] : M[o * n + i]) * v[a][i];
            r[o] = s;
          }
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
          {
            T s = (TRANS ? M[o] : M[o * n]) * v[0][b];
#pragma unroll
            for (int i = 1; i < n; ++i)
              s += (TRANS ? M[i * n + o] : M[o * n + i]) * v[i][b];
            r[o] = s;
          }
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
          {
            T s = D[q * n] * v[a][0];
#pragma unroll
            for (int i = 1; i < n; ++i)
              s += D[q * n + i] * v[a][i];
            f[q] = s * (wfast[q] * wslow[a]);
          }
#pragma unroll
        for (int o = 0; o < n; ++o)
          {
            T s = r[a][o];
#pragma unroll
            for (int q = 0; q < n; ++q)
              s += D[q * n + o] * f[q];
            r[a][o] = s;
          }
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
          {
            T s = D[q * n] * v[0][b];
#pragma unroll
            for (int i = 1; i < n; ++i)
              s += D[q * n + i] * v[i][b];
            f[q] = s * (wfast[b] * wslow[q]);
          }
#pragma unroll
        for (int o = 0; o < n; ++o)
          {
            T s = r[o][b];
#pragma unroll
            for (int q = 0; q < n; ++q)
              s += D[q * n + o] * f[q];
            r[o][b] = s;
          }
      }
  }

  // ---- tile helpers ----------------------------------------------------------------------------------
  // Thread -> tile point mapping: every thread owns "pencils" (px, py) of the tile and walks along pz, so
  // that everything depending on (px, py) (canonical cell, entity codes, offsets, slot addresses) is
  // computed once and the per-point work is a handful of integer instructions.  The 27 compressed indices
  // of the brick's cells are staged in shared memory by one coalesced copy (the cells of a brick are
  // consecutive).
  template <int k>
  __device__ __forceinline__ bool
  point_shared(const BrickDesc &bd, int px, int py, int pz)
  {
    const int ex = bd.b[0] * k, ey = bd.b[1] * k, ez = bd.b[2] * k;
    unsigned  s  = 0;
    s |= (px == 0) ? (bd.shared & 1u) : 0u;
    s |= (px == ex) ? (bd.shared & 2u) : 0u;
    s |= (py == 0) ? (bd.shared & 4u) : 0u;
    s |= (py == ey) ? (bd.shared & 8u) : 0u;
    s |= (pz == 0) ? (bd.shared & 16u) : 0u;
    s |= (pz == ez) ? (bd.shared & 32u) : 0u;
    return s != 0;
  }

  template <int k, int BZ>
  __device__ __forceinline__ void
  brick_stage_cidx(const BrickDesc &bd, const uint32_t *__restrict__ cidx, uint32_t *s_cidx)
  {
    using G          = BrickGeom<k, BZ>;
    const int      n = bd.b[0] * bd.b[1] * bd.b[2] * 27;
    const uint32_t *src = cidx + (size_t)bd.first_cell * 27;
    for (int i = threadIdx.x; i < n; i += G::NT)
      s_cidx[i] = src[i];
  }

  template <int k, int BZ, typename T, typename LoadFn>
  __device__ __forceinline__ void
  brick_load_tile(const BrickDesc &bd, const uint32_t *s_cidx, T *tile, uint32_t *gidx, LoadFn load)
  {
    using G              = BrickGeom<k, BZ>;
    constexpr int NPENC  = G::TX * G::TY;
    constexpr int PITER  = (NPENC + G::NT - 1) / G::NT;
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
        uint32_t  g[G::TZ];
#pragma unroll
        for (int pz = 0; pz < G::TZ; ++pz)
          {
            const int cz = min(pz / k, bd.b[2] - 1), lz = pz - cz * k;
            const int ecz = (lz == 0) ? 0 : ((lz == k) ? 2 : 1), oz = (ecz == 1) ? lz - 1 : 0;
            g[pz]         = DEV_INVALID;
            if (pz < ez)
              {
                const uint32_t st = s_cidx[(cz * czs + cxy) * 27 + exy + 9 * ecz];
                g[pz]             = (st == DEV_INVALID) ? DEV_INVALID : st + oxy + sxy * oz;
              }
          }
        T v[G::TZ];
#pragma unroll
        for (int pz = 0; pz < G::TZ; ++pz)
          v[pz] = (g[pz] == DEV_INVALID) ? T(0) : load(g[pz]);
#pragma unroll
        for (int pz = 0; pz < G::TZ; ++pz)
          if (pz < ez)
            {
              gidx[(pz * G::TY + py) * G::TX + px] = g[pz];
              tile[(pz * G::TY + py) * G::TX + px] = v[pz];
            }
      }
  }

  // The cell results are accumulated into the (re-used) tile by the cell threads themselves, colour by
  // colour: cells of equal parity (cx&1, cy&1, cz&1) share no tile point, so the 8 colour steps need no
  // atomics and the summation order is fixed (deterministic results).
  template <int k, int BZ, typename T>
  __device__ __forceinline__ void
  brick_accumulate(const bool act, const int cx, const int cy, const int cz, const int t, T *tile, const T (&v)[k + 1][k + 1])
  {
    using G         = BrickGeom<k, BZ>;
    constexpr int n = k + 1;
    const int     colour = (cx & 1) | ((cy & 1) << 1) | ((cz & 1) << 2);
    T *           tp     = tile + ((cz * k + t) * G::TY + cy * k) * G::TX + cx * k;
#pragma unroll 1
    for (int col = 0; col < 8; ++col)
      {
        if (act && colour == col)
          {
#pragma unroll
            for (int y = 0; y < n; ++y)
#pragma unroll
              for (int x = 0; x < n; ++x)
                tp[y * G::TX + x] += v[y][x];
          }
        __syncthreads();
      }
  }

  template <int k, int BZ, typename T>
  __device__ __forceinline__ void
  brick_zero_tile(T *tile)
  {
    using G = BrickGeom<k, BZ>;
    for (int p = threadIdx.x; p < G::NPTS; p += G::NT)
      tile[p] = T(0);
  }

  // fused epilogue (private points: plain stores) / accumulation (points on shared faces: red.global.add
  // into the zero-invariant accumulator; finish_shared_kernel completes them: the numbering makes the shared
  // DoFs of a brick one contiguous range, so that pass is a coalesced sweep over ~18 % of the vector).
  // (An in-kernel variant where the last-arriving brick finishes a shared piece - arrival counters,
  // threadfence-reduction pattern - was measured slower: the fence + counter round trips are exposed.)
  template <int k, int BZ, typename T>
  __device__ __forceinline__ void
  brick_store_tile(const BrickDesc &bd, const T *tile, const uint32_t *gidx, T *__restrict__ dst, T *__restrict__ acc,
                   const Epilogue<T> &epi)
  {
    using G             = BrickGeom<k, BZ>;
    constexpr int NPENC = G::TX * G::TY;
    constexpr int PITER = (NPENC + G::NT - 1) / G::NT;
    constexpr int CH    = (G::TZ + 2) / 3; // points along z whose global loads are in flight together
    const int     ex = bd.b[0] * k + 1, ey = bd.b[1] * k + 1, ez = bd.b[2] * k + 1;
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
#pragma unroll 1
        for (int z0 = 0; z0 < ez; z0 += CH)
          {
            uint32_t g[CH];
            T        ea[CH], eb[CH], y[CH];
            bool     sh[CH];
#pragma unroll
            for (int j = 0; j < CH; ++j)
              {
                const int pz = z0 + j;
                g[j]         = DEV_INVALID;
                sh[j]        = false;
                ea[j] = eb[j] = y[j] = T(0);
                if (pz < ez)
                  {
                    g[j]  = gidx[pz * NPENC + base];
                    y[j]  = tile[pz * NPENC + base];
                    sh[j] = (shxy | ((pz == 0) ? (bd.shared & 16u) : 0u) | ((pz == ez - 1) ? (bd.shared & 32u) : 0u)) != 0;
                    if (g[j] != DEV_INVALID && !sh[j])
                      epilogue_load(epi, g[j], ea[j], eb[j]);
                  }
              }
#pragma unroll
            for (int j = 0; j < CH; ++j)
              if (g[j] != DEV_INVALID)
                {
                  if (sh[j])
                    atomic_add(acc + g[j], y[j]);
                  else
                    dst[g[j]] = epilogue_compute(epi, y[j], ea[j], eb[j]);
                }
          }
      }
  }

  // ---- K1 (tuned): Laplace brick kernel -----------------------------------------------------------------
  // GEOM 0: uniform Cartesian;  GEOM 1: merged coefficients geom[cell][6][n^3]
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
                       const int dbg)
  {
    using G         = BrickGeom<k, BZ>;
    constexpr int n = k + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *       tile  = reinterpret_cast<T *>(smem_raw);
    T *       slots = tile + G::NPTS;
    uint32_t *gidx  = reinterpret_cast<uint32_t *>(slots + G::NCELLS * G::CS);
    uint32_t *s_cidx = gidx + G::NPTS;

    const auto &B = BasisOf<T>::template get<k>();
    const int   c = threadIdx.x / n; // cell in brick
    const int   t = threadIdx.x % n; // plane index

    for (int bi = blockIdx.x; bi < n_bricks; bi += gridDim.x)
      {
        const BrickDesc bd     = bricks[bi];
        const int       ncells = bd.b[0] * bd.b[1] * bd.b[2];
        brick_stage_cidx<k, BZ>(bd, cidx, s_cidx);
        __syncthreads();
        if (!(dbg & 1))
          brick_load_tile<k, BZ, T>(bd, s_cidx, tile, gidx, [&](uint32_t g) { return src[g]; });
        __syncthreads();

        const bool act = (c < ncells) && !(dbg & 2);
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
        __syncthreads();
        brick_zero_tile<k, BZ, T>(tile); // the source values are consumed; the tile now collects the results
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
                const T *Gc = geom + (size_t)(bd.first_cell + c) * 6 * G::CS;
#pragma unroll
                for (int z = 0; z < n; ++z)
#pragma unroll
                  for (int x = 0; x < n; ++x)
                    {
                      const int q   = (z * n + t) * n + x;
                      const T   gyv = S[q];
                      const T   a = gx[z][x], cc = gz[z][x];
                      const T   gxx = Gc[q], gxy = Gc[G::CS + q], gxz = Gc[2 * G::CS + q], gyy = Gc[3 * G::CS + q],
                              gyz = Gc[4 * G::CS + q], gzz = Gc[5 * G::CS + q];
                      gx[z][x] = gxx * a + gxy * gyv + gxz * cc;
                      S[q]     = gxy * a + gyy * gyv + gyz * cc;
                      gz[z][x] = gxz * a + gyz * gyv + gzz * cc;
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
        // phase E: plane z = t: N^T in y and x; the result stays in registers and is accumulated into the tile
        {
          T v[n][n];
          if (act)
            {
#pragma unroll
              for (int y = 0; y < n; ++y)
#pragma unroll
                for (int x = 0; x < n; ++x)
                  v[y][x] = S[(t * n + y) * n + x];
              apply_slow<n, T, true>(v, B.N);
              apply_fast<n, T, true>(v, B.N);
            }
          if (!(dbg & 4))
            brick_accumulate<k, BZ, T>(act, cx, cy, cz, t, tile, v);
        }
        if (!(dbg & 8))
          brick_store_tile<k, BZ, T>(bd, tile, gidx, dst, acc, epi);
        __syncthreads();
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
                   const T *__restrict__ cw, // [cell][27] or nullptr
                   const int w_pre,
                   const int w_post)
  {
    using G          = BrickGeom<k, BZ>;
    constexpr int n  = k + 1;
    constexpr int n2 = n * n;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *       tile  = reinterpret_cast<T *>(smem_raw);
    T *       slots = tile + G::NPTS;
    uint32_t *gidx  = reinterpret_cast<uint32_t *>(slots + G::NCELLS * G::CS);
    uint32_t *s_cidx = gidx + G::NPTS;

    const int c = threadIdx.x / n;
    const int t = threadIdx.x % n;

    for (int bi = blockIdx.x; bi < n_bricks; bi += gridDim.x)
      {
        const BrickDesc bd     = bricks[bi];
        const int       ncells = bd.b[0] * bd.b[1] * bd.b[2];
        brick_stage_cidx<k, BZ>(bd, cidx, s_cidx);
        __syncthreads();
        brick_load_tile<k, BZ, T>(bd, s_cidx, tile, gidx, [&](uint32_t g) { return src[g]; });
        __syncthreads();

        const bool     act  = c < ncells;
        const int      cx = c % bd.b[0], cy = (c / bd.b[0]) % bd.b[1], cz = c / (bd.b[0] * bd.b[1]);
        const uint32_t cell = bd.first_cell + c;
        T *            S    = slots + c * G::CS;
        uint32_t       i0 = 0, i1 = 0, i2 = 0;
        if (act)
          {
            i0 = inst[(size_t)cell * 3 + 0];
            i1 = inst[(size_t)cell * 3 + 1];
            i2 = inst[(size_t)cell * 3 + 2];
          }
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
            if (cw != nullptr && w_pre)
              {
                T wloc[9];
#pragma unroll
                for (int e = 0; e < 9; ++e)
                  wloc[e] = cw[(size_t)cell * 27 + e + 9 * et];
#pragma unroll
                for (int y = 0; y < n; ++y)
#pragma unroll
                  for (int x = 0; x < n; ++x)
                    v[y][x] *= wloc[((x == 0) ? 0 : ((x == k) ? 2 : 1)) + 3 * ((y == 0) ? 0 : ((y == k) ? 2 : 1))];
              }
            {
              T M[n2];
#pragma unroll
              for (int i = 0; i < n2; ++i)
                M[i] = Smat[(size_t)i0 * n2 + i];
              apply_fast<n, T, true>(v, M);
            }
            {
              T M[n2];
#pragma unroll
              for (int i = 0; i < n2; ++i)
                M[i] = Smat[(size_t)i1 * n2 + i];
              apply_slow<n, T, true>(v, M);
            }
#pragma unroll
            for (int y = 0; y < n; ++y)
#pragma unroll
              for (int x = 0; x < n; ++x)
                S[(t * n + y) * n + x] = v[y][x];
          }
        __syncthreads();
        brick_zero_tile<k, BZ, T>(tile);
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
#pragma unroll
            for (int i = 0; i < n2; ++i)
              M[i] = Smat[(size_t)i2 * n2 + i];
            apply_slow<n, T, true>(w, M);
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
#pragma unroll
            for (int i = 0; i < n2; ++i)
              M[i] = Smat[(size_t)i0 * n2 + i];
            apply_fast<n, T, false>(w, M);
#pragma unroll
            for (int z = 0; z < n; ++z)
#pragma unroll
              for (int x = 0; x < n; ++x)
                S[(z * n + t) * n + x] = w[z][x];
          }
        __syncthreads();
        // phase C: plane z = t, [y][x]: S1 in y, (post-weights); result accumulated into the tile
        {
          T v[n][n];
          if (act)
            {
#pragma unroll
              for (int y = 0; y < n; ++y)
#pragma unroll
                for (int x = 0; x < n; ++x)
                  v[y][x] = S[(t * n + y) * n + x];
              T M[n2];
#pragma unroll
              for (int i = 0; i < n2; ++i)
                M[i] = Smat[(size_t)i1 * n2 + i];
              apply_slow<n, T, false>(v, M);
              if (cw != nullptr && w_post)
                {
                  T wloc[9];
#pragma unroll
                  for (int e = 0; e < 9; ++e)
                    wloc[e] = cw[(size_t)cell * 27 + e + 9 * et];
#pragma unroll
                  for (int y = 0; y < n; ++y)
#pragma unroll
                    for (int x = 0; x < n; ++x)
                      v[y][x] *= wloc[((x == 0) ? 0 : ((x == k) ? 2 : 1)) + 3 * ((y == 0) ? 0 : ((y == k) ? 2 : 1))];
                }
            }
          brick_accumulate<k, BZ, T>(act, cx, cy, cz, t, tile, v);
        }
        brick_store_tile<k, BZ, T>(bd, tile, gidx, dst, acc, epi);
        __syncthreads();
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
