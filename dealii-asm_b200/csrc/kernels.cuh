// Device kernels of libdasm (sm_100a).  See DESIGN.md for the data layout and the roofline of each.
//
//  K1  laplace cell kernel      gather (27 compressed indices) -> sum factorisation -> scatter-add
//  K4  fdm cell kernel          gather -> (S x S x S)^T -> 1/(lx+ly+lz) -> (S x S x S) -> scatter-add
//  K6  vector epilogues         residual / chebyshev update / scale  (hooks of the reference)
//  K9  diagonal kernel
//
// Generic kernels ("line per thread"): one thread per 1-D line of a cell, n^2 threads per cell,
// CPB cells per thread block; every 1-D contraction reads a line from shared memory, multiplies
// with the n x n matrix held in constant memory (vmult) or shared memory (FDM) and writes the line
// back.  They work for every degree 1..8 and both number types and are the fallback of the tuned
// kernels in kernels_tuned.cuh.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace dasm
{
  constexpr uint32_t DEV_INVALID = 0xFFFFFFFFu;

  template <typename T>
  struct DevBasis
  {
    T N[81];  // N[q*n+i]  nodal basis at Gauss points
    T Dq[81]; // Dq[q*n+p] collocation derivative
    T Dn[81]; // Dn[q*n+i] derivative of nodal basis at Gauss points
    T qw[9];
    T qp[9];  // Gauss points on [0,1]
  };

  static __constant__ DevBasis<double> c_basis_d[9];
  static __constant__ DevBasis<float>  c_basis_f[9];

  template <typename T>
  struct BasisOf;
  template <>
  struct BasisOf<double>
  {
    template <int k>
    static __device__ __forceinline__ const DevBasis<double> &
    get()
    {
      return c_basis_d[k];
    }
  };
  template <>
  struct BasisOf<float>
  {
    template <int k>
    static __device__ __forceinline__ const DevBasis<float> &
    get()
    {
      return c_basis_f[k];
    }
  };

  // ---- index helpers -------------------------------------------------------------------------
  // entity code (0 lower, 1 interior, 2 upper) and offset inside the entity for 1-D position i
  template <int k>
  __device__ __forceinline__ void
  split_1d(int i, int &e, int &o)
  {
    e = (i == 0) ? 0 : ((i == k) ? 2 : 1);
    o = (e == 1) ? (i - 1) : 0;
  }

  // global index of local DoF (x,y,z) of a cell from its 27 compressed indices
  // (standard orientation branch of vector_access_reduced.h:267-405); an entity stored in a lex brick (LEX_FLAG,
  // mesh.h) is expanded with the strides 1, 4k, 16k^2 of the brick's box
  constexpr uint32_t DEV_LEX_FLAG = 0x80000000u;
  template <int k>
  __device__ __forceinline__ uint32_t
  compressed_index(const uint32_t *__restrict__ ci, int x, int y, int z)
  {
    int ex, ey, ez, ox, oy, oz;
    split_1d<k>(x, ex, ox);
    split_1d<k>(y, ey, oy);
    split_1d<k>(z, ez, oz);
    const uint32_t start = ci[ex + 3 * ey + 9 * ez];
    if (start == DEV_INVALID)
      return DEV_INVALID;
    const bool lex = (start & DEV_LEX_FLAG) != 0;
    const int  sx  = lex ? 4 * k : ((ex == 1) ? (k - 1) : 1);
    const int  sy  = lex ? 4 * k : ((ey == 1) ? (k - 1) : 1);
    return (start & ~DEV_LEX_FLAG) + ox + sx * (oy + sy * oz);
  }

  template <typename T>
  __device__ __forceinline__ void
  atomic_add(T *addr, T v)
  {
    atomicAdd(addr, v);
  }

  // ---- 1-D sweeps -------------------------------------------------------------------------------
  // line index helper: DIR 0: along x at (y=a,z=b); 1: along y at (x=a,z=b); 2: along z at (x=a,y=b)
  template <int n, int DIR>
  __device__ __forceinline__ int
  line_idx(int a, int b, int i)
  {
    if (DIR == 0)
      return (b * n + a) * n + i;
    else if (DIR == 1)
      return (b * n + i) * n + a;
    else
      return (i * n + b) * n + a;
  }

  // out_line = M in_line (TRANS: M^T), M[o*n+i] row-major compile-time-indexed (constant memory)
  template <int n, typename T, int DIR, bool TRANS, bool ADD>
  __device__ __forceinline__ void
  sweep_const(const T *M, const T *in, T *out, int a, int b)
  {
    T v[n], r[n];
#pragma unroll
    for (int i = 0; i < n; ++i)
      v[i] = in[line_idx<n, DIR>(a, b, i)];
#pragma unroll
    for (int o = 0; o < n; ++o)
      {
        T s = 0;
#pragma unroll
        for (int i = 0; i < n; ++i)
          s += (TRANS ? M[i * n + o] : M[o * n + i]) * v[i];
        r[o] = s;
      }
#pragma unroll
    for (int o = 0; o < n; ++o)
      {
        const int idx = line_idx<n, DIR>(a, b, o);
        if (ADD)
          out[idx] += r[o];
        else
          out[idx] = r[o];
      }
  }

  // row of a runtime matrix in shared memory into registers; rows of a multiple of 16 bytes (the matrices of the FDM kernel then
  // start on 16-byte boundaries, see fdm_rows_aligned) are read with 128-bit loads: the matrix reads are what occupies the LSU pipe
  template <int m, typename T, bool VEC>
  __device__ __forceinline__ void
  load_row(const T *__restrict__ row, T (&out)[m])
  {
    if constexpr (VEC && sizeof(T) == 8)
      {
#pragma unroll
        for (int j = 0; j < m / 2; ++j)
          {
            const double2 q = reinterpret_cast<const double2 *>(row)[j];
            out[2 * j]      = q.x;
            out[2 * j + 1]  = q.y;
          }
      }
    else if constexpr (VEC && sizeof(T) == 4)
      {
#pragma unroll
        for (int j = 0; j < m / 4; ++j)
          {
            const float4 q = reinterpret_cast<const float4 *>(row)[j];
            out[4 * j]     = q.x;
            out[4 * j + 1] = q.y;
            out[4 * j + 2] = q.z;
            out[4 * j + 3] = q.w;
          }
      }
    else
      {
#pragma unroll
        for (int i = 0; i < m; ++i)
          out[i] = row[i];
      }
  }

  // per-cell shared-memory block of the FDM kernel: U[m^3] S0 S1 S2 [m^2 each] L0 L1 L2 [m each]; rows and blocks 16-byte aligned?
  template <int m, typename T>
  __host__ __device__ constexpr bool
  fdm_rows_aligned()
  {
    return (m * sizeof(T)) % 16 == 0 && ((m * m * m + 3 * m * m + 3 * m) * sizeof(T)) % 16 == 0;
  }

  // same with a runtime matrix in shared memory (row-major M[o*m+i]); the sums run over i in ascending order in both variants
  template <int n, typename T, int DIR, bool TRANS, bool VEC = false>
  __device__ __forceinline__ void
  sweep_smem(const T *__restrict__ M, T *buf, int a, int b)
  {
    T v[n], r[n], row[n];
#pragma unroll
    for (int i = 0; i < n; ++i)
      v[i] = buf[line_idx<n, DIR>(a, b, i)];
    if (TRANS)
      {
#pragma unroll
        for (int o = 0; o < n; ++o)
          r[o] = 0;
#pragma unroll
        for (int i = 0; i < n; ++i)
          {
            load_row<n, T, VEC>(M + i * n, row);
#pragma unroll
            for (int o = 0; o < n; ++o)
              r[o] += row[o] * v[i];
          }
      }
    else
      {
#pragma unroll
        for (int o = 0; o < n; ++o)
          {
            load_row<n, T, VEC>(M + o * n, row);
            T s = 0;
#pragma unroll
            for (int i = 0; i < n; ++i)
              s += row[i] * v[i];
            r[o] = s;
          }
      }
#pragma unroll
    for (int o = 0; o < n; ++o)
      buf[line_idx<n, DIR>(a, b, o)] = r[o];
  }

  template <int k>
  __host__ __device__ constexpr int
  cells_per_block()
  {
    // ~128-256 threads per block
    return (k + 1) * (k + 1) >= 64 ? (k == 8 ? 3 : 4) : (256 / ((k + 1) * (k + 1)) > 16 ? 16 : 256 / ((k + 1) * (k + 1)));
  }

  struct CartesianCoef
  {
    double g[3]; // diag of  det * J^-1 J^-T  (without quadrature weight)
  };

  // "construct q" (operator.h:712-746, 1221-1333): the cell stores the coordinates of its quadrature points, X[e][q]; the Jacobian
  // at a point is the collocation derivative of the coordinate field, G = JxW J^-1 J^-T is rebuilt per point (order xx xy xz yy yz zz)
  template <int k, typename T, typename Basis>
  __device__ __forceinline__ void
  construct_q_coefficients(const T *__restrict__ X, const int qx, const int qy, const int qz, const Basis &B, T (&G)[6])
  {
    constexpr int n = k + 1, n3 = n * n * n;
    T             J[3][3]; // J[e][d] = d x_e / d xi_d
#pragma unroll
    for (int e = 0; e < 3; ++e)
      {
        const T *Xe = X + e * n3;
        T        s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
        for (int i = 0; i < n; ++i)
          {
            s0 += B.Dq[qx * n + i] * Xe[(qz * n + qy) * n + i];
            s1 += B.Dq[qy * n + i] * Xe[(qz * n + i) * n + qx];
            s2 += B.Dq[qz * n + i] * Xe[(i * n + qy) * n + qx];
          }
        J[e][0] = s0;
        J[e][1] = s1;
        J[e][2] = s2;
      }
    const T det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                  J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
    const T id = T(1) / det;
    T       I[3][3]; // inverse: I[d][e] = d xi_d / d x_e
    I[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) * id;
    I[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
    I[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
    I[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) * id;
    I[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
    I[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
    I[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) * id;
    I[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
    I[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
    const T jxw = det * B.qw[qx] * B.qw[qy] * B.qw[qz];
    int     cc  = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int e = d; e < 3; ++e, ++cc)
        G[cc] = jxw * (I[d][0] * I[e][0] + I[d][1] * I[e][1] + I[d][2] * I[e][2]);
  }

  // "quadratic geometry" on unstructured meshes (operator.h:1035-1159: the Jacobian rebuilt per quadrature point from the 27 x 3
  // coefficients of the triquadratic cell map; here the coefficients are the support points themselves, shifted by the first one).
  // A thread works along an x-line of quadrature points at fixed (ty, tz): the sums over the y and z basis functions are formed once
  // per line (P0: d/dxi, P1: d/deta, P2: d/dzeta still open in x), a point then costs 27 FMAs for its Jacobian.
  template <typename T>
  __device__ __forceinline__ void
  q2_basis(const T t, T (&V)[3], T (&D)[3])
  {
    V[0] = (T(1) - T(2) * t) * (T(1) - t);
    V[1] = T(4) * t * (T(1) - t);
    V[2] = t * (T(2) * t - T(1));
    D[0] = T(4) * t - T(3);
    D[1] = T(4) - T(8) * t;
    D[2] = T(4) * t - T(1);
  }

  template <typename T>
  struct SupportLine
  {
    T P0[3][3], P1[3][3], P2[3][3]; // [i][e]
  };

  template <typename T>
  __device__ __forceinline__ void
  support_line(const T *__restrict__ X, const T ty, const T tz, SupportLine<T> &L)
  {
    T Vy[3], Dy[3], Vz[3], Dz[3];
    q2_basis(ty, Vy, Dy);
    q2_basis(tz, Vz, Dz);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int e = 0; e < 3; ++e)
        L.P0[i][e] = L.P1[i][e] = L.P2[i][e] = T(0);
#pragma unroll
    for (int l = 0; l < 3; ++l)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        {
          const T w0 = Vy[j] * Vz[l], w1 = Dy[j] * Vz[l], w2 = Vy[j] * Dz[l];
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int e = 0; e < 3; ++e)
              {
                const T x = X[(9 * l + 3 * j + i) * 3 + e];
                L.P0[i][e] += w0 * x;
                L.P1[i][e] += w1 * x;
                L.P2[i][e] += w2 * x;
              }
        }
  }

  // G = JxW J^-1 J^-T at the point tx of the line (w = product of the three quadrature weights), order xx xy xz yy yz zz
  template <typename T>
  __device__ __forceinline__ void
  support_point_coefficients(const SupportLine<T> &L, const T tx, const T w, T (&G)[6])
  {
    T Vx[3], Dx[3], J[3][3];
    q2_basis(tx, Vx, Dx);
#pragma unroll
    for (int e = 0; e < 3; ++e)
      {
        J[e][0] = Dx[0] * L.P0[0][e] + Dx[1] * L.P0[1][e] + Dx[2] * L.P0[2][e];
        J[e][1] = Vx[0] * L.P1[0][e] + Vx[1] * L.P1[1][e] + Vx[2] * L.P1[2][e];
        J[e][2] = Vx[0] * L.P2[0][e] + Vx[1] * L.P2[1][e] + Vx[2] * L.P2[2][e];
      }
    const T det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                  J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
    const T id = T(1) / det;
    T       I[3][3];
    I[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) * id;
    I[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
    I[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
    I[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) * id;
    I[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
    I[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
    I[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) * id;
    I[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
    I[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
    const T jxw = det * w;
    int     cc  = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int e = d; e < 3; ++e, ++cc)
        G[cc] = jxw * (I[d][0] * I[e][0] + I[d][1] * I[e][1] + I[d][2] * I[e][2]);
  }

  // ---- K1: Laplace cell kernel (generic) ---------------------------------------------------------
  // GEOM 0: uniform Cartesian (3 constants), 1: merged coefficients geom[cell][6][n^3], 2: construct q, geom[cell][3][n^3],
  // 3: support points of the triquadratic cell map geom[cell][27][3] (Jacobian rebuilt per point)
  template <int k, typename T, int GEOM>
  __global__ void __launch_bounds__(cells_per_block<k>() * (k + 1) * (k + 1))
  laplace_generic_kernel(const T *__restrict__ src,
                         T *__restrict__ dst,
                         const uint32_t *__restrict__ cidx,
                         const T *__restrict__ geom,
                         const CartesianCoef cart,
                         const long long     n_cells,
                         const uint32_t *__restrict__ plain = nullptr, // compress_indices = false: n^3 indices per cell
                         const uint32_t *__restrict__ cell_ids = nullptr) // work on the cells cell_ids[0 .. n_cells) (power kernel)
  {
    // GEOM 0 Cartesian, 1 merged coefficients, 2 construct q; 4 mass operator on a Cartesian mesh (cart.g[0] = cell volume),
    // 5 no computation (gather + scatter only): the second operator / the do_computation = false mode of power_kernel_01
    constexpr int n = k + 1, n2 = n * n, n3 = n2 * n, CPB = cells_per_block<k>();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *smem = reinterpret_cast<T *>(smem_raw);

    const int       cl   = threadIdx.x / n2;
    const int       t    = threadIdx.x % n2;
    const int       a    = t % n;
    const int       b    = t / n;
    const long long slot = (long long)blockIdx.x * CPB + cl;
    const bool      act  = slot < n_cells;
    const long long cell = (act && cell_ids != nullptr) ? (long long)cell_ids[slot] : slot;

    T *U  = smem + (size_t)cl * 4 * n3;
    T *GX = U + n3;
    T *GY = GX + n3;
    T *GZ = GY + n3;
    __shared__ uint32_t s_ci[CPB][27];
    __shared__ T        s_X[GEOM == 3 ? CPB : 1][81];
    if (act)
      for (int e = t; e < 27; e += n2)
        s_ci[cl][e] = cidx[cell * 27 + e];
    if (GEOM == 3 && act)
      for (int e = t; e < 81; e += n2)
        s_X[cl][e] = geom[(size_t)cell * 81 + e];
    __syncthreads();

    const auto &B = BasisOf<T>::template get<k>();

    // gather: thread (a,b) loads the x-line at (y=a, z=b)
    if (act)
      {
#pragma unroll
        for (int x = 0; x < n; ++x)
          {
            const uint32_t gi = plain ? plain[cell * n3 + (b * n + a) * n + x] : compressed_index<k>(s_ci[cl], x, a, b);
            U[(b * n + a) * n + x] = (gi == DEV_INVALID) ? T(0) : src[gi];
          }
      }
    __syncthreads();
    if (GEOM != 5)
      {
    // interpolate to Gauss points (in place)
    if (act)
      sweep_const<n, T, 0, false, false>(B.N, U, U, a, b);
    __syncthreads();
    if (act)
      sweep_const<n, T, 1, false, false>(B.N, U, U, a, b);
    __syncthreads();
    if (act)
      sweep_const<n, T, 2, false, false>(B.N, U, U, a, b);
    __syncthreads();
    if (GEOM == 4)
      {
        // mass operator: values times JxW (evaluate(values) / submit_value / integrate(values), power_kernel_01.likwid.cc:423-440)
        if (act)
          {
#pragma unroll
            for (int x = 0; x < n; ++x)
              U[(b * n + a) * n + x] *= T(cart.g[0]) * B.qw[x] * B.qw[a] * B.qw[b];
          }
      }
    else
      {
    // collocation gradients
    if (act)
      {
        sweep_const<n, T, 0, false, false>(B.Dq, U, GX, a, b);
        sweep_const<n, T, 1, false, false>(B.Dq, U, GY, a, b);
        sweep_const<n, T, 2, false, false>(B.Dq, U, GZ, a, b);
      }
    __syncthreads();
    // quadrature-point operation on the x-line (y=a, z=b)
    if (act)
      {
        SupportLine<T> SL;
        if (GEOM == 3)
          support_line<T>(s_X[cl], B.qp[a], B.qp[b], SL);
#pragma unroll
        for (int x = 0; x < n; ++x)
          {
            const int q  = (b * n + a) * n + x;
            const T   gx = GX[q], gy = GY[q], gz = GZ[q];
            if (GEOM == 0)
              {
                const T w = B.qw[x] * B.qw[a] * B.qw[b];
                GX[q]     = T(cart.g[0]) * w * gx;
                GY[q]     = T(cart.g[1]) * w * gy;
                GZ[q]     = T(cart.g[2]) * w * gz;
              }
            else if (GEOM == 1)
              {
                const T *G   = geom + (size_t)cell * 6 * n3 + q;
                const T  gxx = G[0], gxy = G[n3], gxz = G[2 * n3], gyy = G[3 * n3], gyz = G[4 * n3], gzz = G[5 * n3];
                GX[q]        = gxx * gx + gxy * gy + gxz * gz;
                GY[q]        = gxy * gx + gyy * gy + gyz * gz;
                GZ[q]        = gxz * gx + gyz * gy + gzz * gz;
              }
            else if (GEOM == 3)
              {
                T G[6];
                support_point_coefficients<T>(SL, B.qp[x], B.qw[x] * B.qw[a] * B.qw[b], G);
                GX[q] = G[0] * gx + G[1] * gy + G[2] * gz;
                GY[q] = G[1] * gx + G[3] * gy + G[4] * gz;
                GZ[q] = G[2] * gx + G[4] * gy + G[5] * gz;
              }
            else
              {
                T G[6];
                construct_q_coefficients<k, T>(geom + (size_t)cell * 3 * n3, x, a, b, B, G);
                GX[q] = G[0] * gx + G[1] * gy + G[2] * gz;
                GY[q] = G[1] * gx + G[3] * gy + G[4] * gz;
                GZ[q] = G[2] * gx + G[4] * gy + G[5] * gz;
              }
          }
      }
    __syncthreads();
    // integrate: R = Dx^T GX + Dy^T GY + Dz^T GZ  (into U)
    if (act)
      sweep_const<n, T, 0, true, false>(B.Dq, GX, U, a, b);
    __syncthreads();
    if (act)
      sweep_const<n, T, 1, true, true>(B.Dq, GY, U, a, b);
    __syncthreads();
    if (act)
      sweep_const<n, T, 2, true, true>(B.Dq, GZ, U, a, b);
      } // GEOM != 4
    __syncthreads();
    if (act)
      sweep_const<n, T, 2, true, false>(B.N, U, U, a, b);
    __syncthreads();
    if (act)
      sweep_const<n, T, 1, true, false>(B.N, U, U, a, b);
    __syncthreads();
    if (act)
      sweep_const<n, T, 0, true, false>(B.N, U, U, a, b);
    __syncthreads();
      } // GEOM != 5
    // scatter-add
    if (act)
      {
#pragma unroll
        for (int x = 0; x < n; ++x)
          {
            const uint32_t gi = plain ? plain[cell * n3 + (b * n + a) * n + x] : compressed_index<k>(s_ci[cl], x, a, b);
            if (gi != DEV_INVALID)
              atomic_add(dst + gi, U[(b * n + a) * n + x]);
          }
      }
  }

  // ---- K9: diagonal of the cell matrices ---------------------------------------------------------
  template <int k, typename T, int GEOM>
  __global__ void
  laplace_diagonal_kernel(T *__restrict__ diag,
                          const uint32_t *__restrict__ cidx,
                          const T *__restrict__ geom,
                          const CartesianCoef cart,
                          const long long     n_cells,
                          const uint32_t *__restrict__ plain = nullptr) // n^3 indices per cell (plain storage / unstructured meshes)
  {
    constexpr int   n = k + 1, n3 = n * n * n;
    const long long gid  = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long cell = gid / n3;
    if (cell >= n_cells)
      return;
    const int   i  = gid % n3;
    const int   ix = i % n, iy = (i / n) % n, iz = i / (n * n);
    const auto &B  = BasisOf<T>::template get<k>();
    double      s  = 0;
    for (int qz = 0; qz < n; ++qz)
      for (int qy = 0; qy < n; ++qy)
        {
          SupportLine<T> SL;
          if (GEOM == 3)
            support_line<T>(geom + (size_t)cell * 81, B.qp[qy], B.qp[qz], SL);
        for (int qx = 0; qx < n; ++qx)
          {
            const double gx = (double)B.Dn[qx * n + ix] * (double)B.N[qy * n + iy] * (double)B.N[qz * n + iz];
            const double gy = (double)B.N[qx * n + ix] * (double)B.Dn[qy * n + iy] * (double)B.N[qz * n + iz];
            const double gz = (double)B.N[qx * n + ix] * (double)B.N[qy * n + iy] * (double)B.Dn[qz * n + iz];
            if (GEOM == 0)
              {
                const double w = (double)B.qw[qx] * (double)B.qw[qy] * (double)B.qw[qz];
                s += w * (cart.g[0] * gx * gx + cart.g[1] * gy * gy + cart.g[2] * gz * gz);
              }
            else if (GEOM == 1)
              {
                const int q = (qz * n + qy) * n + qx;
                const T * G = geom + (size_t)cell * 6 * n3 + q;
                s += (double)G[0] * gx * gx + (double)G[3 * n3] * gy * gy + (double)G[5 * n3] * gz * gz +
                     2 * ((double)G[n3] * gx * gy + (double)G[2 * n3] * gx * gz + (double)G[4 * n3] * gy * gz);
              }
            else if (GEOM == 3)
              {
                T G[6];
                support_point_coefficients<T>(SL, B.qp[qx], B.qw[qx] * B.qw[qy] * B.qw[qz], G);
                s += (double)G[0] * gx * gx + (double)G[3] * gy * gy + (double)G[5] * gz * gz +
                     2 * ((double)G[1] * gx * gy + (double)G[2] * gx * gz + (double)G[4] * gy * gz);
              }
            else
              {
                T G[6];
                construct_q_coefficients<k, T>(geom + (size_t)cell * 3 * n3, qx, qy, qz, B, G);
                s += (double)G[0] * gx * gx + (double)G[3] * gy * gy + (double)G[5] * gz * gz +
                     2 * ((double)G[1] * gx * gy + (double)G[2] * gx * gz + (double)G[4] * gy * gz);
              }
          }
        }
    const uint32_t gi = plain ? plain[gid] : compressed_index<k>(cidx + cell * 27, ix, iy, iz);
    if (gi != DEV_INVALID)
      atomic_add(diag + gi, (T)s);
  }

  // ---- K4: FDM cell kernel (generic) -------------------------------------------------------------
  // patch size m (1-D).  IDX 0: m == k+1, 27 compressed indices;  IDX 1: explicit list pidx[cell][m^3].
  // weights (runtime WMODE): 0 none, 1 compressed cw[cell][27] (IDX 0 only), 2 per-entry
  //          wl[cell][m^3], 3 gathered from the global weight vector wvec[index]
  template <int m>
  __host__ __device__ constexpr int
  fdm_cells_per_block()
  {
    return m * m >= 64 ? (m * m > 100 ? 2 : 4) : (256 / (m * m) > 16 ? 16 : 256 / (m * m));
  }

  template <int m, typename T, int IDX>
  __global__ void __launch_bounds__(fdm_cells_per_block<m>() * m * m)
  fdm_generic_kernel(const T *__restrict__ src,
                     T *__restrict__ dst,
                     const uint32_t *__restrict__ idx,     // cidx (IDX 0) or pidx (IDX 1)
                     const uint32_t *__restrict__ inst,    // [cell*3+d] instance id
                     const T *__restrict__ Smat,           // [inst][m*m]
                     const T *__restrict__ lam,            // [inst][m]
                     const T *__restrict__ weights,        // per WMODE
                     const int       WMODE,
                     const int       w_pre,
                     const int       w_post,
                     const long long n_cells,
                     const uint32_t *__restrict__ cell_ids = nullptr) // work on the cells cell_ids[0 .. n_cells) (coloured launches)
  {
    constexpr int  m2 = m * m, m3 = m2 * m, CPB = fdm_cells_per_block<m>(), k = m - 1;
    constexpr bool VEC = fdm_rows_aligned<m, T>();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *smem = reinterpret_cast<T *>(smem_raw);

    const int       cl   = threadIdx.x / m2;
    const int       t    = threadIdx.x % m2;
    const int       a    = t % m;
    const int       b    = t / m;
    const long long slot = (long long)blockIdx.x * CPB + cl;
    const bool      act  = slot < n_cells;
    const long long cell = (act && cell_ids != nullptr) ? (long long)cell_ids[slot] : slot;

    T *U  = smem + (size_t)cl * (m3 + 3 * m2 + 3 * m);
    T *S0 = U + m3;
    T *S1 = S0 + m2;
    T *S2 = S1 + m2;
    T *L0 = S2 + m2;
    T *L1 = L0 + m;
    T *L2 = L1 + m;
    __shared__ uint32_t s_ci[CPB][27];
    if (act)
      {
        if (IDX == 0)
          for (int e = t; e < 27; e += m2)
            s_ci[cl][e] = idx[cell * 27 + e];
        const uint32_t i0 = inst[cell * 3 + 0], i1 = inst[cell * 3 + 1], i2 = inst[cell * 3 + 2];
        for (int e = t; e < m2; e += m2)
          {
            S0[e] = Smat[(size_t)i0 * m2 + e];
            S1[e] = Smat[(size_t)i1 * m2 + e];
            S2[e] = Smat[(size_t)i2 * m2 + e];
          }
        if (t < m)
          {
            L0[t] = lam[(size_t)i0 * m + t];
            L1[t] = lam[(size_t)i1 * m + t];
            L2[t] = lam[(size_t)i2 * m + t];
          }
      }
    __syncthreads();

    auto weight_of = [&](int x, int y, int z, uint32_t gi) -> T {
      if (WMODE == 1)
        {
          int ex, ey, ez, o;
          split_1d<k>(x, ex, o);
          split_1d<k>(y, ey, o);
          split_1d<k>(z, ez, o);
          return weights[cell * 27 + ex + 3 * ey + 9 * ez];
        }
      else if (WMODE == 2)
        return weights[(size_t)cell * m3 + (z * m + y) * m + x];
      else if (WMODE == 3)
        return weights[gi];
      return T(1);
    };

    if (act)
      {
#pragma unroll
        for (int x = 0; x < m; ++x)
          {
            uint32_t gi;
            if (IDX == 0)
              gi = compressed_index<k>(s_ci[cl], x, a, b);
            else
              gi = idx[(size_t)cell * m3 + (b * m + a) * m + x];
            T v = (gi == DEV_INVALID) ? T(0) : src[gi];
            if (WMODE != 0 && w_pre && gi != DEV_INVALID)
              v *= weight_of(x, a, b, gi);
            U[(b * m + a) * m + x] = v;
          }
      }
    __syncthreads();
    if (act)
      sweep_smem<m, T, 0, true, VEC>(S0, U, a, b);
    __syncthreads();
    if (act)
      sweep_smem<m, T, 1, true, VEC>(S1, U, a, b);
    __syncthreads();
    if (act)
      {
        // z sweep (transposed), scaling and forward z sweep on the same line (x=a, y=b)
        T v[m], r[m], row[m];
#pragma unroll
        for (int i = 0; i < m; ++i)
          v[i] = U[(i * m + b) * m + a];
#pragma unroll
        for (int o = 0; o < m; ++o)
          r[o] = 0;
#pragma unroll
        for (int i = 0; i < m; ++i)
          {
            load_row<m, T, VEC>(S2 + i * m, row);
#pragma unroll
            for (int o = 0; o < m; ++o)
              r[o] += row[o] * v[i];
          }
#pragma unroll
        for (int o = 0; o < m; ++o)
          r[o] = r[o] / (L0[a] + L1[b] + L2[o]);
#pragma unroll
        for (int o = 0; o < m; ++o)
          {
            load_row<m, T, VEC>(S2 + o * m, row);
            T s = 0;
#pragma unroll
            for (int i = 0; i < m; ++i)
              s += row[i] * r[i];
            U[(o * m + b) * m + a] = s;
          }
      }
    __syncthreads();
    if (act)
      sweep_smem<m, T, 1, false, VEC>(S1, U, a, b);
    __syncthreads();
    if (act)
      sweep_smem<m, T, 0, false, VEC>(S0, U, a, b);
    // the x sweep wrote the line (y=a,z=b) that the same thread scatters: no sync needed
    if (act)
      {
#pragma unroll
        for (int x = 0; x < m; ++x)
          {
            uint32_t gi;
            if (IDX == 0)
              gi = compressed_index<k>(s_ci[cl], x, a, b);
            else
              gi = idx[(size_t)cell * m3 + (b * m + a) * m + x];
            if (gi != DEV_INVALID)
              {
                T v = U[(b * m + a) * m + x];
                if (WMODE != 0 && w_post)
                  v *= weight_of(x, a, b, gi);
                atomic_add(dst + gi, v);
              }
          }
      }
  }

  // ---- K6: vector epilogues ----------------------------------------------------------------------
  template <typename T>
  __global__ void
  vec_residual_kernel(T *__restrict__ dst, const T *__restrict__ b, const long long n)
  {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
      dst[i] = b[i] - dst[i];
  }

  template <typename T>
  __global__ void
  vec_cheb_update_kernel(T *out, const T *dst, const T *x, const T *xold, const T f1, const T f2, const long long n)
  {
    // x+ = x + f1 (x - x_old) + f2 z   (x_old == nullptr means x_old = 0)
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
      {
        const T xv = x[i];
        T       r  = xv + f2 * dst[i];
        if (f1 != T(0))
          r += f1 * (xv - (xold != nullptr ? xold[i] : T(0)));
        out[i] = r;
      }
  }

  template <typename T>
  __global__ void
  vec_scale_kernel(T *out, const T *in, const T f, const long long n)
  {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
      out[i] = f * in[i];
  }

  template <typename T>
  __global__ void
  vec_mul_kernel(T *out, const T *a, const T *b, const long long n)
  {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
      out[i] = a[i] * b[i];
  }

  template <typename T>
  __global__ void
  vec_copy_indexed_kernel(T *dst, const T *src, const uint32_t *idx, const long long n)
  {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
      dst[idx[i]] = src[idx[i]];
  }

  template <typename T>
  __global__ void
  vec_set_indexed_kernel(T *dst, const T v, const uint32_t *idx, const long long n)
  {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
      dst[idx[i]] = v;
  }

  template <typename T>
  __global__ void
  vec_invert_diag_kernel(T *d, const long long n)
  {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
      {
        const T v = d[i];
        d[i]      = (fabs((double)v) > 1e-10) ? T(1) / v : T(1);
      }
  }

  template <typename TO, typename TI>
  __global__ void
  vec_convert_kernel(TO *out, const TI *in, const long long n)
  {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
      out[i] = (TO)in[i];
  }

  // dot product / norms: per-block partial sums in double, finished by a second tiny kernel
  template <typename T>
  __global__ void
  vec_dot_kernel(const T *__restrict__ a, const T *__restrict__ b, double *partial, const long long n)
  {
    __shared__ double sh[32];
    double            s = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      s += (double)a[i] * (double)b[i];
    for (int o = 16; o > 0; o >>= 1)
      s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0)
      sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32)
      {
        s = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1)
          s += __shfl_down_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0)
          partial[blockIdx.x] = s;
      }
  }

  static __global__ void
  reduce_partials_kernel(const double *partial, double *out, const int n)
  {
    __shared__ double sh[32];
    double            s = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      s += partial[i];
    for (int o = 16; o > 0; o >>= 1)
      s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0)
      sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32)
      {
        s = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1)
          s += __shfl_down_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0)
          out[0] = s;
      }
  }

  // pack / unpack of ghost-exchange buffers: ranges (start,len) listed per entity
  template <typename T>
  __global__ void
  pack_kernel(T *__restrict__ buf, const T *__restrict__ vec, const uint32_t *__restrict__ map, const long long n)
  {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
      buf[i] = vec[map[i]];
  }

  template <typename T, bool ADD>
  __global__ void
  unpack_kernel(T *__restrict__ vec, const T *__restrict__ buf, const uint32_t *__restrict__ map, const long long n)
  {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
      {
        if (ADD)
          atomic_add(vec + map[i], buf[i]); // an owned DoF on a partition edge / corner receives from several peers
        else
          vec[map[i]] = buf[i];
      }
  }

  // ---- device-initiated halo exchange over NVLink (peer memory mapped with CUDA IPC) -----------------------------------------
  // push:  every rank writes the values of its send list straight into the receive buffer of the peer (remote stores) and,
  //        when all blocks are done, publishes the sequence number of the message in the peer's flag word;
  // pull:  the receiver waits for the flag words of its peers and unpacks its own receive buffer (copy for a ghost update,
  //        red.add for a compress), then acknowledges.  Receive buffers are double buffered by the parity of the sequence number.
  constexpr int P2P_MAX_PEERS = 32;
  struct P2PPeers
  {
    int                 n;
    void *              buf[P2P_MAX_PEERS];      // peer's receive buffer of the current parity (mapped)
    unsigned long long *flag[P2P_MAX_PEERS];     // peer's flag word for messages from this rank (mapped)
    unsigned long long *ack[P2P_MAX_PEERS];      // peer's acknowledge word for this rank's reads of ITS messages (mapped)
    long long           dst_off[P2P_MAX_PEERS];  // element offset of this rank's segment in the peer's receive buffer
    long long           seg_begin[P2P_MAX_PEERS + 1]; // segments of the local send list per peer
    unsigned long long *my_flag[P2P_MAX_PEERS];  // local flag words written by the peers
    unsigned long long *my_ack[P2P_MAX_PEERS];   // local acknowledge words written by the peers
  };

  __device__ __forceinline__ unsigned long long
  ld_sys(const unsigned long long *p)
  {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
  }
  __device__ __forceinline__ void
  st_sys(unsigned long long *p, const unsigned long long v)
  {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
  }

  template <typename T>
  __global__ void
  p2p_push_kernel(const T *__restrict__ vec, const uint32_t *__restrict__ map, const P2PPeers peers, const unsigned long long seq,
                  unsigned int *done_counter)
  {
    // the peers must have consumed the message that used this half of their receive buffer (sequence number seq - 2)
    if (threadIdx.x == 0 && seq > 2)
      for (int q = 0; q < peers.n; ++q)
        while (ld_sys(peers.my_ack[q]) + 2 < seq)
          ;
    __syncthreads();
    const long long n = peers.seg_begin[peers.n];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      {
        int q = 0;
        while (i >= peers.seg_begin[q + 1])
          ++q;
        reinterpret_cast<T *>(peers.buf[q])[peers.dst_off[q] + (i - peers.seg_begin[q])] = vec[map[i]];
      }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
      {
        const unsigned int prev = atomicAdd(done_counter, 1u);
        if (prev == gridDim.x - 1)
          {
            *done_counter = 0;
            __threadfence_system();
            for (int q = 0; q < peers.n; ++q)
              st_sys(peers.flag[q], seq);
          }
      }
  }

  template <typename T, bool ADD>
  __global__ void
  p2p_pull_kernel(T *__restrict__ vec, const T *__restrict__ buf, const uint32_t *__restrict__ map, const long long n, const P2PPeers peers,
                  const unsigned long long seq, unsigned int *done_counter)
  {
    if (threadIdx.x == 0)
      for (int q = 0; q < peers.n; ++q)
        while (ld_sys(peers.my_flag[q]) < seq)
          ;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      {
        const T v = __ldcv(buf + i);
        if (ADD)
          atomic_add(vec + map[i], v);
        else
          vec[map[i]] = v;
      }
    __syncthreads();
    if (threadIdx.x == 0)
      {
        const unsigned int prev = atomicAdd(done_counter, 1u);
        if (prev == gridDim.x - 1)
          {
            *done_counter = 0;
            __threadfence_system();
            for (int q = 0; q < peers.n; ++q)
              st_sys(peers.ack[q], seq);
          }
      }
  }
} // namespace dasm
