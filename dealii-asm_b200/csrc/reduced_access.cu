// Orientation-aware compressed vector access on the device: ConstraintInfoReduced::read_dof_values / distribute_local_to_global
// (reference include/vector_access_reduced.h:267-548) = expansion of the 3^3 start indices of a cell to (k+1)^3 addresses followed by
// adjust_for_orientation (include/reduced_access.h:528-702) with the packed orientation word (12 line bits + 6 x 3 quad bits).
// The structured meshes of this library only produce the standard orientation (word 0), for which the tuned kernels decode the
// indices inline; this translation unit is the general form an unstructured (ball) mesh layer needs, checked against the reference's
// reduced_access_02.result through tests/test_gpu_parity.py.  The reorientation is applied as an index permutation: thread i of a
// cell computes the standard-layout position whose value lands at local position i.
#include <cuda_runtime.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "dasm.h"

namespace
{
  constexpr uint32_t INVALID = 0xFFFFFFFFu;

  // inverse orientation tables of the quads: inv[flag][p] = q with table[flag][q] = p (ShapeInfo::compute_orientation_table; flag 0
  // standard, 1 transposed - pinned by the reference's golden -, 2..7 as recollected from deal.II)
  struct QuadTables
  {
    unsigned char inv[8][49];
  };

  __device__ __forceinline__ uint32_t
  decode(const uint32_t *__restrict__ ci, const int k, const int x, const int y, const int z)
  {
    const int      ex = (x == 0) ? 0 : ((x == k) ? 2 : 1), ey = (y == 0) ? 0 : ((y == k) ? 2 : 1), ez = (z == 0) ? 0 : ((z == k) ? 2 : 1);
    const int      ox = (ex == 1) ? x - 1 : 0, oy = (ey == 1) ? y - 1 : 0, oz = (ez == 1) ? z - 1 : 0;
    const uint32_t start = ci[ex + 3 * ey + 9 * ez];
    if (start == INVALID)
      return INVALID;
    const int sx = (ex == 1) ? (k - 1) : 1, sy = (ey == 1) ? (k - 1) : 1;
    return start + ox + sx * (oy + sy * oz);
  }

  // standard-layout position (x, y, z) whose value belongs at local position (x, y, z) under the orientation word
  __device__ __forceinline__ void
  oriented_source(const int k, uint32_t orientation, const QuadTables &qt, int &x, int &y, int &z)
  {
    if (orientation == 0u)
      return;
    const int  ex = (x == 0) ? 0 : ((x == k) ? 2 : 1), ey = (y == 0) ? 0 : ((y == k) ? 2 : 1), ez = (z == 0) ? 0 : ((z == k) ? 2 : 1);
    const int  n_interior = (ex == 1) + (ey == 1) + (ez == 1);
    if (n_interior == 1)
      {
        // line: numbering of adjust_for_orientation: lines 0-3 run in x (stride 1) at (y, z) = (0,0) (k,0) (0,k) (k,k); 4-7 in y at
        // (x, z) = (0,0) (k,0) (0,k) (k,k); 8-11 in z at (x, y) = (0,0) (k,0) (0,k) (k,k)
        int l;
        if (ex == 1)
          l = (ey == 2 ? 1 : 0) + (ez == 2 ? 2 : 0);
        else if (ey == 1)
          l = 4 + (ex == 2 ? 1 : 0) + (ez == 2 ? 2 : 0);
        else
          l = 8 + (ex == 2 ? 1 : 0) + (ey == 2 ? 2 : 0);
        if ((orientation >> l) & 1u)
          {
            if (ex == 1)
              x = k - x;
            else if (ey == 1)
              y = k - y;
            else
              z = k - z;
          }
      }
    else if (n_interior == 2)
      {
        // quads: q = 2 d + side, d the normal direction; (i0, i1) run over (y, z) for d = 0, (x, z) for d = 1, (x, y) for d = 2
        const int      d    = (ex != 1) ? 0 : ((ey != 1) ? 1 : 2);
        const int      side = (d == 0 ? ex : (d == 1 ? ey : ez)) == 2 ? 1 : 0;
        const uint32_t flag = (orientation >> (12 + 3 * (2 * d + side))) & 7u;
        if (flag != 0u)
          {
            int &     a = (d == 0) ? y : x;
            int &     b = (d == 2) ? y : z;
            const int m = k - 1, p = (a - 1) + (b - 1) * m, q = qt.inv[flag][p];
            a           = q % m + 1;
            b           = q / m + 1;
          }
      }
  }

  template <typename T, bool DISTRIBUTE>
  __global__ void
  reduced_access_kernel(T *__restrict__ vec, T *__restrict__ local, const uint32_t *__restrict__ cidx, const uint32_t *__restrict__ orientation,
                        const int k, const long long n_cells, const __grid_constant__ QuadTables qt)
  {
    const int       n = k + 1, n3 = n * n * n;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_cells * n3)
      return;
    const long long c = t / n3;
    const int       i = (int)(t % n3);
    int             x = i % n, y = (i / n) % n, z = i / (n * n);
    oriented_source(k, orientation ? orientation[c] : 0u, qt, x, y, z);
    const uint32_t g = decode(cidx + c * 27, k, x, y, z);
    if (DISTRIBUTE)
      {
        if (g != INVALID)
          atomicAdd(vec + g, local[t]);
      }
    else
      local[t] = (g == INVALID) ? T(0) : vec[g];
  }

  QuadTables
  make_tables(const int k)
  {
    QuadTables qt;
    const int  n = k - 1;
    for (int f = 0; f < 8; ++f)
      for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i)
          {
            int p;
            switch (f)
              {
                case 0: p = i + j * n; break;
                case 1: p = j + i * n; break;
                case 2: p = j + (n - 1 - i) * n; break;
                case 3: p = i + (n - 1 - j) * n; break;
                case 4: p = (n - 1 - i) + (n - 1 - j) * n; break;
                case 5: p = (n - 1 - j) + (n - 1 - i) * n; break;
                case 6: p = (n - 1 - j) + i * n; break;
                default: p = (n - 1 - i) + j * n; break;
              }
            // adjust_for_orientation (evaluate): temp[table[f][q]] = v[pos(q)], v[pos(p)] = temp[p]  =>  source of p is q
            qt.inv[f][p] = (unsigned char)(i + j * n);
          }
    return qt;
  }

  template <bool DISTRIBUTE>
  int
  run(int degree, int number_type, const uint32_t *d_cidx, const uint32_t *d_orientation, long long n_cells, void *vec, void *local, void *stream)
  {
    try
      {
        if (degree < 1 || degree > 8)
          throw std::runtime_error("degree must be in 1..8");
        const QuadTables qt    = make_tables(degree);
        const long long  total = n_cells * (degree + 1) * (degree + 1) * (degree + 1);
        const unsigned   nb    = (unsigned)((total + 255) / 256);
        if (total > 0)
          {
            if (number_type == DASM_F64)
              reduced_access_kernel<double, DISTRIBUTE>
                <<<nb, 256, 0, (cudaStream_t)stream>>>((double *)vec, (double *)local, d_cidx, d_orientation, degree, n_cells, qt);
            else
              reduced_access_kernel<float, DISTRIBUTE>
                <<<nb, 256, 0, (cudaStream_t)stream>>>((float *)vec, (float *)local, d_cidx, d_orientation, degree, n_cells, qt);
          }
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess)
          throw std::runtime_error(std::string("reduced_access_kernel: ") + cudaGetErrorString(e));
        return 0;
      }
    catch (const std::exception &e)
      {
        dasm_set_last_error(e.what());
        return 1;
      }
  }
} // namespace

extern "C" int
dasm_reduced_access_read(int degree, int number_type, const uint32_t *d_cidx, const uint32_t *d_orientation, long long n_cells, const void *src,
                         void *local, void *stream)
{
  return run<false>(degree, number_type, d_cidx, d_orientation, n_cells, const_cast<void *>(src), local, stream);
}

extern "C" int
dasm_reduced_access_distribute(int degree, int number_type, const uint32_t *d_cidx, const uint32_t *d_orientation, long long n_cells, void *dst,
                               const void *local, void *stream)
{
  return run<true>(degree, number_type, d_cidx, d_orientation, n_cells, dst, const_cast<void *>(local), stream);
}
