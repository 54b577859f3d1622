/* CPU restatement (plain C, OpenMP) of the smoother hot path - TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * Nothing in the product path may link or call this file; it is used by tests/ (cross-check against the
 * numpy oracle), by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs, where it
 * stands in for the reference's MPI CPU path (deal.II cannot be built here).
 *
 * Restates (file:line into the reference tree):
 *   cell integral evaluate(grad) -> G -> integrate(grad)   include/operator.h:866-875 (merged: 1161-1219)
 *   gather / scatter-add through explicit index lists       include/operator.h:1335-1351
 *   FDM apply_inverse + weights                             include/matrix_free.h:1023-1062
 *   Chebyshev term  x+ = x + f1 (x - x_old) + f2 P^-1 (b - A x)   [deal.II PreconditionChebyshev]
 *
 * Cells are processed colour by colour (cells of one colour share no DoF), each colour in parallel
 * over the host threads; this replaces the reference's one-MPI-rank-per-core partition.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXN 9
#define MAXN3 (MAXN * MAXN * MAXN)

int oracle_c_max_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* out = M(nxn, row-major [o][i]) applied along direction dir of the n^3 tensor (x fastest) */
static void apply_1d(int n, const double *M, int trans, const double *in, double *out, int dir)
{
  const int stride = dir == 0 ? 1 : (dir == 1 ? n : n * n);
  for (int b = 0; b < n; ++b)
    for (int a = 0; a < n; ++a)
      {
        int base;
        if (dir == 0)
          base = (b * n + a) * n;
        else if (dir == 1)
          base = b * n * n + a;
        else
          base = b * n + a;
        double v[MAXN];
        for (int i = 0; i < n; ++i)
          v[i] = in[base + i * stride];
        for (int o = 0; o < n; ++o)
          {
            double s = 0;
            for (int i = 0; i < n; ++i)
              s += (trans ? M[i * n + o] : M[o * n + i]) * v[i];
            out[base + o * stride] = s;
          }
      }
}

/* y += A x over the listed cells.  idx[c*n3+i] global index or 0xFFFFFFFF; G[c][6][n3] merged coefficients */
void oracle_c_vmult_cells(int n, const double *N, const double *Dq, const uint32_t *idx, const double *G,
                          const int64_t *cells, int64_t n_list, const double *x, double *y)
{
  const int n3 = n * n * n;
#pragma omp parallel for schedule(static)
  for (int64_t l = 0; l < n_list; ++l)
    {
      const int64_t   c  = cells[l];
      const uint32_t *ci = idx + c * n3;
      const double *  g  = G + c * 6 * n3;
      double          u[MAXN3], t[MAXN3], gx[MAXN3], gy[MAXN3], gz[MAXN3];
      for (int i = 0; i < n3; ++i)
        u[i] = ci[i] == 0xFFFFFFFFu ? 0.0 : x[ci[i]];
      apply_1d(n, N, 0, u, t, 0);
      apply_1d(n, N, 0, t, u, 1);
      apply_1d(n, N, 0, u, t, 2);
      apply_1d(n, Dq, 0, t, gx, 0);
      apply_1d(n, Dq, 0, t, gy, 1);
      apply_1d(n, Dq, 0, t, gz, 2);
      for (int q = 0; q < n3; ++q)
        {
          const double a = gx[q], b = gy[q], cc = gz[q];
          gx[q] = g[q] * a + g[n3 + q] * b + g[2 * n3 + q] * cc;
          gy[q] = g[n3 + q] * a + g[3 * n3 + q] * b + g[4 * n3 + q] * cc;
          gz[q] = g[2 * n3 + q] * a + g[4 * n3 + q] * b + g[5 * n3 + q] * cc;
        }
      apply_1d(n, Dq, 1, gx, u, 0);
      apply_1d(n, Dq, 1, gy, t, 1);
      for (int i = 0; i < n3; ++i)
        u[i] += t[i];
      apply_1d(n, Dq, 1, gz, t, 2);
      for (int i = 0; i < n3; ++i)
        u[i] += t[i];
      apply_1d(n, N, 1, u, t, 2);
      apply_1d(n, N, 1, t, u, 1);
      apply_1d(n, N, 1, u, t, 0);
      for (int i = 0; i < n3; ++i)
        if (ci[i] != 0xFFFFFFFFu)
          y[ci[i]] += t[i];
    }
}

/* z += sum_c W R^T A_c^-1 R W r over the listed cells (patch = cell closure, n_overlap = 1).
 * S[c][3][n*n] eigenvectors (row-major, column = eigenvector), lam[c][3][n], wl[c][n3] local weights or NULL */
void oracle_c_fdm_cells(int n, const uint32_t *idx, const double *S, const double *lam, const double *wl, int w_pre,
                        int w_post, const int64_t *cells, int64_t n_list, const double *r, double *z)
{
  const int n3 = n * n * n, n2 = n * n;
#pragma omp parallel for schedule(static)
  for (int64_t l = 0; l < n_list; ++l)
    {
      const int64_t   c  = cells[l];
      const uint32_t *ci = idx + c * n3;
      const double *  Sc = S + c * 3 * n2;
      const double *  lc = lam + c * 3 * n;
      const double *  w  = wl ? wl + c * n3 : 0;
      double          u[MAXN3], t[MAXN3];
      for (int i = 0; i < n3; ++i)
        {
          u[i] = ci[i] == 0xFFFFFFFFu ? 0.0 : r[ci[i]];
          if (w && w_pre)
            u[i] *= w[i];
        }
      apply_1d(n, Sc, 1, u, t, 0);
      apply_1d(n, Sc + n2, 1, t, u, 1);
      apply_1d(n, Sc + 2 * n2, 1, u, t, 2);
      for (int k = 0; k < n; ++k)
        for (int j = 0; j < n; ++j)
          for (int i = 0; i < n; ++i)
            t[(k * n + j) * n + i] /= (lc[i] + lc[n + j] + lc[2 * n + k]);
      apply_1d(n, Sc, 0, t, u, 0);
      apply_1d(n, Sc + n2, 0, u, t, 1);
      apply_1d(n, Sc + 2 * n2, 0, t, u, 2);
      for (int i = 0; i < n3; ++i)
        if (ci[i] != 0xFFFFFFFFu)
          z[ci[i]] += (w && w_post) ? u[i] * w[i] : u[i];
    }
}

void oracle_c_residual(int64_t n, const double *b, double *t)
{
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i)
    t[i] = b[i] - t[i];
}

void oracle_c_cheb_update(int64_t n, double f1, double f2, const double *x, const double *xold, const double *z, double *out)
{
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i)
    out[i] = x[i] + f1 * (x[i] - (xold ? xold[i] : 0.0)) + f2 * z[i];
}

void oracle_c_zero(int64_t n, double *v)
{
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i)
    v[i] = 0.0;
}
