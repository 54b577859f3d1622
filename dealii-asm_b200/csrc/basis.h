// 1-D finite element / quadrature tables and small dense linear algebra (host side).
//
// Restates the deal.II pieces the reference relies on for the hot path:
//   FE_Q(k) on Gauss-Lobatto support points, QGauss(k+1), the reference 1-D mass/stiffness
//   matrices (deal.II internal::create_reference_mass_and_stiffness_matrices; used through
//   include/matrix_free.h:350-363 of the reference) and the generalized symmetric eigenproblem
//   behind TensorProductMatrixSymmetricSum (LAPACK sygv in deal.II; a Cholesky + cyclic Jacobi
//   solver here because no LAPACK is available).
#pragma once
#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <vector>

namespace dasm
{
  constexpr int MAX_DEGREE = 8;

  inline void
  gauss_points(int n, std::vector<double> &x, std::vector<double> &w)
  {
    x.assign(n, 0.);
    w.assign(n, 0.);
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int i = 0; i < (n + 1) / 2; ++i)
      {
        long double z = std::cos(pi * (i + 0.75L) / (n + 0.5L)), pp = 0;
        for (int it = 0; it < 100; ++it)
          {
            long double p1 = 1, p2 = 0;
            for (int j = 0; j < n; ++j)
              {
                const long double p3 = p2;
                p2                   = p1;
                p1                   = ((2 * j + 1) * z * p2 - j * p3) / (j + 1);
              }
            pp                   = n * (z * p1 - p2) / (z * z - 1);
            const long double z1 = z;
            z                    = z1 - p1 / pp;
            if (std::fabs((double)(z - z1)) < 1e-19)
              break;
          }
        x[i]         = (double)(0.5L * (1 - z));
        x[n - 1 - i] = (double)(0.5L * (1 + z));
        w[i] = w[n - 1 - i] = (double)(1.0L / ((1 - z * z) * pp * pp));
      }
  }

  // Gauss-Lobatto points on [0,1] (support points of FE_Q(n-1))
  inline std::vector<double>
  gauss_lobatto_points(int n)
  {
    std::vector<double> x(n);
    if (n == 1)
      {
        x[0] = 0.5;
        return x;
      }
    x[0]     = 0;
    x[n - 1] = 1;
    const int         N  = n - 1; // interior points: roots of P'_N
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int i = 1; i < n - 1; ++i)
      {
        long double z = -std::cos(pi * i / N);
        for (int it = 0; it < 200; ++it)
          {
            // P_N, P'_N, P''_N
            long double p0 = 1, p1 = z;
            for (int j = 2; j <= N; ++j)
              {
                const long double p2 = ((2 * j - 1) * z * p1 - (j - 1) * p0) / j;
                p0                   = p1;
                p1                   = p2;
              }
            const long double dp  = N * (z * p1 - p0) / (z * z - 1);
            const long double ddp = (2 * z * dp - (long double)N * (N + 1) * p1) / (1 - z * z);
            const long double dz  = dp / ddp;
            z -= dz;
            if (std::fabs((double)dz) < 1e-19)
              break;
          }
        x[i] = (double)(0.5L * (1 + z));
      }
    for (int i = 0; i < n / 2; ++i)
      { // symmetrise
        const double a = 0.5 * (x[i] + (1 - x[n - 1 - i]));
        x[i]           = a;
        x[n - 1 - i]   = 1 - a;
      }
    if (n % 2)
      x[n / 2] = 0.5;
    return x;
  }

  // V[q*n+i] = l_i(x_q), D[q*n+i] = l_i'(x_q)
  inline void
  lagrange(const std::vector<double> &nodes, const std::vector<double> &x, std::vector<double> &V, std::vector<double> &D)
  {
    const int n = nodes.size(), nq = x.size();
    V.assign(nq * n, 0.);
    D.assign(nq * n, 0.);
    for (int i = 0; i < n; ++i)
      {
        long double denom = 1;
        for (int j = 0; j < n; ++j)
          if (j != i)
            denom *= (long double)nodes[i] - nodes[j];
        for (int q = 0; q < nq; ++q)
          {
            long double v = 1;
            for (int j = 0; j < n; ++j)
              if (j != i)
                v *= (long double)x[q] - nodes[j];
            long double s = 0;
            for (int m = 0; m < n; ++m)
              {
                if (m == i)
                  continue;
                long double p = 1;
                for (int j = 0; j < n; ++j)
                  if (j != i && j != m)
                    p *= (long double)x[q] - nodes[j];
                s += p;
              }
            V[q * n + i] = (double)(v / denom);
            D[q * n + i] = (double)(s / denom);
          }
      }
  }

  struct Basis1D
  {
    int                 k, n;
    std::vector<double> nodes, qp, qw;
    std::vector<double> N, D;   // [q*n+i] nodal basis / derivative at Gauss points
    std::vector<double> Dq;     // [q*n+p] collocation derivative in the Gauss-point basis
    std::vector<double> M_ref, K_ref; // [i*n+j] reference mass / stiffness on the unit interval

    explicit Basis1D(int degree)
      : k(degree)
      , n(degree + 1)
    {
      nodes = gauss_lobatto_points(n);
      gauss_points(n, qp, qw);
      lagrange(nodes, qp, N, D);
      std::vector<double> tmp;
      lagrange(qp, qp, tmp, Dq);
      M_ref.assign(n * n, 0.);
      K_ref.assign(n * n, 0.);
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
          {
            long double m = 0, s = 0;
            for (int q = 0; q < n; ++q)
              {
                m += (long double)N[q * n + i] * N[q * n + j] * qw[q];
                s += (long double)D[q * n + i] * D[q * n + j] * qw[q];
              }
            M_ref[i * n + j] = (double)m;
            K_ref[i * n + j] = (double)s;
          }
    }
  };

  // Generalized symmetric eigenproblem K s = lambda M s with S^T M S = I on the rows whose mass
  // diagonal is non-zero; constrained rows (zero mass diagonal, i.e. cleared Dirichlet rows) get a
  // zero eigenvector row/column and eigenvalue 1 (deal.II TensorProductMatrixSymmetricSum
  // spectral assembly, restated).  S is returned row-major S[i*m+a] (column a = eigenvector a).
  inline void
  generalized_eig(int m, const std::vector<double> &M, const std::vector<double> &K, std::vector<double> &S, std::vector<double> &lam)
  {
    std::vector<int> idx;
    for (int i = 0; i < m; ++i)
      if (M[i * m + i] != 0.0)
        idx.push_back(i);
    const int r = idx.size();
    S.assign(m * m, 0.);
    lam.assign(m, 1.);
    if (r == 0)
      return;
    using ld = long double;
    std::vector<ld> L(r * r, 0), C(r * r, 0), V(r * r, 0);
    // Cholesky M = L L^T
    for (int i = 0; i < r; ++i)
      for (int j = 0; j <= i; ++j)
        {
          ld s = M[idx[i] * m + idx[j]];
          for (int t = 0; t < j; ++t)
            s -= L[i * r + t] * L[j * r + t];
          if (i == j)
            {
              if (s <= 0)
                throw std::runtime_error("generalized_eig: mass matrix not positive definite");
              L[i * r + i] = std::sqrt(s);
            }
          else
            L[i * r + j] = s / L[j * r + j];
        }
    // C = L^-1 K L^-T
    std::vector<ld> T(r * r, 0);
    for (int c = 0; c < r; ++c) // solve L T(:,c) = K(:,c)
      for (int i = 0; i < r; ++i)
        {
          ld s = K[idx[i] * m + idx[c]];
          for (int t = 0; t < i; ++t)
            s -= L[i * r + t] * T[t * r + c];
          T[i * r + c] = s / L[i * r + i];
        }
    for (int i = 0; i < r; ++i) // C^T rows: solve L C(:,i)^T = T(i,:)^T  => C = T L^-T
      for (int c = 0; c < r; ++c)
        {
          ld s = T[i * r + c];
          for (int t = 0; t < c; ++t)
            s -= C[i * r + t] * L[c * r + t];
          C[i * r + c] = s / L[c * r + c];
        }
    for (int i = 0; i < r; ++i)
      for (int j = 0; j < i; ++j)
        C[i * r + j] = C[j * r + i] = 0.5L * (C[i * r + j] + C[j * r + i]);
    // cyclic Jacobi
    for (int i = 0; i < r; ++i)
      V[i * r + i] = 1;
    for (int sweep = 0; sweep < 100; ++sweep)
      {
        ld off = 0, dia = 0;
        for (int i = 0; i < r; ++i)
          for (int j = 0; j < r; ++j)
            (i == j ? dia : off) += C[i * r + j] * C[i * r + j];
        if (off <= 1e-38L * dia || off == 0)
          break;
        for (int p = 0; p < r; ++p)
          for (int q = p + 1; q < r; ++q)
            {
              const ld apq = C[p * r + q];
              if (apq == 0)
                continue;
              const ld theta = (C[q * r + q] - C[p * r + p]) / (2 * apq);
              const ld t     = (theta >= 0 ? 1 : -1) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
              const ld c = 1 / std::sqrt(t * t + 1), s = t * c;
              for (int i = 0; i < r; ++i)
                {
                  const ld a = C[i * r + p], b = C[i * r + q];
                  C[i * r + p] = c * a - s * b;
                  C[i * r + q] = s * a + c * b;
                }
              for (int i = 0; i < r; ++i)
                {
                  const ld a = C[p * r + i], b = C[q * r + i];
                  C[p * r + i] = c * a - s * b;
                  C[q * r + i] = s * a + c * b;
                }
              for (int i = 0; i < r; ++i)
                {
                  const ld a = V[i * r + p], b = V[i * r + q];
                  V[i * r + p] = c * a - s * b;
                  V[i * r + q] = s * a + c * b;
                }
            }
      }
    // sort ascending
    std::vector<int> order(r);
    for (int i = 0; i < r; ++i)
      order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return C[a * r + a] < C[b * r + b]; });
    // S = L^-T V
    for (int a = 0; a < r; ++a)
      {
        const int       col = order[a];
        std::vector<ld> y(r);
        for (int i = r - 1; i >= 0; --i)
          {
            ld s = V[i * r + col];
            for (int t = i + 1; t < r; ++t)
              s -= L[t * r + i] * y[t];
            y[i] = s / L[i * r + i];
          }
        for (int i = 0; i < r; ++i)
          S[idx[i] * m + idx[a]] = (double)y[i];
        lam[idx[a]] = (double)C[col * r + col];
      }
  }
} // namespace dasm
