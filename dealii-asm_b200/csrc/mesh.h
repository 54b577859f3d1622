// Structured hexahedral mesh (Cartesian / sine-deformed / Kershaw), its brick partition over ranks,
// the native DoF numbering with 3^dim-compressed indices, geometry factors and harmonic extents.
//
// Reference behaviour restated (file:line into the reference tree):
//   mesh + mapping of the benchmark driver      matrix_free_loop_08.likwid.cc:160-199
//   Kershaw map                                 include/kershaw.h:4-80
//   3^dim compressed DoF indices per cell       include/vector_access_reduced.h:30-164,
//                                               include/reduced_access.h:154-285
//   merged geometry coefficients                include/operator.h:674-711
//   harmonic cell / patch extents               include/grid_tools.h:11-138
//
// Numbering ("brick-grouped owner-cell numbering", the data-locality numbering of this library; the
// reference uses DoFRenumbering::matrix_free_data_locality, matrix_free_loop_08.likwid.cc:216-222, for
// the same purpose): every mesh entity (vertex/line/quad/hex interior) is owned by the cell for which it
// is a "lower" entity (or an upper-boundary entity of the last cell in a non-periodic direction).  Cells
// are processed brick-major (4x4x4 bricks); inside a brick the entities touched only by cells of the
// brick are numbered first (cell by cell, lexicographic entity order), then the entities on lower brick
// faces shared with neighbouring bricks.  All DoFs of an entity are contiguous, so a cell needs only 27
// start indices, and the DoFs a brick shares with its neighbours are one contiguous range.
//
// "Lex" bricks: a full 4 x 4 x 4 brick whose six faces all have neighbour cells (interior or periodic; hence no
// constrained DoFs) owns the half-open box [0, 4k)^3 of its tile.  Its DoFs are numbered LEXICOGRAPHICALLY in
// that box (x fastest: base + X + 4k Y + 16k^2 Z) and all lex bricks come first in the owned range, so the
// box of brick i is the i-th block of 64 k^3 entries of every vector: a 4-D tensor (x, y, z, brick) that the TMA
// engine moves with box copies (kernels_tma.cuh).  The 27 start indices of a cell stay; an entity that lives in a
// lex brick has bit 31 (LEX_FLAG) set in its start index, which then is the index of its first DoF, and its DoFs
// are expanded with the strides (1, 4k, 16k^2) instead of the entity-contiguous ones.
#pragma once
#include <array>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <set>
#include <stdexcept>
#include <vector>

#include "basis.h"

namespace dasm
{
  constexpr uint32_t INVALID_INDEX = 0xFFFFFFFFu;
  constexpr uint32_t LEX_FLAG      = 0x80000000u; // start index of an entity stored in a lex brick (INVALID_INDEX is tested first)

  // index of local DoF (x, y, z) of a cell from its 27 start indices (host twin of compressed_index<k>, kernels.cuh)
  inline uint32_t
  expand_start_index(const uint32_t *ci, const int k, const int x, const int y, const int z)
  {
    const int      ex = x == 0 ? 0 : (x == k ? 2 : 1), ey = y == 0 ? 0 : (y == k ? 2 : 1), ez = z == 0 ? 0 : (z == k ? 2 : 1);
    const uint32_t st = ci[ex + 3 * ey + 9 * ez];
    if (st == INVALID_INDEX)
      return INVALID_INDEX;
    const int ox = ex == 1 ? x - 1 : 0, oy = ey == 1 ? y - 1 : 0, oz = ez == 1 ? z - 1 : 0;
    if (st & LEX_FLAG)
      return (st & ~LEX_FLAG) + ox + 4 * k * (oy + 4 * k * oz);
    const int sx = ex == 1 ? k - 1 : 1, sy = ey == 1 ? k - 1 : 1;
    return st + ox + sx * (oy + sy * oz);
  }

  enum MapKind
  {
    MAP_CARTESIAN = 0,
    MAP_SINE      = 1,
    MAP_KERSHAW   = 2
  };

  struct MeshParams
  {
    int    nc[3]       = {1, 1, 1};  // global cells per direction
    int    periodic[3] = {0, 0, 0};
    int    dirichlet   = 1;          // non-periodic boundaries: 1 = homogeneous Dirichlet, 0 = natural
    double length[3]   = {1, 1, 1};
    int    map_kind    = MAP_CARTESIAN;
    double map_par[4]  = {0, 0, 0, 0};
    int    part[3]     = {1, 1, 1};  // ranks per direction (brick partition)
    int    rank        = 0;
    int    brick[3]    = {4, 4, 4};  // cells per thread-block brick
  };

  inline double
  kershaw_right(double eps, double x)
  {
    return (x <= 0.5) ? (2 - eps) * x : 1 + eps * (x - 1);
  }
  inline double
  kershaw_left(double eps, double x)
  {
    return 1 - kershaw_right(eps, 1 - x);
  }
  inline double
  kershaw_step(double x)
  {
    if (x <= 0)
      return 0;
    if (x >= 1)
      return 1;
    return ((6 * x - 15) * x + 10) * x * x * x;
  }

  struct ExchangeList
  {
    int                   peer = -1;
    std::vector<uint32_t> send_start, send_len; // owned ranges this rank sends for ghost update
    std::vector<uint32_t> recv_start, recv_len; // ghost ranges filled from the peer
    size_t                n_send = 0, n_recv = 0;
  };

  class Mesh
  {
  public:
    MeshParams p;
    int        lo[3], hi[3], nl[3]; // local cell box (global coordinates) and its size
    int        S[3];                // entity slot lattice size per direction
    size_t     n_cells = 0;         // local cells
    std::vector<std::array<int, 3>> cell_ijk; // processing order -> global cell coordinates
    std::vector<uint32_t>           brick_ptr; // first cell of each brick (+ end)

    explicit Mesh(const MeshParams &params)
      : p(params)
    {
      int r = p.rank;
      for (int d = 0; d < 3; ++d)
        {
          const int pr = r % p.part[d];
          r /= p.part[d];
          lo[d] = (int)((long)p.nc[d] * pr / p.part[d]);
          hi[d] = (int)((long)p.nc[d] * (pr + 1) / p.part[d]);
          nl[d] = hi[d] - lo[d];
          S[d]  = 2 * p.nc[d] + (p.periodic[d] ? 0 : 1);
        }
      n_cells = (size_t)nl[0] * nl[1] * nl[2];
      cell_ijk.reserve(n_cells);
      const int *B = p.brick;
      for (int bz = 0; bz < nl[2]; bz += B[2])
        for (int by = 0; by < nl[1]; by += B[1])
          for (int bx = 0; bx < nl[0]; bx += B[0])
            {
              brick_ptr.push_back(cell_ijk.size());
              for (int z = bz; z < std::min(bz + B[2], nl[2]); ++z)
                for (int y = by; y < std::min(by + B[1], nl[1]); ++y)
                  for (int x = bx; x < std::min(bx + B[0], nl[0]); ++x)
                    cell_ijk.push_back({lo[0] + x, lo[1] + y, lo[2] + z});
            }
      brick_ptr.push_back(cell_ijk.size());
    }

    int
    n_ranks() const
    {
      return p.part[0] * p.part[1] * p.part[2];
    }

    int
    rank_of_cell(const int c[3]) const
    {
      int r = 0;
      for (int d = 2; d >= 0; --d)
        {
          // inverse of lo = nc*pr/part
          int pr = (int)(((long)c[d] * p.part[d]) / p.nc[d]);
          while ((long)p.nc[d] * pr / p.part[d] > c[d])
            --pr;
          while ((long)p.nc[d] * (pr + 1) / p.part[d] <= c[d])
            ++pr;
          r = r * p.part[d] + pr;
        }
      return r;
    }

    bool
    neighbor(const int c[3], int d, int side, int out[3]) const
    {
      out[0] = c[0];
      out[1] = c[1];
      out[2] = c[2];
      out[d] += side ? 1 : -1;
      if (out[d] < 0 || out[d] >= p.nc[d])
        {
          if (!p.periodic[d])
            return false;
          out[d] = (out[d] + p.nc[d]) % p.nc[d];
        }
      return true;
    }

    // ---- geometry -----------------------------------------------------------------------
    void
    map_point(const double X[3], double x[3]) const
    {
      if (p.map_kind == MAP_CARTESIAN)
        {
          x[0] = X[0];
          x[1] = X[1];
          x[2] = X[2];
        }
      else if (p.map_kind == MAP_SINE)
        {
          const double pi = 3.14159265358979323846;
          for (int d = 0; d < 3; ++d)
            x[d] = X[d] + std::sin(2 * pi * X[(d + 1) % 3]) * std::sin(pi * X[d]) * 0.1;
        }
      else
        {
          const double epsy = p.map_par[0], epsz = p.map_par[1];
          const double xx = X[0], y = X[1], z = X[2];
          int          layer = (int)(xx * 6.0);
          if (layer > 5)
            layer = 5;
          const double lambda = (xx - layer / 6.0) * 6;
          double       Y = 0, Z = 0;
          switch (layer)
            {
              case 0:
                Y = kershaw_left(epsy, y);
                Z = kershaw_left(epsz, z);
                break;
              case 1:
              case 4:
                Y = (1 - kershaw_step(lambda)) * kershaw_left(epsy, y) + kershaw_step(lambda) * kershaw_right(epsy, y);
                Z = (1 - kershaw_step(lambda)) * kershaw_left(epsz, z) + kershaw_step(lambda) * kershaw_right(epsz, z);
                break;
              case 2:
                Y = (1 - kershaw_step(lambda / 2)) * kershaw_right(epsy, y) + kershaw_step(lambda / 2) * kershaw_left(epsy, y);
                Z = (1 - kershaw_step(lambda / 2)) * kershaw_right(epsz, z) + kershaw_step(lambda / 2) * kershaw_left(epsz, z);
                break;
              case 3:
                Y = (1 - kershaw_step((1 + lambda) / 2)) * kershaw_right(epsy, y) + kershaw_step((1 + lambda) / 2) * kershaw_left(epsy, y);
                Z = (1 - kershaw_step((1 + lambda) / 2)) * kershaw_right(epsz, z) + kershaw_step((1 + lambda) / 2) * kershaw_left(epsz, z);
                break;
              default:
                Y = kershaw_right(epsy, y);
                Z = kershaw_right(epsz, z);
                break;
            }
          x[0] = xx;
          x[1] = Y;
          x[2] = Z;
        }
    }

    bool
    is_cartesian() const
    {
      return p.map_kind == MAP_CARTESIAN;
    }

    double
    h(int d) const
    {
      return p.length[d] / p.nc[d];
    }

    // 27 support points of the Q2 geometry interpolant of a cell (MappingQCache degree 2)
    // linear = true: the trilinear map through the 8 mapped vertices (the geometry "linear geometry" of the
    // reference sees, operator.h:512-591), sampled at the same 27 nodes
    void
    cell_support_points(const int c[3], double X[27][3], const bool linear = false) const
    {
      static const double nodes[3] = {0., 0.5, 1.};
      for (int k = 0; k < 3; ++k)
        for (int j = 0; j < 3; ++j)
          for (int i = 0; i < 3; ++i)
            {
              const double ref[3] = {(c[0] + nodes[i]) * h(0), (c[1] + nodes[j]) * h(1), (c[2] + nodes[k]) * h(2)};
              map_point(ref, X[9 * k + 3 * j + i]);
            }
      if (linear)
        for (int k = 0; k < 3; ++k)
          for (int j = 0; j < 3; ++j)
            for (int i = 0; i < 3; ++i)
              if (i == 1 || j == 1 || k == 1)
                for (int d = 0; d < 3; ++d)
                  {
                    double v = 0;
                    for (int c2 = 0; c2 < 2; ++c2)
                      for (int b2 = 0; b2 < 2; ++b2)
                        for (int a2 = 0; a2 < 2; ++a2)
                          {
                            const double w = (a2 ? nodes[i] : 1 - nodes[i]) * (b2 ? nodes[j] : 1 - nodes[j]) * (c2 ? nodes[k] : 1 - nodes[k]);
                            v += w * X[9 * (2 * c2) + 3 * (2 * b2) + 2 * a2][d];
                          }
                    X[9 * k + 3 * j + i][d] = v;
                  }
    }

    // 27 x 3 monomial coefficients of the triquadratic cell map x(xi) = sum c[9k+3j+i] xi^i eta^j zeta^k
    // (cell_quadratic_coefficients, operator.h:592-673): out[(9k+3j+i)*3 + d]
    void
    quadratic_coefficients(const int c[3], const bool linear, double *out) const
    {
      double X[27][3];
      cell_support_points(c, X, linear);
      static const double T[3][3] = {{1, 0, 0}, {-3, 4, -1}, {2, -4, 2}}; // nodal values at (0, 1/2, 1) -> monomials
      double             A[27][3], Bm[27][3];
      for (int k = 0; k < 3; ++k)
        for (int j = 0; j < 3; ++j)
          for (int i = 0; i < 3; ++i)
            for (int d = 0; d < 3; ++d)
              {
                double v = 0;
                for (int a = 0; a < 3; ++a)
                  v += T[i][a] * X[9 * k + 3 * j + a][d];
                A[9 * k + 3 * j + i][d] = v;
              }
      for (int k = 0; k < 3; ++k)
        for (int j = 0; j < 3; ++j)
          for (int i = 0; i < 3; ++i)
            for (int d = 0; d < 3; ++d)
              {
                double v = 0;
                for (int a = 0; a < 3; ++a)
                  v += T[j][a] * A[9 * k + 3 * a + i][d];
                Bm[9 * k + 3 * j + i][d] = v;
              }
      for (int k = 0; k < 3; ++k)
        for (int j = 0; j < 3; ++j)
          for (int i = 0; i < 3; ++i)
            for (int d = 0; d < 3; ++d)
              {
                double v = 0;
                for (int a = 0; a < 3; ++a)
                  v += T[k][a] * Bm[9 * a + 3 * j + i][d];
                out[(9 * k + 3 * j + i) * 3 + d] = v;
              }
    }

    // average distance between opposite faces in direction d (grid_tools.h:11-50), Gauss n x n rule
    double
    harmonic_extent(const int c[3], int d, const Basis1D &b) const
    {
      if (is_cartesian())
        return h(d);
      double X[27][3];
      cell_support_points(c, X);
      const std::vector<double> q2nodes = {0., 0.5, 1.};
      std::vector<double>       V, D;
      lagrange(q2nodes, b.qp, V, D); // [q*3+i]
      const int d1 = (d + 1) % 3, d2 = (d + 2) % 3;
      double    ext = 0;
      for (int qa = 0; qa < b.n; ++qa)
        for (int qb = 0; qb < b.n; ++qb)
          {
            double x0[3] = {0, 0, 0}, x1[3] = {0, 0, 0};
            for (int ia = 0; ia < 3; ++ia)
              for (int ib = 0; ib < 3; ++ib)
                {
                  const double w = V[qa * 3 + ia] * V[qb * 3 + ib];
                  int          idx0[3], idx1[3];
                  idx0[d] = 0;
                  idx1[d] = 2;
                  idx0[d1] = idx1[d1] = ia;
                  idx0[d2] = idx1[d2] = ib;
                  const double *P0 = X[9 * idx0[2] + 3 * idx0[1] + idx0[0]];
                  const double *P1 = X[9 * idx1[2] + 3 * idx1[1] + idx1[0]];
                  for (int e = 0; e < 3; ++e)
                    {
                      x0[e] += w * P0[e];
                      x1[e] += w * P1[e];
                    }
                }
            const double dist = std::sqrt((x0[0] - x1[0]) * (x0[0] - x1[0]) + (x0[1] - x1[1]) * (x0[1] - x1[1]) +
                                          (x0[2] - x1[2]) * (x0[2] - x1[2]));
            ext += dist * b.qw[qa] * b.qw[qb];
          }
      return ext;
    }

    // JxW at the n^3 Gauss points of a cell: out[q]
    void
    jxw(const int c[3], const Basis1D &b, double *out, const bool linear = false) const
    {
      const int n = b.n;
      double    X[27][3];
      cell_support_points(c, X, linear);
      const std::vector<double> q2nodes = {0., 0.5, 1.};
      std::vector<double>       V, D;
      lagrange(q2nodes, b.qp, V, D);
      for (int qz = 0; qz < n; ++qz)
        for (int qy = 0; qy < n; ++qy)
          for (int qx = 0; qx < n; ++qx)
            {
              double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
              for (int k = 0; k < 3; ++k)
                for (int j = 0; j < 3; ++j)
                  for (int i = 0; i < 3; ++i)
                    {
                      const double *P  = X[9 * k + 3 * j + i];
                      const double  gx = D[qx * 3 + i] * V[qy * 3 + j] * V[qz * 3 + k];
                      const double  gy = V[qx * 3 + i] * D[qy * 3 + j] * V[qz * 3 + k];
                      const double  gz = V[qx * 3 + i] * V[qy * 3 + j] * D[qz * 3 + k];
                      for (int dd = 0; dd < 3; ++dd)
                        {
                          J[dd][0] += P[dd] * gx;
                          J[dd][1] += P[dd] * gy;
                          J[dd][2] += P[dd] * gz;
                        }
                    }
              const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                                 J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
              out[(qz * n + qy) * n + qx] = det * b.qw[qx] * b.qw[qy] * b.qw[qz];
            }
    }

    // coordinates of the n^3 Gauss points of a cell in the Q2 geometry: out[e*n3 + q] ("construct q", operator.h:712-746)
    void
    quadrature_points(const int c[3], const Basis1D &b, double *out) const
    {
      const int n = b.n, n3 = n * n * n;
      double    X[27][3];
      cell_support_points(c, X, false);
      const std::vector<double> q2nodes = {0., 0.5, 1.};
      std::vector<double>       V, D;
      lagrange(q2nodes, b.qp, V, D);
      for (int qz = 0; qz < n; ++qz)
        for (int qy = 0; qy < n; ++qy)
          for (int qx = 0; qx < n; ++qx)
            {
              double x[3] = {0, 0, 0};
              for (int k = 0; k < 3; ++k)
                for (int j = 0; j < 3; ++j)
                  for (int i = 0; i < 3; ++i)
                    {
                      const double w = V[qx * 3 + i] * V[qy * 3 + j] * V[qz * 3 + k];
                      for (int e = 0; e < 3; ++e)
                        x[e] += w * X[9 * k + 3 * j + i][e];
                    }
              for (int e = 0; e < 3; ++e)
                out[e * n3 + (qz * n + qy) * n + qx] = x[e];
            }
    }

    // merged coefficients JxW * J^-1 J^-T at the n^3 Gauss points: out[comp*n3 + q], comp order
    // xx,xy,xz,yy,yz,zz (operator.h:696-704)
    void
    merged_coefficients(const int c[3], const Basis1D &b, double *out, const bool linear = false) const
    {
      const int n = b.n, n3 = n * n * n;
      double    X[27][3];
      cell_support_points(c, X, linear);
      const std::vector<double> q2nodes = {0., 0.5, 1.};
      std::vector<double>       V, D;
      lagrange(q2nodes, b.qp, V, D);
      for (int qz = 0; qz < n; ++qz)
        for (int qy = 0; qy < n; ++qy)
          for (int qx = 0; qx < n; ++qx)
            {
              double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}; // J[d][e] = dx_d/dxi_e
              for (int k = 0; k < 3; ++k)
                for (int j = 0; j < 3; ++j)
                  for (int i = 0; i < 3; ++i)
                    {
                      const double *P  = X[9 * k + 3 * j + i];
                      const double  gx = D[qx * 3 + i] * V[qy * 3 + j] * V[qz * 3 + k];
                      const double  gy = V[qx * 3 + i] * D[qy * 3 + j] * V[qz * 3 + k];
                      const double  gz = V[qx * 3 + i] * V[qy * 3 + j] * D[qz * 3 + k];
                      for (int dd = 0; dd < 3; ++dd)
                        {
                          J[dd][0] += P[dd] * gx;
                          J[dd][1] += P[dd] * gy;
                          J[dd][2] += P[dd] * gz;
                        }
                    }
              const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                                 J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
              double I[3][3];
              I[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / det;
              I[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
              I[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
              I[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) / det;
              I[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
              I[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
              I[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / det;
              I[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
              I[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
              const double jxw = det * b.qw[qx] * b.qw[qy] * b.qw[qz];
              const int    q   = (qz * n + qy) * n + qx;
              int          cc  = 0;
              for (int dd = 0; dd < 3; ++dd)
                for (int e = dd; e < 3; ++e, ++cc)
                  {
                    double s = 0;
                    for (int f = 0; f < 3; ++f)
                      s += I[dd][f] * I[e][f];
                    out[cc * n3 + q] = jxw * s;
                  }
            }
    }

    // ---- DoF numbering ------------------------------------------------------------------
    struct Numbering
    {
      int                   k        = 0;
      uint32_t              n_owned  = 0; // locally owned DoFs
      uint32_t              n_ghost  = 0; // ghost DoFs appended after the owned range
      std::vector<uint32_t> cidx;         // [cell*27+e] start index, INVALID if constrained (Dirichlet)
      std::vector<uint32_t> cidx_plain;   // same, constrained entities keep their index
      std::vector<uint32_t> constrained;  // list of constrained owned DoFs
      std::vector<ExchangeList> exchange; // per neighbouring rank
      std::vector<char>     brick_lex;    // per mesh brick: lexicographic box numbering
      std::vector<uint32_t> brick_base;   // per mesh brick: first owned DoF
      uint32_t              n_lex = 0;    // number of lex bricks: their boxes are [i 64k^3, (i+1) 64k^3), i < n_lex
    };

    // a brick whose own DoFs form the full box [0, 4k)^3: 4 x 4 x 4 cells, neighbour cells across all six faces
    bool
    brick_is_lex(const size_t b) const
    {
      static const bool off = getenv("DASM_NO_LEX") && getenv("DASM_NO_LEX")[0] == '1';
      if (off || p.brick[0] != 4 || p.brick[1] != 4 || p.brick[2] != 4 || brick_ptr[b + 1] - brick_ptr[b] != 64)
        return false;
      const auto &c0 = cell_ijk[brick_ptr[b]], &c1 = cell_ijk[brick_ptr[b + 1] - 1];
      const int   lo_c[3] = {c0[0], c0[1], c0[2]}, hi_c[3] = {c1[0], c1[1], c1[2]};
      for (int d = 0; d < 3; ++d)
        {
          int nbc[3];
          if (hi_c[d] - lo_c[d] != 3 || !neighbor(lo_c, d, 0, nbc) || !neighbor(hi_c, d, 1, nbc))
            return false;
        }
      return true;
    }

    static int
    entity_size(int e, int k)
    {
      int s = 1;
      for (int d = 0; d < 3; ++d, e /= 3)
        if (e % 3 == 1)
          s *= (k - 1);
      return s;
    }

    // owner cell of slot s (global)
    void
    owner_cell(const int s[3], int c[3]) const
    {
      for (int d = 0; d < 3; ++d)
        c[d] = std::min(s[d] / 2, p.nc[d] - 1);
    }

    bool
    slot_on_dirichlet_boundary(const int s[3]) const
    {
      if (!p.dirichlet)
        return false;
      for (int d = 0; d < 3; ++d)
        if (!p.periodic[d] && (s[d] == 0 || s[d] == S[d] - 1))
          return true;
      return false;
    }

    void
    cell_slot(const int c[3], int e, int s[3]) const
    {
      for (int d = 0; d < 3; ++d, e /= 3)
        {
          s[d] = 2 * c[d] + (e % 3);
          if (p.periodic[d])
            s[d] %= S[d];
        }
    }

    Numbering
    number_dofs(int k) const
    {
      Numbering nb;
      nb.k = k;
      // processing index of local cells by local lexicographic position
      std::vector<uint32_t> proc_of_local(n_cells);
      for (size_t i = 0; i < n_cells; ++i)
        {
          const auto &c = cell_ijk[i];
          proc_of_local[((size_t)(c[2] - lo[2]) * nl[1] + (c[1] - lo[1])) * nl[0] + (c[0] - lo[0])] = i;
        }
      // start index of every owned entity: bricks in order; inside a brick first the entities touched only
      // by cells of this brick ("private", cell by cell in lexicographic entity order), then the entities on
      // the brick's lower faces that are shared with neighbouring bricks - so that the DoFs a brick shares
      // with its neighbours form one contiguous range
      std::vector<uint32_t> own_start(n_cells * 27, INVALID_INDEX);
      {
        const size_t n_bricks = brick_ptr.size() - 1;
        nb.brick_lex.assign(n_bricks, 0);
        nb.brick_base.assign(n_bricks, 0);
        uint32_t next = 0;
        // lex bricks first: box [0, 4k)^3 in lexicographic order, start index = first DoF of the entity | LEX_FLAG
        for (size_t b = 0; b < n_bricks; ++b)
          {
            if (!brick_is_lex(b))
              continue;
            nb.brick_lex[b]  = 1;
            nb.brick_base[b] = next;
            ++nb.n_lex;
            const size_t first = brick_ptr[b];
            for (size_t i = first; i < brick_ptr[b + 1]; ++i)
              {
                const int cc[3] = {cell_ijk[i][0] - cell_ijk[first][0], cell_ijk[i][1] - cell_ijk[first][1], cell_ijk[i][2] - cell_ijk[first][2]};
                for (int e = 0; e < 27; ++e)
                  {
                    const int ee[3] = {e % 3, (e / 3) % 3, e / 9};
                    if (ee[0] == 2 || ee[1] == 2 || ee[2] == 2)
                      continue; // upper entities belong to the neighbour cell
                    const uint32_t X = cc[0] * k + (ee[0] ? 1 : 0), Y = cc[1] * k + (ee[1] ? 1 : 0), Z = cc[2] * k + (ee[2] ? 1 : 0);
                    own_start[i * 27 + e] = (next + X + 4 * k * (Y + 4 * k * Z)) | LEX_FLAG;
                  }
              }
            next += 64u * k * k * k;
          }
        for (size_t b = 0; b + 1 < brick_ptr.size(); ++b)
          {
            if (nb.brick_lex[b])
              continue;
            nb.brick_base[b] = next;
            const size_t first = brick_ptr[b], last = brick_ptr[b + 1];
            const int    org[3] = {cell_ijk[first][0], cell_ijk[first][1], cell_ijk[first][2]};
            for (int pass = 0; pass < 2; ++pass)
              for (size_t i = first; i < last; ++i)
                {
                  const int c[3] = {cell_ijk[i][0], cell_ijk[i][1], cell_ijk[i][2]};
                  for (int e = 0; e < 27; ++e)
                    {
                      bool owned = true, shared = false;
                      for (int d = 0, ee = e; d < 3; ++d, ee /= 3)
                        {
                          const int ed = ee % 3;
                          if (ed == 2 && !(!p.periodic[d] && c[d] == p.nc[d] - 1))
                            owned = false;
                          int nbc[3];
                          if (ed == 0 && c[d] == org[d] && neighbor(c, d, 0, nbc))
                            shared = true;
                        }
                      if (!owned || (shared ? 1 : 0) != pass)
                        continue;
                      own_start[i * 27 + e] = next;
                      next += entity_size(e, k);
                    }
                }
          }
        nb.n_owned = next;
        if (next >= LEX_FLAG)
          throw std::runtime_error("too many DoFs per rank for 31-bit start indices");
      }

      const int my_rank = p.rank;
      // ghost entities: key (owner rank, owner cell lexicographic global id, entity code) -> index
      std::map<std::array<long, 3>, uint32_t> ghost_map;
      struct Ref
      {
        size_t cell;
        int    e;
        std::array<long, 3> key;
      };
      std::vector<Ref> ghost_refs;
      nb.cidx.assign(n_cells * 27, INVALID_INDEX);
      nb.cidx_plain.assign(n_cells * 27, INVALID_INDEX);
      for (size_t i = 0; i < n_cells; ++i)
        {
          const int c[3] = {cell_ijk[i][0], cell_ijk[i][1], cell_ijk[i][2]};
          for (int e = 0; e < 27; ++e)
            {
              int s[3], oc[3];
              cell_slot(c, e, s);
              owner_cell(s, oc);
              int eo = 0; // entity code relative to the owner cell
              for (int d = 2; d >= 0; --d)
                eo = eo * 3 + (s[d] - 2 * oc[d]);
              const int orank = rank_of_cell(oc);
              if (orank == my_rank)
                {
                  const size_t oi = proc_of_local[((size_t)(oc[2] - lo[2]) * nl[1] + (oc[1] - lo[1])) * nl[0] + (oc[0] - lo[0])];
                  const uint32_t idx = own_start[oi * 27 + eo];
                  nb.cidx_plain[i * 27 + e] = idx;
                  if (!slot_on_dirichlet_boundary(s))
                    nb.cidx[i * 27 + e] = idx;
                  else if (oi == i)
                    for (int t = 0; t < entity_size(e, k); ++t)
                      nb.constrained.push_back(idx + t);
                }
              else
                {
                  const long gid = ((long)oc[2] * p.nc[1] + oc[1]) * p.nc[0] + oc[0];
                  ghost_refs.push_back({i, e, {orank, gid, eo}});
                  ghost_map[{orank, gid, eo}] = 0;
                }
            }
        }
      // number ghosts grouped by owner rank, then owner cell id, then entity code
      uint32_t next = nb.n_owned;
      {
        int           cur_rank = -1;
        ExchangeList *cur      = nullptr;
        for (auto &kv : ghost_map)
          {
            const int sz = entity_size((int)kv.first[2], k);
            kv.second    = next;
            if ((int)kv.first[0] != cur_rank)
              {
                cur_rank = (int)kv.first[0];
                nb.exchange.emplace_back();
                cur       = &nb.exchange.back();
                cur->peer = cur_rank;
              }
            if (sz > 0)
              {
                cur->recv_start.push_back(next);
                cur->recv_len.push_back(sz);
                cur->n_recv += sz;
              }
            next += sz;
          }
      }
      nb.n_ghost = next - nb.n_owned;
      for (const auto &r : ghost_refs)
        {
          const uint32_t idx = ghost_map[r.key];
          int            s[3];
          const int      c[3] = {cell_ijk[r.cell][0], cell_ijk[r.cell][1], cell_ijk[r.cell][2]};
          cell_slot(c, r.e, s);
          nb.cidx_plain[r.cell * 27 + r.e] = idx;
          if (!slot_on_dirichlet_boundary(s))
            nb.cidx[r.cell * 27 + r.e] = idx;
        }
      // send lists: entities owned by this rank that are touched by cells of other ranks.  A cell of
      // rank q touches an entity of my cell oc iff it is one of the <= 26 neighbours "below" the
      // slot; enumerate, for every local cell and owned entity, the cells touching the slot.
      if (n_ranks() > 1)
        {
          std::map<int, std::map<std::array<long, 3>, uint32_t>> send; // peer -> key -> start index (with LEX_FLAG)
          for (size_t i = 0; i < n_cells; ++i)
            {
              const int c[3] = {cell_ijk[i][0], cell_ijk[i][1], cell_ijk[i][2]};
              bool      near_boundary = false;
              for (int d = 0; d < 3; ++d)
                if (c[d] - lo[d] < 1 || hi[d] - c[d] <= 1)
                  near_boundary = true;
              if (!near_boundary)
                continue;
              for (int e = 0; e < 27; ++e)
                {
                  // owned by cell i?
                  bool owned = true;
                  for (int d = 0, ee = e; d < 3; ++d, ee /= 3)
                    if (ee % 3 == 2 && !(!p.periodic[d] && c[d] == p.nc[d] - 1))
                      owned = false;
                  if (!owned || entity_size(e, k) == 0)
                    continue;
                  int s[3];
                  cell_slot(c, e, s);
                  // cells touching slot s: per direction, odd slot -> cell s/2 only; even slot -> cells s/2-1 and s/2
                  int cand[3][2], ncand[3];
                  for (int d = 0; d < 3; ++d)
                    {
                      ncand[d] = 0;
                      if (s[d] % 2 == 1)
                        cand[d][ncand[d]++] = s[d] / 2;
                      else
                        {
                          int a = s[d] / 2 - 1, b2 = s[d] / 2;
                          if (p.periodic[d])
                            {
                              a  = (a + p.nc[d]) % p.nc[d];
                              b2 = b2 % p.nc[d];
                            }
                          if (a >= 0 && a < p.nc[d])
                            cand[d][ncand[d]++] = a;
                          if (b2 >= 0 && b2 < p.nc[d] && (ncand[d] == 0 || cand[d][0] != b2))
                            cand[d][ncand[d]++] = b2;
                        }
                    }
                  const long gid = ((long)c[2] * p.nc[1] + c[1]) * p.nc[0] + c[0];
                  for (int a = 0; a < ncand[0]; ++a)
                    for (int b2 = 0; b2 < ncand[1]; ++b2)
                      for (int g = 0; g < ncand[2]; ++g)
                        {
                          const int tc[3] = {cand[0][a], cand[1][b2], cand[2][g]};
                          const int q     = rank_of_cell(tc);
                          if (q != my_rank)
                            send[q][{(long)my_rank, gid, (long)e}] = own_start[i * 27 + e];
                        }
                }
            }
          for (auto &pr : send)
            {
              ExchangeList *ex = nullptr;
              for (auto &x : nb.exchange)
                if (x.peer == pr.first)
                  ex = &x;
              if (!ex)
                {
                  nb.exchange.emplace_back();
                  ex       = &nb.exchange.back();
                  ex->peer = pr.first;
                }
              for (auto &kv : pr.second)
                {
                  // runs in the lexicographic order of the DoFs inside the entity (the receiver's ghost copy is contiguous)
                  const int      e  = (int)kv.first[2];
                  const uint32_t st = kv.second;
                  const int      ex1 = (e % 3 == 1) ? k - 1 : 1, ey1 = ((e / 3) % 3 == 1) ? k - 1 : 1, ez1 = (e / 9 == 1) ? k - 1 : 1;
                  if (st & LEX_FLAG)
                    {
                      for (int l = 0; l < ez1; ++l)
                        for (int j = 0; j < ey1; ++j)
                          {
                            ex->send_start.push_back((st & ~LEX_FLAG) + 4 * k * (j + 4 * k * l));
                            ex->send_len.push_back((uint32_t)ex1);
                          }
                    }
                  else
                    {
                      ex->send_start.push_back(st);
                      ex->send_len.push_back((uint32_t)(ex1 * ey1 * ez1));
                    }
                  ex->n_send += (size_t)ex1 * ey1 * ez1;
                }
            }
        }
      return nb;
    }

    // ---- enlarged ghost layout (include/matrix_free.h:154-213 of the reference: the partitioner of a preconditioner with
    // overlapping patches holds, besides the ghost DoFs of the operator, all DoFs of the cells around the rank's cells) ------------
    // Halo cells = cells of other ranks adjacent (face, edge, corner) to a local cell.  Their 27 start indices point to owned DoFs,
    // to the ghost DoFs of `nb` or to NEW ghost DoFs appended behind them (grouped by owner rank, ordered by owner cell and entity
    // like the ghosts of number_dofs).  `exchange` lists ALL ghosts (old and new) per peer, both sides ordered by the same key.
    struct HaloNumbering
    {
      std::vector<std::array<int, 3>> cells;      // global coordinates of the halo cells
      std::vector<uint32_t>           cidx;       // [halo cell * 27 + e], INVALID if constrained
      std::vector<uint32_t>           cidx_plain;
      uint32_t                        n_ghost_ext = 0; // new ghost DoFs (behind n_owned + n_ghost)
      std::vector<ExchangeList>       exchange;
    };

    HaloNumbering
    halo_numbering(const Numbering &nb) const
    {
      HaloNumbering h;
      if (n_ranks() == 1)
        return h;
      const int k       = nb.k;
      const int my_rank = p.rank;
      std::vector<uint32_t> proc_of_local(n_cells);
      for (size_t i = 0; i < n_cells; ++i)
        {
          const auto &c = cell_ijk[i];
          proc_of_local[((size_t)(c[2] - lo[2]) * nl[1] + (c[1] - lo[1])) * nl[0] + (c[0] - lo[0])] = i;
        }
      auto is_local = [&](const int c[3]) {
        return c[0] >= lo[0] && c[0] < hi[0] && c[1] >= lo[1] && c[1] < hi[1] && c[2] >= lo[2] && c[2] < hi[2];
      };
      auto local_index = [&](const int c[3]) {
        return (size_t)proc_of_local[((size_t)(c[2] - lo[2]) * nl[1] + (c[1] - lo[1])) * nl[0] + (c[0] - lo[0])];
      };
      // owner cell, entity code relative to it and owner rank of entity e of cell c
      auto owner_of = [&](const int c[3], const int e, int s[3], int oc[3], int &eo) {
        cell_slot(c, e, s);
        owner_cell(s, oc);
        eo = 0;
        for (int d = 2; d >= 0; --d)
          eo = eo * 3 + (s[d] - 2 * oc[d]);
        return rank_of_cell(oc);
      };
      auto gid_of = [&](const int c[3]) { return ((long)c[2] * p.nc[1] + c[1]) * p.nc[0] + c[0]; };
      // ghosts of nb by key
      std::map<std::array<long, 3>, uint32_t> ghosts;
      for (size_t i = 0; i < n_cells; ++i)
        {
          const int c[3] = {cell_ijk[i][0], cell_ijk[i][1], cell_ijk[i][2]};
          for (int e = 0; e < 27; ++e)
            {
              int       s[3], oc[3], eo;
              const int orank = owner_of(c, e, s, oc, eo);
              if (orank != my_rank)
                ghosts[{(long)orank, gid_of(oc), (long)eo}] = nb.cidx_plain[i * 27 + e];
            }
        }
      // halo cells
      std::set<std::array<int, 3>> halo;
      for (int z = lo[2] - 1; z <= hi[2]; ++z)
        for (int y = lo[1] - 1; y <= hi[1]; ++y)
          for (int x = lo[0] - 1; x <= hi[0]; ++x)
            {
              int  c[3] = {x, y, z};
              bool ok   = true;
              for (int d = 0; d < 3 && ok; ++d)
                if (c[d] < 0 || c[d] >= p.nc[d])
                  {
                    if (!p.periodic[d])
                      ok = false;
                    else
                      c[d] = (c[d] + p.nc[d]) % p.nc[d];
                  }
              if (ok && !is_local(c))
                halo.insert({c[0], c[1], c[2]});
            }
      h.cells.assign(halo.begin(), halo.end());
      // new ghosts
      std::map<std::array<long, 3>, uint32_t> fresh;
      for (const auto &hc : h.cells)
        {
          const int c[3] = {hc[0], hc[1], hc[2]};
          for (int e = 0; e < 27; ++e)
            {
              int       s[3], oc[3], eo;
              const int orank = owner_of(c, e, s, oc, eo);
              if (orank == my_rank)
                continue;
              const std::array<long, 3> key = {(long)orank, gid_of(oc), (long)eo};
              if (!ghosts.count(key))
                fresh[key] = 0;
            }
        }
      uint32_t next = nb.n_owned + nb.n_ghost;
      for (auto &kv : fresh)
        {
          kv.second = next;
          next += entity_size((int)kv.first[2], k);
        }
      h.n_ghost_ext = next - (nb.n_owned + nb.n_ghost);
      if (next >= LEX_FLAG)
        throw std::runtime_error("too many DoFs per rank for 31-bit start indices");
      // index rows of the halo cells
      h.cidx.assign(h.cells.size() * 27, INVALID_INDEX);
      h.cidx_plain.assign(h.cells.size() * 27, INVALID_INDEX);
      for (size_t i = 0; i < h.cells.size(); ++i)
        {
          const int c[3] = {h.cells[i][0], h.cells[i][1], h.cells[i][2]};
          for (int e = 0; e < 27; ++e)
            {
              int       s[3], oc[3], eo;
              const int orank = owner_of(c, e, s, oc, eo);
              uint32_t  idx;
              if (orank == my_rank)
                idx = nb.cidx_plain[local_index(oc) * 27 + eo];
              else
                {
                  const std::array<long, 3> key = {(long)orank, gid_of(oc), (long)eo};
                  const auto                it  = ghosts.find(key);
                  idx                           = it != ghosts.end() ? it->second : fresh[key];
                }
              h.cidx_plain[i * 27 + e] = idx;
              if (!slot_on_dirichlet_boundary(s))
                h.cidx[i * 27 + e] = idx;
            }
        }
      // receive lists: all ghosts per owner rank in key order
      std::map<std::array<long, 3>, uint32_t> all(ghosts);
      all.insert(fresh.begin(), fresh.end());
      {
        int           cur_rank = -1;
        ExchangeList *cur      = nullptr;
        for (const auto &kv : all)
          {
            const int sz = entity_size((int)kv.first[2], k);
            if ((int)kv.first[0] != cur_rank)
              {
                cur_rank = (int)kv.first[0];
                h.exchange.emplace_back();
                cur       = &h.exchange.back();
                cur->peer = cur_rank;
              }
            if (sz > 0)
              {
                cur->recv_start.push_back(kv.second);
                cur->recv_len.push_back(sz);
                cur->n_recv += sz;
              }
          }
      }
      // send lists: owned entities touched by a cell inside the grown box of another rank
      std::map<int, std::map<std::array<long, 3>, uint32_t>> send;
      for (size_t i = 0; i < n_cells; ++i)
        {
          const int c[3] = {cell_ijk[i][0], cell_ijk[i][1], cell_ijk[i][2]};
          bool      near_boundary = false;
          for (int d = 0; d < 3; ++d)
            if (c[d] - lo[d] < 2 || hi[d] - c[d] <= 2)
              near_boundary = true;
          if (!near_boundary)
            continue;
          for (int e = 0; e < 27; ++e)
            {
              bool owned = true;
              for (int d = 0, ee = e; d < 3; ++d, ee /= 3)
                if (ee % 3 == 2 && !(!p.periodic[d] && c[d] == p.nc[d] - 1))
                  owned = false;
              if (!owned || entity_size(e, k) == 0)
                continue;
              int s[3];
              cell_slot(c, e, s);
              // cells within one cell of a cell touching the slot: per direction the range [first touching - 1, last touching + 1]
              int cand[3][4], ncand[3];
              for (int d = 0; d < 3; ++d)
                {
                  ncand[d]     = 0;
                  const int t0 = (s[d] % 2 == 1) ? s[d] / 2 : s[d] / 2 - 1, t1 = s[d] / 2;
                  for (int v = t0 - 1; v <= t1 + 1; ++v)
                    {
                      int w = v;
                      if (p.periodic[d])
                        w = ((w % p.nc[d]) + p.nc[d]) % p.nc[d];
                      else if ((s[d] % 2 == 0) && (v == t1) && t1 >= p.nc[d])
                        continue; // (upper boundary slot: the cell above does not exist)
                      if (w < 0 || w >= p.nc[d])
                        continue;
                      bool dup = false;
                      for (int q = 0; q < ncand[d]; ++q)
                        if (cand[d][q] == w)
                          dup = true;
                      if (!dup)
                        cand[d][ncand[d]++] = w;
                    }
                }
              const long gid = gid_of(c);
              for (int a = 0; a < ncand[0]; ++a)
                for (int b2 = 0; b2 < ncand[1]; ++b2)
                  for (int g = 0; g < ncand[2]; ++g)
                    {
                      const int tc[3] = {cand[0][a], cand[1][b2], cand[2][g]};
                      const int q     = rank_of_cell(tc);
                      if (q != my_rank)
                        send[q][{(long)my_rank, gid, (long)e}] = nb.cidx_plain[i * 27 + e];
                    }
            }
        }
      for (auto &pr : send)
        {
          ExchangeList *ex = nullptr;
          for (auto &x : h.exchange)
            if (x.peer == pr.first)
              ex = &x;
          if (!ex)
            {
              h.exchange.emplace_back();
              ex       = &h.exchange.back();
              ex->peer = pr.first;
            }
          for (auto &kv : pr.second)
            {
              const int      e   = (int)kv.first[2];
              const uint32_t st  = kv.second;
              const int      ex1 = (e % 3 == 1) ? k - 1 : 1, ey1 = ((e / 3) % 3 == 1) ? k - 1 : 1, ez1 = (e / 9 == 1) ? k - 1 : 1;
              if (st & LEX_FLAG)
                {
                  for (int l = 0; l < ez1; ++l)
                    for (int j = 0; j < ey1; ++j)
                      {
                        ex->send_start.push_back((st & ~LEX_FLAG) + 4 * k * (j + 4 * k * l));
                        ex->send_len.push_back((uint32_t)ex1);
                      }
                }
              else
                {
                  ex->send_start.push_back(st);
                  ex->send_len.push_back((uint32_t)(ex1 * ey1 * ez1));
                }
              ex->n_send += (size_t)ex1 * ey1 * ez1;
            }
        }
      return h;
    }
  };
} // namespace dasm
