// Unstructured hexahedral meshes (host side): the ball of BASELINE configs[3] (element_centered_preconditioners_01.cc:398-402,
// experiments/ball.py) and any other conforming all-hex mesh given by arrays.
//
// What the reference takes from deal.II for such a mesh and what is restated here:
//   * entity connectivity and the orientation of a cell's lines / quads relative to the mesh entity -> the packed orientation word
//     of ConstraintInfoReduced (12 line bits + 6 x 3 quad bits, include/reduced_access.h:154-285, 528-702): line l of a cell is
//     flipped when the cell runs it against the entity's direction; the 3-bit code f of a quad says by which of the 8 symmetries of
//     the square the entity's (k-1)^2 DoFs appear in the cell's lexicographic face layout (csrc/reduced_access.cu make_tables:
//     entity DoF (i, j) sits at local (a, b) = T_f(i, j));
//   * the 3^3 compressed start indices per cell (include/vector_access_reduced.h:30-164): vertices, then lines (k-1 DoFs each),
//     quads ((k-1)^2), cell interiors ((k-1)^3), each entity's DoFs contiguous in the entity's own frame; entities on the boundary
//     are constrained (homogeneous Dirichlet, element_centered_preconditioners_01.cc:404-413) and read invalid;
//   * the geometry of MappingQCache(2): 27 support points per cell; merged coefficients / quadrature points / JxW from them
//     (include/operator.h:674-746);
//   * compute_harmonic_patch_extend (include/grid_tools.h:11-138): mean distance of opposite faces per cell and direction, the
//     neighbour's extent normal to the shared face.
// The frame of an entity is that of the first cell (in cell order) that contains it.  One rank only.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include "basis.h"

namespace dasm
{
  struct UMesh
  {
    static constexpr uint32_t INVALID = 0xFFFFFFFFu;

    long long             n_vertices = 0, n_cells = 0, n_lines = 0, n_quads = 0;
    bool                  dirichlet = true;
    std::vector<double>   coords;  // [vertex][3]
    std::vector<uint32_t> cells;   // [cell][8] vertices in lexicographic order (x fastest)
    std::vector<double>   support; // [cell][27][3] support points of the triquadratic cell map, lexicographic
    std::vector<uint32_t> entity;  // [cell][27] number of the entity within its class (vertex / line / quad / cell)
    std::vector<uint32_t> orientation; // [cell] packed word
    std::vector<int32_t>  quad_cell;   // [quad][2] cells on the two sides (-1: boundary)
    std::vector<int8_t>   quad_face;   // [quad][2] local face (2 d + side) in those cells
    std::vector<char>     vertex_bnd, line_bnd, quad_bnd;

    // local vertex (0..7) of entity e = ex + 3 ey + 9 ez for the corner (i0, i1) of its own lexicographic frame
    static int
    entity_corner(const int e, const int i0, const int i1)
    {
      const int c[3] = {e % 3, (e / 3) % 3, e / 9};
      int       v = 0, t = 0;
      for (int d = 0; d < 3; ++d)
        {
          int bit;
          if (c[d] == 1)
            bit = (t++ == 0) ? i0 : i1;
          else
            bit = c[d] / 2;
          v |= bit << d;
        }
      return v;
    }

    // position of the entity DoF (i, j) (n x n grid) in the local n x n face layout under quad code f
    static void
    quad_transform(const int f, const int n, const int i, const int j, int &a, int &b)
    {
      switch (f)
        {
          case 0: a = i, b = j; break;
          case 1: a = j, b = i; break;
          case 2: a = j, b = n - 1 - i; break;
          case 3: a = i, b = n - 1 - j; break;
          case 4: a = n - 1 - i, b = n - 1 - j; break;
          case 5: a = n - 1 - j, b = n - 1 - i; break;
          case 6: a = n - 1 - j, b = i; break;
          default: a = n - 1 - i, b = j; break;
        }
    }

    static int
    line_number(const int e)
    {
      const int ex = e % 3, ey = (e / 3) % 3, ez = e / 9;
      if (ex == 1)
        return (ey == 2 ? 1 : 0) + (ez == 2 ? 2 : 0);
      if (ey == 1)
        return 4 + (ex == 2 ? 1 : 0) + (ez == 2 ? 2 : 0);
      return 8 + (ex == 2 ? 1 : 0) + (ey == 2 ? 2 : 0);
    }

    static int
    quad_number(const int e)
    {
      const int ex = e % 3, ey = (e / 3) % 3, ez = e / 9;
      const int d  = (ex != 1) ? 0 : ((ey != 1) ? 1 : 2);
      const int s  = (d == 0 ? ex : (d == 1 ? ey : ez)) == 2 ? 1 : 0;
      return 2 * d + s;
    }

    static int
    entity_dim(const int e)
    {
      return (e % 3 == 1) + ((e / 3) % 3 == 1) + (e / 9 == 1);
    }

    void
    build()
    {
      if ((long long)cells.size() != 8 * n_cells || (long long)coords.size() != 3 * n_vertices)
        throw std::runtime_error("unstructured mesh: array sizes do not match the vertex / cell counts");
      for (const uint32_t v : cells)
        if ((long long)v >= n_vertices)
          throw std::runtime_error("unstructured mesh: vertex index out of range");
      if (support.empty())
        {
          // trilinear cells: support points from the vertices
          support.resize((size_t)n_cells * 81);
          for (long long c = 0; c < n_cells; ++c)
            for (int z = 0; z < 3; ++z)
              for (int y = 0; y < 3; ++y)
                for (int x = 0; x < 3; ++x)
                  for (int d = 0; d < 3; ++d)
                    {
                      double s = 0;
                      for (int v = 0; v < 8; ++v)
                        {
                          const double wx = (v & 1) ? 0.5 * x : 1 - 0.5 * x, wy = (v & 2) ? 0.5 * y : 1 - 0.5 * y, wz = (v & 4) ? 0.5 * z : 1 - 0.5 * z;
                          s += wx * wy * wz * coords[3 * (size_t)cells[8 * c + v] + d];
                        }
                      support[((size_t)c * 27 + 9 * z + 3 * y + x) * 3 + d] = s;
                    }
        }
      if ((long long)support.size() != 81 * n_cells)
        throw std::runtime_error("unstructured mesh: 27 x 3 support point coordinates per cell expected");

      entity.assign((size_t)n_cells * 27, 0);
      orientation.assign((size_t)n_cells, 0);
      std::map<std::array<uint32_t, 2>, uint32_t> line_of;
      std::map<std::array<uint32_t, 4>, uint32_t> quad_of;
      std::vector<std::array<uint32_t, 2>>        line_frame; // (start, end) vertex
      std::vector<std::array<uint32_t, 4>>        quad_frame; // corners in the frame's lexicographic order
      for (long long c = 0; c < n_cells; ++c)
        {
          const uint32_t *cv   = cells.data() + 8 * c;
          uint32_t        word = 0;
          for (int e = 0; e < 27; ++e)
            {
              const int dim = entity_dim(e);
              if (dim == 0)
                entity[c * 27 + e] = cv[entity_corner(e, 0, 0)];
              else if (dim == 1)
                {
                  const uint32_t a = cv[entity_corner(e, 0, 0)], b = cv[entity_corner(e, 1, 0)];
                  if (a == b)
                    throw std::runtime_error("unstructured mesh: degenerate line");
                  const std::array<uint32_t, 2> key = {std::min(a, b), std::max(a, b)};
                  auto                          it  = line_of.find(key);
                  if (it == line_of.end())
                    {
                      it = line_of.emplace(key, (uint32_t)line_frame.size()).first;
                      line_frame.push_back({a, b});
                    }
                  entity[c * 27 + e] = it->second;
                  if (line_frame[it->second][0] != a)
                    word |= 1u << line_number(e);
                }
              else if (dim == 2)
                {
                  std::array<uint32_t, 4> loc, key;
                  for (int p = 0; p < 4; ++p)
                    loc[p] = cv[entity_corner(e, p & 1, p >> 1)];
                  key = loc;
                  std::sort(key.begin(), key.end());
                  auto it = quad_of.find(key);
                  const int q = quad_number(e);
                  if (it == quad_of.end())
                    {
                      it = quad_of.emplace(key, (uint32_t)quad_frame.size()).first;
                      quad_frame.push_back(loc);
                      quad_cell.push_back((int32_t)c);
                      quad_cell.push_back(-1);
                      quad_face.push_back((int8_t)q);
                      quad_face.push_back(-1);
                    }
                  else
                    {
                      if (quad_cell[2 * it->second + 1] != -1)
                        throw std::runtime_error("unstructured mesh: a face belongs to more than two cells");
                      quad_cell[2 * it->second + 1] = (int32_t)c;
                      quad_face[2 * it->second + 1] = (int8_t)q;
                    }
                  entity[c * 27 + e] = it->second;
                  const auto &fr     = quad_frame[it->second];
                  int         code   = -1;
                  for (int f = 0; f < 8 && code < 0; ++f)
                    {
                      bool ok = true;
                      for (int p = 0; p < 4 && ok; ++p)
                        {
                          int a, b;
                          quad_transform(f, 2, p & 1, p >> 1, a, b);
                          ok = loc[a + 2 * b] == fr[p];
                        }
                      if (ok)
                        code = f;
                    }
                  if (code < 0)
                    throw std::runtime_error("unstructured mesh: the two cells of a face do not see the same quadrilateral");
                  word |= (uint32_t)code << (12 + 3 * q);
                }
              else
                entity[c * 27 + e] = (uint32_t)c;
            }
          orientation[c] = word;
        }
      n_lines = (long long)line_frame.size();
      n_quads = (long long)quad_frame.size();
      vertex_bnd.assign((size_t)n_vertices, 0);
      line_bnd.assign((size_t)n_lines, 0);
      quad_bnd.assign((size_t)n_quads, 0);
      for (long long q = 0; q < n_quads; ++q)
        if (quad_cell[2 * q + 1] == -1)
          {
            quad_bnd[q] = 1;
            // its vertices and lines through the cell that holds it
            const long long c  = quad_cell[2 * q];
            const int       fq = quad_face[2 * q], d = fq / 2, s = fq % 2;
            for (int e = 0; e < 27; ++e)
              {
                const int ce[3] = {e % 3, (e / 3) % 3, e / 9};
                if (ce[d] != 2 * s)
                  continue;
                const int dim = entity_dim(e);
                if (dim == 0)
                  vertex_bnd[entity[c * 27 + e]] = 1;
                else if (dim == 1)
                  line_bnd[entity[c * 27 + e]] = 1;
              }
          }
    }

    struct Numbering
    {
      long long             n_dofs = 0;
      std::vector<uint32_t> cidx, cidx_plain, constrained, plain; // plain: [cell][(k+1)^3] with the orientation applied
    };

    // (x, y, z) of the standard layout whose value belongs at the local position (x, y, z): adjust_for_orientation as an index map
    static void
    oriented_source(const int k, const uint32_t word, int &x, int &y, int &z)
    {
      if (word == 0u)
        return;
      const int ex = (x == 0) ? 0 : ((x == k) ? 2 : 1), ey = (y == 0) ? 0 : ((y == k) ? 2 : 1), ez = (z == 0) ? 0 : ((z == k) ? 2 : 1);
      const int e = ex + 3 * ey + 9 * ez, dim = entity_dim(e);
      if (dim == 1)
        {
          if ((word >> line_number(e)) & 1u)
            {
              if (ex == 1)
                x = k - x;
              else if (ey == 1)
                y = k - y;
              else
                z = k - z;
            }
        }
      else if (dim == 2)
        {
          const int q = quad_number(e), d = q / 2, f = (int)((word >> (12 + 3 * q)) & 7u);
          if (f != 0)
            {
              int &     a = (d == 0) ? y : x;
              int &     b = (d == 2) ? y : z;
              const int m = k - 1;
              // find the entity DoF (i, j) that sits at local (a - 1, b - 1)
              for (int j = 0; j < m; ++j)
                for (int i = 0; i < m; ++i)
                  {
                    int ta, tb;
                    quad_transform(f, m, i, j, ta, tb);
                    if (ta == a - 1 && tb == b - 1)
                      {
                        a = i + 1;
                        b = j + 1;
                        return;
                      }
                  }
            }
        }
    }

    Numbering
    number_dofs(const int k) const
    {
      Numbering       nb;
      const long long per_line = k - 1, per_quad = (long long)(k - 1) * (k - 1), per_hex = per_quad * (k - 1);
      const long long line0 = n_vertices, quad0 = line0 + n_lines * per_line, hex0 = quad0 + n_quads * per_quad;
      nb.n_dofs = hex0 + n_cells * per_hex;
      if (nb.n_dofs >= (1ll << 31))
        throw std::runtime_error("unstructured mesh: more than 2^31 DoFs");
      nb.cidx.assign((size_t)n_cells * 27, INVALID);
      nb.cidx_plain.assign((size_t)n_cells * 27, INVALID);
      for (long long c = 0; c < n_cells; ++c)
        for (int e = 0; e < 27; ++e)
          {
            const int      dim = entity_dim(e);
            const uint32_t id  = entity[c * 27 + e];
            long long      start;
            bool           bnd;
            if (dim == 0)
              start = id, bnd = vertex_bnd[id];
            else if (dim == 1)
              start = line0 + id * per_line, bnd = line_bnd[id];
            else if (dim == 2)
              start = quad0 + id * per_quad, bnd = quad_bnd[id];
            else
              start = hex0 + id * per_hex, bnd = false;
            nb.cidx_plain[c * 27 + e] = (uint32_t)start;
            if (!(bnd && dirichlet))
              nb.cidx[c * 27 + e] = (uint32_t)start;
          }
      if (dirichlet)
        {
          for (long long v = 0; v < n_vertices; ++v)
            if (vertex_bnd[v])
              nb.constrained.push_back((uint32_t)v);
          for (long long l = 0; l < n_lines; ++l)
            if (line_bnd[l])
              for (long long i = 0; i < per_line; ++i)
                nb.constrained.push_back((uint32_t)(line0 + l * per_line + i));
          for (long long q = 0; q < n_quads; ++q)
            if (quad_bnd[q])
              for (long long i = 0; i < per_quad; ++i)
                nb.constrained.push_back((uint32_t)(quad0 + q * per_quad + i));
        }
      // (k+1)^3 addresses per cell
      const int n = k + 1, n3 = n * n * n;
      nb.plain.resize((size_t)n_cells * n3);
      for (long long c = 0; c < n_cells; ++c)
        {
          const uint32_t *ci = nb.cidx.data() + c * 27;
          for (int z = 0; z < n; ++z)
            for (int y = 0; y < n; ++y)
              for (int x = 0; x < n; ++x)
                {
                  int sx = x, sy = y, sz = z;
                  oriented_source(k, orientation[c], sx, sy, sz);
                  const int      ex = (sx == 0) ? 0 : ((sx == k) ? 2 : 1), ey = (sy == 0) ? 0 : ((sy == k) ? 2 : 1), ez = (sz == 0) ? 0 : ((sz == k) ? 2 : 1);
                  const uint32_t st = ci[ex + 3 * ey + 9 * ez];
                  uint32_t       g  = INVALID;
                  if (st != INVALID)
                    {
                      const int ox = (ex == 1) ? sx - 1 : 0, oy = (ey == 1) ? sy - 1 : 0, oz = (ez == 1) ? sz - 1 : 0;
                      const int stx = (ex == 1) ? (k - 1) : 1, sty = (ey == 1) ? (k - 1) : 1;
                      g             = st + ox + stx * (oy + sty * oz);
                    }
                  nb.plain[(size_t)c * n3 + (z * n + y) * n + x] = g;
                }
        }
      return nb;
    }

    // ---- geometry of the triquadratic cell map --------------------------------------------------------------------------------------
    // Jacobian J[d][e] = dx_d / dxi_e and the point x at the reference point (V, D = values / derivatives of the three quadratic
    // Lagrange polynomials per direction)
    void
    map_at(const long long c, const double *Vx, const double *Dx, const double *Vy, const double *Dy, const double *Vz, const double *Dz,
           double x[3], double J[3][3]) const
    {
      const double *X = support.data() + (size_t)c * 81;
      for (int d = 0; d < 3; ++d)
        {
          x[d] = 0;
          J[d][0] = J[d][1] = J[d][2] = 0;
        }
      for (int kz = 0; kz < 3; ++kz)
        for (int j = 0; j < 3; ++j)
          for (int i = 0; i < 3; ++i)
            {
              const double *P = X + 3 * (9 * kz + 3 * j + i);
              const double  v = Vx[i] * Vy[j] * Vz[kz], gx = Dx[i] * Vy[j] * Vz[kz], gy = Vx[i] * Dy[j] * Vz[kz], gz = Vx[i] * Vy[j] * Dz[kz];
              for (int d = 0; d < 3; ++d)
                {
                  x[d] += P[d] * v;
                  J[d][0] += P[d] * gx;
                  J[d][1] += P[d] * gy;
                  J[d][2] += P[d] * gz;
                }
            }
    }

    // per Gauss point of cell c: the point (xq[e*n3+q]), JxW (jxw[q]) and the merged coefficients JxW J^-1 J^-T (G[comp*n3+q], comp
    // xx xy xz yy yz zz, operator.h:696-704); null outputs are skipped
    void
    cell_geometry(const long long c, const Basis1D &b, double *xq, double *jxw, double *G) const
    {
      const int                 n = b.n, n3 = n * n * n;
      const std::vector<double> q2nodes = {0., 0.5, 1.};
      std::vector<double>       V, D;
      lagrange(q2nodes, b.qp, V, D);
      for (int qz = 0; qz < n; ++qz)
        for (int qy = 0; qy < n; ++qy)
          for (int qx = 0; qx < n; ++qx)
            {
              double x[3], J[3][3];
              map_at(c, &V[3 * qx], &D[3 * qx], &V[3 * qy], &D[3 * qy], &V[3 * qz], &D[3 * qz], x, J);
              const int    q   = (qz * n + qy) * n + qx;
              const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                                 J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
              if (!(det > 0))
                throw std::runtime_error("unstructured mesh: cell " + std::to_string(c) + " is inverted (det J <= 0 at a quadrature point)");
              const double w = det * b.qw[qx] * b.qw[qy] * b.qw[qz];
              if (xq)
                for (int e = 0; e < 3; ++e)
                  xq[e * n3 + q] = x[e];
              if (jxw)
                jxw[q] = w;
              if (G)
                {
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
                  int cc = 0;
                  for (int dd = 0; dd < 3; ++dd)
                    for (int e = dd; e < 3; ++e, ++cc)
                      G[cc * n3 + q] = w * (I[dd][0] * I[e][0] + I[dd][1] * I[e][1] + I[dd][2] * I[e][2]);
                }
            }
    }

    // mean distance between the faces 2 d and 2 d + 1 of cell c (grid_tools.h:11-50), Gauss n x n rule on the face
    double
    cell_extent(const long long c, const int d, const Basis1D &b) const
    {
      const int                 n = b.n;
      const std::vector<double> q2nodes = {0., 0.5, 1.};
      std::vector<double>       V, D;
      lagrange(q2nodes, b.qp, V, D);
      const double V0[3] = {1, 0, 0}, V1[3] = {0, 0, 1}, Z[3] = {0, 0, 0};
      const int    d1 = (d + 1) % 3, d2 = (d + 2) % 3;
      double       ext = 0;
      for (int qa = 0; qa < n; ++qa)
        for (int qb = 0; qb < n; ++qb)
          {
            const double *Vd[3];
            double        x0[3], x1[3], J[3][3];
            Vd[d1] = &V[3 * qa];
            Vd[d2] = &V[3 * qb];
            Vd[d]  = V0;
            map_at(c, Vd[0], Z, Vd[1], Z, Vd[2], Z, x0, J);
            Vd[d] = V1;
            map_at(c, Vd[0], Z, Vd[1], Z, Vd[2], Z, x1, J);
            ext += std::sqrt((x0[0] - x1[0]) * (x0[0] - x1[0]) + (x0[1] - x1[1]) * (x0[1] - x1[1]) + (x0[2] - x1[2]) * (x0[2] - x1[2])) *
                   b.qw[qa] * b.qw[qb];
          }
      return ext;
    }

    // compute_harmonic_patch_extend (grid_tools.h:54-138): out[(c*3+d)*3 + {0,1,2}] = extent of the neighbour across face 2d (0 at
    // the boundary), own extent, neighbour across face 2d+1; the neighbour's extent is the one normal to the shared face
    std::vector<double>
    patch_extents(const Basis1D &b) const
    {
      std::vector<double> ext((size_t)n_cells * 3), face((size_t)n_quads, 0.), out((size_t)n_cells * 9);
      static const int    face_entity[6] = {12, 14, 10, 16, 4, 22}; // e of the quads 2 d + s
      for (long long c = 0; c < n_cells; ++c)
        for (int d = 0; d < 3; ++d)
          {
            ext[c * 3 + d] = cell_extent(c, d, b);
            face[entity[c * 27 + face_entity[2 * d]]] += ext[c * 3 + d];
            face[entity[c * 27 + face_entity[2 * d + 1]]] += ext[c * 3 + d];
          }
      for (long long c = 0; c < n_cells; ++c)
        for (int d = 0; d < 3; ++d)
          {
            const double own        = ext[c * 3 + d];
            out[(c * 3 + d) * 3 + 0] = quad_bnd[entity[c * 27 + face_entity[2 * d]]] ? 0. : face[entity[c * 27 + face_entity[2 * d]]] - own;
            out[(c * 3 + d) * 3 + 1] = own;
            out[(c * 3 + d) * 3 + 2] = quad_bnd[entity[c * 27 + face_entity[2 * d + 1]]] ? 0. : face[entity[c * 27 + face_entity[2 * d + 1]]] - own;
          }
      return out;
    }

    bool
    face_at_boundary(const long long c, const int face) const
    {
      static const int face_entity[6] = {12, 14, 10, 16, 4, 22};
      return quad_bnd[entity[c * 27 + face_entity[face]]] != 0;
    }
  };
} // namespace dasm
