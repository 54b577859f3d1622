// Unstructured all-hex meshes as arrays for LaplaceOperatorMatrixFree(const UnstructuredMesh &, ...) (dasm_op_create_unstructured).
//
// hyper_ball(): the ball of the reference's driver (GridGenerator::hyper_ball_balanced + refine_global + MappingQCache(2),
// element_centered_preconditioners_01.cc:398-402, 415-425).  deal.II is not part of the reference tree, so the generator is this
// library's (the same as dealii-asm_b200/grid.py): the coarse topology of hyper_ball_balanced in 3-D (32 cells: a 2 x 2 x 2 inner
// block and 4 cells over each of its 6 faces, outer faces on the sphere); a refinement splits every cell into 8, new points follow
// the coarse cell's map (inner cells trilinear; outer cells: linear blend between the inner face and the radial projection of the
// outer face); every cell carries the 27 support points of its triquadratic map.  The children of a cell keep its frame, which is
// what ball_parents() encodes for the geometric two-level transfer.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <functional>
#include <map>
#include <numeric>
#include <vector>

namespace dasm
{
  struct UnstructuredMesh
  {
    std::vector<double>        vertices; // [V][3]
    std::vector<std::uint32_t> cells;    // [C][8] lexicographic
    std::vector<double>        support;  // [C][27][3]
    long long                  n_cells() const { return (long long)cells.size() / 8; }
    long long                  n_vertices() const { return (long long)vertices.size() / 3; }
  };

  namespace GridGenerator
  {
    namespace internal
    {
      using Point = std::array<double, 3>;
      using Map   = std::function<Point(double, double, double)>;

      inline std::vector<Map>
      ball_coarse_maps(const double radius)
      {
        const double mag[4] = {0.0, 0.528, 0.4533, 0.3752};
        const auto   lattice = [=](int i, int j, int l) {
          const double m = radius * mag[std::abs(i) + std::abs(j) + std::abs(l)];
          return Point{{i * m, j * m, l * m}};
        };
        const auto sphere = [=](int i, int j, int l) {
          const double n = std::sqrt((double)(i * i + j * j + l * l));
          return Point{{radius * i / n, radius * j / n, radius * l / n}};
        };
        std::vector<Map> maps;
        for (int l0 = -1; l0 <= 0; ++l0)
          for (int j0 = -1; j0 <= 0; ++j0)
            for (int i0 = -1; i0 <= 0; ++i0)
              {
                std::array<Point, 8> c;
                for (int v = 0; v < 8; ++v)
                  c[v] = lattice(i0 + (v & 1), j0 + ((v >> 1) & 1), l0 + (v >> 2));
                maps.push_back([c](double x, double y, double z) {
                  Point p{{0, 0, 0}};
                  for (int v = 0; v < 8; ++v)
                    {
                      const double w = ((v & 1) ? x : 1 - x) * ((v & 2) ? y : 1 - y) * ((v & 4) ? z : 1 - z);
                      for (int d = 0; d < 3; ++d)
                        p[d] += w * c[v][d];
                    }
                  return p;
                });
              }
        for (int d = 0; d < 3; ++d)
          for (int s = -1; s <= 1; s += 2)
            {
              const int d1 = s == 1 ? (d + 1) % 3 : (d + 2) % 3, d2 = s == 1 ? (d + 2) % 3 : (d + 1) % 3; // right-handed (xi, eta, radial)
              for (int b0 = -1; b0 <= 0; ++b0)
                for (int a0 = -1; a0 <= 0; ++a0)
                  {
                    std::array<Point, 4> in, out;
                    for (int v = 0; v < 4; ++v)
                      {
                        int idx[3];
                        idx[d]  = s;
                        idx[d1] = a0 + (v & 1);
                        idx[d2] = b0 + (v >> 1);
                        in[v]   = lattice(idx[0], idx[1], idx[2]);
                        out[v]  = sphere(idx[0], idx[1], idx[2]);
                      }
                    maps.push_back([in, out, radius](double x, double y, double z) {
                      Point pi{{0, 0, 0}}, po{{0, 0, 0}};
                      for (int v = 0; v < 4; ++v)
                        {
                          const double w = ((v & 1) ? x : 1 - x) * ((v & 2) ? y : 1 - y);
                          for (int e = 0; e < 3; ++e)
                            {
                              pi[e] += w * in[v][e];
                              po[e] += w * out[v][e];
                            }
                        }
                      const double n = std::sqrt(po[0] * po[0] + po[1] * po[1] + po[2] * po[2]);
                      Point        p;
                      for (int e = 0; e < 3; ++e)
                        p[e] = (1 - z) * pi[e] + z * radius * po[e] / n;
                      return p;
                    });
                  }
            }
        return maps;
      }

      // vertex numbers of the cell corners: points equal up to tol get one number (per-axis clustering of the sorted coordinates, so
      // that no rounding boundary splits two evaluations of the same point), numbered in order of first appearance
      inline void
      merge_vertices(const std::vector<double> &corner_points, std::vector<double> &vertices, std::vector<std::uint32_t> &ids, const double tol = 1e-10)
      {
        const std::size_t                       N = corner_points.size() / 3;
        std::vector<std::array<long long, 3>> key(N);
        std::vector<std::size_t>                order(N);
        for (int d = 0; d < 3; ++d)
          {
            std::iota(order.begin(), order.end(), std::size_t(0));
            std::stable_sort(order.begin(), order.end(), [&](std::size_t a, std::size_t b) { return corner_points[3 * a + d] < corner_points[3 * b + d]; });
            long long cluster = 0;
            for (std::size_t i = 0; i < N; ++i)
              {
                if (i > 0 && corner_points[3 * order[i] + d] - corner_points[3 * order[i - 1] + d] > tol)
                  ++cluster;
                key[order[i]][d] = cluster;
              }
          }
        std::map<std::array<long long, 3>, std::uint32_t> number;
        ids.resize(N);
        vertices.clear();
        for (std::size_t i = 0; i < N; ++i)
          {
            auto it = number.find(key[i]);
            if (it == number.end())
              {
                it = number.emplace(key[i], (std::uint32_t)number.size()).first;
                vertices.insert(vertices.end(), corner_points.begin() + 3 * i, corner_points.begin() + 3 * i + 3);
              }
            ids[i] = it->second;
          }
      }
    } // namespace internal

    // 32 * 8^n_refinements cells
    inline UnstructuredMesh
    hyper_ball(const unsigned int n_refinements = 0, const double radius = 1.0)
    {
      const auto       maps = internal::ball_coarse_maps(radius);
      const int        m    = 1 << n_refinements;
      UnstructuredMesh mesh;
      mesh.support.reserve(maps.size() * (std::size_t)m * m * m * 81);
      std::vector<double> corners;
      corners.reserve(maps.size() * (std::size_t)m * m * m * 24);
      for (const auto &f : maps)
        for (int l = 0; l < m; ++l)
          for (int j = 0; j < m; ++j)
            for (int i = 0; i < m; ++i)
              for (int z = 0; z < 3; ++z)
                for (int y = 0; y < 3; ++y)
                  for (int x = 0; x < 3; ++x)
                    {
                      const auto p = f((i + 0.5 * x) / m, (j + 0.5 * y) / m, (l + 0.5 * z) / m);
                      mesh.support.insert(mesh.support.end(), p.begin(), p.end());
                      if (x != 1 && y != 1 && z != 1)
                        corners.insert(corners.end(), p.begin(), p.end());
                    }
      internal::merge_vertices(corners, mesh.vertices, mesh.cells);
      return mesh;
    }

    // parent[fine cell] = coarse cell | child position << 28 between hyper_ball(n_refinements) and hyper_ball(n_refinements - 1)
    inline std::vector<std::uint32_t>
    ball_parents(const unsigned int n_refinements)
    {
      const long long            mf = 1ll << n_refinements, mc = mf / 2;
      std::vector<std::uint32_t> parent((std::size_t)(32 * mf * mf * mf));
      for (long long f = 0; f < (long long)parent.size(); ++f)
        {
          const long long q = f / (mf * mf * mf), r = f % (mf * mf * mf), i = r % mf, j = (r / mf) % mf, l = r / (mf * mf);
          const long long p = q * mc * mc * mc + ((l / 2) * mc + j / 2) * mc + i / 2;
          parent[f]         = (std::uint32_t)p | (std::uint32_t)(((i & 1) | ((j & 1) << 1) | ((l & 1) << 2)) << 28);
        }
      return parent;
    }
  } // namespace GridGenerator
} // namespace dasm
