// Warp-specialised, TMA-fed brick kernels (sm_100a) for the "lex" bricks of a mesh (mesh.h): full 4 x 4 x 4 bricks whose
// own DoFs are the lexicographic box [0, 4k)^3 = block i of 64 k^3 entries of every vector, i.e. a 4-D tensor
// (x, y, z, brick) that the TMA engine moves with box copies (cp.async.bulk.tensor, SASS UTMALDG).
//
// Work item of a thread block = a z-slab of BZ cell layers of a brick (BZ = 4: the whole brick, k <= 3; BZ = 2: half a brick,
// k = 4, so that TWO blocks are resident per SM and the barrier / latency stalls of one overlap with the work of the other):
//
//   tile of an item = own box {R, R, RZ} (R = 4k, RZ = BZ k)                         one tensor copy
//                   + the faces / edges / corner towards +x, +y, +z                   7 tensor copies: {XW,R,RZ} (x = 0 plane of
//                     the +x brick, XW = 16 bytes wide), {R,1,RZ}, {R,R,1}, {XW,1,RZ}, {XW,R,1}, {R,1,1}, {XW,1,1}
//   all issued by ONE thread on one mbarrier (complete_tx); no index tables, no LSU work, no per-brick index traffic.
//   Bricks with an upper neighbour that is not a local lex brick (partition boundary: ghost entries; irregular bricks)
//   fetch those points through an index list with cp.async instead (mode 1).
//   The epilogue operands (b; x and x_old) of the own box are staged by tensor copies as well.
//
// Compute: thread (cell, plane t), lanes = cells; phases, even-odd contractions and the merge by shuffles are described in
// kernels_fast.cuh.  After the merge every point of the item's closure is final in exactly one thread, which runs the fused
// vector epilogue straight from its registers: plain 16-byte global stores for the private points, red.global.add for the own
// points on the faces shared with bricks processed elsewhere and for the points owned by the upper neighbours.
// Items are walked in chunks of +x neighbours (and the slabs of a brick bottom-up): the contributions to the face between
// two consecutive items are handed over through shared memory, so these faces need neither red.add nor zeroing.
#pragma once
#include <cuda.h>

#include "kernels_fast.cuh"

namespace dasm
{
  // descriptor of a lex brick in the processing order of a kernel launch
  struct TmaBrick
  {
    uint32_t base;     // first DoF of the own box (a multiple of 64 k^3)
    uint32_t flags;    // TMA_MODE1 | TMA_CARRY_IN | TMA_CARRY_OUT | TMA_LAST
    uint32_t nb[7];    // bases of the upper neighbours +x, +y, +z, +xy, +xz, +yz, +xyz (mode 0)
    uint32_t list_off; // mode 1: offset of the brick's foreign index list (NFP entries, face order)
  };
  constexpr int TMA_DW = sizeof(TmaBrick) / 4;
  enum : uint32_t
  {
    TMA_MODE1     = 1u, // an upper neighbour is not a local lex brick: foreign points through the index list
    TMA_CARRY_IN  = 2u, // the previous brick of the chunk is the -x neighbour: its X = R contributions arrive through shared memory,
                        // the own face X = 0 is complete after this brick (plain stores, no red.add / zeroing)
    TMA_CARRY_OUT = 4u, // the next brick of the chunk is the +x neighbour: X = R contributions are handed over, not red.add-ed
    TMA_LAST      = 8u, // last brick of its chunk
    TMA_NONE      = 0xFFFFFFFFu
  };

  // tensor maps of one vector (kernel parameters): own box, +x face, +y face, +z face, edges, corner
  struct TmaMaps
  {
    CUtensorMap main, fx, fy, fz, exy, exz, eyz, cxyz;
  };

  // Bricks are processed in chunks of consecutive +x neighbours: thread block b walks the chunks b, b + grid, ... and the
  // bricks [chunk_start[c], chunk_start[c + 1]) of a chunk in order.
  struct TmaList
  {
    const TmaBrick *bricks;      // whole list
    const uint32_t *foreign;     // index lists of the mode-1 bricks
    const uint32_t *chunk_start; // [n_chunks + 1] of the chunks this launch processes (indices into bricks)
    int             n_chunks;
    int             any_mode1;
  };

  struct TmaWalk
  {
    uint32_t idx;        // current brick
    uint32_t next_start; // first brick of this block's next chunk (TMA_NONE: none)
    int      chunk;
  };

  // geometry of the foreign index list of a brick (whole brick, face order):
  //   fx[Z][Y] | fy[Z][X] | fz[Y][X] | exy[Z] | exz[Y] | eyz[X] | c
  template <int k>
  struct TmaListGeom
  {
    static constexpr int R     = 4 * k;
    static constexpr int J_FX  = 0;
    static constexpr int J_FY  = R * R;
    static constexpr int J_FZ  = 2 * R * R;
    static constexpr int J_EXY = 3 * R * R;
    static constexpr int J_EXZ = J_EXY + R;
    static constexpr int J_EYZ = J_EXZ + R;
    static constexpr int J_C   = J_EYZ + R;
    static constexpr int NFOR  = J_C + 1;
    static constexpr int NFP   = (NFOR + 3) / 4 * 4;
  };

  template <int k, typename T, int BZ>
  struct TmaGeom
  {
    static constexpr int n      = k + 1;
    static constexpr int R      = 4 * k;  // box edge in x and y
    static constexpr int RZ     = BZ * k; // box edge in z of one item
    static constexpr int NH     = 4 / BZ; // items per brick
    static constexpr int XW     = 16 / (int)sizeof(T);
    static constexpr int NB     = R * R * R;  // DoFs of a brick
    static constexpr int NBI    = R * R * RZ; // DoFs of an item
    static constexpr int NCELLS = 16 * BZ;
    static constexpr int NCT    = NCELLS * n;
    static constexpr int NT     = NCT;
    static constexpr int CS     = (n * n * n) | 1;
    __host__ __device__ static constexpr int
    pad(const int e)
    {
      return (e * (int)sizeof(T) + 127) / 128 * 128 / (int)sizeof(T);
    }
    // tile regions (element offsets, 128-byte aligned: TMA destinations); Z is local to the item (0 .. RZ)
    static constexpr int O_MAIN = 0;                       // [Z][Y][X] (layout TmaLayout)
    static constexpr int O_FX   = O_MAIN + pad(NBI);       // X = R:          [Z][Y][XW]
    static constexpr int O_FY   = O_FX + pad(RZ * R * XW); // Y = R:          [Z][X]
    static constexpr int O_FZ   = O_FY + pad(RZ * R);      // Z = RZ:         [Y][X]
    static constexpr int O_EXY  = O_FZ + pad(R * R);       // X = R, Y = R:   [Z][XW]
    static constexpr int O_EXZ  = O_EXY + pad(RZ * XW);    // X = R, Z = RZ:  [Y][XW]
    static constexpr int O_EYZ  = O_EXZ + pad(R * XW);     // Y = R, Z = RZ:  [X]
    static constexpr int O_C    = O_EYZ + pad(R);          // corner
    static constexpr int TILE   = O_C + pad(XW);
    static constexpr unsigned BYTES_MAIN = (unsigned)(NBI * sizeof(T));
    static constexpr unsigned BYTES_FOR  = (unsigned)((RZ * R * XW + RZ * R + R * R + RZ * XW + R * XW + R + XW) * sizeof(T));
    // foreign points of an item in its own face order: fx[Z][Y] | fy[Z][X] | fz[Y][X] | exy[Z] | exz[Y] | eyz[X] | c
    static constexpr int I_FY  = RZ * R;
    static constexpr int I_FZ  = 2 * RZ * R;
    static constexpr int I_EXY = I_FZ + R * R;
    static constexpr int I_EXZ = I_EXY + RZ;
    static constexpr int I_EYZ = I_EXZ + R;
    static constexpr int I_C   = I_EYZ + R;
    static constexpr int NFOR  = I_C + 1;
    static constexpr int NFT   = (NFOR + NCT - 1) / NCT;
  };

  // Layout of a box [0, R) x [0, R) x [0, RZ) in shared memory (tile main region, operand boxes).  Natural: row (Y, Z) of R
  // elements at (Y + R Z) R.  k = 4: a cell row is 4 elements and the rows of the four cells (cx, cy = 0..3) a quarter-warp
  // reads with one 16-byte load per lane must cover all 32 banks:
  //   PERM  the rows are stored in the order rho = ((Y >> 2) & 1) + 2 (Y & 3) + 8 (Y >> 3) + 16 Z, i.e. the rows Y and Y + 4
  //         (cells cy and cy + 1) are adjacent; the tensor map enumerates the box as (x, (Y>>2)&1, Y&3, (Y>>3) + 2 Z) to make
  //         the TMA engine write it that way.  float: the two adjacent 64-byte rows are one 128-byte line: conflict-free.
  //   SWZ   double (128-byte rows): additionally the 16-byte chunks of a row are XOR-ed with rho & 7 (the 128-byte swizzle
  //         mode of the tensor map, CuTe Swizzle<3,4,3>): the lanes (cx, cy & 1) hit the 8 different chunks.
  template <int k, typename T>
  struct TmaLayout
  {
    static constexpr bool PERM = (k == 4);
    static constexpr bool SWZ  = (k == 4 && sizeof(T) == 8);
    static constexpr int  R    = 4 * k;
    __host__ __device__ static constexpr int
    row(const int Y, const int Z)
    {
      return PERM ? ((Y >> 2) & 1) + 2 * (Y & 3) + 8 * (Y >> 3) + 16 * Z : Y + R * Z;
    }
    __host__ __device__ static constexpr int
    swz(const int rho)
    {
      return SWZ ? (rho & 7) : 0;
    }
  };

  __device__ __forceinline__ void
  walk_init(TmaWalk &w, const TmaList &list)
  {
    w.chunk      = blockIdx.x;
    w.idx        = ldg_early(list.chunk_start + w.chunk);
    w.next_start = (w.chunk + (int)gridDim.x < list.n_chunks) ? ldg_early(list.chunk_start + w.chunk + gridDim.x) : (uint32_t)TMA_NONE;
  }
  __device__ __forceinline__ uint32_t
  walk_peek(const TmaWalk &w, const bool last)
  {
    return last ? w.next_start : w.idx + 1u;
  }
  __device__ __forceinline__ void
  walk_advance(TmaWalk &w, const bool last, const TmaList &list)
  {
    if (!last)
      {
        ++w.idx;
        return;
      }
    w.chunk += (int)gridDim.x;
    w.idx        = w.next_start;
    w.next_start = (w.chunk + (int)gridDim.x < list.n_chunks) ? ldg_early(list.chunk_start + w.chunk + gridDim.x) : (uint32_t)TMA_NONE;
  }

  // ---- TMA primitives ------------------------------------------------------------------------------------------------------
  __device__ __forceinline__ void
  tma_load_4d(const unsigned dst, const CUtensorMap *map, const int c0, const int c1, const int c2, const int c3, const unsigned mbar)
  {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(mbar)
                 : "memory");
  }

  // own box of item (brick at `base`, slab hz) through the `main` map of a vector
  template <int k, typename T, int BZ>
  __device__ __forceinline__ void
  tma_load_box(const unsigned dst, const CUtensorMap *map, const uint32_t base, const int hz, const unsigned mbar)
  {
    using G       = TmaGeom<k, T, BZ>;
    const int brk = (int)(base / G::NB), z0 = hz * G::RZ;
    if (TmaLayout<k, T>::PERM)
      tma_load_4d(dst, map, 0, 0, 0, 32 * brk + 2 * z0, mbar);
    else
      tma_load_4d(dst, map, 0, 0, z0, brk, mbar);
  }

  // all tensor copies of the tile of one item (one elected thread)
  template <int k, typename T, int BZ>
  __device__ __forceinline__ void
  tma_issue_tile(T *tile, const TmaMaps &maps, const uint32_t *desc, const int hz, const unsigned mbar)
  {
    using G                = TmaGeom<k, T, BZ>;
    const unsigned t0      = (unsigned)__cvta_generic_to_shared(tile);
    const uint32_t base    = desc[0];
    const bool     mode0   = (desc[1] & TMA_MODE1) == 0u;
    constexpr unsigned ES  = (unsigned)sizeof(T);
    mbar_expect_tx(mbar, mode0 ? G::BYTES_MAIN + G::BYTES_FOR : G::BYTES_MAIN);
    tma_load_box<k, T, BZ>(t0 + G::O_MAIN * ES, &maps.main, base, hz, mbar);
    if (mode0)
      {
        // the +z plane of the item lies in the bricks themselves (next slab) or in their +z neighbours (top slab)
        const bool top = (hz == G::NH - 1);
        const int  z0 = hz * G::RZ, zc = top ? 0 : z0 + G::RZ;
        const int  b_x = (int)(desc[2] / G::NB), b_y = (int)(desc[3] / G::NB), b_xy = (int)(desc[5] / G::NB);
        const int  b_z = (int)((top ? desc[4] : base) / G::NB), b_xz = top ? (int)(desc[6] / G::NB) : b_x;
        const int  b_yz = top ? (int)(desc[7] / G::NB) : b_y, b_xyz = top ? (int)(desc[8] / G::NB) : b_xy;
        tma_load_4d(t0 + G::O_FX * ES, &maps.fx, 0, 0, z0, b_x, mbar);
        tma_load_4d(t0 + G::O_FY * ES, &maps.fy, 0, 0, z0, b_y, mbar);
        tma_load_4d(t0 + G::O_FZ * ES, &maps.fz, 0, 0, zc, b_z, mbar);
        tma_load_4d(t0 + G::O_EXY * ES, &maps.exy, 0, 0, z0, b_xy, mbar);
        tma_load_4d(t0 + G::O_EXZ * ES, &maps.exz, 0, 0, zc, b_xz, mbar);
        tma_load_4d(t0 + G::O_EYZ * ES, &maps.eyz, 0, 0, zc, b_yz, mbar);
        tma_load_4d(t0 + G::O_C * ES, &maps.cxyz, 0, 0, zc, b_xyz, mbar);
      }
  }

  // global index of the tile point (X, Y, Zl) of item (desc, hz) with X = R, Y = R or Zl = RZ (a point the item does not own)
  template <int k, typename T, int BZ>
  __device__ __forceinline__ uint32_t
  tma_foreign_index(const uint32_t *desc, const uint32_t *__restrict__ lists, const int hz, const int X, const int Y, const int Zl)
  {
    using G            = TmaGeom<k, T, BZ>;
    using LG           = TmaListGeom<k>;
    constexpr int R    = G::R;
    const bool    top  = (hz == G::NH - 1);
    const bool    xR = (X == R), yR = (Y == R), zR = (Zl == G::RZ);
    const int     Z    = hz * G::RZ + Zl; // within the brick (R for the plane above the top slab)
    const bool    zf   = zR && top;       // the point lies in a +z neighbour
    if ((desc[1] & TMA_MODE1) == 0u)
      {
        // owner among self / the 7 upper neighbours and the offset in its box
        const int      q   = (xR ? 1 : 0) | (yR ? 2 : 0) | (zf ? 4 : 0); // bits: x, y, z
        const uint32_t own = q == 0 ? desc[0] :
                                      (q == 1 ? desc[2] :
                                                (q == 2 ? desc[3] : (q == 3 ? desc[5] : (q == 4 ? desc[4] : (q == 5 ? desc[6] : (q == 6 ? desc[7] : desc[8]))))));
        return own + (uint32_t)((xR ? 0 : X) + R * ((yR ? 0 : Y) + R * (zf ? 0 : Z)));
      }
    if (!xR && !yR && !zf)
      return desc[0] + (uint32_t)(X + R * (Y + R * Z)); // own brick (the plane between two slabs)
    const int j = zf ? (xR ? (yR ? LG::J_C : LG::J_EXZ + Y) : (yR ? LG::J_EYZ + X : LG::J_FZ + Y * R + X)) :
                       (xR ? (yR ? LG::J_EXY + Z : LG::J_FX + Z * R + Y) : LG::J_FY + Z * R + X);
    return __ldg(lists + desc[9] + j);
  }

  // mode 1: foreign points of the tile through the index list (cp.async by the compute threads)
  template <int k, typename T, int BZ>
  __device__ __forceinline__ void
  tma_foreign_gather(T *tile, const uint32_t *desc, const uint32_t *__restrict__ lists, const int hz, const T *__restrict__ src, const int tid)
  {
    using G         = TmaGeom<k, T, BZ>;
    constexpr int R = G::R, RZ = G::RZ, XW = G::XW;
    // all index loads first (independent, one round of latency), then the copies
    uint32_t gi[G::NFT];
    int      oo[G::NFT];
#pragma unroll
    for (int jj = 0; jj < G::NFT; ++jj)
      {
        const int j = tid + jj * G::NCT;
        oo[jj]      = -1;
        gi[jj]      = 0;
        if (j < G::NFOR)
          {
            int X, Y, Zl, o;
            if (j < G::I_FY)
              X = R, Y = j % R, Zl = j / R, o = G::O_FX + j * XW;
            else if (j < G::I_FZ)
              X = (j - G::I_FY) % R, Y = R, Zl = (j - G::I_FY) / R, o = G::O_FY + (j - G::I_FY);
            else if (j < G::I_EXY)
              X = (j - G::I_FZ) % R, Y = (j - G::I_FZ) / R, Zl = RZ, o = G::O_FZ + (j - G::I_FZ);
            else if (j < G::I_EXZ)
              X = R, Y = R, Zl = j - G::I_EXY, o = G::O_EXY + (j - G::I_EXY) * XW;
            else if (j < G::I_EYZ)
              X = R, Y = j - G::I_EXZ, Zl = RZ, o = G::O_EXZ + (j - G::I_EXZ) * XW;
            else if (j < G::I_C)
              X = j - G::I_EYZ, Y = R, Zl = RZ, o = G::O_EYZ + (j - G::I_EYZ);
            else
              X = R, Y = R, Zl = RZ, o = G::O_C;
            oo[jj] = o;
            gi[jj] = tma_foreign_index<k, T, BZ>(desc, lists, hz, X, Y, Zl);
          }
      }
#pragma unroll
    for (int jj = 0; jj < G::NFT; ++jj)
      if (oo[jj] >= 0)
        cp_async_value(tile + oo[jj], src + gi[jj]);
  }

  // per-thread addressing of a cell plane in the tile: rows r = 0..k (row r < k starts at r0 + r rs, row k at rk; the x
  // position inside a row depends on the layout and the row's swizzle (s0 ^ (r SX)) & sm, sk), and the point x = k of the
  // cells cx = 3 (owned by the +x neighbours) through a second set of offsets
  struct PlaneAddr
  {
    int r0, rs, rk;
    int x0, xs, xk;
    int s0, sm, sk;
  };

  // Laplace phase A: plane y = t of cell (cx, cy, cz), rows z   (cz: cell layer inside the item)
  template <int k, typename T, int BZ>
  __device__ __forceinline__ PlaneAddr
  tile_plane_y(const int cy, const int cz, const int t)
  {
    using G         = TmaGeom<k, T, BZ>;
    using L         = TmaLayout<k, T>;
    constexpr int R = G::R, XW = G::XW;
    const int     Y = k * cy + t, Z0 = k * cz;
    const bool    yR = (Y == R), zR = (cz == BZ - 1);
    PlaneAddr     a;
    a.r0 = yR ? G::O_FY + Z0 * R : G::O_MAIN + L::row(Y, Z0) * R;
    a.rs = yR ? R : R * R;
    a.s0 = yR ? 0 : L::swz(L::row(Y, 0));
    a.sm = yR ? 0 : 7;
    a.rk = zR ? (yR ? G::O_EYZ : G::O_FZ + Y * R) : a.r0 + k * a.rs;
    a.sk = zR ? 0 : a.s0;
    a.x0 = yR ? G::O_EXY + Z0 * XW : G::O_FX + (Z0 * R + Y) * XW;
    a.xs = yR ? XW : R * XW;
    a.xk = zR ? (yR ? G::O_C : G::O_EXZ + Y * XW) : a.x0 + k * a.xs;
    return a;
  }

  // FDM phase A: plane z = t of cell (cx, cy, cz), rows y
  template <int k, typename T, int BZ>
  __device__ __forceinline__ PlaneAddr
  tile_plane_z(const int cy, const int cz, const int t)
  {
    using G         = TmaGeom<k, T, BZ>;
    using L         = TmaLayout<k, T>;
    constexpr int R = G::R, XW = G::XW;
    const int     Z = k * cz + t, Y0 = k * cy;
    const bool    zR = (Z == G::RZ), yR = (cy == 3);
    PlaneAddr     a;
    a.r0 = zR ? G::O_FZ + Y0 * R : G::O_MAIN + L::row(Y0, Z) * R;
    a.rs = zR ? R : (L::PERM ? 2 * R : R);
    a.s0 = zR ? 0 : L::swz(L::row(Y0, 0));
    a.sm = zR ? 0 : 7;
    a.rk = yR ? (zR ? G::O_EYZ : G::O_FY + Z * R) : (zR ? a.r0 + k * R : G::O_MAIN + L::row(Y0 + k, Z) * R);
    a.sk = (yR || zR) ? 0 : L::swz(L::row(Y0 + k, 0));
    a.x0 = zR ? G::O_EXZ + Y0 * XW : G::O_FX + (Z * R + Y0) * XW;
    a.xs = XW;
    a.xk = yR ? (zR ? G::O_C : G::O_EXY + Z * XW) : a.x0 + k * XW;
    return a;
  }

  // the n values of row r of a cell plane (SX: swizzle increment per row: 0 for rows z, 2 for rows y)
  template <int k, typename T, int SX>
  __device__ __forceinline__ void
  plane_load_row(T (&v)[k + 1], const T *tile, const PlaneAddr &a, const int r, const int cx)
  {
    using L        = TmaLayout<k, T>;
    const int rowo = r < k ? a.r0 + r * a.rs : a.rk;
    const int sw   = r < k ? ((a.s0 ^ (r * SX)) & a.sm) : a.sk;
    if constexpr (L::SWZ)
      {
        const double2 p0 = *reinterpret_cast<const double2 *>(tile + rowo + (((2 * cx) ^ sw) << 1));
        const double2 p1 = *reinterpret_cast<const double2 *>(tile + rowo + (((2 * cx + 1) ^ sw) << 1));
        v[0]             = p0.x;
        v[1]             = p0.y;
        v[2]             = p1.x;
        v[3]             = p1.y;
        // x = k is x = 0 of the +x neighbour cell (lane + 1), or a point of the +x face of the brick
        const double nbv = __shfl_down_sync(0xffffffffu, p0.x, 1);
        v[4]             = nbv;
        if (cx == 3)
          v[4] = tile[r < k ? a.x0 + r * a.xs : a.xk];
      }
    else if constexpr (L::PERM)
      {
        const float4 p0 = *reinterpret_cast<const float4 *>(tile + rowo + 4 * cx);
        v[0]            = p0.x;
        v[1]            = p0.y;
        v[2]            = p0.z;
        v[3]            = p0.w;
        const float nbv = __shfl_down_sync(0xffffffffu, p0.x, 1);
        v[4]            = nbv;
        if (cx == 3)
          v[4] = tile[r < k ? a.x0 + r * a.xs : a.xk];
      }
    else
      {
        const T *row = tile + rowo + k * cx;
#pragma unroll
        for (int x = 0; x < k; ++x)
          v[x] = row[x];
        const T nbv = __shfl_down_sync(0xffffffffu, v[0], 1);
        v[k]        = nbv;
        if (cx == 3)
          v[k] = tile[r < k ? a.x0 + r * a.xs : a.xk];
      }
  }

  // ---- fused vector epilogue from the registers of the compute threads ------------------------------------------------------
  // All epilogues are one affine form: dst = sa a + cy y + f1 (a - b)
  //   EPI_STORE (0, 1, 0)   EPI_RESIDUAL (1, -1, 0)   EPI_CHEB (1, f2, f1)   EPI_SCALE (0, f2, 0)
  // with the operands a (b) read only if NOPS >= 1 (NOPS == 2).
  template <typename T>
  struct EpiCoef
  {
    T cy, f1;
  };

  template <typename T>
  __device__ __forceinline__ EpiCoef<T>
  epi_coef(const Epilogue<T> &epi)
  {
    EpiCoef<T> c;
    c.cy = (epi.kind == EPI_RESIDUAL) ? T(-1) : ((epi.kind == EPI_CHEB || epi.kind == EPI_SCALE) ? epi.f2 : T(1));
    c.f1 = (epi.kind == EPI_CHEB) ? epi.f1 : T(0); // (v1 == nullptr: b = 0, the f1 term stays)
    return c;
  }

  // carries of an item (shared memory):
  //   cx_in / cx_out  [RZ + 1][R + 1] values of the plane X = R (Z slow, Y fast) handed to the same slab of the +x brick
  //   cz_in / cz_out  [R + 1][R + 1]  values of the plane Zl = RZ (Y slow, X fast) handed to the next slab of the same brick
  template <typename T>
  struct TmaCarry
  {
    const T *cx_in;
    T *      cx_out;
    const T *cz_in;  // written by the slab below (double buffer by slab parity: a middle slab reads one and writes the other)
    T *      cz_out;
  };

  // After the merge every thread (cell, plane z = t) holds the FINAL values of its exclusive points x < k (or x <= k for
  // cx = 3), y < k (or y <= k for cy = 3) of the plane Zl = k cz + t of the item.
  // NOPS: number of epilogue operands staged in shared memory (0: store / scale, 1: residual or update without x_old, 2: update)
  // HOIST (items with an index list, mode 1): the global indices of the points other bricks own are loaded up front, so that the
  // loads of a thread are in flight together instead of one round trip in front of every red.add
  template <int k, typename T, int BZ, int NOPS, bool HOIST>
  __device__ __forceinline__ void
  tma_epilogue(T (&r)[k + 1][k + 1], const T *ops0, const T *ops1, const TmaCarry<T> &cr, T *__restrict__ dst, T *__restrict__ sh_dst,
               const bool direct, T *__restrict__ ni_out, const EpiCoef<T> &ec, const uint32_t *desc, const uint32_t *__restrict__ lists,
               const int hz, const int cx, const int cy, const int cz, const int t)
  {
    using G         = TmaGeom<k, T, BZ>;
    using L         = TmaLayout<k, T>;
    constexpr int R = G::R, RZ = G::RZ;
    const uint32_t  base = desc[0];
    // (only in the fused sequences: with the accumulator protocol finish_shared_kernel owns every DoF of the shared-DoF list)
    const bool      cin = direct && (desc[1] & TMA_CARRY_IN) != 0u, cout = direct && (desc[1] & TMA_CARRY_OUT) != 0u;
    const int       Zl = k * cz + t;   // plane inside the item (RZ: the plane above it)
    const int       Z  = hz * RZ + Zl; // plane inside the brick
    const bool      zcarry_in  = (hz > 0) && (Zl == 0);          // the slab below (processed just before) contributed to this plane
    const bool      zcarry_out = (hz < G::NH - 1) && (Zl == RZ); // this plane belongs to the next slab of the brick
    const T         sh_a = direct ? ec.cy : T(1);                // factor of y in the red.add of a point another brick owns
    const int       Y0 = k * cy, X0 = k * cx;
    // contributions of the slab below to the plane Zl = 0 (all points of the plane, X = R / Y = R included)
    if (zcarry_in)
      {
        const T *c = cr.cz_in + Y0 * (R + 1) + X0;
#pragma unroll
        for (int y = 0; y <= k; ++y)
#pragma unroll
          for (int x = 0; x <= k; ++x)
            if ((x < k || cx == 3) && (y < k || cy == 3))
              r[y][x] += c[y * (R + 1) + x];
      }
    // contributions of the -x neighbour (the same slab of the previous brick of this block) to the plane X = 0
    // (not for a plane that is handed to the next slab: its x-neighbour contributions are merged there)
    if (cin && cx == 0 && !zcarry_out)
      {
        const T *c = cr.cx_in + Zl * (R + 1) + Y0;
#pragma unroll
        for (int y = 0; y < k; ++y)
          r[y][0] += c[y];
        if (cy == 3)
          r[k][0] += c[k];
      }
    // the plane above the item goes to the next slab as a whole
    if (zcarry_out)
      {
        T *c = cr.cz_out + Y0 * (R + 1) + X0;
#pragma unroll
        for (int y = 0; y <= k; ++y)
#pragma unroll
          for (int x = 0; x <= k; ++x)
            if ((x < k || cx == 3) && (y < k || cy == 3))
              c[y * (R + 1) + x] = r[y][x];
        return;
      }
    // contributions to the plane X = R: handed to the same slab of the next brick of this block
    if (cout && cx == 3)
      {
        T *c = cr.cx_out + Zl * (R + 1) + Y0;
#pragma unroll
        for (int y = 0; y < k; ++y)
          c[y] = r[y][k];
        if (cy == 3)
          c[k] = r[k][k];
      }
    const bool xred = (cx == 3) && !cout; // X = R points go to the neighbours with red.add
    const bool x0sh = (cx == 0) && !cin;  // the own face X = 0 is shared with a brick processed elsewhere
    uint32_t gf[HOIST ? k + 1 : 1][HOIST ? k + 1 : 1];
    if (!HOIST)
      {
      }
    else if (Zl < RZ)
      {
        if (xred)
          {
#pragma unroll
            for (int y = 0; y < k; ++y)
              gf[y][k] = tma_foreign_index<k, T, BZ>(desc, lists, hz, R, Y0 + y, Zl);
          }
        if (cy == 3)
          {
#pragma unroll
            for (int x = 0; x < k; ++x)
              gf[k][x] = tma_foreign_index<k, T, BZ>(desc, lists, hz, X0 + x, R, Zl);
            if (xred)
              gf[k][k] = tma_foreign_index<k, T, BZ>(desc, lists, hz, R, R, Zl);
          }
      }
    else
      {
#pragma unroll
        for (int y = 0; y < k; ++y)
          {
#pragma unroll
            for (int x = 0; x < k; ++x)
              gf[y][x] = tma_foreign_index<k, T, BZ>(desc, lists, hz, X0 + x, Y0 + y, RZ);
            if (xred)
              gf[y][k] = tma_foreign_index<k, T, BZ>(desc, lists, hz, R, Y0 + y, RZ);
          }
        if (cy == 3)
          {
#pragma unroll
            for (int x = 0; x < k; ++x)
              gf[k][x] = tma_foreign_index<k, T, BZ>(desc, lists, hz, X0 + x, R, RZ);
            if (xred)
              gf[k][k] = tma_foreign_index<k, T, BZ>(desc, lists, hz, R, R, RZ);
          }
      }
    if (Zl < RZ)
      {
        const bool     zsh = (Z == 0);
        const uint32_t g0  = base + (uint32_t)((Z * R + Y0) * R + X0);
        // operand rows: row y at rowo0 + y rstep, swizzle sw0 ^ (2 y)
        const int     rho0 = L::row(Y0, Zl), rowo0 = rho0 * R, sw0 = L::swz(rho0);
        constexpr int rstep = (L::PERM ? 2 : 1) * R;
#pragma unroll
        for (int y = 0; y < k; ++y)
          {
            const uint32_t g  = g0 + (uint32_t)(y * R);
            const bool     sh = zsh || (Y0 + y == 0);
            T              a[k], b[k], res[k];
            if constexpr (NOPS >= 1)
              {
                const int rowo = rowo0 + y * rstep, sw = sw0 ^ (L::SWZ ? 2 * y : 0);
                if constexpr (L::SWZ)
                  {
                    const double2 a0 = *reinterpret_cast<const double2 *>(ops0 + rowo + (((2 * cx) ^ sw) << 1));
                    const double2 a1 = *reinterpret_cast<const double2 *>(ops0 + rowo + (((2 * cx + 1) ^ sw) << 1));
                    a[0] = a0.x, a[1] = a0.y, a[2] = a1.x, a[3] = a1.y;
                    if constexpr (NOPS == 2)
                      {
                        const double2 b0 = *reinterpret_cast<const double2 *>(ops1 + rowo + (((2 * cx) ^ sw) << 1));
                        const double2 b1 = *reinterpret_cast<const double2 *>(ops1 + rowo + (((2 * cx + 1) ^ sw) << 1));
                        b[0] = b0.x, b[1] = b0.y, b[2] = b1.x, b[3] = b1.y;
                      }
                  }
                else if constexpr (L::PERM)
                  {
                    const float4 a0 = *reinterpret_cast<const float4 *>(ops0 + rowo + 4 * cx);
                    a[0] = a0.x, a[1] = a0.y, a[2] = a0.z, a[3] = a0.w;
                    if constexpr (NOPS == 2)
                      {
                        const float4 b0 = *reinterpret_cast<const float4 *>(ops1 + rowo + 4 * cx);
                        b[0] = b0.x, b[1] = b0.y, b[2] = b0.z, b[3] = b0.w;
                      }
                  }
                else
                  {
#pragma unroll
                    for (int x = 0; x < k; ++x)
                      {
                        a[x] = ops0[rowo + X0 + x];
                        if constexpr (NOPS == 2)
                          b[x] = ops1[rowo + X0 + x];
                      }
                  }
              }
#pragma unroll
            for (int x = 0; x < k; ++x)
              {
                if constexpr (NOPS == 2)
                  res[x] = a[x] + ec.cy * r[y][x] + ec.f1 * (a[x] - b[x]);
                else if constexpr (NOPS == 1)
                  res[x] = a[x] + ec.cy * r[y][x] + ec.f1 * a[x];
                else
                  res[x] = ec.cy * r[y][x];
              }
            if (sh)
              {
                // own DoFs on the shared faces Y = 0 / Z = 0: this brick adds the full epilogue value; the next kernel's
                // destination is zeroed there
#pragma unroll
                for (int x = 0; x < k; ++x)
                  {
                    atomic_add(sh_dst + g + x, direct ? res[x] : r[y][x]);
                    if (ni_out != nullptr)
                      ni_out[g + x] = T(0);
                  }
              }
            else
              {
                if (x0sh)
                  {
                    atomic_add(sh_dst + g, direct ? res[0] : r[y][0]);
                    if (ni_out != nullptr)
                      ni_out[g] = T(0);
#pragma unroll
                    for (int x = 1; x < k; ++x)
                      dst[g + x] = res[x];
                  }
                else if constexpr (k == 4 && sizeof(T) == 8)
                  {
                    *reinterpret_cast<double2 *>(dst + g)     = make_double2(res[0], res[1]);
                    *reinterpret_cast<double2 *>(dst + g + 2) = make_double2(res[2], res[3]);
                  }
                else if constexpr (k == 4 && sizeof(T) == 4)
                  *reinterpret_cast<float4 *>(dst + g) = make_float4(res[0], res[1], res[2], res[3]);
                else if constexpr (k == 2 && sizeof(T) == 8)
                  *reinterpret_cast<double2 *>(dst + g) = make_double2(res[0], res[1]);
                else
                  {
#pragma unroll
                    for (int x = 0; x < k; ++x)
                      dst[g + x] = res[x];
                  }
              }
            if (xred) // X = R: face of the +x neighbour
              atomic_add(sh_dst + (HOIST ? gf[HOIST ? y : 0][HOIST ? k : 0] : tma_foreign_index<k, T, BZ>(desc, lists, hz, R, Y0 + y, Zl)), sh_a * r[y][k]);
          }
        if (cy == 3) // Y = R: face of the +y neighbour, edge of the +xy neighbour
          {
#pragma unroll
            for (int x = 0; x < k; ++x)
              atomic_add(sh_dst + (HOIST ? gf[HOIST ? k : 0][HOIST ? x : 0] : tma_foreign_index<k, T, BZ>(desc, lists, hz, X0 + x, R, Zl)), sh_a * r[k][x]);
            if (xred)
              atomic_add(sh_dst + (HOIST ? gf[HOIST ? k : 0][HOIST ? k : 0] : tma_foreign_index<k, T, BZ>(desc, lists, hz, R, R, Zl)), sh_a * r[k][k]);
          }
      }
    else
      {
        // Zl = RZ of the top slab: face of the +z neighbour, edges of +xz, +yz, corner of +xyz
#pragma unroll
        for (int y = 0; y < k; ++y)
          {
#pragma unroll
            for (int x = 0; x < k; ++x)
              atomic_add(sh_dst + (HOIST ? gf[HOIST ? y : 0][HOIST ? x : 0] : tma_foreign_index<k, T, BZ>(desc, lists, hz, X0 + x, Y0 + y, Zl)), sh_a * r[y][x]);
            if (xred)
              atomic_add(sh_dst + (HOIST ? gf[HOIST ? y : 0][HOIST ? k : 0] : tma_foreign_index<k, T, BZ>(desc, lists, hz, R, Y0 + y, Zl)), sh_a * r[y][k]);
          }
        if (cy == 3)
          {
#pragma unroll
            for (int x = 0; x < k; ++x)
              atomic_add(sh_dst + (HOIST ? gf[HOIST ? k : 0][HOIST ? x : 0] : tma_foreign_index<k, T, BZ>(desc, lists, hz, X0 + x, R, Zl)), sh_a * r[k][x]);
            if (xred)
              atomic_add(sh_dst + (HOIST ? gf[HOIST ? k : 0][HOIST ? k : 0] : tma_foreign_index<k, T, BZ>(desc, lists, hz, R, R, Zl)), sh_a * r[k][k]);
          }
      }
  }

  // operand boxes of an item (one elected compute thread, after the barrier that ends the previous epilogue)
  template <int k, typename T, int BZ>
  __device__ __forceinline__ void
  tma_issue_ops(T *ops0, T *ops1, const CUtensorMap *map0, const CUtensorMap *map1, const bool need1, const uint32_t base, const int hz,
                const unsigned mbar)
  {
    using G                  = TmaGeom<k, T, BZ>;
    constexpr unsigned bytes = (unsigned)(G::NBI * sizeof(T));
    mbar_expect_tx(mbar, need1 ? 2 * bytes : bytes);
    tma_load_box<k, T, BZ>((unsigned)__cvta_generic_to_shared(ops0), map0, base, hz, mbar);
    if (need1)
      tma_load_box<k, T, BZ>((unsigned)__cvta_generic_to_shared(ops1), map1, base, hz, mbar);
  }

  // shared memory of the kernels: header (mbarriers, two brick descriptors) | tile | X slots | operand boxes | carry planes
  template <int k, typename T, int BZ>
  struct TmaSmem
  {
    using G = TmaGeom<k, T, BZ>;
    static constexpr int
    pad1k(const int e)
    {
      return (e * (int)sizeof(T) + 1023) / 1024 * 1024 / (int)sizeof(T);
    }
    static constexpr int TILE   = pad1k(G::TILE);
    static constexpr int XSLOT  = pad1k(G::NCELLS * G::CS);
    static constexpr int OPS    = pad1k(G::NBI);
    static constexpr int CARRYX = G::pad((G::RZ + 1) * (G::R + 1)); // per slab and brick parity
    static constexpr int CARRYZ = G::pad((G::R + 1) * (G::R + 1));
    static constexpr int NCZ    = (G::NH > 2) ? 2 : 1; // z-carry buffers: a middle slab reads one and writes the other
    static constexpr int CARRY  = pad1k(2 * G::NH * CARRYX + NCZ * CARRYZ);
    static constexpr size_t
    bytes(const int n_x, const int n_ops)
    {
      return 2048 + (size_t)(TILE + n_x * XSLOT + n_ops * OPS + CARRY) * sizeof(T);
    }
  };

  // Work item and residency per degree and number type:
  //   k <= 3   whole bricks (BZ = 4), two resident blocks per SM (two independent blocks overlap the barrier / latency stalls of one
  //            with the work of the other: +34 % (double), +49 % (float) at k = 3, profiles/r02l_k3_two_blocks.txt)
  //   k  = 4   half bricks (BZ = 2), two blocks per SM
  //   k  = 5   double: single cell layers (BZ = 1, 110 KB), float: half bricks; two blocks per SM
  //   k  = 6   single cell layers; double: one block per SM (170 KB), float: two
  // (shared memory: tile + 2 exchange slots of NCELLS n^3 + operand box + carries, TmaSmem)
  template <int k, typename T>
  struct TmaItem
  {
    static constexpr int BZ   = (k <= 3) ? 4 : (k == 4 ? 2 : ((k == 5 && sizeof(T) == 4) ? 2 : 1));
    static constexpr int MINB = (k == 6 && sizeof(T) == 8) ? 1 : 2;
  };
  // host-side view (chunking and grid size): resident blocks per SM
  inline int
  tma_min_blocks(const int k, const int esize)
  {
    return (k == 6 && esize == 8) ? 1 : 2;
  }
  inline int
  tma_item_layers(const int k, const int esize)
  {
    return (k <= 3) ? 4 : (k == 4 ? 2 : ((k == 5 && esize == 4) ? 2 : 1));
  }

  // barrier over all threads of the block (all threads are compute threads; bar.sync with a thread count needs a multiple of 32)
  template <int NCT>
  __device__ __forceinline__ void
  block_sync()
  {
    if constexpr (NCT % 32 == 0)
      bar_sync(FB_COMPUTE, NCT);
    else
      __syncthreads();
  }

  // state of the walk over the items of a block
  struct TmaCursor
  {
    TmaWalk w;
    int     hz;  // slab of the current brick
    int     par; // parity of the current brick (descriptor / x-carry ping-pong)
  };

  // first tile of a block: descriptor -> s_desc[0], tensor copies (and the index-list gather of a mode-1 brick)
  template <int k, typename T, int BZ, bool M1>
  __device__ __forceinline__ void
  tma_prologue(uint32_t *s_desc, T *tile, const TmaMaps &tmaps, const TmaList &list, const uint32_t idx, const T *__restrict__ src,
               const unsigned mb_tile, const int tid)
  {
    using G            = TmaGeom<k, T, BZ>;
    const uint32_t *dw = reinterpret_cast<const uint32_t *>(list.bricks);
    if (tid < TMA_DW)
      s_desc[tid] = ldg_early(dw + (size_t)idx * TMA_DW + tid);
    block_sync<G::NCT>();
    if (tid == 0)
      tma_issue_tile<k, T, BZ>(tile, tmaps, s_desc, 0, mb_tile);
    if (M1 && (s_desc[1] & TMA_MODE1))
      tma_foreign_gather<k, T, BZ>(tile, s_desc, list.foreign, 0, src, tid);
  }

  // after the barrier behind phase A: the tile is dead, fetch the next item into it
  template <int k, typename T, int BZ, bool M1>
  __device__ __forceinline__ void
  tma_fetch_next(const uint32_t *dn, const int hz_next, T *tile, const TmaMaps &tmaps, const TmaList &list, const T *__restrict__ src,
                 const unsigned mb_tile, const int tid)
  {
    if (tid == 0)
      tma_issue_tile<k, T, BZ>(tile, tmaps, dn, hz_next, mb_tile);
    if (M1 && (dn[1] & TMA_MODE1))
      tma_foreign_gather<k, T, BZ>(tile, dn, list.foreign, hz_next, src, tid);
  }

  // ---- Laplace, uniform Cartesian geometry --------------------------------------------------------------------------------------
  // M1: the launch contains mode-1 bricks (index lists): gather of the foreign points with cp.async and an epilogue that loads
  // the indices up front; launches without such bricks run the leaner instantiation
  template <int k, typename T, int NOPS, bool M1>
  __global__ void __launch_bounds__((TmaGeom<k, T, TmaItem<k, T>::BZ>::NT), (TmaItem<k, T>::MINB))
  laplace_tma_kernel(const T *__restrict__ src,
                     T *__restrict__ dst,
                     T *__restrict__ acc,
                     const Epilogue<T> epi,
                     const __grid_constant__ FastLaplaceMats<T, k + 1> mats,
                     const __grid_constant__ TmaMaps tmaps,
                     const __grid_constant__ CUtensorMap omap0,
                     const int         shared_mode,
                     const NextInit<T> ni,
                     const TmaList     list,
                     const FastMaps    dbgmaps)
  {
    constexpr int BZ = TmaItem<k, T>::BZ;
    using G          = TmaGeom<k, T, BZ>;
    using SM         = TmaSmem<k, T, BZ>;
    constexpr int n  = k + 1;
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    // the tile and the operand boxes are written by the TMA engine with the 128-byte swizzle pattern: 1024-byte alignment
    unsigned char *smem_raw = smem_dyn + ((1024u - ((unsigned)__cvta_generic_to_shared(smem_dyn) & 1023u)) & 1023u);
    // [0] mbarrier of the tile, [8] mbarrier of the operand staging, [64..] descriptors of two bricks (ping-pong)
    const unsigned mb_tile = (unsigned)__cvta_generic_to_shared(smem_raw);
    const unsigned mb_ops  = mb_tile + 8;
    uint32_t *     s_desc  = reinterpret_cast<uint32_t *>(smem_raw + 64);
    T *            tile    = reinterpret_cast<T *>(smem_raw + 1024);
    T *            Xq      = tile + SM::TILE;
    T *            Xp      = Xq + SM::XSLOT;
    T *            ops0    = Xp + SM::XSLOT;
    T *            carry   = ops0 + SM::OPS;
    if ((int)blockIdx.x >= list.n_chunks)
      return;
    if (threadIdx.x == 0)
      {
        mbar_init(mb_tile, 1);
        mbar_init(mb_ops, 1);
      }
    __syncthreads();

    const int        tid = threadIdx.x;
    const int        c = tid % G::NCELLS, t = tid / G::NCELLS;
    const int        cx = c & 3, cy = (c >> 2) & 3, cz = c >> 4;
    // BZ = 4: the warp (t = k, layers 0 / 1) has nothing to do in the last phase (its planes belong to the cells above)
    const bool       skip_last = (BZ == 4 && (t == k) && (cz < 2)) || (dbgmaps.dbg & 1);
    const PlaneAddr  pa = tile_plane_y<k, T, BZ>(cy, cz, t);
    T *              xq = Xq + c * G::CS, *xp = Xp + c * G::CS;
    const EpiCoef<T> ec     = epi_coef(epi);
    constexpr bool   need0  = NOPS >= 1;
    const bool       direct = (shared_mode == SHARED_DIRECT);
    const uint32_t * dw     = reinterpret_cast<const uint32_t *>(list.bricks);
    unsigned         tphase = 0, ophase = 0;
    TmaCursor        cur;
    cur.hz  = 0;
    cur.par = 0; // s_desc[par]: this brick, s_desc[par ^ 1]: the next one
    walk_init(cur.w, list);
    tma_prologue<k, T, BZ, M1>(s_desc, tile, tmaps, list, cur.w.idx, src, mb_tile, tid);
    for (;;)
      {
        const uint32_t *desc      = s_desc + cur.par * 16;
        const bool      last      = (desc[1] & TMA_LAST) != 0u;
        const bool      last_slab = (cur.hz == G::NH - 1);
        const uint32_t  nidx      = walk_peek(cur.w, last);
        const bool      has_next  = !last_slab || nidx != TMA_NONE;
        mbar_wait(mb_tile, tphase); // the tensor copies of this item's tile have landed
        tphase ^= 1u;
        if (M1)
          cp_async_wait_all();
        // foreign points gathered by the other threads (mode 1); all phase B reads of the exchange slots of the previous item are
        // done before phase A overwrites them; all threads have left the epilogue of the previous item
        block_sync<G::NCT>();
        if (need0 && tid == 0)
          tma_issue_ops<k, T, BZ>(ops0, ops0, &omap0, &omap0, false, desc[0], cur.hz, mb_ops);
        // descriptor of the next brick -> shared memory (fire and forget, awaited before the barrier after phase A; its buffer held
        // the descriptor of the previous brick)
        if (cur.hz == 0 && nidx != TMA_NONE && tid < TMA_DW)
          cp_async_4(s_desc + (cur.par ^ 1) * 16 + tid, dw + (size_t)nidx * TMA_DW + tid);
        // phase A: plane y = t, [z][x]: q = Mx Mz v, p = (g0 Kx Mz + g2 Mx Kz) v
        if constexpr (n >= 7)
          {
            // two passes over the plane (one n x n array in registers instead of two): first q = Mx Mz v and p = g2 Mx Kz v,
            // then p += g0 Kx Mz v
            if (!(dbgmaps.dbg & 1))
              {
                T a[n][n];
#pragma unroll
                for (int z = 0; z < n; ++z)
                  {
                    T v[n];
                    plane_load_row<k, T, 0>(v, tile, pa, z, cx);
                    mat_vec<n, T, true, true, false>(a[z], mats.M, v);
                  }
#pragma unroll
                for (int x = 0; x < n; ++x)
                  {
                    T ca[n], q[n], p[n];
#pragma unroll
                    for (int z = 0; z < n; ++z)
                      ca[z] = a[z][x];
                    mat_vec<n, T, true, true, false>(q, mats.M, ca);
                    mat_vec<n, T, true, true, false>(p, mats.K2, ca);
#pragma unroll
                    for (int z = 0; z < n; ++z)
                      {
                        xq[(z * n + t) * n + x] = q[z];
                        xp[(z * n + t) * n + x] = p[z];
                      }
                  }
#pragma unroll
                for (int z = 0; z < n; ++z)
                  {
                    T v[n];
                    plane_load_row<k, T, 0>(v, tile, pa, z, cx);
                    mat_vec<n, T, true, true, false>(a[z], mats.K0, v);
                  }
#pragma unroll
                for (int x = 0; x < n; ++x)
                  {
                    T cb[n], p[n];
#pragma unroll
                    for (int z = 0; z < n; ++z)
                      cb[z] = a[z][x];
                    mat_vec<n, T, true, true, false>(p, mats.M, cb);
#pragma unroll
                    for (int z = 0; z < n; ++z)
                      xp[(z * n + t) * n + x] += p[z];
                  }
              }
          }
        else if (!(dbgmaps.dbg & 1))
          {
            T a[n][n], b[n][n];
#pragma unroll
            for (int z = 0; z < n; ++z)
              {
                T v[n];
                plane_load_row<k, T, 0>(v, tile, pa, z, cx);
                mat_vec<n, T, true, true, false>(a[z], mats.M, v);
                mat_vec<n, T, true, true, false>(b[z], mats.K0, v);
              }
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
        if (cur.hz == 0 && nidx != TMA_NONE && tid < TMA_DW)
          cp_async_wait_all();
        block_sync<G::NCT>();
        if (has_next)
          {
            if (last_slab)
              tma_fetch_next<k, T, BZ, M1>(s_desc + (cur.par ^ 1) * 16, 0, tile, tmaps, list, src, mb_tile, tid);
            else
              tma_fetch_next<k, T, BZ, M1>(desc, cur.hz + 1, tile, tmaps, list, src, mb_tile, tid);
          }
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
        // fused epilogue from the registers
        if (need0)
          mbar_wait(mb_ops, ophase); // the operand box has landed
        ophase ^= 1u;
        if (!skip_last && (t < k || cz == BZ - 1) && !(dbgmaps.dbg & 2))
          {
            const TmaCarry<T> cr = {carry + (cur.hz * 2 + cur.par) * SM::CARRYX, carry + (cur.hz * 2 + (cur.par ^ 1)) * SM::CARRYX,
                                    carry + 2 * G::NH * SM::CARRYX + (SM::NCZ == 2 ? ((cur.hz + 1) & 1) : 0) * SM::CARRYZ,
                                    carry + 2 * G::NH * SM::CARRYX + (SM::NCZ == 2 ? (cur.hz & 1) : 0) * SM::CARRYZ};
            tma_epilogue<k, T, BZ, NOPS, M1>(r, ops0, ops0, cr, dst, direct ? dst : acc, direct, ni.out, ec, desc, list.foreign, cur.hz, cx, cy,
                                                 cz, t);
          }
        if (!has_next)
          break;
        if (last_slab)
          {
            walk_advance(cur.w, last, list);
            cur.hz = 0;
            cur.par ^= 1;
          }
        else
          ++cur.hz;
      }
  }

  // ---- FDM, one eigen-decomposition triple, tensor-product weights folded into the matrices ----------------------------------
  template <int k, typename T, int NOPS, bool M1>
  __global__ void __launch_bounds__((TmaGeom<k, T, TmaItem<k, T>::BZ>::NT), (TmaItem<k, T>::MINB))
  fdm_tma_kernel(const T *__restrict__ src,
                 T *__restrict__ dst,
                 T *__restrict__ acc,
                 const Epilogue<T> epi,
                 const __grid_constant__ FastFdmMats<T, k + 1> mats,
                 const __grid_constant__ TmaMaps tmaps,
                 const __grid_constant__ CUtensorMap omap0,
                 const __grid_constant__ CUtensorMap omap1,
                 const int         shared_mode,
                 const NextInit<T> ni,
                 const TmaList     list,
                 const FastMaps    dbgmaps)
  {
    constexpr int BZ = TmaItem<k, T>::BZ;
    using G          = TmaGeom<k, T, BZ>;
    using SM         = TmaSmem<k, T, BZ>;
    constexpr int n  = k + 1;
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    unsigned char *smem_raw = smem_dyn + ((1024u - ((unsigned)__cvta_generic_to_shared(smem_dyn) & 1023u)) & 1023u);
    const unsigned mb_tile = (unsigned)__cvta_generic_to_shared(smem_raw);
    const unsigned mb_ops  = mb_tile + 8;
    uint32_t *     s_desc  = reinterpret_cast<uint32_t *>(smem_raw + 64);
    T *            tile    = reinterpret_cast<T *>(smem_raw + 1024);
    T *            X       = tile + SM::TILE;
    T *            ops0    = X + SM::XSLOT;
    T *            ops1    = ops0 + SM::OPS;
    T *            carry   = ops1 + SM::OPS;
    __shared__ T   s_inv[n * n * n];
    if ((int)blockIdx.x >= list.n_chunks)
      return;
    for (int i = threadIdx.x; i < n * n * n; i += G::NT)
      s_inv[i] = mats.inv[i];
    if (threadIdx.x == 0)
      {
        mbar_init(mb_tile, 1);
        mbar_init(mb_ops, 1);
      }
    __syncthreads();

    const int        tid = threadIdx.x;
    const int        c = tid % G::NCELLS, t = tid / G::NCELLS;
    const int        cx = c & 3, cy = (c >> 2) & 3, cz = c >> 4;
    const bool       skip_last = (BZ == 4 && (t == k) && (cz < 2)) || (dbgmaps.dbg & 1);
    const PlaneAddr  pa = tile_plane_z<k, T, BZ>(cy, cz, t);
    T *              xs = X + c * G::CS;
    const T *        inv = s_inv + t * n; // row (z, y = t) of this thread's plane in phase B: broadcast reads
    const EpiCoef<T> ec     = epi_coef(epi);
    constexpr bool   need0 = NOPS >= 1, need1 = NOPS == 2;
    const bool       direct = (shared_mode == SHARED_DIRECT);
    const uint32_t * dw     = reinterpret_cast<const uint32_t *>(list.bricks);
    unsigned         tphase = 0, ophase = 0;
    TmaCursor        cur;
    cur.hz  = 0;
    cur.par = 0;
    walk_init(cur.w, list);
    tma_prologue<k, T, BZ, M1>(s_desc, tile, tmaps, list, cur.w.idx, src, mb_tile, tid);
    for (;;)
      {
        const uint32_t *desc      = s_desc + cur.par * 16;
        const bool      last      = (desc[1] & TMA_LAST) != 0u;
        const bool      last_slab = (cur.hz == G::NH - 1);
        const uint32_t  nidx      = walk_peek(cur.w, last);
        const bool      has_next  = !last_slab || nidx != TMA_NONE;
        mbar_wait(mb_tile, tphase);
        tphase ^= 1u;
        if (M1)
          cp_async_wait_all();
        // foreign points gathered by the other threads (mode 1); all phase C reads of the exchange slot of the previous item are
        // done before phase A overwrites it; all threads have left the epilogue of the previous item
        block_sync<G::NCT>();
        if (need0 && tid == 0)
          tma_issue_ops<k, T, BZ>(ops0, ops1, &omap0, &omap1, need1, desc[0], cur.hz, mb_ops);
        if (cur.hz == 0 && nidx != TMA_NONE && tid < TMA_DW)
          cp_async_4(s_desc + (cur.par ^ 1) * 16 + tid, dw + (size_t)nidx * TMA_DW + tid);
        // phase A: plane z = t, [y][x]: Ax in x, Ay in y
        if (!(dbgmaps.dbg & 1))
          {
            T a[n][n];
#pragma unroll
            for (int y = 0; y < n; ++y)
              {
                T v[n];
                plane_load_row<k, T, 2>(v, tile, pa, y, cx);
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
        if (cur.hz == 0 && nidx != TMA_NONE && tid < TMA_DW)
          cp_async_wait_all();
        block_sync<G::NCT>();
        if (has_next)
          {
            if (last_slab)
              tma_fetch_next<k, T, BZ, M1>(s_desc + (cur.par ^ 1) * 16, 0, tile, tmaps, list, src, mb_tile, tid);
            else
              tma_fetch_next<k, T, BZ, M1>(desc, cur.hz + 1, tile, tmaps, list, src, mb_tile, tid);
          }
        // phase B: plane y = t, [z][x]: Az, scale, Bz in z; Bx in x
        if (!(dbgmaps.dbg & 1))
          {
            T wv[n][n];
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
                  wv[z][x] = col[z];
              }
#pragma unroll
            for (int z = 0; z < n; ++z)
              {
                T u[n];
                mat_vec<n, T, false, true, false>(u, mats.Bx, wv[z]);
#pragma unroll
                for (int x = 0; x < n; ++x)
                  xs[(z * n + t) * n + x] = u[x];
              }
          }
        block_sync<G::NCT>();
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
        if (need0)
          mbar_wait(mb_ops, ophase);
        ophase ^= 1u;
        if (!skip_last && (t < k || cz == BZ - 1) && !(dbgmaps.dbg & 2))
          {
            const TmaCarry<T> cr = {carry + (cur.hz * 2 + cur.par) * SM::CARRYX, carry + (cur.hz * 2 + (cur.par ^ 1)) * SM::CARRYX,
                                    carry + 2 * G::NH * SM::CARRYX + (SM::NCZ == 2 ? ((cur.hz + 1) & 1) : 0) * SM::CARRYZ,
                                    carry + 2 * G::NH * SM::CARRYX + (SM::NCZ == 2 ? (cur.hz & 1) : 0) * SM::CARRYZ};
            tma_epilogue<k, T, BZ, NOPS, M1>(r, ops0, ops1, cr, dst, direct ? dst : acc, direct, ni.out, ec, desc, list.foreign, cur.hz, cx, cy,
                                                 cz, t);
          }
        if (!has_next)
          break;
        if (last_slab)
          {
            walk_advance(cur.w, last, list);
            cur.hz = 0;
            cur.par ^= 1;
          }
        else
          ++cur.hz;
      }
  }
} // namespace dasm
