// Warp-specialised, TMA-fed brick kernels (sm_100a) for the "lex" bricks of a mesh (mesh.h): full 4 x 4 x 4 bricks whose
// own DoFs are the lexicographic box [0, 4k)^3 = block i of 64 k^3 entries of every vector, i.e. a 4-D tensor
// (x, y, z, brick) that the TMA engine moves with box copies (cp.async.bulk.tensor, SASS UTMALDG):
//
//   tile of a brick = own box {R, R, R} (R = 4k)                                  one tensor copy
//                   + the faces / edges / corner owned by the 7 upper neighbours   7 tensor copies: {XW,R,R} (x = 0 plane of the
//                     +x brick, XW = 16 bytes wide), {R,1,R}, {R,R,1}, {XW,1,R}, {XW,R,1}, {R,1,1}, {XW,1,1}
//   all issued by ONE thread on one mbarrier (complete_tx); no index tables, no LSU work, no per-brick index traffic.
//   Bricks with an upper neighbour that is not a local lex brick (partition boundary: ghost entries; irregular bricks)
//   fetch the foreign points through an index list with cp.async instead (mode 1).
//   epilogue operands (b; x and x_old) of the own box: 1-D bulk copies (the box is contiguous in global memory).
//
// Compute phases, even-odd contractions, the merge by shuffles and the named-barrier hand-over to the mover warps are
// those of kernels_fast.cuh (which remains the path for the old numbering, DASM_NO_LEX=1).  The mover warps run the
// fused vector epilogue in the order of the box: 16-byte shared-memory loads and global stores, red.global.add for the
// box points on the shared lower faces (X = 0, Y = 0 or Z = 0) and for the points owned by the upper neighbours.
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
                        // the own face X = 0 is complete after this brick (plain stores, no red.add / pre-initialisation)
    TMA_CARRY_OUT = 4u, // the next brick of the chunk is the +x neighbour: X = R contributions are handed over, not red.add-ed
    TMA_LAST      = 8u, // last brick of its chunk
    TMA_NONE      = 0xFFFFFFFFu
  };

  // tensor maps of one vector (kernel parameters)
  struct TmaMaps
  {
    CUtensorMap main, fx, fy, fz, exy, exz, eyz, cxyz;
  };

  // Bricks are processed in chunks of consecutive +x neighbours: thread block b walks the chunks b, b + grid, ... and the
  // bricks [chunk_start[c], chunk_start[c + 1]) of a chunk in order, so that the contributions to the face between two bricks
  // of a chunk never leave the SM.
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

  template <int k, typename T>
  struct TmaGeom
  {
    static constexpr int n      = k + 1;
    static constexpr int R      = 4 * k;
    static constexpr int XW     = 16 / (int)sizeof(T);
    static constexpr int NB     = R * R * R;
    static constexpr int NCELLS = 64;
    static constexpr int NCT    = NCELLS * n;
    static constexpr int NT     = NCT;
    static constexpr int CS     = (n * n * n) | 1;
    static constexpr int V      = 16 / (int)sizeof(T); // elements per 16-byte vector
    __host__ __device__ static constexpr int
    pad(const int e)
    {
      return (e * (int)sizeof(T) + 127) / 128 * 128 / (int)sizeof(T);
    }
    // tile regions (element offsets, 128-byte aligned: TMA destinations)
    static constexpr int O_MAIN = 0;
    static constexpr int O_FX   = O_MAIN + pad(NB);
    static constexpr int O_FY   = O_FX + pad(R * R * XW);
    static constexpr int O_FZ   = O_FY + pad(R * R);
    static constexpr int O_EXY  = O_FZ + pad(R * R);
    static constexpr int O_EXZ  = O_EXY + pad(R * XW);
    static constexpr int O_EYZ  = O_EXZ + pad(R * XW);
    static constexpr int O_C    = O_EYZ + pad(R);
    static constexpr int TILE   = O_C + pad(XW);
    static constexpr unsigned BYTES_MAIN = (unsigned)(NB * sizeof(T));
    static constexpr unsigned BYTES_FOR  = (unsigned)((R * R * XW + 2 * R * R + 2 * R * XW + R + XW) * sizeof(T));
    // points owned by the upper neighbours in face order: fx[Z][Y] | fy[Z][X] | fz[Y][X] | exy[Z] | exz[Y] | eyz[X] | c
    static constexpr int J_FY   = R * R;
    static constexpr int J_FZ   = 2 * R * R;
    static constexpr int J_EXY  = 3 * R * R;
    static constexpr int J_EXZ  = J_EXY + R;
    static constexpr int J_EYZ  = J_EXZ + R;
    static constexpr int J_C    = J_EYZ + R;
    static constexpr int NFOR   = J_C + 1;
    static constexpr int NFP    = (NFOR + 3) / 4 * 4;
    static constexpr int NFORP  = pad(NFOR);
    static constexpr int NSH    = NB - (R - 1) * (R - 1) * (R - 1); // own DoFs on the shared lower faces
    static constexpr int NFT    = (NFOR + NCT - 1) / NCT;
    static constexpr int XSLOT  = pad(NCELLS * CS);
    // tile offset of the j-th foreign point
    __host__ __device__ static constexpr int
    foreign_tile_offset(const int j)
    {
      return j < J_FY ? O_FX + j * XW :
                        (j < J_FZ ? O_FY + (j - J_FY) :
                                    (j < J_EXY ? O_FZ + (j - J_FZ) :
                                                 (j < J_EXZ ? O_EXY + (j - J_EXY) * XW :
                                                              (j < J_EYZ ? O_EXZ + (j - J_EXZ) * XW : (j < J_C ? O_EYZ + (j - J_EYZ) : O_C)))));
    }
    // offset of the j-th foreign point relative to the base of the neighbour brick that owns it, and which neighbour
    __host__ __device__ static constexpr int
    foreign_owner(const int j)
    {
      return j < J_FY ? 0 : (j < J_FZ ? 1 : (j < J_EXY ? 2 : (j < J_EXZ ? 3 : (j < J_EYZ ? 4 : (j < J_C ? 5 : 6)))));
    }
    __host__ __device__ static constexpr int
    foreign_box_offset(const int j)
    {
      // fx (Z, Y): R Y + R^2 Z = R j;  fy (Z, X): X + R^2 Z;  fz (Y, X): X + R Y;  exy (Z): R^2 Z;  exz (Y): R Y;  eyz (X): X
      return j < J_FY ? R * j :
                        (j < J_FZ ? ((j - J_FY) % R) + R * R * ((j - J_FY) / R) :
                                    (j < J_EXY ? (j - J_FZ) : (j < J_EXZ ? R * R * (j - J_EXY) : (j < J_EYZ ? R * (j - J_EXZ) : (j < J_C ? (j - J_EYZ) : 0)))));
    }
    // box index of the s-th own DoF on the shared lower faces: plane Z = 0, then per Z >= 1 the row Y = 0 and the column X = 0
    __host__ __device__ static constexpr int
    shared_box_index(const int s)
    {
      if (s < R * R)
        return s;
      const int q = s - R * R, Z = 1 + q / (2 * R - 1), w = q % (2 * R - 1);
      return Z * R * R + (w < R ? w : (w - R + 1) * R);
    }
  };

  // Layout of a box [0, R)^3 in shared memory (tile main region, output box).  Natural: row (Y, Z) of R elements at
  // (Y + R Z) R.  k = 4: a cell row is 4 elements and the rows of the four cells (cx, cy = 0..3) a quarter-warp reads with
  // one 16-byte load per lane must cover all 32 banks:
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
    row_y(const int rho) // inverse: Y of a row
    {
      return PERM ? (((rho & 1) << 2) | ((rho >> 1) & 3) | (rho & 8)) : rho % R;
    }
    __host__ __device__ static constexpr int
    row_z(const int rho)
    {
      return PERM ? (rho >> 4) : rho / R;
    }
    __host__ __device__ static constexpr int
    swz(const int rho)
    {
      return SWZ ? (rho & 7) : 0;
    }
    // element offset of column X in a row with swizzle s
    __host__ __device__ static constexpr int
    col(const int X, const int s)
    {
      return SWZ ? ((((X >> 1) ^ s) << 1) | (X & 1)) : X;
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

  // all tensor copies of one tile (one elected thread)
  template <int k, typename T>
  __device__ __forceinline__ void
  tma_issue_tile(T *tile, const TmaMaps &maps, const uint32_t *desc, const unsigned mbar)
  {
    using G                = TmaGeom<k, T>;
    const unsigned t0      = (unsigned)__cvta_generic_to_shared(tile);
    const uint32_t base    = desc[0];
    const bool     mode0   = (desc[1] & TMA_MODE1) == 0u;
    constexpr unsigned ES  = (unsigned)sizeof(T);
    mbar_expect_tx(mbar, mode0 ? G::BYTES_MAIN + G::BYTES_FOR : G::BYTES_MAIN);
    tma_load_4d(t0 + G::O_MAIN * ES, &maps.main, 0, 0, 0, (int)(base / G::NB) * (TmaLayout<k, T>::PERM ? 32 : 1), mbar);
    if (mode0)
      {
        tma_load_4d(t0 + G::O_FX * ES, &maps.fx, 0, 0, 0, (int)(desc[2] / G::NB), mbar);
        tma_load_4d(t0 + G::O_FY * ES, &maps.fy, 0, 0, 0, (int)(desc[3] / G::NB), mbar);
        tma_load_4d(t0 + G::O_FZ * ES, &maps.fz, 0, 0, 0, (int)(desc[4] / G::NB), mbar);
        tma_load_4d(t0 + G::O_EXY * ES, &maps.exy, 0, 0, 0, (int)(desc[5] / G::NB), mbar);
        tma_load_4d(t0 + G::O_EXZ * ES, &maps.exz, 0, 0, 0, (int)(desc[6] / G::NB), mbar);
        tma_load_4d(t0 + G::O_EYZ * ES, &maps.eyz, 0, 0, 0, (int)(desc[7] / G::NB), mbar);
        tma_load_4d(t0 + G::O_C * ES, &maps.cxyz, 0, 0, 0, (int)(desc[8] / G::NB), mbar);
      }
  }

  // mode 1: foreign points of the tile through the index list (cp.async by the compute threads)
  template <int k, typename T>
  __device__ __forceinline__ void
  tma_foreign_gather(T *tile, const TmaList &list, const uint32_t list_off, const T *__restrict__ src, const int tid)
  {
    using G = TmaGeom<k, T>;
    uint32_t gf[G::NFT];
#pragma unroll
    for (int jj = 0; jj < G::NFT; ++jj)
      {
        const int j = tid + jj * G::NCT;
        gf[jj]      = (j < G::NFOR) ? ldg_early(list.foreign + list_off + j) : 0u;
      }
#pragma unroll
    for (int jj = 0; jj < G::NFT; ++jj)
      {
        const int j = tid + jj * G::NCT;
        if (j < G::NFOR)
          cp_async_value(tile + G::foreign_tile_offset(j), src + gf[jj]);
      }
  }

  // per-thread addressing of a cell plane in the tile / output box: rows r = 0..k (row r < k starts at r0 + r rs, row k at rk;
  // the x position inside a row depends on the layout and the row's swizzle (s0 ^ (r SX)) & sm, sk), and the point x = k of
  // the cells cx = 3 (owned by the +x neighbours) through a second set of offsets
  struct PlaneAddr
  {
    int r0, rs, rk;
    int x0, xs, xk;
    int s0, sm, sk;
  };

  // Laplace phase A: plane y = t of cell (cx, cy, cz), rows z
  template <int k, typename T>
  __device__ __forceinline__ PlaneAddr
  tile_plane_y(const int cx, const int cy, const int cz, const int t)
  {
    using G       = TmaGeom<k, T>;
    using L       = TmaLayout<k, T>;
    constexpr int R = G::R, XW = G::XW;
    const int     Y = k * cy + t, Z0 = k * cz;
    const bool    yR = (Y == R), zR = (cz == 3);
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
    (void)cx;
    return a;
  }

  // FDM phase A: plane z = t of cell (cx, cy, cz), rows y
  template <int k, typename T>
  __device__ __forceinline__ PlaneAddr
  tile_plane_z(const int cx, const int cy, const int cz, const int t)
  {
    using G       = TmaGeom<k, T>;
    using L       = TmaLayout<k, T>;
    constexpr int R = G::R, XW = G::XW;
    const int     Z = k * cz + t, Y0 = k * cy;
    const bool    zR = (Z == R), yR = (cy == 3);
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
    (void)cx;
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
  // After the merge every thread (cell, plane z = t) holds the FINAL values of its exclusive points x < k (or x <= k for
  // cx = 3), y < k (or y <= k for cy = 3) of the plane Z = k cz + t.  Points inside the own box and off the shared lower
  // faces are private: epilogue with the operands from the (TMA-staged) operand boxes in shared memory and plain global
  // stores straight from the registers.  Points on X = 0, Y = 0 or Z = 0 of the own box, and the points owned by the 7 upper
  // neighbours (X = R, Y = R or Z = R), get red.global.add of alpha y (shared-face protocol of kernels_brick.cuh).
  // All epilogues are one affine form: dst = sa a + cy y + f1 (a - b)
  //   EPI_STORE (0, 1, 0)   EPI_RESIDUAL (1, -1, 0)   EPI_CHEB (1, f2, f1)   EPI_SCALE (0, f2, 0)
  // with the operands a (b) read only if need0 (need1).
  template <typename T>
  struct EpiCoef
  {
    T    sa, cy, f1;
    bool need0, need1;
  };

  template <typename T>
  __device__ __forceinline__ EpiCoef<T>
  epi_coef(const Epilogue<T> &epi)
  {
    EpiCoef<T> c;
    c.need0 = (epi.kind == EPI_RESIDUAL || epi.kind == EPI_CHEB);
    c.need1 = (epi.kind == EPI_CHEB && epi.f1 != T(0) && epi.v1 != nullptr);
    c.sa    = c.need0 ? T(1) : T(0);
    c.cy    = (epi.kind == EPI_RESIDUAL) ? T(-1) : ((epi.kind == EPI_CHEB || epi.kind == EPI_SCALE) ? epi.f2 : T(1));
    c.f1    = (epi.kind == EPI_CHEB) ? epi.f1 : T(0); // (v1 == nullptr: b = 0, the f1 term stays)
    return c;
  }

  // global index of a point owned by upper neighbour q (0 +x, 1 +y, 2 +z, 3 +xy, 4 +xz, 5 +yz, 6 +xyz): box offset `off`
  // in the neighbour's box (mode 0) or entry j of the brick's index list (mode 1)
  __device__ __forceinline__ uint32_t
  foreign_index(const uint32_t *desc, const uint32_t *__restrict__ lists, const int q, const int off, const int j)
  {
    return (desc[1] & TMA_MODE1) == 0u ? desc[2 + q] + (uint32_t)off : __ldg(lists + desc[9] + j);
  }

  // carry: [R + 1][R + 1] values of the plane X = R (Z slow, Y fast) handed from a brick to its +x neighbour
  // NOPS: number of epilogue operands staged in shared memory (0: store / scale, 1: residual or update without x_old, 2: update)
  template <int k, typename T, int NOPS>
  __device__ __forceinline__ void
  tma_epilogue(T (&r)[k + 1][k + 1], const T *ops0, const T *ops1, const T *carry_in, T *carry_out, T *__restrict__ dst, T *__restrict__ sh_dst,
               const bool direct, T *__restrict__ ni_out, const EpiCoef<T> &ec, const uint32_t *desc, const uint32_t *__restrict__ lists,
               const int cx, const int cy, const int cz, const int t)
  {
    using G         = TmaGeom<k, T>;
    using L         = TmaLayout<k, T>;
    constexpr int R = G::R;
    const uint32_t  base = desc[0];
    const bool      cin = (desc[1] & TMA_CARRY_IN) != 0u, cout = (desc[1] & TMA_CARRY_OUT) != 0u;
    const int       Z = k * cz + t;
    const T         sh_a = direct ? ec.cy : T(1); // factor of y in the red.add of a point another brick owns
    // (two carry planes, alternating per brick: no thread writes the plane another one still reads)
    const T *       c_in  = carry_in + Z * (R + 1) + k * cy;
    T *             c_out = carry_out + Z * (R + 1) + k * cy;
    // contributions of the -x neighbour (the previous brick of this block) to the plane X = 0
    if (cin && cx == 0)
      {
#pragma unroll
        for (int y = 0; y < k; ++y)
          r[y][0] += c_in[y];
        if (cy == 3)
          r[k][0] += c_in[k];
      }
    // contributions to the plane X = R: handed to the next brick of this block
    if (cout && cx == 3)
      {
#pragma unroll
        for (int y = 0; y < k; ++y)
          c_out[y] = r[y][k];
        if (cy == 3)
          c_out[k] = r[k][k];
      }
    const bool xred = (cx == 3) && !cout; // X = R points go to the neighbours with red.add
    const bool x0sh = (cx == 0) && !cin;  // the own face X = 0 is shared with a brick processed elsewhere
    if (Z < R)
      {
        const bool     zsh = (Z == 0);
        const int      Y0 = k * cy;
        const uint32_t g0 = base + (uint32_t)((Z * R + Y0) * R + k * cx);
        // operand rows: row y at rowo0 + y rstep, swizzle sw0 ^ (2 y)
        const int rho0 = L::row(Y0, Z), rowo0 = rho0 * R, sw0 = L::swz(rho0);
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
                        a[x] = ops0[rowo + k * cx + x];
                        if constexpr (NOPS == 2)
                          b[x] = ops1[rowo + k * cx + x];
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
              atomic_add(sh_dst + foreign_index(desc, lists, 0, R * (Y0 + y + R * Z), Z * R + Y0 + y), sh_a * r[y][k]);
          }
        if (cy == 3) // Y = R: face of the +y neighbour, edge of the +xy neighbour
          {
#pragma unroll
            for (int x = 0; x < k; ++x)
              atomic_add(sh_dst + foreign_index(desc, lists, 1, k * cx + x + R * R * Z, G::J_FY + Z * R + k * cx + x), sh_a * r[k][x]);
            if (xred)
              atomic_add(sh_dst + foreign_index(desc, lists, 3, R * R * Z, G::J_EXY + Z), sh_a * r[k][k]);
          }
      }
    else
      {
        // Z = R (cz = 3, t = k): face of the +z neighbour, edges of +xz, +yz, corner of +xyz
#pragma unroll
        for (int y = 0; y < k; ++y)
          {
            const int Y = k * cy + y;
#pragma unroll
            for (int x = 0; x < k; ++x)
              atomic_add(sh_dst + foreign_index(desc, lists, 2, k * cx + x + R * Y, G::J_FZ + Y * R + k * cx + x), sh_a * r[y][x]);
            if (xred)
              atomic_add(sh_dst + foreign_index(desc, lists, 4, R * Y, G::J_EXZ + Y), sh_a * r[y][k]);
          }
        if (cy == 3)
          {
#pragma unroll
            for (int x = 0; x < k; ++x)
              atomic_add(sh_dst + foreign_index(desc, lists, 5, k * cx + x, G::J_EYZ + k * cx + x), sh_a * r[k][x]);
            if (xred)
              atomic_add(sh_dst + foreign_index(desc, lists, 6, 0, G::J_C), sh_a * r[k][k]);
          }
      }
  }

  // operand boxes of a brick (one elected compute thread, after the barrier that ends the previous epilogue)
  template <int k, typename T>
  __device__ __forceinline__ void
  tma_issue_ops(T *ops0, T *ops1, const CUtensorMap *map0, const CUtensorMap *map1, const bool need1, const uint32_t base, const unsigned mbar)
  {
    using G                  = TmaGeom<k, T>;
    constexpr unsigned bytes = (unsigned)(G::NB * sizeof(T));
    const int          c3    = (int)(base / G::NB) * (TmaLayout<k, T>::PERM ? 32 : 1);
    mbar_expect_tx(mbar, need1 ? 2 * bytes : bytes);
    tma_load_4d((unsigned)__cvta_generic_to_shared(ops0), map0, 0, 0, 0, c3, mbar);
    if (need1)
      tma_load_4d((unsigned)__cvta_generic_to_shared(ops1), map1, 0, 0, 0, c3, mbar);
  }

  // shared memory of the kernels: header (mbarriers, two brick descriptors) | tile | X slots | operand boxes | carry plane
  template <int k, typename T>
  struct TmaSmem
  {
    using G = TmaGeom<k, T>;
    static constexpr int
    pad1k(const int e)
    {
      return (e * (int)sizeof(T) + 1023) / 1024 * 1024 / (int)sizeof(T);
    }
    static constexpr int TILE  = pad1k(G::TILE);
    static constexpr int XSLOT = pad1k(G::NCELLS * G::CS);
    static constexpr int OPS   = pad1k(G::NB);
    static constexpr int CARRY = pad1k((G::R + 1) * (G::R + 1));
    static constexpr size_t
    bytes(const int n_x, const int n_ops)
    {
      return 2048 + (size_t)(TILE + n_x * XSLOT + n_ops * OPS + 2 * CARRY) * sizeof(T);
    }
  };

  // common pipeline pieces of the two kernels ------------------------------------------------------------------------------
  // first tile of a block: descriptor -> s_desc[0], tensor copies (and the index-list gather of a mode-1 brick)
  template <int k, typename T>
  __device__ __forceinline__ void
  tma_prologue(uint32_t *s_desc, T *tile, const TmaMaps &tmaps, const TmaList &list, const uint32_t idx, const T *__restrict__ src,
               const unsigned mb_tile, const int tid)
  {
    using G            = TmaGeom<k, T>;
    const uint32_t *dw = reinterpret_cast<const uint32_t *>(list.bricks);
    if (tid < TMA_DW)
      s_desc[tid] = ldg_early(dw + (size_t)idx * TMA_DW + tid);
    bar_sync(FB_COMPUTE, G::NCT);
    if (tid == 0)
      tma_issue_tile<k, T>(tile, tmaps, s_desc, mb_tile);
    if (list.any_mode1 && (s_desc[1] & TMA_MODE1))
      tma_foreign_gather<k, T>(tile, list, s_desc[9], src, tid);
  }

  // after the barrier behind phase A: the tile is dead, fetch the next brick into it
  template <int k, typename T>
  __device__ __forceinline__ void
  tma_fetch_next(const uint32_t *dn, T *tile, const TmaMaps &tmaps, const TmaList &list, const T *__restrict__ src, const unsigned mb_tile,
                 const int tid)
  {
    if (tid == 0)
      tma_issue_tile<k, T>(tile, tmaps, dn, mb_tile);
    if (list.any_mode1 && (dn[1] & TMA_MODE1))
      tma_foreign_gather<k, T>(tile, list, dn[9], src, tid);
  }

// resident thread blocks per SM: two for k <= 3 (shared memory and registers allow it; two independent blocks overlap the barrier /
// latency stalls of one with the work of the other: +34 % (double), +49 % (float) at k = 3, profiles/r02l_k3_two_blocks.txt)
#ifndef TMA_MINB
#define TMA_MINB(k) ((k) <= 3 ? 2 : 1)
#endif
  // ---- Laplace, uniform Cartesian geometry --------------------------------------------------------------------------------------
  template <int k, typename T, int NOPS>
  __global__ void __launch_bounds__(TmaGeom<k, T>::NT, TMA_MINB(k))
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
    using G           = TmaGeom<k, T>;
    using SM          = TmaSmem<k, T>;
    constexpr int n   = k + 1;
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

    const int       tid = threadIdx.x;
    const int       c = tid % G::NCELLS, t = tid / G::NCELLS;
    const int       cx = c & 3, cy = (c >> 2) & 3, cz = c >> 4;
    const bool      skip_last = ((t == k) && (cz < 2)) || (dbgmaps.dbg & 1); // whole warp: its plane z = k belongs to the cell above
    const PlaneAddr pa = tile_plane_y<k, T>(cx, cy, cz, t);
    T *             xq = Xq + c * G::CS, *xp = Xp + c * G::CS;
    const EpiCoef<T> ec     = epi_coef(epi);
    constexpr bool   need0  = NOPS >= 1;
    const bool       direct = (shared_mode == SHARED_DIRECT);
    const uint32_t * dw     = reinterpret_cast<const uint32_t *>(list.bricks);
    unsigned        tphase = 0, ophase = 0;
    int             par = 0; // s_desc[par]: this brick, s_desc[par ^ 1]: the next one
    TmaWalk         w;
    walk_init(w, list);
    tma_prologue<k, T>(s_desc, tile, tmaps, list, w.idx, src, mb_tile, tid);
    for (;; par ^= 1)
      {
        const uint32_t *desc     = s_desc + par * 16;
        const bool      last     = (desc[1] & TMA_LAST) != 0u;
        const uint32_t  nidx     = walk_peek(w, last);
        const bool      has_next = nidx != TMA_NONE;
        mbar_wait(mb_tile, tphase); // the tensor copies of this brick's tile have landed
        tphase ^= 1u;
        if (list.any_mode1)
          cp_async_wait_all();
        // foreign points gathered by the other threads (mode 1); all phase B reads of the exchange slots of the previous brick are
        // done before phase A overwrites them
        bar_sync(FB_COMPUTE, G::NCT);
        // all threads have left the epilogue of the previous brick: its operand box may be overwritten
        if (need0 && tid == 0)
          tma_issue_ops<k, T>(ops0, ops0, &omap0, &omap0, false, desc[0], mb_ops);
        // descriptor of the next brick -> shared memory (fire and forget, awaited before the barrier after phase A; its buffer held
        // the descriptor of the previous brick)
        if (has_next && tid < TMA_DW)
          cp_async_4(s_desc + (par ^ 1) * 16 + tid, dw + (size_t)nidx * TMA_DW + tid);
        // phase A: plane y = t, [z][x]: q = Mx Mz v, p = (g0 Kx Mz + g2 Mx Kz) v
        if (!(dbgmaps.dbg & 1))
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
        if (has_next && tid < TMA_DW)
          cp_async_wait_all();
        bar_sync(FB_COMPUTE, G::NCT);
        if (has_next)
          tma_fetch_next<k, T>(s_desc + (par ^ 1) * 16, tile, tmaps, list, src, mb_tile, tid);
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
        if (!skip_last && (t < k || cz == 3) && !(dbgmaps.dbg & 2))
          tma_epilogue<k, T, NOPS>(r, ops0, ops0, carry + par * SM::CARRY, carry + (par ^ 1) * SM::CARRY, dst, direct ? dst : acc, direct, ni.out, ec, desc,
                             list.foreign, cx, cy, cz, t);
        if (!has_next)
          break;
        walk_advance(w, last, list);
      }
  }

  // ---- FDM, one eigen-decomposition triple, tensor-product weights folded into the matrices ----------------------------------
  template <int k, typename T, int NOPS>
  __global__ void __launch_bounds__(TmaGeom<k, T>::NT, TMA_MINB(k))
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
    using G           = TmaGeom<k, T>;
    using SM          = TmaSmem<k, T>;
    constexpr int n   = k + 1;
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

    const int       tid = threadIdx.x;
    const int       c = tid % G::NCELLS, t = tid / G::NCELLS;
    const int       cx = c & 3, cy = (c >> 2) & 3, cz = c >> 4;
    const bool      skip_last = ((t == k) && (cz < 2)) || (dbgmaps.dbg & 1);
    const PlaneAddr pa = tile_plane_z<k, T>(cx, cy, cz, t);
    T *             xs = X + c * G::CS;
    const T *       inv = s_inv + t * n; // row (z, y = t) of this thread's plane in phase B: broadcast reads
    const EpiCoef<T> ec     = epi_coef(epi);
    constexpr bool   need0 = NOPS >= 1, need1 = NOPS == 2;
    const bool       direct = (shared_mode == SHARED_DIRECT);
    const uint32_t * dw     = reinterpret_cast<const uint32_t *>(list.bricks);
    unsigned        tphase = 0, ophase = 0;
    int             par = 0;
    TmaWalk         w;
    walk_init(w, list);
    tma_prologue<k, T>(s_desc, tile, tmaps, list, w.idx, src, mb_tile, tid);
    for (;; par ^= 1)
      {
        const uint32_t *desc     = s_desc + par * 16;
        const bool      last     = (desc[1] & TMA_LAST) != 0u;
        const uint32_t  nidx     = walk_peek(w, last);
        const bool      has_next = nidx != TMA_NONE;
        mbar_wait(mb_tile, tphase);
        tphase ^= 1u;
        if (list.any_mode1)
          cp_async_wait_all();
        // foreign points gathered by the other threads (mode 1); all phase C reads of the exchange slot of the previous brick are
        // done before phase A overwrites it
        bar_sync(FB_COMPUTE, G::NCT);
        if (need0 && tid == 0)
          tma_issue_ops<k, T>(ops0, ops1, &omap0, &omap1, need1, desc[0], mb_ops);
        if (has_next && tid < TMA_DW)
          cp_async_4(s_desc + (par ^ 1) * 16 + tid, dw + (size_t)nidx * TMA_DW + tid);
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
        if (has_next && tid < TMA_DW)
          cp_async_wait_all();
        bar_sync(FB_COMPUTE, G::NCT);
        if (has_next)
          tma_fetch_next<k, T>(s_desc + (par ^ 1) * 16, tile, tmaps, list, src, mb_tile, tid);
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
        bar_sync(FB_COMPUTE, G::NCT);
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
        if (!skip_last && (t < k || cz == 3) && !(dbgmaps.dbg & 2))
          tma_epilogue<k, T, NOPS>(r, ops0, ops1, carry + par * SM::CARRY, carry + (par ^ 1) * SM::CARRY, dst, direct ? dst : acc, direct, ni.out, ec, desc,
                             list.foreign, cx, cy, cz, t);
        if (!has_next)
          break;
        walk_advance(w, last, list);
      }
  }
} // namespace dasm
