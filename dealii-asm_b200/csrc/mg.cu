// Multigrid V-cycle on the device: two-level transfers and the cycle of PreconditionerGMG (reference include/multigrid.h:109-537;
// the arithmetic is deal.II's MGTwoLevelTransfer / Multigrid::level_v_step / PreconditionMG, set up in
// element_centered_preconditioners_01.cc:540-740).  Built on the public C ABI (include/dasm.h) only: level operators, smoothers and
// the ghost exchange are the objects of dasm_lib.cu.
//
// Transfer kernels: one thread block per FINE cell.  The (k_c + 1)^3 values of the parent coarse cell are gathered through its 27
// compressed start indices, expanded by three 1-D contractions with the (k_f + 1) x (k_c + 1) embedding matrix of the child
// position, and added to the fine vector with the inverse valence of the fine DoF as weight (so that the contributions of the cells
// sharing a DoF sum to the interpolated value); the restriction is the transpose with atomic adds into the coarse vector.
// HBM traffic: one read + one atomic update of the fine vector, the coarse values stay in L2 (8 children per coarse cell).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "basis.h"
#include "dasm.h"

namespace
{
  constexpr uint32_t INVALID  = 0xFFFFFFFFu;
  constexpr uint32_t LEX_FLAG = 0x80000000u;
  constexpr int      MAXN     = 9; // k <= 8

#define MG_CUDA_CHECK(x)                                                                                     \
  do                                                                                                         \
    {                                                                                                        \
      cudaError_t e_ = (x);                                                                                  \
      if (e_ != cudaSuccess)                                                                                 \
        throw std::runtime_error(std::string("CUDA error ") + cudaGetErrorString(e_) + " at " + __FILE__ + ":" + \
                                 std::to_string(__LINE__));                                                  \
    }                                                                                                        \
  while (0)
#define MG_CALL(x)                                   \
  do                                                 \
    {                                                \
      if ((x) != 0)                                  \
        throw std::runtime_error(dasm_last_error()); \
    }                                                \
  while (0)
#define MG_API_BEGIN try {
#define MG_API_END                          \
  return 0;                                 \
  }                                         \
  catch (const std::exception &e)           \
  {                                         \
    dasm_set_last_error(e.what());          \
    return 1;                               \
  }

  // global index of local DoF (x, y, z) of a cell of degree k from its 27 compressed start indices (kernels.cuh compressed_index
  // with a run-time degree)
  __device__ __forceinline__ uint32_t
  decode(const uint32_t *__restrict__ ci, const int k, const int x, const int y, const int z)
  {
    const int      ex = (x == 0) ? 0 : ((x == k) ? 2 : 1), ey = (y == 0) ? 0 : ((y == k) ? 2 : 1), ez = (z == 0) ? 0 : ((z == k) ? 2 : 1);
    const int      ox = (ex == 1) ? x - 1 : 0, oy = (ey == 1) ? y - 1 : 0, oz = (ez == 1) ? z - 1 : 0;
    const uint32_t start = ci[ex + 3 * ey + 9 * ez];
    if (start == INVALID)
      return INVALID;
    const bool lex = (start & LEX_FLAG) != 0;
    const int  sx = lex ? 4 * k : ((ex == 1) ? (k - 1) : 1), sy = lex ? 4 * k : ((ey == 1) ? (k - 1) : 1);
    return (start & ~LEX_FLAG) + ox + sx * (oy + sy * oz);
  }

  struct TransferArgs
  {
    const uint32_t *cidx_f, *cidx_c; // 27 start indices per cell
    const uint32_t *parent;          // per fine cell: parent coarse cell (local index) | child code << 28 (bits x, y, z)
    const uint8_t * valence;         // per fine cell: 27 entity valences
    const uint32_t *plain_f = nullptr, *plain_c = nullptr; // unstructured meshes: (k+1)^3 oriented addresses per cell
    int             kf, kc;
    long long       n_cells_f;
  };

  // P[child][i_f * nc + i_c] in constant-size kernel parameter space
  template <typename T>
  struct TransferMats
  {
    T P[2][MAXN * MAXN];
  };

  // contraction of u (size a x b x c along the x, y, z axes of a [z][y][x] array) along `dir` with M (no x ni, or its transpose)
  template <typename T, bool TRANSPOSE>
  __device__ __forceinline__ void
  contract(const T *M, const int no, const int ni, const T *in, T *out, const int nx, const int ny, const int nz, const int dir)
  {
    // sizes of the OUTPUT array
    const int ox = dir == 0 ? no : nx, oy = dir == 1 ? no : ny, oz = dir == 2 ? no : nz;
    const int stride = dir == 0 ? 1 : (dir == 1 ? nx : nx * ny);
    for (int o = threadIdx.x; o < ox * oy * oz; o += blockDim.x)
      {
        const int x = o % ox, y = (o / ox) % oy, z = o / (ox * oy);
        const int j = dir == 0 ? x : (dir == 1 ? y : z);
        const int base = (dir == 0 ? 0 : x) + nx * ((dir == 1 ? 0 : y) + ny * (dir == 2 ? 0 : z));
        T         s = 0;
        for (int i = 0; i < ni; ++i)
          s += (TRANSPOSE ? M[i * no + j] : M[j * ni + i]) * in[base + i * stride];
        out[o] = s;
      }
  }

  template <typename T, bool RESTRICT>
  __global__ void __launch_bounds__(128)
  transfer_kernel(T *__restrict__ dst, const T *__restrict__ src, const TransferArgs a, const __grid_constant__ TransferMats<T> mats)
  {
    __shared__ T        buf0[MAXN * MAXN * MAXN], buf1[MAXN * MAXN * MAXN];
    __shared__ uint32_t ci_f[27], ci_c[27];
    __shared__ T        wv[27];
    const long long     cell = blockIdx.x;
    const uint32_t      pc = a.parent[cell];
    const uint32_t      par = pc & 0x0FFFFFFFu;
    const int           child[3] = {(int)((pc >> 28) & 1u), (int)((pc >> 29) & 1u), (int)((pc >> 30) & 1u)};
    const int           nf = a.kf + 1, nc = a.kc + 1;
    if (threadIdx.x < 27)
      {
        ci_f[threadIdx.x] = a.cidx_f[cell * 27 + threadIdx.x];
        ci_c[threadIdx.x] = a.cidx_c[(long long)par * 27 + threadIdx.x];
        const int v       = a.valence[cell * 27 + threadIdx.x];
        wv[threadIdx.x]   = T(1) / T(v > 0 ? v : 1);
      }
    __syncthreads();
    if (!RESTRICT)
      {
        for (int i = threadIdx.x; i < nc * nc * nc; i += blockDim.x)
          {
            const uint32_t g = a.plain_c ? a.plain_c[(long long)par * nc * nc * nc + i] : decode(ci_c, a.kc, i % nc, (i / nc) % nc, i / (nc * nc));
            buf0[i]          = (g == INVALID) ? T(0) : src[g];
          }
        __syncthreads();
        contract<T, false>(mats.P[child[0]], nf, nc, buf0, buf1, nc, nc, nc, 0);
        __syncthreads();
        contract<T, false>(mats.P[child[1]], nf, nc, buf1, buf0, nf, nc, nc, 1);
        __syncthreads();
        contract<T, false>(mats.P[child[2]], nf, nc, buf0, buf1, nf, nf, nc, 2);
        __syncthreads();
        for (int i = threadIdx.x; i < nf * nf * nf; i += blockDim.x)
          {
            const int      x = i % nf, y = (i / nf) % nf, z = i / (nf * nf);
            const uint32_t g = a.plain_f ? a.plain_f[cell * nf * nf * nf + i] : decode(ci_f, a.kf, x, y, z);
            const int      e = ((x == 0) ? 0 : ((x == a.kf) ? 2 : 1)) + 3 * ((y == 0) ? 0 : ((y == a.kf) ? 2 : 1)) +
                          9 * ((z == 0) ? 0 : ((z == a.kf) ? 2 : 1));
            if (g != INVALID)
              atomicAdd(dst + g, wv[e] * buf1[i]);
          }
      }
    else
      {
        for (int i = threadIdx.x; i < nf * nf * nf; i += blockDim.x)
          {
            const int      x = i % nf, y = (i / nf) % nf, z = i / (nf * nf);
            const uint32_t g = a.plain_f ? a.plain_f[cell * nf * nf * nf + i] : decode(ci_f, a.kf, x, y, z);
            const int      e = ((x == 0) ? 0 : ((x == a.kf) ? 2 : 1)) + 3 * ((y == 0) ? 0 : ((y == a.kf) ? 2 : 1)) +
                          9 * ((z == 0) ? 0 : ((z == a.kf) ? 2 : 1));
            buf0[i] = (g == INVALID) ? T(0) : wv[e] * src[g];
          }
        __syncthreads();
        contract<T, true>(mats.P[child[0]], nc, nf, buf0, buf1, nf, nf, nf, 0);
        __syncthreads();
        contract<T, true>(mats.P[child[1]], nc, nf, buf1, buf0, nc, nf, nf, 1);
        __syncthreads();
        contract<T, true>(mats.P[child[2]], nc, nf, buf0, buf1, nc, nc, nf, 2);
        __syncthreads();
        for (int i = threadIdx.x; i < nc * nc * nc; i += blockDim.x)
          {
            const uint32_t g = a.plain_c ? a.plain_c[(long long)par * nc * nc * nc + i] : decode(ci_c, a.kc, i % nc, (i / nc) % nc, i / (nc * nc));
            if (g != INVALID)
              atomicAdd(dst + g, buf1[i]);
          }
      }
  }

  template <typename TO, typename TI>
  __global__ void
  convert_kernel(TO *out, const TI *in, const long long n)
  {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
      out[i] = (TO)in[i];
  }
} // namespace

struct dasm_transfer
{
  dasm_op *    fine = nullptr, *coarse = nullptr;
  cudaStream_t stream  = nullptr;
  int          ntype   = DASM_F64;
  bool         geometric = false;
  uint32_t *   d_parent  = nullptr;
  uint8_t *    d_valence = nullptr;
  double       P[2][MAXN * MAXN];
  TransferArgs args;
};

// two-level transfer between operators on unstructured meshes (the ball): the same cell-wise embedding through the oriented
// addresses of both levels; weights = 1 / number of cells sharing the fine entity
extern "C" int
dasm_transfer_create_unstructured(dasm_op *fine, dasm_op *coarse, const uint32_t *parent_in, dasm_transfer **out)
{
  MG_API_BEGIN
  if (!dasm_op_is_unstructured(fine) || !dasm_op_is_unstructured(coarse))
    throw std::runtime_error("transfer: both operators must live on unstructured meshes");
  if (dasm_op_ctx(fine) != dasm_op_ctx(coarse))
    throw std::runtime_error("transfer: both operators must live on one context");
  if (dasm_op_number_type(fine) != dasm_op_number_type(coarse))
    throw std::runtime_error("transfer: both operators must have one number type");
  const long long nf = dasm_op_n_cells(fine), nc = dasm_op_n_cells(coarse);
  const int       kf = dasm_op_degree(fine), kc = dasm_op_degree(coarse);
  auto            t  = new dasm_transfer;
  t->fine      = fine;
  t->coarse    = coarse;
  t->stream    = (cudaStream_t)dasm_ctx_stream(dasm_op_ctx(fine));
  t->ntype     = dasm_op_number_type(fine);
  t->geometric = parent_in != nullptr;
  if (t->geometric ? (nf != 8 * nc || kf != kc) : (nf != nc || kc > kf))
    {
      delete t;
      throw std::runtime_error("transfer: the levels must be related by global 2:1 coarsening (same degree) or share the mesh (lower degree)");
    }
  const std::vector<double> xc = dasm::gauss_lobatto_points(kc + 1), xf0 = dasm::gauss_lobatto_points(kf + 1);
  for (int ch = 0; ch < 2; ++ch)
    {
      std::vector<double> xf(xf0), V, D;
      if (t->geometric)
        for (auto &x : xf)
          x = 0.5 * (x + ch);
      dasm::lagrange(xc, xf, V, D);
      for (int i = 0; i < (kf + 1) * (kc + 1); ++i)
        t->P[ch][i] = std::fabs(V[i]) < 1e-15 ? 0. : V[i];
    }
  std::vector<uint32_t> parent(nf);
  for (long long c = 0; c < nf; ++c)
    {
      parent[c] = t->geometric ? parent_in[c] : (uint32_t)c;
      if ((long long)(parent[c] & 0x0FFFFFFFu) >= nc)
        {
          delete t;
          throw std::runtime_error("transfer: parent cell out of range");
        }
    }
  std::vector<uint8_t> valence((size_t)nf * 27);
  MG_CALL(dasm_op_entity_valence(fine, valence.data()));
  MG_CUDA_CHECK(cudaMalloc(&t->d_parent, std::max<size_t>(1, parent.size()) * sizeof(uint32_t)));
  MG_CUDA_CHECK(cudaMalloc(&t->d_valence, std::max<size_t>(1, valence.size())));
  MG_CUDA_CHECK(cudaMemcpy(t->d_parent, parent.data(), parent.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  MG_CUDA_CHECK(cudaMemcpy(t->d_valence, valence.data(), valence.size(), cudaMemcpyHostToDevice));
  t->args.cidx_f    = dasm_op_device_indices(fine);
  t->args.cidx_c    = dasm_op_device_indices(coarse);
  t->args.plain_f   = dasm_op_device_plain_indices(fine);
  t->args.plain_c   = dasm_op_device_plain_indices(coarse);
  t->args.parent    = t->d_parent;
  t->args.valence   = t->d_valence;
  t->args.kf        = kf;
  t->args.kc        = kc;
  t->args.n_cells_f = nf;
  *out              = t;
  MG_API_END
}

extern "C" int
dasm_transfer_create(dasm_op *fine, dasm_op *coarse, dasm_transfer **out)
{
  if (dasm_op_is_unstructured(fine) || dasm_op_is_unstructured(coarse))
    return dasm_transfer_create_unstructured(fine, coarse, nullptr, out); // same mesh, polynomial transfer
  MG_API_BEGIN
  if (dasm_op_ctx(fine) != dasm_op_ctx(coarse))
    throw std::runtime_error("transfer: both operators must live on one context");
  if (dasm_op_number_type(fine) != dasm_op_number_type(coarse))
    throw std::runtime_error("transfer: both operators must have one number type");
  auto t    = new dasm_transfer;
  t->fine   = fine;
  t->coarse = coarse;
  t->stream = (cudaStream_t)dasm_ctx_stream(dasm_op_ctx(fine));
  t->ntype  = dasm_op_number_type(fine);
  int ncf[3], ncc[3], perf[3], perc[3];
  MG_CALL(dasm_mesh_global_size(dasm_op_mesh(fine), ncf, perf));
  MG_CALL(dasm_mesh_global_size(dasm_op_mesh(coarse), ncc, perc));
  const int  kf = dasm_op_degree(fine), kc = dasm_op_degree(coarse);
  const bool same = ncf[0] == ncc[0] && ncf[1] == ncc[1] && ncf[2] == ncc[2];
  const bool half = ncf[0] == 2 * ncc[0] && ncf[1] == 2 * ncc[1] && ncf[2] == 2 * ncc[2];
  if (!(same && kc <= kf) && !(half && kc == kf))
    throw std::runtime_error("transfer: the levels must be related by global 2:1 coarsening (same degree) or share the mesh (lower degree)");
  for (int d = 0; d < 3; ++d)
    if (perf[d] != perc[d])
      throw std::runtime_error("transfer: periodicity of the two meshes differs");
  t->geometric = half && !same;
  // 1-D embedding matrices: coarse basis at the fine support points of child 0 / 1 (or of the same cell)
  const std::vector<double> xc = dasm::gauss_lobatto_points(kc + 1), xf0 = dasm::gauss_lobatto_points(kf + 1);
  for (int ch = 0; ch < 2; ++ch)
    {
      std::vector<double> xf(xf0), V, D;
      if (t->geometric)
        for (auto &x : xf)
          x = 0.5 * (x + ch);
      dasm::lagrange(xc, xf, V, D);
      for (int i = 0; i < (kf + 1) * (kc + 1); ++i)
        t->P[ch][i] = std::fabs(V[i]) < 1e-15 ? 0. : V[i];
    }
  // parent cell and child position of every local fine cell; entity valences from the topology
  const long long  nf = dasm_mesh_n_cells(dasm_op_mesh(fine)), nc = dasm_mesh_n_cells(dasm_op_mesh(coarse));
  std::vector<int> cf((size_t)nf * 3), cc((size_t)nc * 3);
  MG_CALL(dasm_mesh_cell_coordinates(dasm_op_mesh(fine), cf.data()));
  MG_CALL(dasm_mesh_cell_coordinates(dasm_op_mesh(coarse), cc.data()));
  std::map<long long, uint32_t> coarse_of;
  for (long long c = 0; c < nc; ++c)
    coarse_of[((long long)cc[3 * c + 2] * ncc[1] + cc[3 * c + 1]) * ncc[0] + cc[3 * c]] = (uint32_t)c;
  std::vector<uint32_t> parent(nf);
  std::vector<uint8_t>  valence((size_t)nf * 27);
  for (long long c = 0; c < nf; ++c)
    {
      int      p[3];
      uint32_t code = 0;
      for (int d = 0; d < 3; ++d)
        {
          p[d] = t->geometric ? cf[3 * c + d] / 2 : cf[3 * c + d];
          if (t->geometric && (cf[3 * c + d] & 1))
            code |= 1u << d;
        }
      const auto it = coarse_of.find(((long long)p[2] * ncc[1] + p[1]) * ncc[0] + p[0]);
      if (it == coarse_of.end())
        throw std::runtime_error("transfer: the parent of a local fine cell is not a local coarse cell (partitions of the levels must nest)");
      parent[c] = it->second | (code << 28);
      for (int e = 0; e < 27; ++e)
        {
          int       v      = 1;
          const int ed[3] = {e % 3, (e / 3) % 3, e / 9};
          for (int d = 0; d < 3; ++d)
            if (ed[d] != 1)
              {
                const bool at_boundary = (ed[d] == 0) ? (cf[3 * c + d] == 0) : (cf[3 * c + d] == ncf[d] - 1);
                if (perf[d] || !at_boundary)
                  v *= 2;
              }
          valence[(size_t)c * 27 + e] = (uint8_t)v;
        }
    }
  MG_CUDA_CHECK(cudaMalloc(&t->d_parent, std::max<size_t>(1, parent.size()) * sizeof(uint32_t)));
  MG_CUDA_CHECK(cudaMalloc(&t->d_valence, std::max<size_t>(1, valence.size())));
  MG_CUDA_CHECK(cudaMemcpy(t->d_parent, parent.data(), parent.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  MG_CUDA_CHECK(cudaMemcpy(t->d_valence, valence.data(), valence.size(), cudaMemcpyHostToDevice));
  t->args.cidx_f    = dasm_op_device_indices(fine);
  t->args.cidx_c    = dasm_op_device_indices(coarse);
  t->args.parent    = t->d_parent;
  t->args.valence   = t->d_valence;
  t->args.kf        = kf;
  t->args.kc        = kc;
  t->args.n_cells_f = nf;
  *out              = t;
  MG_API_END
}

extern "C" int
dasm_transfer_destroy(dasm_transfer *t)
{
  if (t)
    {
      cudaFree(t->d_parent);
      cudaFree(t->d_valence);
      delete t;
    }
  return 0;
}

// the ghost part of a destination starts from zero: the contributions of the local cells to DoFs of other ranks are collected
// there and added to their owners by the compress
static void
zero_ghosts(dasm_op *op, void *vec, cudaStream_t s)
{
  const long long no = dasm_op_n_dofs(op), nv = dasm_op_vec_size(op);
  const size_t    es = dasm_op_number_type(op) == DASM_F64 ? 8 : 4;
  if (nv > no)
    MG_CUDA_CHECK(cudaMemsetAsync((char *)vec + (size_t)no * es, 0, (size_t)(nv - no) * es, s));
}

template <typename T, bool RESTRICT>
static void
transfer_launch(dasm_transfer *t, T *dst, const T *src)
{
  TransferMats<T> m;
  for (int ch = 0; ch < 2; ++ch)
    for (int i = 0; i < MAXN * MAXN; ++i)
      m.P[ch][i] = (T)t->P[ch][i];
  if (t->args.n_cells_f > 0)
    transfer_kernel<T, RESTRICT><<<(unsigned)t->args.n_cells_f, 128, 0, t->stream>>>(dst, src, t->args, m);
  MG_CUDA_CHECK(cudaGetLastError());
}

extern "C" int
dasm_transfer_prolongate_and_add(dasm_transfer *t, void *dst_fine, const void *src_coarse)
{
  MG_API_BEGIN
  // coarse ghost values are read; every owned fine DoF is touched by local cells only through local contributions plus the
  // contributions of the neighbour ranks' cells: add them up (compress), like the weighted distribute of MGTwoLevelTransfer
  MG_CALL(dasm_op_update_ghost_values(t->coarse, const_cast<void *>(src_coarse)));
  zero_ghosts(t->fine, dst_fine, t->stream);
  if (t->ntype == DASM_F64)
    transfer_launch<double, false>(t, (double *)dst_fine, (const double *)src_coarse);
  else
    transfer_launch<float, false>(t, (float *)dst_fine, (const float *)src_coarse);
  MG_CALL(dasm_op_compress_add(t->fine, dst_fine));
  MG_API_END
}

extern "C" int
dasm_transfer_restrict_and_add(dasm_transfer *t, void *dst_coarse, const void *src_fine)
{
  MG_API_BEGIN
  MG_CALL(dasm_op_update_ghost_values(t->fine, const_cast<void *>(src_fine)));
  zero_ghosts(t->coarse, dst_coarse, t->stream);
  if (t->ntype == DASM_F64)
    transfer_launch<double, true>(t, (double *)dst_coarse, (const double *)src_fine);
  else
    transfer_launch<float, true>(t, (float *)dst_coarse, (const float *)src_fine);
  MG_CALL(dasm_op_compress_add(t->coarse, dst_coarse));
  MG_API_END
}

// ---- V-cycle ---------------------------------------------------------------------------------------------------------------------
struct dasm_mg
{
  int                          L = 0; // finest level
  int                          ntype = DASM_F64;
  bool                         one_sided = false;
  cudaStream_t                 stream = nullptr;
  std::vector<dasm_op *>       ops;
  std::vector<dasm_cheb *>     smoothers;
  std::vector<dasm_transfer *> transfers; // [l]: between l and l - 1
  std::vector<bool>            borrowed;  // transfer owned by the caller
  std::vector<void *>          defect, solution, t;
  std::vector<long long>       vec_size, n_owned;
  size_t                       esize = 8;
};

extern "C" int
dasm_mg_create(int n_levels, dasm_op **level_ops, dasm_cheb **smoothers, int one_sided_v_cycle, dasm_mg **out)
{
  return dasm_mg_create_with_transfers(n_levels, level_ops, smoothers, nullptr, one_sided_v_cycle, out);
}

// same with caller-built transfers (transfers[l] between the levels l and l - 1; NULL entries are created here): the geometric levels
// of an unstructured mesh need the parent map of dasm_transfer_create_unstructured.  The caller keeps the ownership of its transfers.
extern "C" int
dasm_mg_create_with_transfers(int n_levels, dasm_op **level_ops, dasm_cheb **smoothers, dasm_transfer **transfers, int one_sided_v_cycle,
                              dasm_mg **out)
{
  MG_API_BEGIN
  if (n_levels < 1)
    throw std::runtime_error("multigrid: at least one level is needed");
  auto mg       = new dasm_mg;
  mg->L         = n_levels - 1;
  mg->ntype     = dasm_op_number_type(level_ops[0]);
  mg->esize     = mg->ntype == DASM_F64 ? 8 : 4;
  mg->one_sided = one_sided_v_cycle != 0;
  mg->stream    = (cudaStream_t)dasm_ctx_stream(dasm_op_ctx(level_ops[0]));
  mg->transfers.assign(n_levels, nullptr);
  mg->borrowed.assign(n_levels, false);
  for (int l = 0; l < n_levels; ++l)
    {
      if (dasm_op_number_type(level_ops[l]) != mg->ntype)
        throw std::runtime_error("multigrid: all level operators must have one number type");
      if (smoothers[l] == nullptr || dasm_cheb_op(smoothers[l]) != level_ops[l])
        throw std::runtime_error("multigrid: smoother of level " + std::to_string(l) + " does not belong to the level operator");
      mg->ops.push_back(level_ops[l]);
      mg->smoothers.push_back(smoothers[l]);
      mg->vec_size.push_back(dasm_op_vec_size(level_ops[l]));
      mg->n_owned.push_back(dasm_op_n_dofs(level_ops[l]));
      void *v[3];
      for (auto &p : v)
        {
          MG_CALL(dasm_op_vec_alloc(level_ops[l], &p));
          MG_CUDA_CHECK(cudaMemsetAsync(p, 0, (size_t)mg->vec_size[l] * mg->esize, mg->stream));
        }
      mg->defect.push_back(v[0]);
      mg->solution.push_back(v[1]);
      mg->t.push_back(v[2]);
      if (l > 0 && transfers != nullptr && transfers[l] != nullptr)
        {
          if (transfers[l]->fine != level_ops[l] || transfers[l]->coarse != level_ops[l - 1])
            throw std::runtime_error("multigrid: transfer of level " + std::to_string(l) + " does not connect the level operators");
          mg->transfers[l] = transfers[l];
          mg->borrowed[l]  = true;
        }
      else if (l > 0)
        MG_CALL(dasm_transfer_create(level_ops[l], level_ops[l - 1], &mg->transfers[l]));
    }
  *out = mg;
  MG_API_END
}

extern "C" int
dasm_mg_destroy(dasm_mg *mg)
{
  if (mg)
    {
      cudaDeviceSynchronize();
      for (size_t l = 0; l < mg->ops.size(); ++l)
        {
          // (plain frees: the level operators may already have been destroyed by the caller)
          cudaFree(mg->defect[l]);
          cudaFree(mg->solution[l]);
          cudaFree(mg->t[l]);
          if (!mg->borrowed[l])
            dasm_transfer_destroy(mg->transfers[l]);
        }
      delete mg;
    }
  return 0;
}

// Multigrid::level_v_step
static void
mg_level(dasm_mg *mg, const int l)
{
  if (l == 0)
    {
      MG_CALL(dasm_cheb_vmult(mg->smoothers[0], mg->solution[0], mg->defect[0]));
      return;
    }
  // pre-smoothing from a zero start (MGSmootherRelaxation::apply: vmult)
  MG_CALL(dasm_cheb_vmult(mg->smoothers[l], mg->solution[l], mg->defect[l]));
  // t = defect - A solution
  dasm_hook pre  = {DASM_HOOK_ZERO_DST, 0., 0., nullptr, nullptr};
  dasm_hook post = {DASM_HOOK_RESIDUAL, 0., 0., mg->defect[l], nullptr};
  MG_CALL(dasm_op_vmult_hooks(mg->ops[l], mg->t[l], mg->solution[l], &pre, &post));
  MG_CUDA_CHECK(cudaMemsetAsync(mg->defect[l - 1], 0, (size_t)mg->vec_size[l - 1] * mg->esize, mg->stream));
  MG_CALL(dasm_transfer_restrict_and_add(mg->transfers[l], mg->defect[l - 1], mg->t[l]));
  mg_level(mg, l - 1);
  MG_CALL(dasm_transfer_prolongate_and_add(mg->transfers[l], mg->solution[l], mg->solution[l - 1]));
  if (!mg->one_sided)
    MG_CALL(dasm_cheb_step(mg->smoothers[l], mg->solution[l], mg->defect[l]));
}

extern "C" int
dasm_mg_vmult(dasm_mg *mg, void *dst, const void *src)
{
  return dasm_mg_vmult_outer(mg, dst, src, mg->ntype);
}

extern "C" int
dasm_mg_vmult_outer(dasm_mg *mg, void *dst, const void *src, int outer_number_type)
{
  MG_API_BEGIN
  const int       L = mg->L;
  const long long n = mg->n_owned[L];
  const unsigned  nb = (unsigned)((n + 255) / 256);
  if (outer_number_type == mg->ntype)
    MG_CUDA_CHECK(cudaMemcpyAsync(mg->defect[L], src, (size_t)n * mg->esize, cudaMemcpyDeviceToDevice, mg->stream));
  else if (outer_number_type == DASM_F64)
    convert_kernel<float, double><<<nb, 256, 0, mg->stream>>>((float *)mg->defect[L], (const double *)src, n);
  else
    convert_kernel<double, float><<<nb, 256, 0, mg->stream>>>((double *)mg->defect[L], (const float *)src, n);
  mg_level(mg, L);
  if (outer_number_type == mg->ntype)
    MG_CUDA_CHECK(cudaMemcpyAsync(dst, mg->solution[L], (size_t)n * mg->esize, cudaMemcpyDeviceToDevice, mg->stream));
  else if (outer_number_type == DASM_F64)
    convert_kernel<double, float><<<nb, 256, 0, mg->stream>>>((double *)dst, (const float *)mg->solution[L], n);
  else
    convert_kernel<float, double><<<nb, 256, 0, mg->stream>>>((float *)dst, (const double *)mg->solution[L], n);
  MG_CUDA_CHECK(cudaGetLastError());
  MG_API_END
}
