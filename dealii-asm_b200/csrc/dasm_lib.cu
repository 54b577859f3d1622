// libdasm: C ABI (include/dasm.h) + host orchestration of the CUDA kernels.
//
// Host orchestration mirrors the reference's cell-loop schedule MFRunner::loop
// (include/matrix_free_internal.h:309-359): pre-operation, ghost update, cell kernel, compress(add),
// post-operation; constrained DoFs get dst = src when a post operation is given (226-255).
#include <cuda_runtime.h>
#include <nccl.h>

#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/dasm.h"
#include "kernels.cuh"
#include "kernels_brick.cuh"
#include "kernels_fast.cuh"
#include "mesh.h"
#include "unstructured.h"
#include "tma_launch.h"

using namespace dasm;

// ------------------------------------------------------------------------------------------------
// error handling
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

#define CUDA_CHECK(expr)                                                                                      \
  do                                                                                                          \
    {                                                                                                         \
      cudaError_t _e = (expr);                                                                                \
      if (_e != cudaSuccess)                                                                                  \
        throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + __FILE__ + ":" + \
                                 std::to_string(__LINE__));                                                   \
    }                                                                                                         \
  while (0)

#define NCCL_CHECK(expr)                                                                                      \
  do                                                                                                          \
    {                                                                                                         \
      ncclResult_t _e = (expr);                                                                               \
      if (_e != ncclSuccess)                                                                                  \
        throw std::runtime_error(std::string("NCCL error: ") + ncclGetErrorString(_e) + " at " + __FILE__ + ":" + \
                                 std::to_string(__LINE__));                                                   \
    }                                                                                                         \
  while (0)

#define DASM_API_BEGIN try {
#define DASM_API_END                      \
  return 0;                               \
  }                                       \
  catch (const std::exception &e)         \
  {                                       \
    g_last_error = e.what();              \
    return 1;                             \
  }                                       \
  catch (...)                             \
  {                                       \
    g_last_error = "unknown exception";   \
    return 1;                             \
  }

#define DASM_REQUIRE(cond, msg)             \
  do                                        \
    {                                       \
      if (!(cond))                          \
        throw std::runtime_error(msg);      \
    }                                       \
  while (0)

// ------------------------------------------------------------------------------------------------
// objects
// ------------------------------------------------------------------------------------------------
struct dasm_ctx
{
  int          device   = 0;
  cudaStream_t stream   = nullptr;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t  ev_a = nullptr, ev_b = nullptr;
  long long    launches = 0;
  ncclComm_t   comm     = nullptr;
  int          n_ranks = 1, rank = 0;
  double *     d_partial = nullptr; // [1024]
  double *     d_scalar  = nullptr; // [8]
  double *     h_scalar  = nullptr; // pinned [8]
  // optional per-kernel-class timing with CUDA events on `stream`
  bool                                  timing = false;
  std::vector<std::array<cudaEvent_t, 2>> ev_pool;
  std::vector<int>                      ev_class;
  size_t                                ev_used = 0;
};

// kernel classes for dasm_ctx_kernel_time
enum
{
  KC_LAPLACE = 0,
  KC_FDM     = 1,
  KC_VECTOR  = 2,
  KC_EXCHANGE = 3,
  KC_COUNT   = 4
};

struct KernelTimer
{
  dasm_ctx *ctx;
  size_t    slot = 0;
  bool      on;
  KernelTimer(dasm_ctx *c, int klass)
    : ctx(c)
    , on(c->timing)
  {
    if (!on)
      return;
    if (ctx->ev_used == ctx->ev_pool.size())
      {
        std::array<cudaEvent_t, 2> e;
        cudaEventCreate(&e[0]);
        cudaEventCreate(&e[1]);
        ctx->ev_pool.push_back(e);
        ctx->ev_class.push_back(klass);
      }
    slot                = ctx->ev_used++;
    ctx->ev_class[slot] = klass;
    cudaEventRecord(ctx->ev_pool[slot][0], ctx->stream);
  }
  ~KernelTimer()
  {
    if (on)
      cudaEventRecord(ctx->ev_pool[slot][1], ctx->stream);
  }
};

struct dasm_mesh
{
  dasm_ctx *            ctx;
  std::unique_ptr<Mesh> mesh;
};

template <typename T>
static T *
dev_alloc(size_t n)
{
  T *p = nullptr;
  CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
  return p;
}

template <typename T>
static T *
dev_upload(const std::vector<T> &v, cudaStream_t)
{
  T *p = dev_alloc<T>(v.size());
  if (!v.empty())
    CUDA_CHECK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return p;
}

// ghost exchange (replaces VectorDataExchange, matrix_free_internal.h:3-109)
//   default: device-initiated transfers over NVLink: the receive buffers are mapped into the peers with CUDA IPC, the pack kernel
//            writes into the peer's buffer and publishes a sequence number, the unpack kernel waits for it (kernels.cuh p2p_*);
//   DASM_P2P=0 or when the mapping fails: pack kernel -> grouped ncclSend / ncclRecv -> unpack kernel.
struct Exchange
{
  dasm_ctx *            ctx = nullptr;
  std::vector<int>      peers;
  std::vector<size_t>   send_off, recv_off; // per peer offsets (+ end) in elements
  uint32_t *            d_send_map = nullptr, *d_recv_map = nullptr;
  void *                d_send_buf = nullptr, *d_recv_buf = nullptr;
  size_t                n_send = 0, n_recv = 0;
  size_t                elem_size = 8;
  // peer-to-peer state
  int                   p2p_state = 0; // 0 not tried, 1 ready, -1 unavailable (NCCL path)
  char *                d_ipc     = nullptr; // [2 halves of cap elements | flags[n_ranks] | acks[n_ranks]]
  size_t                cap       = 0;
  unsigned long long    seq       = 0;
  unsigned int *        d_counters = nullptr;
  std::vector<char *>   peer_base;               // per peer: mapped IPC block
  std::vector<size_t>   peer_cap;                // per peer: elements per half
  std::vector<long long> peer_off_update, peer_off_compress; // per peer: offset of this rank's segment in the peer's buffer

  struct Info
  {
    cudaIpcMemHandle_t handle;
    long long          cap;
    long long          off_update[64];   // by world rank: offset of the segment received from that rank in a ghost update (-1: none)
    long long          off_compress[64]; // ... in a compress
  };

  void
  init(dasm_ctx *c, const std::vector<ExchangeList> &lists, size_t esize)
  {
    ctx       = c;
    elem_size = esize;
    std::vector<uint32_t> smap, rmap;
    send_off.push_back(0);
    recv_off.push_back(0);
    for (const auto &l : lists)
      {
        peers.push_back(l.peer);
        for (size_t i = 0; i < l.send_start.size(); ++i)
          for (uint32_t t = 0; t < l.send_len[i]; ++t)
            smap.push_back(l.send_start[i] + t);
        for (size_t i = 0; i < l.recv_start.size(); ++i)
          for (uint32_t t = 0; t < l.recv_len[i]; ++t)
            rmap.push_back(l.recv_start[i] + t);
        send_off.push_back(smap.size());
        recv_off.push_back(rmap.size());
      }
    n_send     = smap.size();
    n_recv     = rmap.size();
    d_send_map = dev_upload(smap, nullptr);
    d_recv_map = dev_upload(rmap, nullptr);
    CUDA_CHECK(cudaMalloc(&d_send_buf, std::max<size_t>(std::max(n_send, n_recv), 1) * elem_size));
    CUDA_CHECK(cudaMalloc(&d_recv_buf, std::max<size_t>(std::max(n_send, n_recv), 1) * elem_size));
  }

  bool
  active() const
  {
    return !peers.empty();
  }

  // collective over all ranks of the communicator: map the receive buffers of the peers
  void
  p2p_setup(cudaStream_t s)
  {
    p2p_state = -1;
    if (const char *e = getenv("DASM_P2P"))
      if (e[0] == '0')
        return;
    const int R = ctx->n_ranks;
    if (R > 64 || (int)peers.size() > P2P_MAX_PEERS)
      return;
    cap                    = std::max<size_t>(std::max(n_send, n_recv), 1);
    const size_t data_bytes = (2 * cap * elem_size + 255) / 256 * 256;
    if (cudaMalloc(&d_ipc, data_bytes + 2 * 64 * sizeof(unsigned long long)) != cudaSuccess)
      {
        cudaGetLastError();
        return;
      }
    CUDA_CHECK(cudaMemsetAsync(d_ipc, 0, data_bytes + 2 * 64 * sizeof(unsigned long long), s));
    CUDA_CHECK(cudaMalloc(&d_counters, 2 * sizeof(unsigned int)));
    CUDA_CHECK(cudaMemsetAsync(d_counters, 0, 2 * sizeof(unsigned int), s));
    Info mine;
    memset(&mine, 0, sizeof(mine));
    bool ok  = cudaIpcGetMemHandle(&mine.handle, d_ipc) == cudaSuccess;
    mine.cap = ok ? (long long)cap : -1;
    for (int r = 0; r < 64; ++r)
      mine.off_update[r] = mine.off_compress[r] = -1;
    for (size_t q = 0; q < peers.size(); ++q)
      {
        mine.off_update[peers[q]]   = (long long)recv_off[q];
        mine.off_compress[peers[q]] = (long long)send_off[q];
      }
    // all-gather of the descriptors through NCCL (device staging)
    Info *d_all = nullptr;
    CUDA_CHECK(cudaMalloc(&d_all, sizeof(Info) * (size_t)R));
    CUDA_CHECK(cudaMemcpyAsync(d_all + ctx->rank, &mine, sizeof(Info), cudaMemcpyHostToDevice, s));
    NCCL_CHECK(ncclAllGather(d_all + ctx->rank, d_all, sizeof(Info), ncclChar, ctx->comm, s));
    std::vector<Info> all(R);
    CUDA_CHECK(cudaMemcpyAsync(all.data(), d_all, sizeof(Info) * (size_t)R, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    cudaFree(d_all);
    for (int r = 0; r < R; ++r)
      if (all[r].cap < 0)
        ok = false; // some rank could not export its buffer: everybody uses NCCL
    if (ok)
      for (size_t q = 0; q < peers.size() && ok; ++q)
        {
          void *ptr = nullptr;
          if (cudaIpcOpenMemHandle(&ptr, all[peers[q]].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
            {
              cudaGetLastError();
              ok = false;
              break;
            }
          peer_base.push_back((char *)ptr);
          peer_cap.push_back((size_t)all[peers[q]].cap);
          peer_off_update.push_back(all[peers[q]].off_update[ctx->rank]);
          peer_off_compress.push_back(all[peers[q]].off_compress[ctx->rank]);
        }
    // agree on the outcome (a rank that failed to map a peer makes everybody fall back)
    int *d_ok = nullptr, h_ok = ok ? 1 : 0;
    CUDA_CHECK(cudaMalloc(&d_ok, sizeof(int)));
    CUDA_CHECK(cudaMemcpyAsync(d_ok, &h_ok, sizeof(int), cudaMemcpyHostToDevice, s));
    NCCL_CHECK(ncclAllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, ctx->comm, s));
    CUDA_CHECK(cudaMemcpyAsync(&h_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    cudaFree(d_ok);
    if (h_ok)
      p2p_state = 1;
    else if (getenv("DASM_VERBOSE"))
      fprintf(stderr, "[dasm] rank %d: peer-to-peer halo exchange unavailable, using NCCL send/recv\n", ctx->rank);
  }

  template <typename T>
  void
  run_p2p(T *vec, bool compress, cudaStream_t s)
  {
    ++seq;
    const size_t    par   = (size_t)(seq & 1ull);
    const size_t    n_out = compress ? n_recv : n_send;
    const size_t    n_in  = compress ? n_send : n_recv;
    const uint32_t *omap  = compress ? d_recv_map : d_send_map;
    const uint32_t *imap  = compress ? d_send_map : d_recv_map;
    const auto &    ooff  = compress ? recv_off : send_off;
    const size_t    data_bytes = (2 * cap * elem_size + 255) / 256 * 256;
    P2PPeers        P;
    P.n = (int)peers.size();
    for (int q = 0; q < P.n; ++q)
      {
        const size_t pdata = (2 * peer_cap[q] * elem_size + 255) / 256 * 256;
        P.buf[q]           = peer_base[q] + par * peer_cap[q] * elem_size;
        P.flag[q]          = reinterpret_cast<unsigned long long *>(peer_base[q] + pdata) + ctx->rank;
        P.ack[q]           = reinterpret_cast<unsigned long long *>(peer_base[q] + pdata) + 64 + ctx->rank;
        P.dst_off[q]       = compress ? peer_off_compress[q] : peer_off_update[q];
        P.seg_begin[q]     = (long long)ooff[q];
        P.my_flag[q]       = reinterpret_cast<unsigned long long *>(d_ipc + data_bytes) + peers[q];
        P.my_ack[q]        = reinterpret_cast<unsigned long long *>(d_ipc + data_bytes) + 64 + peers[q];
      }
    P.seg_begin[P.n] = (long long)ooff[P.n];
    const unsigned g_out = (unsigned)std::min<size_t>(1184, std::max<size_t>(1, (n_out + 255) / 256));
    p2p_push_kernel<T><<<g_out, 256, 0, s>>>(vec, omap, P, seq, d_counters);
    const unsigned g_in = (unsigned)std::min<size_t>(1184, std::max<size_t>(1, (n_in + 255) / 256));
    const T *      buf  = reinterpret_cast<const T *>(d_ipc + par * cap * elem_size);
    if (compress)
      p2p_pull_kernel<T, true><<<g_in, 256, 0, s>>>(vec, buf, imap, (long long)n_in, P, seq, d_counters + 1);
    else
      p2p_pull_kernel<T, false><<<g_in, 256, 0, s>>>(vec, buf, imap, (long long)n_in, P, seq, d_counters + 1);
    ctx->launches += 2;
  }

  // owner -> ghost (update_ghost_values) or ghost -> owner with add (compress)
  template <typename T>
  void
  run(T *vec, bool compress, cudaStream_t on_stream = nullptr)
  {
    if (!active())
      return;
    DASM_REQUIRE(ctx->comm != nullptr, "multi-rank mesh needs dasm_ctx_comm_init before any operator application");
    cudaStream_t    s     = on_stream ? on_stream : ctx->stream;
    if (p2p_state == 0)
      p2p_setup(s);
    if (p2p_state == 1)
      {
        run_p2p<T>(vec, compress, s);
        return;
      }
    const size_t    n_out = compress ? n_recv : n_send;
    const size_t    n_in  = compress ? n_send : n_recv;
    const uint32_t *omap  = compress ? d_recv_map : d_send_map;
    const uint32_t *imap  = compress ? d_send_map : d_recv_map;
    const auto &    ooff  = compress ? recv_off : send_off;
    const auto &    ioff  = compress ? send_off : recv_off;
    if (n_out)
      {
        pack_kernel<T><<<(unsigned)((n_out + 255) / 256), 256, 0, s>>>((T *)d_send_buf, vec, omap, (long long)n_out);
        ctx->launches++;
      }
    NCCL_CHECK(ncclGroupStart());
    for (size_t pidx = 0; pidx < peers.size(); ++pidx)
      {
        const size_t so = ooff[pidx], sn = ooff[pidx + 1] - so;
        const size_t ro = ioff[pidx], rn = ioff[pidx + 1] - ro;
        if (sn)
          NCCL_CHECK(ncclSend((const char *)d_send_buf + so * sizeof(T), sn * sizeof(T), ncclChar, peers[pidx], ctx->comm, s));
        if (rn)
          NCCL_CHECK(ncclRecv((char *)d_recv_buf + ro * sizeof(T), rn * sizeof(T), ncclChar, peers[pidx], ctx->comm, s));
      }
    NCCL_CHECK(ncclGroupEnd());
    if (n_in)
      {
        if (compress)
          unpack_kernel<T, true><<<(unsigned)((n_in + 255) / 256), 256, 0, s>>>(vec, (const T *)d_recv_buf, imap, (long long)n_in);
        else
          unpack_kernel<T, false><<<(unsigned)((n_in + 255) / 256), 256, 0, s>>>(vec, (const T *)d_recv_buf, imap, (long long)n_in);
        ctx->launches++;
      }
  }

  void
  destroy()
  {
    cudaFree(d_send_map);
    cudaFree(d_recv_map);
    cudaFree(d_send_buf);
    cudaFree(d_recv_buf);
    for (char *pb : peer_base)
      cudaIpcCloseMemHandle(pb);
    cudaFree(d_ipc);
    cudaFree(d_counters);
  }
};

struct dasm_op
{
  dasm_ctx *        ctx;
  dasm_mesh *       mesh;
  int               k;
  int               ntype;
  bool              compress_indices;
  std::string       mapping_type;
  Basis1D           basis;
  Mesh::Numbering   nb;
  long long         n_cells;
  long long         n_owned, n_ghost, n_vec;
  long long         n_global_dofs;
  uint32_t *        d_cidx        = nullptr;
  uint32_t *        d_plain       = nullptr; // compress_indices = false: (k+1)^3 indices per cell (operator.h:1343-1350, plain branch)
  // unstructured hex mesh (csrc/unstructured.h; mesh == nullptr): 27 start indices + orientation word per cell, expanded to d_plain
  std::unique_ptr<UMesh> umesh;
  std::vector<uint32_t>  h_plain;
  // DASM_DETERMINISTIC=1: bitwise reproducible results.  The cells are coloured so that no two cells of a colour share a DoF; the
  // generic kernels run colour by colour, so every vector entry receives its contributions in a fixed order
  bool                   deterministic = false;
  std::vector<long long> color_ptr;               // cells of colour c: d_color_cells[color_ptr[c] .. color_ptr[c + 1])
  uint32_t *             d_color_cells = nullptr;
  uint32_t *        d_constrained = nullptr;
  long long         n_constrained = 0;
  int               geom_mode     = 0; // 0 cartesian, 1 merged, 2 quadratic / linear coefficients (brick kernel)
  void *            d_geom        = nullptr; // merged coefficients [cell][6][n^3]
  void *            d_qcoef       = nullptr; // monomial coefficients of the cell map [cell][27][3]
  bool              linear_geometry = false;
  CartesianCoef     cart;
  Exchange          exchange;
  // enlarged ghost layout for preconditioners with overlapping patches on several ranks (matrix_free.h:154-213; the operator and
  // the preconditioner share ONE vector layout, which is what set_partitioner, operator.h:780-849, establishes in the reference)
  Mesh::HaloNumbering halo;
  Exchange          exchange_ext; // all ghosts (those of `exchange` + the DoFs of the halo cells)
  long long         n_ghost_ext = 0;
  std::vector<void *> scratch; // owned by op, freed at destroy
  std::vector<void *> krylov_pool; // work vectors of dasm_solve, kept between solves (entries are also listed in scratch)
  // tuned brick path
  bool              use_brick = false;
  bool              tma_only  = false; // degrees 5 and 6: every brick is a lex brick and runs the TMA-fed kernels (no brick kernels)
  int               brick_bz  = 4;
  BrickDesc *       d_bricks  = nullptr;
  int               n_bricks  = 0;
  void *            d_acc     = nullptr; // zero-invariant accumulator of shared-face DoFs (n_vec)
  uint32_t *        d_shared_list = nullptr; // owned DoFs on brick faces shared with other bricks
  long long         n_shared  = 0;
  bool              shared_ranges_ok = false; // every brick's own shared DoFs are one contiguous range
  BrickMaps         maps = {}; // tile maps (coalesced gather / store)
  void *            d_map_bufs[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int               n_sm      = 148;
  int               max_smem  = 48 * 1024; // opt-in dynamic shared memory per block
  // warp-specialised kernels (kernels_fast.cuh) for the regular bricks; the other bricks go through the brick kernels
  bool              fast_ok     = false;
  uint32_t *        d_fast_ids  = nullptr; // regular bricks
  uint32_t *        d_slow_ids  = nullptr; // all other bricks
  int               n_fast = 0, n_slow = 0;
  std::vector<uint32_t> h_fast_ids;
  // overlap of the halo exchange with the interior bricks: d_fast_ids = [bricks on the partition boundary | interior bricks]
  int               n_fast_boundary = 0;
  std::vector<char> h_brick_boundary;
  const void *      last_compressed = nullptr; // vector whose partition-boundary values are final on the comm stream
  double            lap_P[4][25], lap_Q[4][25]; // even-odd blocks (kernels_fast.cuh EOMat) of M, g0 K, g1 K, g2 K
  // TMA-fed kernels (kernels_tma.cuh) on the lex bricks: the fast lists above then hold lex bricks
  bool                  tma_ok = false;
  std::vector<TmaBrick> h_tma;           // per kernel brick (valid for lex bricks)
  TmaBrick *            d_tma_lap     = nullptr; // descriptors in the order of d_fast_ids
  uint32_t *            d_tma_lap_chunks = nullptr; // chunk_start of the chunks of +x neighbours (kernels_tma.cuh)
  int                   tma_lap_n_chunks = 0, tma_lap_n_chunks_boundary = 0;
  uint32_t *            d_tma_foreign = nullptr; // foreign index lists of the mode-1 bricks
  int                   tma_any_mode1 = 0, tma_any_mode1_interior = 0;
  std::unordered_map<const void *, TmaMaps> tma_cache; // tensor maps per vector

  dasm_op(int degree)
    : basis(degree)
  {}
  size_t
  esize() const
  {
    return ntype == DASM_F64 ? 8 : 4;
  }
};

struct dasm_fdm
{
  dasm_op * op;
  int       n_overlap, weight_type, weight_sequence, element_centric;
  bool      use_ext = false; // several ranks + overlapping patches: the enlarged ghost layout and its exchange
  int       m;          // 1-D patch size
  long long n_instances; // unique 1-D (S, lambda) instances
  uint32_t *d_inst   = nullptr;
  void *    d_S      = nullptr;
  void *    d_lam    = nullptr;
  void *    d_wvec   = nullptr; // global weight vector (owned+ghost)
  void *    d_cw     = nullptr; // compressed weights [cell][27] or per-entry [cell][m^3]
  uint8_t * d_cwcode = nullptr; // brick kernel: weight codes [cell][32] (code = valence, 0 = weight 0)
  double    wtab[16] = {0};     // weight value per code
  uint4 *   d_brick_tri = nullptr; // per kernel brick: instance triple of the first cell + uniform flag
  uint32_t *d_pidx   = nullptr; // explicit patch index list for n_overlap > 1
  std::vector<long long> color_ptr; // deterministic mode with explicit patch lists: colouring of the patches
  uint32_t *d_color_cells = nullptr;
  int       wmode    = 0;       // kernel weight mode
  bool      w_pre = false, w_post = false;
  // warp-specialised kernel: bricks with the most frequent instance triple and weight pattern
  bool      fast_ok    = false;
  uint32_t *d_fast_ids = nullptr, *d_slow_ids = nullptr;
  int       n_fast = 0, n_slow = 0, n_fast_boundary = 0;
  TmaBrick *d_tma_list = nullptr; // TMA-fed kernel: descriptors in the order of d_fast_ids
  uint32_t *d_tma_chunks = nullptr;
  int       tma_n_chunks = 0, tma_n_chunks_boundary = 0;
  int       tma_any_mode1 = 0, tma_any_mode1_interior = 0;
  double    fast_P[6][25], fast_Q[6][25]; // even-odd blocks of Ax Ay Az Bx By Bz
  double    fast_inv[729];
  std::vector<double>   h_S, h_lam; // double copies for inspection
  std::vector<uint32_t> h_inst;
  std::vector<double>   h_weights;
};

struct dasm_cheb
{
  dasm_op * op;
  dasm_fdm *fdm;
  int       degree, poly, ev_algo, optimize, n_ev_it;
  double    smoothing_range;
  bool      ev_ready = false;
  double    min_ev = 0, max_ev = 0, delta = 0, theta = 0;
  void *    d_inv_diag = nullptr;
  void *    t1 = nullptr, *t1b = nullptr, *t2 = nullptr, *xold = nullptr, *xin = nullptr, *bin = nullptr;
  void *    d_stage = nullptr; // single precision: double staging buffer of the host entry points
  // pipelined host entry point (dasm_cheb_step_host_batch): two sets of device vectors, one copy stream per direction
  struct Pipe
  {
    bool         ready = false;
    cudaStream_t h2d = nullptr, d2h = nullptr;
    cudaEvent_t  ev_h2d[2], ev_run[2], ev_d2h[2];
    void *       x[2] = {nullptr, nullptr}, *b[2] = {nullptr, nullptr};
    double *     sx[2] = {nullptr, nullptr}, *sb[2] = {nullptr, nullptr};
  } pipe;
  int       t1_zero_idx = -1; // residual buffer (0 / 1) whose shared DoFs the last kernel of the previous fused call left zero
};

// ------------------------------------------------------------------------------------------------
// dispatch helpers
// ------------------------------------------------------------------------------------------------
#define DISPATCH_DEGREE(k, ...)                                         \
  switch (k)                                                            \
    {                                                                   \
      case 1: { constexpr int K = 1; __VA_ARGS__; } break;              \
      case 2: { constexpr int K = 2; __VA_ARGS__; } break;              \
      case 3: { constexpr int K = 3; __VA_ARGS__; } break;              \
      case 4: { constexpr int K = 4; __VA_ARGS__; } break;              \
      case 5: { constexpr int K = 5; __VA_ARGS__; } break;              \
      case 6: { constexpr int K = 6; __VA_ARGS__; } break;              \
      case 7: { constexpr int K = 7; __VA_ARGS__; } break;              \
      case 8: { constexpr int K = 8; __VA_ARGS__; } break;              \
      default: throw std::runtime_error("degree must be in 1..8");      \
    }

#define DISPATCH_PATCH(m, ...)                                                        \
  switch (m)                                                                          \
    {                                                                                 \
      case 2: { constexpr int M = 2; __VA_ARGS__; } break;                            \
      case 3: { constexpr int M = 3; __VA_ARGS__; } break;                            \
      case 4: { constexpr int M = 4; __VA_ARGS__; } break;                            \
      case 5: { constexpr int M = 5; __VA_ARGS__; } break;                            \
      case 6: { constexpr int M = 6; __VA_ARGS__; } break;                            \
      case 7: { constexpr int M = 7; __VA_ARGS__; } break;                            \
      case 8: { constexpr int M = 8; __VA_ARGS__; } break;                            \
      case 9: { constexpr int M = 9; __VA_ARGS__; } break;                            \
      case 10: { constexpr int M = 10; __VA_ARGS__; } break;                          \
      case 11: { constexpr int M = 11; __VA_ARGS__; } break;                          \
      default: throw std::runtime_error("patch size " + std::to_string(m) + " not instantiated"); \
    }

#define DISPATCH_TYPE(ntype, ...)                 \
  if (ntype == DASM_F64)                          \
    {                                             \
      using T = double;                           \
      __VA_ARGS__;                                \
    }                                             \
  else                                            \
    {                                             \
      using T = float;                            \
      __VA_ARGS__;                                \
    }

template <int k>
__global__ void expand_compressed_kernel(uint32_t *out, const uint32_t *cidx, const long long n_cells);

// ---- even-odd blocks of the 1-D matrices of the warp-specialised kernels (EOMat in kernels_fast.cuh) -------------------
// centrosymmetric matrix A[o][i] = A[n-1-o][n-1-i] (mass / stiffness on symmetric nodes), nodal -> nodal
static bool
eo_pack_centrosymmetric(const int n, const double *A, double *P, double *Q)
{
  const int m = (n + 1) / 2, h = n / 2;
  double    amax = 0;
  for (int i = 0; i < n * n; ++i)
    amax = std::max(amax, std::fabs(A[i]));
  for (int o = 0; o < n; ++o)
    for (int i = 0; i < n; ++i)
      if (std::fabs(A[o * n + i] - A[(n - 1 - o) * n + (n - 1 - i)]) > 1e-12 * amax)
        return false;
  for (int o = 0; o < m; ++o)
    {
      for (int i = 0; i < h; ++i)
        P[o * m + i] = 0.5 * (A[o * n + i] + A[o * n + n - 1 - i]);
      if (m > h)
        P[o * m + h] = A[o * n + h];
    }
  for (int o = 0; o < h; ++o)
    for (int i = 0; i < h; ++i)
      Q[o * h + i] = 0.5 * (A[o * n + i] - A[o * n + n - 1 - i]);
  return true;
}

// forward matrix B[a][i] (rows: eigen index in even-first order, columns: nodal) with B[a][n-1-i] = +-B[a][i]
static bool
eo_pack_forward(const int n, const double *B, double *P, double *Q)
{
  const int m = (n + 1) / 2, h = n / 2;
  double    bmax = 0;
  for (int i = 0; i < n * n; ++i)
    bmax = std::max(bmax, std::fabs(B[i]));
  for (int a = 0; a < n; ++a)
    for (int i = 0; i < n; ++i)
      if (std::fabs(B[a * n + n - 1 - i] - (a < m ? 1. : -1.) * B[a * n + i]) > 1e-11 * bmax)
        return false;
  for (int a = 0; a < m; ++a)
    for (int i = 0; i < m; ++i)
      P[a * m + i] = B[a * n + i];
  for (int a = 0; a < h; ++a)
    for (int i = 0; i < h; ++i)
      Q[a * h + i] = B[(m + a) * n + i];
  return true;
}

// backward matrix C[o][a] (rows: nodal, columns: eigen index in even-first order) with C[n-1-o][a] = +-C[o][a]
static bool
eo_pack_backward(const int n, const double *C, double *P, double *Q)
{
  const int m = (n + 1) / 2, h = n / 2;
  double    cmax = 0;
  for (int i = 0; i < n * n; ++i)
    cmax = std::max(cmax, std::fabs(C[i]));
  for (int o = 0; o < n; ++o)
    for (int a = 0; a < n; ++a)
      if (std::fabs(C[(n - 1 - o) * n + a] - (a < m ? 1. : -1.) * C[o * n + a]) > 1e-11 * cmax)
        return false;
  for (int o = 0; o < m; ++o)
    for (int a = 0; a < m; ++a)
      P[o * m + a] = C[o * n + a];
  for (int o = 0; o < h; ++o)
    for (int a = 0; a < h; ++a)
      Q[o * h + a] = C[o * n + m + a];
  return true;
}

template <typename T, int n>
static void
eo_fill(EOMat<T, n> &E, const double *P, const double *Q)
{
  constexpr int m = (n + 1) / 2, h = n / 2;
  for (int i = 0; i < m * m; ++i)
    E.P[i] = (T)P[i];
  for (int i = 0; i < h * h; ++i)
    E.Q[i] = (T)Q[i];
}

static inline unsigned
nblocks(long long n, int bs = 256)
{
  return (unsigned)((n + bs - 1) / bs);
}

// ------------------------------------------------------------------------------------------------
// vector helpers
// ------------------------------------------------------------------------------------------------
template <typename T>
static double
device_dot(dasm_ctx *ctx, const T *a, const T *b, long long n)
{
  const int nb = (int)std::min<long long>(1024, std::max<long long>(1, (n + 255) / 256));
  vec_dot_kernel<T><<<nb, 256, 0, ctx->stream>>>(a, b, ctx->d_partial, n);
  reduce_partials_kernel<<<1, 256, 0, ctx->stream>>>(ctx->d_partial, ctx->d_scalar, nb);
  ctx->launches += 2;
  if (ctx->n_ranks > 1 && ctx->comm)
    NCCL_CHECK(ncclAllReduce(ctx->d_scalar, ctx->d_scalar, 1, ncclDouble, ncclSum, ctx->comm, ctx->stream));
  CUDA_CHECK(cudaMemcpyAsync(ctx->h_scalar, ctx->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  return ctx->h_scalar[0];
}

template <typename T>
static void
apply_hook_post(dasm_op *op, T *dst, const T *src, const dasm_hook *post)
{
  dasm_ctx *ctx = op->ctx;
  if (post == nullptr || post->kind == DASM_HOOK_NONE)
    return;
  // unit-matrix operation on constrained DoFs (matrix_free_internal.h:226-229, 247-255)
  if (op->n_constrained > 0)
    {
      vec_copy_indexed_kernel<T><<<nblocks(op->n_constrained), 256, 0, ctx->stream>>>(dst, src, op->d_constrained, op->n_constrained);
      ctx->launches++;
    }
  const long long n = op->n_owned;
  switch (post->kind)
    {
      case DASM_HOOK_RESIDUAL:
        vec_residual_kernel<T><<<nblocks(n), 256, 0, ctx->stream>>>(dst, (const T *)post->v0, n);
        break;
      case DASM_HOOK_CHEB_UPDATE:
        vec_cheb_update_kernel<T><<<nblocks(n), 256, 0, ctx->stream>>>(dst, dst, (const T *)post->v0, (const T *)post->v1, (T)post->f1, (T)post->f2, n);
        break;
      case DASM_HOOK_SCALE:
        vec_scale_kernel<T><<<nblocks(n), 256, 0, ctx->stream>>>(dst, dst, (T)post->f2, n);
        break;
      default:
        throw std::runtime_error("unsupported post hook kind " + std::to_string(post->kind));
    }
  ctx->launches++;
}

// ------------------------------------------------------------------------------------------------
// operator application
// ------------------------------------------------------------------------------------------------
static bool
deterministic_env()
{
  const char *e = getenv("DASM_DETERMINISTIC");
  return e != nullptr && e[0] == '1';
}

// greedy colouring of the cells by shared vector entries: idx[cell][per_cell] (invalid entries skipped); cells of one colour touch
// disjoint entries.  ptr / cells: the cells grouped by colour.
static void
greedy_coloring(const uint32_t *idx, const long long n_cells, const int per_cell, const long long n_vec, std::vector<long long> &ptr,
                std::vector<uint32_t> &cells)
{
  std::vector<uint64_t> used((size_t)std::max<long long>(n_vec, 1), 0);
  std::vector<uint8_t>  color((size_t)n_cells);
  int                   n_colors = 0;
  for (long long c = 0; c < n_cells; ++c)
    {
      uint64_t m = 0;
      for (int i = 0; i < per_cell; ++i)
        {
          const uint32_t g = idx[c * per_cell + i];
          if (g != INVALID_INDEX)
            m |= used[g & ~LEX_FLAG];
        }
      int col = 0;
      while (col < 64 && ((m >> col) & 1))
        ++col;
      if (col >= 64)
        throw std::runtime_error("deterministic mode: more than 64 colours needed");
      color[c] = (uint8_t)col;
      n_colors = std::max(n_colors, col + 1);
      for (int i = 0; i < per_cell; ++i)
        {
          const uint32_t g = idx[c * per_cell + i];
          if (g != INVALID_INDEX)
            used[g & ~LEX_FLAG] |= (uint64_t)1 << col;
        }
    }
  ptr.assign(n_colors + 1, 0);
  for (long long c = 0; c < n_cells; ++c)
    ptr[color[c] + 1]++;
  for (int i = 0; i < n_colors; ++i)
    ptr[i + 1] += ptr[i];
  cells.resize((size_t)n_cells);
  std::vector<long long> fill(ptr.begin(), ptr.end() - 1);
  for (long long c = 0; c < n_cells; ++c)
    cells[fill[color[c]]++] = (uint32_t)c;
}

static void
setup_deterministic(dasm_op *op)
{
  op->deterministic = true;
  std::vector<uint32_t> cells;
  greedy_coloring(op->nb.cidx.data(), op->n_cells, 27, op->n_vec, op->color_ptr, cells);
  op->d_color_cells = dev_upload(cells, op->ctx->stream);
}

template <typename T>
static void
launch_laplace(dasm_op *op, T *dst, const T *src)
{
  dasm_ctx *ctx = op->ctx;
  KernelTimer timer(ctx, KC_LAPLACE);
  const int n_pass = op->deterministic ? (int)op->color_ptr.size() - 1 : 1;
  for (int pass = 0; pass < n_pass; ++pass)
    {
      const long long count = op->deterministic ? op->color_ptr[pass + 1] - op->color_ptr[pass] : op->n_cells;
      const uint32_t *ids   = op->deterministic ? op->d_color_cells + op->color_ptr[pass] : nullptr;
      if (count == 0)
        continue;
      DISPATCH_DEGREE(op->k, {
        constexpr int n = K + 1, CPB = cells_per_block<K>();
        const size_t  smem = (size_t)CPB * 4 * n * n * n * sizeof(T);
        const unsigned grid = (unsigned)((count + CPB - 1) / CPB);
        if (op->geom_mode == 0)
          {
            auto kern = laplace_generic_kernel<K, T, 0>;
            CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, CPB * n * n, smem, ctx->stream>>>(src, dst, op->d_cidx, (const T *)nullptr, op->cart, count, op->d_plain, ids);
          }
        else if (op->geom_mode == 3)
          {
            auto kern = laplace_generic_kernel<K, T, 2>;
            CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, CPB * n * n, smem, ctx->stream>>>(src, dst, op->d_cidx, (const T *)op->d_geom, op->cart, count, op->d_plain, ids);
          }
        else if (op->geom_mode == 5)
          {
            auto kern = laplace_generic_kernel<K, T, 3>;
            CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, CPB * n * n, smem, ctx->stream>>>(src, dst, op->d_cidx, (const T *)op->d_geom, op->cart, count, op->d_plain, ids);
          }
        else
          {
            auto kern = laplace_generic_kernel<K, T, 1>;
            CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, CPB * n * n, smem, ctx->stream>>>(src, dst, op->d_cidx, (const T *)op->d_geom, op->cart, count, op->d_plain, ids);
          }
      });
      ctx->launches++;
    }
  CUDA_CHECK(cudaGetLastError());
}

template <typename T>
static Epilogue<T>
epilogue_from_hook(const dasm_hook *post)
{
  Epilogue<T> e;
  e.kind = EPI_STORE;
  e.f1   = 0;
  e.f2   = 0;
  e.v0   = nullptr;
  e.v1   = nullptr;
  if (post == nullptr || post->kind == DASM_HOOK_NONE)
    return e;
  switch (post->kind)
    {
      case DASM_HOOK_RESIDUAL:
        e.kind = EPI_RESIDUAL;
        e.v0   = (const T *)post->v0;
        break;
      case DASM_HOOK_CHEB_UPDATE:
        e.kind = EPI_CHEB;
        e.f1   = (T)post->f1;
        e.f2   = (T)post->f2;
        e.v0   = (const T *)post->v0;
        e.v1   = (const T *)post->v1;
        break;
      case DASM_HOOK_SCALE:
        e.kind = EPI_SCALE;
        e.f2   = (T)post->f2;
        break;
      default:
        throw std::runtime_error("unsupported post hook kind " + std::to_string(post->kind));
    }
  return e;
}

// multi-rank: ghost values of the source are updated before a brick kernel (update_ghost_values,
// matrix_free_internal.h:321-324); contributions to ghost DoFs are sent to and added at their owners after it
// (compress(add), 350-352).  In SHARED_DIRECT mode the owners hold base + local contributions in dst, so the
// ghost part of dst starts from zero; in SHARED_ACC mode the accumulator is exchanged and its ghost part
// re-zeroed (zero-invariant).
template <typename T>
static void
brick_pre_exchange(dasm_op *op, T *dst, const T *src, const int shared_mode)
{
  if (!op->exchange.active())
    return;
  KernelTimer timer(op->ctx, KC_EXCHANGE);
  op->exchange.run<T>(const_cast<T *>(src), false);
  if (shared_mode == SHARED_DIRECT && op->n_ghost > 0)
    CUDA_CHECK(cudaMemsetAsync(dst + op->n_owned, 0, (size_t)op->n_ghost * sizeof(T), op->ctx->stream));
}

template <typename T>
static void
brick_post_exchange(dasm_op *op, T *dst, const int shared_mode, const bool needs_compression)
{
  if (!op->exchange.active())
    return;
  KernelTimer timer(op->ctx, KC_EXCHANGE);
  T *         vec = (shared_mode == SHARED_DIRECT) ? dst : (T *)op->d_acc;
  if (needs_compression)
    op->exchange.run<T>(vec, true);
  if (shared_mode == SHARED_ACC && op->n_ghost > 0)
    CUDA_CHECK(cudaMemsetAsync((T *)op->d_acc + op->n_owned, 0, (size_t)op->n_ghost * sizeof(T), op->ctx->stream));
}

template <int K, int BZ, typename T, typename Kern>
static int
brick_grid(dasm_op *op, Kern kern, size_t smem)
{
  CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BrickGeom<K, BZ>::NT, smem));
  occ = std::max(occ, 1);
  return std::min(op->n_bricks, op->n_sm * occ);
}

// finish the shared-face DoFs (dst = epilogue(acc), acc = 0) and run the epilogue on constrained DoFs
template <int K, int BZ, typename T>
static void
brick_finish(dasm_op *op, T *dst, const T *src_for_constrained, const Epilogue<T> &epi, const int shared_mode)
{
  dasm_ctx *ctx = op->ctx;
  if (op->n_shared > 0 && shared_mode == SHARED_ACC)
    {
      KernelTimer timer(ctx, KC_VECTOR);
      finish_shared_kernel<T><<<nblocks(op->n_shared), 256, 0, ctx->stream>>>(dst, (T *)op->d_acc, epi, op->d_shared_list, op->n_shared);
      ctx->launches++;
    }
  if (op->n_constrained > 0)
    {
      epilogue_indexed_kernel<T><<<nblocks(op->n_constrained), 256, 0, ctx->stream>>>(dst, src_for_constrained, epi, op->d_constrained,
                                                                                      op->n_constrained);
      ctx->launches++;
    }
}

template <typename T>
static int
epilogue_n_operands(const Epilogue<T> &epi)
{
  if (epi.kind == EPI_RESIDUAL)
    return 1;
  if (epi.kind == EPI_CHEB)
    return (epi.f1 != T(0) && epi.v1 != nullptr) ? 2 : 1;
  return 0;
}

// timing experiments with the warp-specialised kernels (DASM_FAST_DBG, results invalid; kernels_fast.cuh FastMaps)
static int
fast_dbg()
{
  static const int v = getenv("DASM_FAST_DBG") ? atoi(getenv("DASM_FAST_DBG")) : 0;
  return v;
}

// ---- overlap of the halo exchange with the interior bricks (fused SHARED_DIRECT sequences on several ranks) --------------
// main stream:  [wait ghost update] boundary bricks + irregular bricks | interior bricks                 | [wait compress]
// comm stream:  ghost update(src)                                     | compress(dst)  (-> ghost update of the next sweep)
// The source of a sweep is the destination of the previous one: its values on the partition boundary are final once the
// previous compress has run on the comm stream, so the ghost update of sweep j+1 overlaps with the interior bricks of sweep j.
static bool
overlap_enabled(const dasm_op *op, const int shared_mode, const int n_fast_boundary, const int n_fast)
{
  // default on (DASM_OVERLAP=0 switches it off): with the device-initiated exchange (pack kernel writes into the peer, no NCCL
  // kernels) the exchange runs next to the interior bricks: 5.96 instead of 6.20 ms per step at 2 GPUs (profiles/r02k); with the
  // NCCL send / receive fallback the gain was 1 % (r01f).  Parity: tests/multi_gpu_parity.py at 2, 4 and 8 ranks.
  static const bool on = !(getenv("DASM_OVERLAP") && getenv("DASM_OVERLAP")[0] == '0');
  return on && op->exchange.active() && shared_mode == SHARED_DIRECT && n_fast > n_fast_boundary;
}

// SMs left free by the (persistent, one block per SM) interior launch so that the NCCL send / receive kernels of the
// comm stream can run next to it
static int
overlap_reserved_sms()
{
  static const int v = getenv("DASM_OVERLAP_RESERVE_SMS") ? atoi(getenv("DASM_OVERLAP_RESERVE_SMS")) : 0;
  return v;
}

template <typename T>
static void
overlap_pre(dasm_op *op, T *dst, const T *src)
{
  dasm_ctx *ctx = op->ctx;
  if ((const void *)src != op->last_compressed)
    {
      // the source was produced on the main stream
      CUDA_CHECK(cudaEventRecord(ctx->ev_a, ctx->stream));
      CUDA_CHECK(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_a, 0));
    }
  if (op->n_ghost > 0)
    CUDA_CHECK(cudaMemsetAsync(dst + op->n_owned, 0, (size_t)op->n_ghost * sizeof(T), ctx->stream));
  op->exchange.run<T>(const_cast<T *>(src), false, ctx->comm_stream);
  CUDA_CHECK(cudaEventRecord(ctx->ev_b, ctx->comm_stream));
  CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_b, 0)); // the boundary bricks read the ghost values
}

// after the boundary (and irregular) bricks have been launched on the main stream
template <typename T>
static void
overlap_mid(dasm_op *op, T *dst, const bool needs_compression)
{
  dasm_ctx *ctx = op->ctx;
  CUDA_CHECK(cudaEventRecord(ctx->ev_a, ctx->stream));
  CUDA_CHECK(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_a, 0));
  if (needs_compression)
    op->exchange.run<T>(dst, true, ctx->comm_stream);
  CUDA_CHECK(cudaEventRecord(ctx->ev_b, ctx->comm_stream));
  op->last_compressed = dst;
}

// after the interior bricks have been launched: later work on the main stream sees the compressed destination
static void
overlap_post(dasm_op *op)
{
  CUDA_CHECK(cudaStreamWaitEvent(op->ctx->stream, op->ctx->ev_b, 0));
}

// descriptors of the lex bricks `ids` (boundary bricks first: [0, n_boundary)) in processing order, grouped into chunks of
// consecutive +x neighbours (kernels_tma.cuh TmaList)
struct TmaChunked
{
  std::vector<TmaBrick> descs;
  std::vector<uint32_t> chunk_start;
  int                   n_chunks_boundary = 0;
  int                   any_mode1         = 0;
  int                   any_mode1_interior = 0; // ... among the chunks behind the boundary chunks
};

static TmaChunked
tma_build_list_L(const dasm_op *op, const std::vector<uint32_t> &ids, const int n_boundary, const int L)
{
  TmaChunked out;
  const int  n   = (int)ids.size();
  int        len = 0;
  for (int i = 0; i < n; ++i)
    {
      TmaBrick t = op->h_tma[ids[i]];
      out.any_mode1 |= (int)(t.flags & TMA_MODE1);
      if (i >= n_boundary)
        out.any_mode1_interior |= (int)(t.flags & TMA_MODE1);
      bool cont = false; // continues the chunk of the previous brick
      if (i > 0 && i != n_boundary && len < L)
        {
          const TmaBrick &p = out.descs.back();
          cont = !(p.flags & TMA_MODE1) && !(t.flags & TMA_MODE1) && p.nb[0] == t.base;
        }
      if (cont)
        {
          out.descs.back().flags |= TMA_CARRY_OUT;
          t.flags |= TMA_CARRY_IN;
          ++len;
        }
      else
        {
          if (i > 0)
            out.descs.back().flags |= TMA_LAST;
          if (i == n_boundary)
            out.n_chunks_boundary = (int)out.chunk_start.size();
          out.chunk_start.push_back((uint32_t)i);
          len = 1;
        }
      out.descs.push_back(t);
    }
  if (n > 0)
    out.descs.back().flags |= TMA_LAST;
  if (n_boundary >= n)
    out.n_chunks_boundary = (int)out.chunk_start.size();
  out.chunk_start.push_back((uint32_t)n);
  return out;
}

// Chunk length: long chunks save the red.add / zeroing of the faces between their bricks (measured: time ~ 1 + 0.235 / L), but the
// chunks are dealt out statically to the resident blocks, so their number should fill the last round: the length 1..16 with the
// best product of both effects (DASM_TMA_CHUNK overrides).
static TmaChunked
tma_build_list(const dasm_op *op, const std::vector<uint32_t> &ids, const int n_boundary)
{
  if (const char *e = getenv("DASM_TMA_CHUNK"))
    return tma_build_list_L(op, ids, n_boundary, std::max(1, atoi(e)));
  const int  blocks = std::max(1, op->n_sm * tma_min_blocks(op->k, (int)op->esize()));
  if (getenv("DASM_VERBOSE"))
    fprintf(stderr, "[dasm] chunking %zu lex bricks over %d resident blocks\n", ids.size(), blocks);
  TmaChunked best;
  double     best_score = -1;
  for (int L = 16; L >= 1; --L)
    {
      TmaChunked   c  = tma_build_list_L(op, ids, n_boundary, L);
      const int    nc = (int)c.chunk_start.size() - 1;
      // static deal: block b processes the chunks b, b + blocks, ...; efficiency = mean load / maximum load (in bricks)
      double eff = 1.;
      if (nc > 0)
        {
          const int             nb_used = std::min(nc, blocks);
          std::vector<uint32_t> load(nb_used, 0);
          for (int q = 0; q < nc; ++q)
            load[q % nb_used] += c.chunk_start[q + 1] - c.chunk_start[q];
          const uint32_t mx = *std::max_element(load.begin(), load.end());
          eff               = (double)ids.size() / ((double)blocks * mx);
        }
      // average number of bricks per chunk (chunks are cut at the ends of the brick rows and at non-lex neighbours)
      const double len = nc == 0 ? 1. : (double)ids.size() / nc;
      const double score = eff / (1. + 0.235 / len);
      if (getenv("DASM_VERBOSE"))
        fprintf(stderr, "[dasm]   L = %d: %d chunks, efficiency %.3f, score %.3f\n", L, nc, eff, score);
      if (score > best_score)
        {
          best_score = score;
          best       = std::move(c);
        }
    }
  return best;
}

// chunk range of a launch over the bricks [first, first + count) of a fast list (whole list, boundary part or interior part)
static bool
tma_chunk_range(const int first, const int count, const int n_fast, const int n_fast_boundary, const int n_chunks, const int n_chunks_boundary,
                int &c_first, int &c_count)
{
  if (first == 0 && count == n_fast)
    c_first = 0, c_count = n_chunks;
  else if (first == 0 && count == n_fast_boundary)
    c_first = 0, c_count = n_chunks_boundary;
  else if (first == n_fast_boundary && count == n_fast - n_fast_boundary)
    c_first = n_chunks_boundary, c_count = n_chunks - n_chunks_boundary;
  else
    return false;
  return true;
}

// thread blocks per SM of the TMA-fed kernels (kernels_tma.cuh TMA_MINB; DASM_TMA_CTAS overrides for experiments)
static int
tma_ctas_per_sm(const int k, const int esize)
{
  static const int v = getenv("DASM_TMA_CTAS") ? std::max(1, atoi(getenv("DASM_TMA_CTAS"))) : 0;
  return v > 0 ? v : tma_min_blocks(k, esize);
}

// tensor maps of a vector for the TMA-fed kernels (cached per pointer); nullptr: not usable (alignment)
static const TmaMaps *
tma_maps_for(dasm_op *op, const void *vec)
{
  auto it = op->tma_cache.find(vec);
  if (it != op->tma_cache.end())
    return &it->second;
  TmaMaps     m;
  std::string err;
  if (!tma_encode_maps(m, vec, op->k, (int)op->esize(), (long long)op->nb.n_lex, err))
    return nullptr;
  if (op->tma_cache.size() > 64)
    op->tma_cache.clear();
  return &op->tma_cache.emplace(vec, m).first->second;
}

static bool
tma_aligned(const void *a, const void *b, const void *c)
{
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) == 0;
}

// warp-specialised Laplace kernel over the regular bricks; false: not available (shared memory), nothing launched
template <int K, typename T>
static bool
launch_laplace_fast(dasm_op *op, T *dst, const T *src, const Epilogue<T> &epi, const int shared_mode, const NextInit<T> &ni, const int first = 0,
                    int count = -1, const int reserve_sms = 0)
{
  if (count < 0)
    count = op->n_fast - first;
  if (count == 0)
    return true;
  if (!op->tma_ok || epilogue_n_operands(epi) > 1) // one operand box (b of the residual epilogue)
    return false;
  const TmaMaps *tm = tma_maps_for(op, src);
  if (tm == nullptr || !tma_aligned(dst, epi.v0, epi.v1) || tma_laplace_smem(K, (int)sizeof(T)) > (size_t)op->max_smem)
    return false;
  int c_first = 0, c_count = 0;
  if (!tma_chunk_range(first, count, op->n_fast, op->n_fast_boundary, op->tma_lap_n_chunks, op->tma_lap_n_chunks_boundary, c_first, c_count))
    return false;
  if (c_count == 0)
    return true;
  const int      grid = std::min(c_count, std::max(1, op->n_sm * tma_ctas_per_sm(K, (int)sizeof(T)) - reserve_sms));
  // (a launch over the interior bricks only - the overlapped schedule - usually contains no brick with an index list)
  const bool     interior_only = first > 0 && first == op->n_fast_boundary;
  const TmaList  list = {op->d_tma_lap, op->d_tma_foreign, op->d_tma_lap_chunks + c_first, c_count,
                         interior_only ? op->tma_any_mode1_interior : op->tma_any_mode1};
  const TmaMaps *o0   = epi.v0 ? tma_maps_for(op, epi.v0) : tm;
  if (o0 == nullptr)
    return false;
  launch_laplace_tma<T>(K, op->ctx->stream, grid, src, dst, (T *)op->d_acc, epi, op->lap_P, op->lap_Q, *tm, o0->main, shared_mode, ni, list,
                        fast_dbg());
  op->ctx->launches++;
  return true;
}

template <int K, int BZ, typename T>
static void
launch_laplace_brick(dasm_op *op, T *dst, const T *src, const Epilogue<T> &epi, bool copy_constrained, const int shared_mode,
                     const NextInit<T> &ni)
{
  dasm_ctx *   ctx   = op->ctx;
  const int    n_ops = epilogue_n_operands(epi);
  const size_t smem  = BrickGeom<K, BZ>::template smem_bytes<T>(n_ops, BZ == 4);
  if constexpr (BZ == 4 && K >= 2 && K <= 4)
    {
      if (op->geom_mode == 0 && op->fast_ok && n_ops <= 1 && tma_laplace_smem(K, (int)sizeof(T)) <= (size_t)op->max_smem &&
          overlap_enabled(op, shared_mode, op->n_fast_boundary, op->n_fast))
        {
          // halo exchange overlapped with the interior bricks (see overlap_pre)
          overlap_pre<T>(op, dst, src);
          {
            KernelTimer timer(ctx, KC_LAPLACE);
            launch_laplace_fast<K, T>(op, dst, src, epi, shared_mode, ni, 0, op->n_fast_boundary);
            if (op->n_slow > 0)
              {
                auto      kern = laplace_brick_kernel<K, T, BZ, 0>;
                const int grid = std::min(op->n_slow, brick_grid<K, BZ, T>(op, kern, smem));
                kern<<<grid, BrickGeom<K, BZ>::NT, smem, ctx->stream>>>(src, dst, (T *)op->d_acc, epi, op->d_cidx, op->d_bricks, op->n_slow,
                                                                         (const T *)nullptr, op->cart, n_ops, shared_mode, ni, op->maps, op->d_slow_ids);
                ctx->launches++;
              }
            overlap_mid<T>(op, dst, true);
            launch_laplace_fast<K, T>(op, dst, src, epi, shared_mode, ni, op->n_fast_boundary, op->n_fast - op->n_fast_boundary, overlap_reserved_sms());
          }
          overlap_post(op);
          CUDA_CHECK(cudaGetLastError());
          brick_finish<K, BZ, T>(op, dst, copy_constrained ? src : nullptr, epi, shared_mode);
          CUDA_CHECK(cudaGetLastError());
          return;
        }
    }
  brick_pre_exchange<T>(op, dst, src, shared_mode);
  {
    KernelTimer timer(ctx, KC_LAPLACE);
    if (op->geom_mode == 0)
      {
        // regular bricks: warp-specialised kernel; the rest (if any): brick kernel over the list of the other bricks
        const uint32_t *order    = nullptr;
        int             n_bricks = op->n_bricks;
        if constexpr (BZ == 4 && K >= 2 && K <= 4)
          {
            if (op->fast_ok && launch_laplace_fast<K, T>(op, dst, src, epi, shared_mode, ni))
              {
                order    = op->d_slow_ids;
                n_bricks = op->n_slow;
              }
          }
        if (n_bricks > 0)
          {
            auto      kern = laplace_brick_kernel<K, T, BZ, 0>;
            const int grid = std::min(n_bricks, brick_grid<K, BZ, T>(op, kern, smem));
            kern<<<grid, BrickGeom<K, BZ>::NT, smem, ctx->stream>>>(src, dst, (T *)op->d_acc, epi, op->d_cidx, op->d_bricks, n_bricks,
                                                                     (const T *)nullptr, op->cart, n_ops, shared_mode, ni, op->maps, order);
            ctx->launches++;
          }
      }
    else if (op->geom_mode == 1)
      {
        auto      kern = laplace_brick_kernel<K, T, BZ, 1>;
        const int grid = brick_grid<K, BZ, T>(op, kern, smem);
        kern<<<grid, BrickGeom<K, BZ>::NT, smem, ctx->stream>>>(src, dst, (T *)op->d_acc, epi, op->d_cidx, op->d_bricks, op->n_bricks,
                                                                 (const T *)op->d_geom, op->cart, n_ops, shared_mode, ni, op->maps, (const uint32_t *)nullptr);
        ctx->launches++;
      }
    else
      {
        auto      kern = laplace_brick_kernel<K, T, BZ, 2>;
        const int grid = brick_grid<K, BZ, T>(op, kern, smem);
        kern<<<grid, BrickGeom<K, BZ>::NT, smem, ctx->stream>>>(src, dst, (T *)op->d_acc, epi, op->d_cidx, op->d_bricks, op->n_bricks,
                                                                 (const T *)op->d_qcoef, op->cart, n_ops, shared_mode, ni, op->maps, (const uint32_t *)nullptr);
        ctx->launches++;
      }
  }
  CUDA_CHECK(cudaGetLastError());
  brick_post_exchange<T>(op, dst, shared_mode, true);
  brick_finish<K, BZ, T>(op, dst, copy_constrained ? src : nullptr, epi, shared_mode);
  CUDA_CHECK(cudaGetLastError());
}

// degrees 5 and 6 on meshes whose bricks are all lex bricks (one rank, periodic or interior only): the TMA-fed kernel is the only
// cell kernel; there are no constrained DoFs and no ghost entries
template <int K, typename T>
static void
launch_laplace_tma_only(dasm_op *op, T *dst, const T *src, const Epilogue<T> &epi, const int shared_mode, const NextInit<T> &ni)
{
  {
    KernelTimer timer(op->ctx, KC_LAPLACE);
    if (!launch_laplace_fast<K, T>(op, dst, src, epi, shared_mode, ni))
      throw std::runtime_error("laplace_tma_kernel: vectors must be 16-byte aligned device vectors of the operator's layout");
  }
  CUDA_CHECK(cudaGetLastError());
  brick_finish<K, 4, T>(op, dst, (const T *)nullptr, epi, shared_mode);
  CUDA_CHECK(cudaGetLastError());
}

template <typename T>
static NextInit<T>
no_next_init()
{
  NextInit<T> ni;
  ni.out = nullptr;
  ni.v0  = nullptr;
  ni.v1  = nullptr;
  ni.f1  = 0;
  return ni;
}

template <typename T>
static void
op_vmult_brick(dasm_op *op, T *dst, const T *src, const dasm_hook *post, const int shared_mode = SHARED_ACC,
               const NextInit<T> &ni = no_next_init<T>())
{
  const Epilogue<T> epi  = epilogue_from_hook<T>(post);
  const bool        copy = (post != nullptr && post->kind != DASM_HOOK_NONE);
  switch (op->k)
    {
      case 1: launch_laplace_brick<1, 4, T>(op, dst, src, epi, copy, shared_mode, ni); break;
      case 2: launch_laplace_brick<2, 4, T>(op, dst, src, epi, copy, shared_mode, ni); break;
      case 3: launch_laplace_brick<3, 4, T>(op, dst, src, epi, copy, shared_mode, ni); break;
      case 4:
        if (op->brick_bz == 2)
          launch_laplace_brick<4, 2, T>(op, dst, src, epi, copy, shared_mode, ni);
        else
          launch_laplace_brick<4, 4, T>(op, dst, src, epi, copy, shared_mode, ni);
        break;
      case 5:
        if (op->tma_only)
          launch_laplace_tma_only<5, T>(op, dst, src, epi, shared_mode, ni);
        else
          launch_laplace_brick<5, 2, T>(op, dst, src, epi, copy, shared_mode, ni);
        break;
      case 6:
        if (op->tma_only)
          {
            launch_laplace_tma_only<6, T>(op, dst, src, epi, shared_mode, ni);
            break;
          }
        // fall through
      default: throw std::runtime_error("internal: brick path for unsupported degree");
    }
}

// stand-alone pre-initialisation of a destination on all shared DoFs
template <typename T>
static void
init_shared(dasm_op *op, const NextInit<T> &ni)
{
  if (op->n_shared == 0)
    return;
  KernelTimer timer(op->ctx, KC_VECTOR);
  init_shared_kernel<T><<<nblocks(op->n_shared), 256, 0, op->ctx->stream>>>(ni, op->d_shared_list, op->n_shared);
  op->ctx->launches++;
}

template <typename T>
static void
op_vmult(dasm_op *op, T *dst, const T *src, const dasm_hook *pre, const dasm_hook *post)
{
  dasm_ctx *ctx = op->ctx;
  if (pre != nullptr && pre->kind != DASM_HOOK_NONE && pre->kind != DASM_HOOK_ZERO_DST)
    throw std::runtime_error("LaplaceOperatorMatrixFree::vmult: only the zeroing pre-operation is supported");
  DASM_REQUIRE((const void *)dst != (const void *)src, "vmult: dst and src must not alias");
  bool brick_path = op->use_brick;
  if (op->tma_only)
    {
      // the TMA-fed kernel is the only brick kernel of these degrees: two-operand epilogues and unaligned user vectors go
      // through the generic kernels
      const Epilogue<T> epi = epilogue_from_hook<T>(post);
      brick_path            = epilogue_n_operands(epi) <= 1 && tma_aligned(dst, epi.v0, epi.v1) && tma_aligned(src, nullptr, nullptr);
    }
  if (brick_path)
    {
      op_vmult_brick<T>(op, dst, src, post);
      return;
    }
  // pre: dst = 0 (the reference's default pre-operation, operator.h:1356-1363; the cell kernel
  // accumulates, so dst must start from zero in all cases)
  CUDA_CHECK(cudaMemsetAsync(dst, 0, (size_t)op->n_vec * sizeof(T), ctx->stream));
  op->exchange.run<T>(const_cast<T *>(src), false); // update_ghost_values(src)
  launch_laplace<T>(op, dst, src);
  op->exchange.run<T>(dst, true); // compress(add)
  apply_hook_post<T>(op, dst, src, post);
}

// ------------------------------------------------------------------------------------------------
// FDM application
// ------------------------------------------------------------------------------------------------
template <int M, typename T>
static void
launch_fdm_m(dasm_fdm *f, T *dst, const T *src)
{
  dasm_op *      op   = f->op;
  dasm_ctx *     ctx  = op->ctx;
  KernelTimer    timer(ctx, KC_FDM);
  constexpr int  CPB  = fdm_cells_per_block<M>();
  const size_t   smem = (size_t)CPB * (M * M * M + 3 * M * M + 3 * M) * sizeof(T);
  const void *   w    = (f->wmode == 3) ? f->d_wvec : f->d_cw;
  // deterministic mode: colour by colour (patches with explicit lists have their own colouring)
  const std::vector<long long> &cptr = f->d_color_cells ? f->color_ptr : op->color_ptr;
  const uint32_t *              call = f->d_color_cells ? f->d_color_cells : op->d_color_cells;
  const int                     n_pass = op->deterministic ? (int)cptr.size() - 1 : 1;
  for (int pass = 0; pass < n_pass; ++pass)
    {
      const long long count = op->deterministic ? cptr[pass + 1] - cptr[pass] : op->n_cells;
      const uint32_t *ids   = op->deterministic ? call + cptr[pass] : nullptr;
      if (count == 0)
        continue;
      const unsigned grid = (unsigned)((count + CPB - 1) / CPB);
      if (f->d_pidx == nullptr)
        {
          if (M - 1 != op->k)
            throw std::runtime_error("internal: compressed FDM patch needs m == k+1");
          if constexpr (M <= 9)
            {
              auto kern = fdm_generic_kernel<M, T, 0>;
              CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
              kern<<<grid, CPB * M * M, smem, ctx->stream>>>(src, dst, op->d_cidx, f->d_inst, (const T *)f->d_S, (const T *)f->d_lam,
                                                             (const T *)w, f->wmode, (int)f->w_pre, (int)f->w_post, count, ids);
            }
        }
      else
        {
          if constexpr (M >= 3)
            {
              auto kern = fdm_generic_kernel<M, T, 1>;
              CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
              kern<<<grid, CPB * M * M, smem, ctx->stream>>>(src, dst, f->d_pidx, f->d_inst, (const T *)f->d_S, (const T *)f->d_lam,
                                                             (const T *)w, f->wmode, (int)f->w_pre, (int)f->w_post, count, ids);
            }
          else
            throw std::runtime_error("explicit patch lists are instantiated for patch sizes >= 3");
        }
      ctx->launches++;
    }
  CUDA_CHECK(cudaGetLastError());
}

template <typename T>
static void
launch_fdm(dasm_fdm *f, T *dst, const T *src)
{
  DISPATCH_PATCH(f->m, launch_fdm_m<M, T>(f, dst, src));
}

// The reference skips compress(add) for RAS only after it has verified that the patch that keeps a DoF (weight 1) is a locally
// owned cell for EVERY owned DoF, and throws otherwise (matrix_free.h:654-668).  Here the winner of an entity is the touching
// cell with the smallest global id, i.e. on a partition interface a cell of the LOWER rank, while the DoF is owned by the upper
// one: the result sits in a ghost slot and must be sent.  Several ranks therefore keep the compression.
static bool
fdm_needs_compression(const dasm_fdm *f)
{
  return f->weight_type != DASM_WEIGHT_RAS || (f->op->mesh != nullptr && f->op->mesh->mesh->n_ranks() > 1);
}

template <int K, typename T>
static bool
launch_fdm_fast(dasm_fdm *f, T *dst, const T *src, const Epilogue<T> &epi, const int shared_mode, const NextInit<T> &ni, const int first = 0,
                int count = -1, const int reserve_sms = 0)
{
  dasm_op *op = f->op;
  if (count < 0)
    count = f->n_fast - first;
  if (count == 0)
    return true;
  if (!op->tma_ok || f->d_tma_list == nullptr)
    return false;
  const TmaMaps *tm = tma_maps_for(op, src);
  if (tm == nullptr || !tma_aligned(dst, epi.v0, epi.v1) || tma_fdm_smem(K, (int)sizeof(T)) > (size_t)op->max_smem)
    return false;
  int c_first = 0, c_count = 0;
  if (!tma_chunk_range(first, count, f->n_fast, f->n_fast_boundary, f->tma_n_chunks, f->tma_n_chunks_boundary, c_first, c_count))
    return false;
  if (c_count == 0)
    return true;
  const int      grid = std::min(c_count, std::max(1, op->n_sm * tma_ctas_per_sm(K, (int)sizeof(T)) - reserve_sms));
  const bool     interior_only = first > 0 && first == f->n_fast_boundary;
  const TmaList  list = {f->d_tma_list, op->d_tma_foreign, f->d_tma_chunks + c_first, c_count,
                         interior_only ? f->tma_any_mode1_interior : f->tma_any_mode1};
  const TmaMaps *o0 = epi.v0 ? tma_maps_for(op, epi.v0) : tm, *o1 = epi.v1 ? tma_maps_for(op, epi.v1) : tm;
  if (o0 == nullptr || o1 == nullptr)
    return false;
  launch_fdm_tma<T>(K, op->ctx->stream, grid, src, dst, (T *)op->d_acc, epi, f->fast_P, f->fast_Q, f->fast_inv, *tm, o0->main, o1->main,
                    shared_mode, ni, list, fast_dbg());
  op->ctx->launches++;
  return true;
}

template <int K, int BZ, typename T>
static void
launch_fdm_brick(dasm_fdm *f, T *dst, const T *src, const Epilogue<T> &epi, const int shared_mode, const NextInit<T> &ni)
{
  dasm_op *    op    = f->op;
  dasm_ctx *   ctx   = op->ctx;
  const int    n_ops = epilogue_n_operands(epi);
  const size_t smem  = BrickGeom<K, BZ>::template smem_bytes<T>(n_ops, BZ == 4);
  static const int dbg = getenv("DASM_DEBUG_SKIP") ? atoi(getenv("DASM_DEBUG_SKIP")) : 0; // timing experiments only
  WeightTable<T>   wt;
  for (int i = 0; i < 16; ++i)
    wt.v[i] = (T)f->wtab[i];
  if constexpr (BZ == 4 && K >= 2 && K <= 4)
    {
      if (f->fast_ok && tma_fdm_smem(K, (int)sizeof(T)) <= (size_t)op->max_smem &&
          overlap_enabled(op, shared_mode, f->n_fast_boundary, f->n_fast))
        {
          // halo exchange overlapped with the interior bricks (see overlap_pre)
          overlap_pre<T>(op, dst, src);
          {
            KernelTimer timer(ctx, KC_FDM);
            launch_fdm_fast<K, T>(f, dst, src, epi, shared_mode, ni, 0, f->n_fast_boundary);
            if (f->n_slow > 0)
              {
                auto      kern = fdm_brick_kernel<K, T, BZ>;
                const int grid = std::min(f->n_slow, brick_grid<K, BZ, T>(op, kern, smem));
                kern<<<grid, BrickGeom<K, BZ>::NT, smem, ctx->stream>>>(src, dst, (T *)op->d_acc, epi, op->d_cidx, op->d_bricks, f->n_slow, f->d_inst,
                                                                         (const T *)f->d_S, (const T *)f->d_lam, (const uint8_t *)(f->wmode == 1 ? f->d_cwcode : nullptr), wt, f->d_brick_tri,
                                                                         (int)f->w_pre, (int)f->w_post, n_ops, shared_mode, ni, op->maps, dbg, f->d_slow_ids);
                ctx->launches++;
              }
            overlap_mid<T>(op, dst, fdm_needs_compression(f));
            launch_fdm_fast<K, T>(f, dst, src, epi, shared_mode, ni, f->n_fast_boundary, f->n_fast - f->n_fast_boundary, overlap_reserved_sms());
          }
          overlap_post(op);
          CUDA_CHECK(cudaGetLastError());
          brick_finish<K, BZ, T>(op, dst, (const T *)nullptr, epi, shared_mode);
          CUDA_CHECK(cudaGetLastError());
          return;
        }
    }
  brick_pre_exchange<T>(op, dst, src, shared_mode);
  {
    KernelTimer     timer(ctx, KC_FDM);
    const uint32_t *order    = nullptr;
    int             n_bricks = op->n_bricks;
    if constexpr (BZ == 4 && K >= 2 && K <= 4)
      {
        if (f->fast_ok && launch_fdm_fast<K, T>(f, dst, src, epi, shared_mode, ni))
          {
            order    = f->d_slow_ids;
            n_bricks = f->n_slow;
          }
      }
    if (n_bricks > 0)
      {
        auto      kern = fdm_brick_kernel<K, T, BZ>;
        const int grid = std::min(n_bricks, brick_grid<K, BZ, T>(op, kern, smem));
        kern<<<grid, BrickGeom<K, BZ>::NT, smem, ctx->stream>>>(src, dst, (T *)op->d_acc, epi, op->d_cidx, op->d_bricks, n_bricks, f->d_inst,
                                                                 (const T *)f->d_S, (const T *)f->d_lam, (const uint8_t *)(f->wmode == 1 ? f->d_cwcode : nullptr), wt, f->d_brick_tri,
                                                                 (int)f->w_pre, (int)f->w_post, n_ops, shared_mode, ni, op->maps, dbg, order);
        ctx->launches++;
      }
  }
  CUDA_CHECK(cudaGetLastError());
  brick_post_exchange<T>(op, dst, shared_mode, fdm_needs_compression(f));
  brick_finish<K, BZ, T>(op, dst, (const T *)nullptr, epi, shared_mode);
  CUDA_CHECK(cudaGetLastError());
}

static bool
fdm_uses_brick(const dasm_fdm *f)
{
  if (f->op->tma_only) // no brick kernels behind the TMA-fed kernel: every brick must qualify for it
    return f->d_pidx == nullptr && f->fast_ok && f->n_slow == 0 && f->d_tma_list != nullptr;
  return f->op->use_brick && f->d_pidx == nullptr && (f->wmode == 0 || f->wmode == 1);
}

template <int K, typename T>
static void
launch_fdm_tma_only(dasm_fdm *f, T *dst, const T *src, const Epilogue<T> &epi, const int shared_mode, const NextInit<T> &ni)
{
  dasm_op *op = f->op;
  {
    KernelTimer timer(op->ctx, KC_FDM);
    if (!launch_fdm_fast<K, T>(f, dst, src, epi, shared_mode, ni))
      throw std::runtime_error("fdm_tma_kernel: vectors must be 16-byte aligned device vectors of the operator's layout");
  }
  CUDA_CHECK(cudaGetLastError());
  brick_finish<K, 4, T>(op, dst, (const T *)nullptr, epi, shared_mode);
  CUDA_CHECK(cudaGetLastError());
}

template <typename T>
static void
fdm_vmult_brick(dasm_fdm *f, T *dst, const T *src, const dasm_hook *post, const int shared_mode = SHARED_ACC,
                const NextInit<T> &ni = no_next_init<T>())
{
  dasm_op *         op  = f->op;
  const Epilogue<T> epi = epilogue_from_hook<T>(post);
  switch (op->k)
    {
      case 1: launch_fdm_brick<1, 4, T>(f, dst, src, epi, shared_mode, ni); break;
      case 2: launch_fdm_brick<2, 4, T>(f, dst, src, epi, shared_mode, ni); break;
      case 3: launch_fdm_brick<3, 4, T>(f, dst, src, epi, shared_mode, ni); break;
      case 4:
        if (op->brick_bz == 2)
          launch_fdm_brick<4, 2, T>(f, dst, src, epi, shared_mode, ni);
        else
          launch_fdm_brick<4, 4, T>(f, dst, src, epi, shared_mode, ni);
        break;
      case 5:
        if (op->tma_only)
          launch_fdm_tma_only<5, T>(f, dst, src, epi, shared_mode, ni);
        else
          launch_fdm_brick<5, 2, T>(f, dst, src, epi, shared_mode, ni);
        break;
      case 6:
        if (op->tma_only)
          {
            launch_fdm_tma_only<6, T>(f, dst, src, epi, shared_mode, ni);
            break;
          }
        // fall through
      default: throw std::runtime_error("internal: brick path for unsupported degree");
    }
}

template <typename T>
static void
fdm_vmult(dasm_fdm *f, T *dst, const T *src, const dasm_hook *pre, const dasm_hook *post)
{
  dasm_op * op  = f->op;
  dasm_ctx *ctx = op->ctx;
  if (pre != nullptr && pre->kind != DASM_HOOK_NONE && pre->kind != DASM_HOOK_ZERO_DST)
    throw std::runtime_error("ASPoissonPreconditioner::vmult: only the zeroing pre-operation is supported");
  DASM_REQUIRE((const void *)dst != (const void *)src, "vmult: dst and src must not alias");
  bool brick_path = fdm_uses_brick(f);
  if (brick_path && op->tma_only)
    {
      const Epilogue<T> epi = epilogue_from_hook<T>(post);
      brick_path            = tma_aligned(dst, epi.v0, epi.v1) && tma_aligned(src, nullptr, nullptr);
    }
  if (brick_path)
    {
      fdm_vmult_brick<T>(f, dst, src, post);
      return;
    }
  CUDA_CHECK(cudaMemsetAsync(dst, 0, (size_t)op->n_vec * sizeof(T), ctx->stream));
  Exchange &ex = f->use_ext ? op->exchange_ext : op->exchange; // overlapping patches reach into the cells of the neighbour ranks
  ex.run<T>(const_cast<T *>(src), false);
  launch_fdm<T>(f, dst, src);
  if (fdm_needs_compression(f))
    ex.run<T>(dst, true);
  // post hook without the constrained-DoF copy: the preconditioner leaves constrained DoFs at zero
  if (post != nullptr && post->kind != DASM_HOOK_NONE)
    {
      const long long n = op->n_owned;
      switch (post->kind)
        {
          case DASM_HOOK_RESIDUAL:
            vec_residual_kernel<T><<<nblocks(n), 256, 0, ctx->stream>>>(dst, (const T *)post->v0, n);
            break;
          case DASM_HOOK_CHEB_UPDATE:
            vec_cheb_update_kernel<T><<<nblocks(n), 256, 0, ctx->stream>>>(dst, dst, (const T *)post->v0, (const T *)post->v1, (T)post->f1, (T)post->f2, n);
            break;
          case DASM_HOOK_SCALE:
            vec_scale_kernel<T><<<nblocks(n), 256, 0, ctx->stream>>>(dst, dst, (T)post->f2, n);
            break;
          default:
            throw std::runtime_error("unsupported post hook kind");
        }
      ctx->launches++;
    }
}

// ------------------------------------------------------------------------------------------------
// Degrees 5 and 6: TMA-fed kernels only.  Applies when every mesh brick is a lex brick (full 4 x 4 x 4 bricks with neighbour cells
// across all faces: the periodic meshes of matrix_free_loop_08) on one rank: then there are no constrained DoFs, no ghost entries and
// no irregular bricks, so the brick kernels (instantiated up to degree 4 / 5) are not needed behind the TMA-fed kernels.
// ------------------------------------------------------------------------------------------------
static bool
setup_tma_only(dasm_op *op)
{
  const Mesh &M   = *op->mesh->mesh;
  dasm_ctx *  ctx = op->ctx;
  const int   k = op->k, n = k + 1;
  const size_t n_mesh_bricks = M.brick_ptr.size() - 1;
  if (M.n_ranks() != 1 || op->geom_mode != 0 || op->n_constrained != 0 || op->n_ghost != 0 || op->nb.n_lex == 0 ||
      (size_t)op->nb.n_lex != n_mesh_bricks || (size_t)op->n_cells != 64 * n_mesh_bricks)
    return false;
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, ctx->device));
  op->n_sm     = prop.multiProcessorCount;
  op->max_smem = (int)prop.sharedMemPerBlockOptin;
  if (tma_laplace_smem(k, (int)op->esize()) > (size_t)op->max_smem)
    return false;
  // 1-D matrices of the Kronecker form in even-odd form
  {
    std::vector<double> A(n * n);
    bool                eo = eo_pack_centrosymmetric(n, op->basis.M_ref.data(), op->lap_P[0], op->lap_Q[0]);
    for (int d = 0; d < 3; ++d)
      {
        for (int i = 0; i < n * n; ++i)
          A[i] = op->cart.g[d] * op->basis.K_ref[i];
        eo = eo_pack_centrosymmetric(n, A.data(), op->lap_P[1 + d], op->lap_Q[1 + d]) && eo;
      }
    if (!eo)
      return false;
  }
  const int                 R = 4 * k;
  std::vector<BrickDesc>    bricks(n_mesh_bricks);
  std::vector<uint32_t>     shared_list, fast_ids(n_mesh_bricks);
  std::map<long long, uint32_t> brick_at; // global coordinates of the first cell / 4 -> brick
  auto key = [&](const int c0, const int c1, const int c2) { return ((long long)(c2 / 4) * (M.p.nc[1] / 4 + 1) + c1 / 4) * (M.p.nc[0] / 4 + 1) + c0 / 4; };
  for (size_t b = 0; b < n_mesh_bricks; ++b)
    {
      const auto &c0 = M.cell_ijk[M.brick_ptr[b]];
      if (c0[0] % 4 != 0 || c0[1] % 4 != 0 || c0[2] % 4 != 0 || !op->nb.brick_lex[b])
        return false;
      brick_at[key(c0[0], c0[1], c0[2])] = (uint32_t)b;
    }
  op->h_tma.assign(n_mesh_bricks, TmaBrick());
  static const int off7[7][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}};
  for (size_t b = 0; b < n_mesh_bricks; ++b)
    {
      BrickDesc &bd = bricks[b];
      bd.first_cell = M.brick_ptr[b];
      bd.b[0] = bd.b[1] = bd.b[2] = 4;
      bd.shared   = 0x3F | BRICK_LEX;
      bd.base     = op->nb.brick_base[b];
      bd.sh_base  = bd.base;
      bd.sh_count = 0;
      bd.npriv    = 0;
      bd.variant  = 0xFFFFu;
      if (bd.base % (uint32_t)(64 * k * k * k) != 0)
        return false;
      fast_ids[b] = (uint32_t)b;
      // own DoFs on the lower faces of the box (shared with the lower neighbour bricks)
      for (int Z = 0; Z < R; ++Z)
        for (int Y = 0; Y < R; ++Y)
          for (int X = 0; X < R; ++X)
            if (X == 0 || Y == 0 || Z == 0)
              shared_list.push_back(bd.base + (uint32_t)(X + R * (Y + R * Z)));
      TmaBrick &t  = op->h_tma[b];
      t.base       = bd.base;
      t.flags      = 0;
      t.list_off   = 0;
      const auto &o = M.cell_ijk[bd.first_cell];
      for (int q = 0; q < 7; ++q)
        {
          int c2[3];
          for (int d = 0; d < 3; ++d)
            {
              c2[d] = o[d] + 4 * off7[q][d];
              if (c2[d] >= M.p.nc[d])
                c2[d] -= M.p.nc[d];
            }
          const auto it = brick_at.find(key(c2[0], c2[1], c2[2]));
          if (it == brick_at.end())
            return false;
          t.nb[q] = op->nb.brick_base[it->second];
        }
    }
  op->tma_only         = true;
  op->use_brick        = true;
  op->brick_bz         = 4;
  op->n_bricks         = (int)n_mesh_bricks;
  op->d_bricks         = dev_upload(bricks, ctx->stream);
  op->h_brick_boundary.assign(n_mesh_bricks, 0);
  op->n_shared         = (long long)shared_list.size();
  op->d_shared_list    = dev_upload(shared_list, ctx->stream);
  op->shared_ranges_ok = true;
  op->n_fast_boundary  = 0;
  op->tma_ok           = true; // (tma_build_list reads the block count of the kernels)
  const TmaChunked ch  = tma_build_list(op, fast_ids, 0);
  op->fast_ok          = true;
  op->tma_any_mode1    = ch.any_mode1;
  op->tma_any_mode1_interior = ch.any_mode1_interior;
  op->d_tma_lap        = dev_upload(ch.descs, ctx->stream);
  op->d_tma_lap_chunks = dev_upload(ch.chunk_start, ctx->stream);
  op->tma_lap_n_chunks = (int)ch.chunk_start.size() - 1;
  op->tma_lap_n_chunks_boundary = ch.n_chunks_boundary;
  std::vector<uint32_t> none(4, 0);
  op->d_tma_foreign    = dev_upload(none, ctx->stream);
  op->d_fast_ids       = dev_upload(fast_ids, ctx->stream);
  op->d_slow_ids       = dev_upload(none, ctx->stream);
  op->n_fast           = (int)fast_ids.size();
  op->n_slow           = 0;
  op->h_fast_ids       = fast_ids;
  CUDA_CHECK(cudaMalloc(&op->d_acc, std::max<size_t>(1, (size_t)op->n_vec) * op->esize()));
  CUDA_CHECK(cudaMemset(op->d_acc, 0, std::max<size_t>(1, (size_t)op->n_vec) * op->esize()));
  return true;
}

// ------------------------------------------------------------------------------------------------
// C ABI: context
// ------------------------------------------------------------------------------------------------
extern "C" const char *
dasm_last_error(void)
{
  return g_last_error.c_str();
}

extern "C" const char *
dasm_version(void)
{
  return "libdasm 0.1 (sm_100a)";
}

static void
upload_basis_tables()
{
  DevBasis<double> hd[9];
  DevBasis<float>  hf[9];
  memset(hd, 0, sizeof(hd));
  memset(hf, 0, sizeof(hf));
  for (int k = 1; k <= MAX_DEGREE; ++k)
    {
      Basis1D   b(k);
      const int n = k + 1;
      // Dn = Dq * N (derivative of the nodal basis at the Gauss points) equals b.D
      for (int i = 0; i < n * n; ++i)
        {
          hd[k].N[i]  = b.N[i];
          hd[k].Dq[i] = b.Dq[i];
          hd[k].Dn[i] = b.D[i];
          hf[k].N[i]  = (float)b.N[i];
          hf[k].Dq[i] = (float)b.Dq[i];
          hf[k].Dn[i] = (float)b.D[i];
        }
      for (int i = 0; i < n; ++i)
        {
          hd[k].qw[i] = b.qw[i];
          hf[k].qw[i] = (float)b.qw[i];
          hd[k].qp[i] = b.qp[i];
          hf[k].qp[i] = (float)b.qp[i];
        }
    }
  CUDA_CHECK(cudaMemcpyToSymbol(c_basis_d, hd, sizeof(hd)));
  CUDA_CHECK(cudaMemcpyToSymbol(c_basis_f, hf, sizeof(hf)));
}

extern "C" int
dasm_ctx_create(int device, dasm_ctx **out)
{
  DASM_API_BEGIN
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    throw std::runtime_error("libdasm needs a CUDA device (sm_100a); there is no CPU fallback");
  DASM_REQUIRE(device >= 0 && device < count, "invalid device ordinal");
  CUDA_CHECK(cudaSetDevice(device));
  auto ctx    = new dasm_ctx;
  ctx->device = device;
  CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
  CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_a, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_b, cudaEventDisableTiming));
  CUDA_CHECK(cudaMalloc(&ctx->d_partial, 1024 * sizeof(double)));
  CUDA_CHECK(cudaMalloc(&ctx->d_scalar, 8 * sizeof(double)));
  CUDA_CHECK(cudaMallocHost(&ctx->h_scalar, 8 * sizeof(double)));
  upload_basis_tables();
  *out = ctx;
  DASM_API_END
}

extern "C" int
dasm_ctx_destroy(dasm_ctx *ctx)
{
  DASM_API_BEGIN
  if (ctx == nullptr)
    return 0;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->comm)
    ncclCommDestroy(ctx->comm);
  cudaFree(ctx->d_partial);
  cudaFree(ctx->d_scalar);
  cudaFreeHost(ctx->h_scalar);
  cudaEventDestroy(ctx->ev_a);
  cudaEventDestroy(ctx->ev_b);
  cudaStreamDestroy(ctx->stream);
  cudaStreamDestroy(ctx->comm_stream);
  delete ctx;
  DASM_API_END
}

extern "C" int
dasm_ctx_sync(dasm_ctx *ctx)
{
  DASM_API_BEGIN
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  DASM_API_END
}

extern "C" long long
dasm_ctx_launch_count(const dasm_ctx *ctx)
{
  return ctx->launches;
}

extern "C" int
dasm_ctx_enable_kernel_timing(dasm_ctx *ctx, int on)
{
  DASM_API_BEGIN
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  ctx->timing  = on != 0;
  ctx->ev_used = 0;
  DASM_API_END
}

extern "C" int
dasm_ctx_kernel_time(dasm_ctx *ctx, int klass, double *ms, long long *count)
{
  DASM_API_BEGIN
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  double    t = 0;
  long long n = 0;
  for (size_t i = 0; i < ctx->ev_used; ++i)
    if (ctx->ev_class[i] == klass)
      {
        float e = 0;
        CUDA_CHECK(cudaEventElapsedTime(&e, ctx->ev_pool[i][0], ctx->ev_pool[i][1]));
        t += e;
        ++n;
      }
  *ms    = t;
  *count = n;
  DASM_API_END
}

extern "C" void *
dasm_ctx_stream(dasm_ctx *ctx)
{
  return (void *)ctx->stream;
}

extern "C" int
dasm_nccl_unique_id(void *id128)
{
  DASM_API_BEGIN
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  NCCL_CHECK(ncclGetUniqueId(&id));
  memcpy(id128, &id, 128);
  DASM_API_END
}

extern "C" int
dasm_ctx_comm_init(dasm_ctx *ctx, int n_ranks, int rank, const void *id128)
{
  DASM_API_BEGIN
  CUDA_CHECK(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  NCCL_CHECK(ncclCommInitRank(&ctx->comm, n_ranks, id, rank));
  ctx->n_ranks = n_ranks;
  ctx->rank    = rank;
  DASM_API_END
}

// ------------------------------------------------------------------------------------------------
// C ABI: mesh
// ------------------------------------------------------------------------------------------------
extern "C" int
dasm_decompose_balanced(int s, int *n_refine, int subdivisions[3])
{
  DASM_API_BEGIN
  // include/grid_generator.h:107-135
  int       nr        = s / 6;
  const int remainder = s % 6;
  int       sub[3]    = {1, 1, 1};
  if (remainder == 1 && s > 1)
    {
      sub[0] = 3;
      sub[1] = 2;
      sub[2] = 2;
      nr -= 1;
    }
  if (remainder == 2)
    sub[0] = 2;
  else if (remainder == 3)
    sub[0] = 3;
  else if (remainder == 4)
    sub[0] = sub[1] = 2;
  else if (remainder == 5)
    {
      sub[0] = 3;
      sub[1] = 2;
    }
  *n_refine = nr;
  for (int d = 0; d < 3; ++d)
    subdivisions[d] = sub[d];
  DASM_API_END
}

extern "C" int
dasm_mesh_create_structured(dasm_ctx *ctx, const int n_cells[3], const int periodic[3], int dirichlet, const double length[3],
                            int map_kind, const double map_params[4], const int partition[3], int rank, dasm_mesh **out)
{
  DASM_API_BEGIN
  MeshParams p;
  for (int d = 0; d < 3; ++d)
    {
      DASM_REQUIRE(n_cells[d] >= 1, "n_cells must be positive");
      p.nc[d]       = n_cells[d];
      p.periodic[d] = periodic ? periodic[d] : 0;
      p.length[d]   = length ? length[d] : 1.0;
      p.part[d]     = partition ? partition[d] : 1;
      DASM_REQUIRE(p.part[d] >= 1 && p.part[d] <= p.nc[d], "invalid partition");
      if (p.periodic[d] && p.part[d] > 1)
        DASM_REQUIRE(p.nc[d] / p.part[d] >= 2, "periodic partitioned direction needs >= 2 cells per rank");
    }
  p.dirichlet = dirichlet;
  p.map_kind  = map_kind;
  DASM_REQUIRE(map_kind >= 0 && map_kind <= 2, "unknown map kind");
  if (map_params)
    for (int i = 0; i < 4; ++i)
      p.map_par[i] = map_params[i];
  p.rank = rank;
  DASM_REQUIRE(rank >= 0 && rank < p.part[0] * p.part[1] * p.part[2], "rank outside partition");
  auto m  = new dasm_mesh;
  m->ctx  = ctx;
  m->mesh = std::make_unique<Mesh>(p);
  *out    = m;
  DASM_API_END
}

extern "C" int
dasm_mesh_host_numbering(const dasm_mesh *mesh, int degree, long long sizes[6], unsigned int *cidx_plain, int *peers, long long *send_count,
                         long long *recv_count, unsigned int *send_idx, unsigned int *recv_idx)
{
  DASM_API_BEGIN
  DASM_REQUIRE(mesh != nullptr && degree >= 1 && degree <= MAX_DEGREE, "invalid arguments");
  const Mesh::Numbering nb = mesh->mesh->number_dofs(degree);
  long long             ns = 0, nr = 0;
  for (const ExchangeList &l : nb.exchange)
    {
      ns += (long long)l.n_send;
      nr += (long long)l.n_recv;
    }
  if (sizes)
    {
      sizes[0] = nb.n_owned;
      sizes[1] = nb.n_ghost;
      sizes[2] = (long long)mesh->mesh->n_cells;
      sizes[3] = (long long)nb.exchange.size();
      sizes[4] = ns;
      sizes[5] = nr;
    }
  if (cidx_plain)
    std::copy(nb.cidx_plain.begin(), nb.cidx_plain.end(), cidx_plain);
  size_t so = 0, ro = 0;
  for (size_t p = 0; p < nb.exchange.size(); ++p)
    {
      const ExchangeList &l = nb.exchange[p];
      if (peers)
        peers[p] = l.peer;
      if (send_count)
        send_count[p] = (long long)l.n_send;
      if (recv_count)
        recv_count[p] = (long long)l.n_recv;
      if (send_idx)
        for (size_t r = 0; r < l.send_start.size(); ++r)
          for (uint32_t i = 0; i < l.send_len[r]; ++i)
            send_idx[so++] = l.send_start[r] + i;
      if (recv_idx)
        for (size_t r = 0; r < l.recv_start.size(); ++r)
          for (uint32_t i = 0; i < l.recv_len[r]; ++i)
            recv_idx[ro++] = l.recv_start[r] + i;
    }
  DASM_API_END
}

// Host-only view of the enlarged ghost layout (Mesh::halo_numbering): halo cells, their index rows and the exchange lists of ALL
// ghosts.  sizes = {n_owned, n_ghost (old + new), n_local_cells, n_halo_cells, n_peers, n_send_total, n_recv_total}; cidx_plain has
// (n_local_cells + n_halo_cells) * 27 entries (local cells first).
extern "C" int
dasm_mesh_host_halo_numbering(const dasm_mesh *mesh, int degree, long long sizes[7], int *halo_cells, unsigned int *cidx_plain, int *peers,
                              long long *send_count, long long *recv_count, unsigned int *send_idx, unsigned int *recv_idx)
{
  DASM_API_BEGIN
  DASM_REQUIRE(mesh != nullptr && degree >= 1 && degree <= MAX_DEGREE, "invalid arguments");
  const Mesh::Numbering     nb = mesh->mesh->number_dofs(degree);
  const Mesh::HaloNumbering h  = mesh->mesh->halo_numbering(nb);
  long long                 ns = 0, nr = 0;
  for (const ExchangeList &l : h.exchange)
    {
      ns += (long long)l.n_send;
      nr += (long long)l.n_recv;
    }
  if (sizes)
    {
      sizes[0] = nb.n_owned;
      sizes[1] = (long long)nb.n_ghost + h.n_ghost_ext;
      sizes[2] = (long long)mesh->mesh->n_cells;
      sizes[3] = (long long)h.cells.size();
      sizes[4] = (long long)h.exchange.size();
      sizes[5] = ns;
      sizes[6] = nr;
    }
  if (halo_cells)
    for (size_t i = 0; i < h.cells.size(); ++i)
      for (int d = 0; d < 3; ++d)
        halo_cells[3 * i + d] = h.cells[i][d];
  if (cidx_plain)
    {
      std::copy(nb.cidx_plain.begin(), nb.cidx_plain.end(), cidx_plain);
      std::copy(h.cidx_plain.begin(), h.cidx_plain.end(), cidx_plain + nb.cidx_plain.size());
    }
  size_t so = 0, ro = 0;
  for (size_t p = 0; p < h.exchange.size(); ++p)
    {
      const ExchangeList &l = h.exchange[p];
      if (peers)
        peers[p] = l.peer;
      if (send_count)
        send_count[p] = (long long)l.n_send;
      if (recv_count)
        recv_count[p] = (long long)l.n_recv;
      if (send_idx)
        for (size_t r = 0; r < l.send_start.size(); ++r)
          for (uint32_t i = 0; i < l.send_len[r]; ++i)
            send_idx[so++] = l.send_start[r] + i;
      if (recv_idx)
        for (size_t r = 0; r < l.recv_start.size(); ++r)
          for (uint32_t i = 0; i < l.recv_len[r]; ++i)
            recv_idx[ro++] = l.recv_start[r] + i;
    }
  DASM_API_END
}

extern "C" int
dasm_test_eo_pack(int n, int kind, const double *A, double *P, double *Q)
{
  DASM_API_BEGIN
  DASM_REQUIRE(n >= 2 && n <= 9 && kind >= 0 && kind <= 2 && A && P && Q, "invalid arguments");
  const bool ok = kind == 0 ? eo_pack_centrosymmetric(n, A, P, Q) : (kind == 1 ? eo_pack_forward(n, A, P, Q) : eo_pack_backward(n, A, P, Q));
  DASM_REQUIRE(ok, "matrix does not have the even-odd structure");
  DASM_API_END
}

extern "C" int
dasm_mesh_destroy(dasm_mesh *mesh)
{
  delete mesh;
  return 0;
}

extern "C" long long
dasm_mesh_n_cells(const dasm_mesh *mesh)
{
  return (long long)mesh->mesh->n_cells;
}

extern "C" long long
dasm_mesh_n_global_cells(const dasm_mesh *mesh)
{
  const auto &p = mesh->mesh->p;
  return (long long)p.nc[0] * p.nc[1] * p.nc[2];
}

extern "C" int
dasm_mesh_cell_coordinates(const dasm_mesh *mesh, int *out)
{
  for (size_t i = 0; i < mesh->mesh->n_cells; ++i)
    for (int d = 0; d < 3; ++d)
      out[3 * i + d] = mesh->mesh->cell_ijk[i][d];
  return 0;
}

// ------------------------------------------------------------------------------------------------
// C ABI: operator
// ------------------------------------------------------------------------------------------------
extern "C" int
dasm_op_create(dasm_mesh *mesh, int degree, int number_type, const char *mapping_type, int compress_indices, dasm_op **out)
{
  DASM_API_BEGIN
  DASM_REQUIRE(degree >= 1 && degree <= MAX_DEGREE, "degree must be in 1..8");
  DASM_REQUIRE(number_type == DASM_F64 || number_type == DASM_F32, "unknown number type");
  const std::string mt = mapping_type ? mapping_type : "";
  if (mt != "" && mt != "merged" && mt != "quadratic geometry" && mt != "linear geometry" && mt != "construct q")
    throw std::runtime_error("Mapping type <" + mt + "> is not known!"); // operator.h:747-752
  dasm_ctx *ctx = mesh->ctx;
  CUDA_CHECK(cudaSetDevice(ctx->device));
  auto op              = new dasm_op(degree);
  op->ctx              = ctx;
  op->mesh             = mesh;
  op->k                = degree;
  op->ntype            = number_type;
  op->compress_indices = compress_indices != 0;
  op->mapping_type     = mt;
  const Mesh &M        = *mesh->mesh;
  op->nb               = M.number_dofs(degree);
  op->n_cells          = (long long)M.n_cells;
  op->n_owned          = op->nb.n_owned;
  op->n_ghost          = op->nb.n_ghost;
  if (M.n_ranks() > 1 && !(getenv("DASM_NO_HALO") && getenv("DASM_NO_HALO")[0] == '1'))
    {
      op->halo        = M.halo_numbering(op->nb);
      op->n_ghost_ext = op->halo.n_ghost_ext;
    }
  op->n_vec            = op->n_owned + op->n_ghost + op->n_ghost_ext;
  {
    long long g = 1;
    for (int d = 0; d < 3; ++d)
      g *= (long long)M.p.nc[d] * degree + (M.p.periodic[d] ? 0 : 1);
    op->n_global_dofs = g;
  }
  op->d_cidx        = dev_upload(op->nb.cidx, ctx->stream);
  op->d_constrained = dev_upload(op->nb.constrained, ctx->stream);
  op->n_constrained = (long long)op->nb.constrained.size();
  op->exchange.init(ctx, op->nb.exchange, op->esize());
  if (!op->halo.exchange.empty())
    op->exchange_ext.init(ctx, op->halo.exchange, op->esize());
  // geometry
  const int n3 = (degree + 1) * (degree + 1) * (degree + 1);
  if (M.is_cartesian() && mt == "")
    {
      op->geom_mode   = 0;
      const double h0 = M.h(0), h1 = M.h(1), h2 = M.h(2), det = h0 * h1 * h2;
      op->cart.g[0] = det / (h0 * h0);
      op->cart.g[1] = det / (h1 * h1);
      op->cart.g[2] = det / (h2 * h2);
    }
  else
    {
      op->geom_mode       = (mt == "quadratic geometry" || mt == "linear geometry") ? 2 : (mt == "construct q" ? 3 : 1);
      op->linear_geometry = (mt == "linear geometry");
      const bool lingeo   = op->linear_geometry;
      if (op->geom_mode == 3)
        {
          // "construct q" (operator.h:712-746): 3 n^3 coordinates of the quadrature points per cell; the Jacobians are rebuilt in
          // the kernels by collocation differentiation (generic kernels; the brick kernels have no such geometry mode)
          std::vector<double> xq((size_t)op->n_cells * 3 * n3);
          for (long long c = 0; c < op->n_cells; ++c)
            {
              const int cc[3] = {M.cell_ijk[c][0], M.cell_ijk[c][1], M.cell_ijk[c][2]};
              M.quadrature_points(cc, op->basis, xq.data() + (size_t)c * 3 * n3);
            }
          if (number_type == DASM_F64)
            op->d_geom = dev_upload(xq, ctx->stream);
          else
            {
              std::vector<float> xf(xq.begin(), xq.end());
              op->d_geom = dev_upload(xf, ctx->stream);
            }
          op->cart.g[0] = op->cart.g[1] = op->cart.g[2] = 0;
        }
      else
        {
      if (op->geom_mode == 2)
        {
          std::vector<double> qc((size_t)op->n_cells * 81);
          for (long long c = 0; c < op->n_cells; ++c)
            {
              const int cc[3] = {M.cell_ijk[c][0], M.cell_ijk[c][1], M.cell_ijk[c][2]};
              M.quadratic_coefficients(cc, lingeo, qc.data() + (size_t)c * 81);
            }
          if (number_type == DASM_F64)
            op->d_qcoef = dev_upload(qc, ctx->stream);
          else
            {
              std::vector<float> qf(qc.begin(), qc.end());
              op->d_qcoef = dev_upload(qf, ctx->stream);
            }
        }
      std::vector<double> tmp(6 * n3);
      if (number_type == DASM_F64)
        {
          std::vector<double> g((size_t)op->n_cells * 6 * n3);
          for (long long c = 0; c < op->n_cells; ++c)
            {
              const int cc[3] = {M.cell_ijk[c][0], M.cell_ijk[c][1], M.cell_ijk[c][2]};
              M.merged_coefficients(cc, op->basis, g.data() + (size_t)c * 6 * n3, lingeo);
            }
          op->d_geom = dev_upload(g, ctx->stream);
        }
      else
        {
          std::vector<float> g((size_t)op->n_cells * 6 * n3);
          for (long long c = 0; c < op->n_cells; ++c)
            {
              const int cc[3] = {M.cell_ijk[c][0], M.cell_ijk[c][1], M.cell_ijk[c][2]};
              M.merged_coefficients(cc, op->basis, tmp.data(), lingeo);
              for (int i = 0; i < 6 * n3; ++i)
                g[(size_t)c * 6 * n3 + i] = (float)tmp[i];
            }
          op->d_geom = dev_upload(g, ctx->stream);
        }
      op->cart.g[0] = op->cart.g[1] = op->cart.g[2] = 0;
        }
    }
  // tuned brick path: degrees 1..5, one rank (the multi-rank path uses the generic kernels for now)
  {
    const char *force = getenv("DASM_FORCE_GENERIC");
    // (degree 5 runs faster through the generic kernels: 1.70e10 vs 1.24e10 DoFs/s per Chebyshev term, profiles/r01c_secondary.log;
    // DASM_BRICK_K5=1 selects the 4x4x2 brick kernels)
    const char *k5    = getenv("DASM_BRICK_K5");
    op->use_brick     = (degree <= 4 || (degree == 5 && k5 && k5[0] == '1')) && !((force && force[0] == '1') || deterministic_env()) && op->geom_mode != 3 &&
                    op->compress_indices;
    if (!op->compress_indices)
      {
        // plain index storage: (k+1)^3 indices per cell read by the generic kernels (the compressed / tile formats of the tuned kernels
        // are what compress_indices = true stands for)
        const long long ne = op->n_cells * n3;
        op->d_plain        = dev_alloc<uint32_t>(std::max<long long>(ne, 1));
        DISPATCH_DEGREE(degree, expand_compressed_kernel<K><<<nblocks(ne), 256, 0, ctx->stream>>>(op->d_plain, op->d_cidx, op->n_cells));
        ctx->launches++;
      }
    const char *nofast_hi = getenv("DASM_NO_FAST");
    if (!op->use_brick && op->compress_indices && (degree == 5 || degree == 6) && !((force && force[0] == '1') || deterministic_env()) && !(nofast_hi && nofast_hi[0] == '1'))
      setup_tma_only(op);
    if (op->use_brick && !op->tma_only)
      {
        op->brick_bz = (degree <= 4) ? 4 : 2;
        if (const char *bz = getenv("DASM_BRICK_BZ"))
          if (degree == 4 && (bz[0] == '2' || bz[0] == '4'))
            op->brick_bz = bz[0] - '0';
        std::vector<BrickDesc> bricks;
        // mesh bricks are 4x4x4 boxes of consecutive cells (x fastest); kernel bricks are z-slabs of them
        const int *B = M.p.brick;
        size_t     first = 0, mesh_brick = 0;
        int        ibz_base = 0, nbz_total = 0;
        for (int bz = 0; bz < M.nl[2]; bz += B[2])
          {
            const int dz      = std::min(B[2], M.nl[2] - bz);
            const int n_slabs = (dz + op->brick_bz - 1) / op->brick_bz;
            for (int by = 0; by < M.nl[1]; by += B[1])
              for (int bx = 0; bx < M.nl[0]; bx += B[0], ++mesh_brick)
                {
                  const int dx = std::min(B[0], M.nl[0] - bx), dy = std::min(B[1], M.nl[1] - by);
                  for (int z0 = 0, sl = 0; z0 < dz; z0 += op->brick_bz, ++sl)
                    {
                      const int sz = std::min(op->brick_bz, dz - z0);
                      BrickDesc bd;
                      bd.first_cell = (uint32_t)(first + (size_t)z0 * dx * dy);
                      bd.b[0]       = dx;
                      bd.b[1]       = dy;
                      bd.b[2]       = sz;
                      bd.shared     = 0;
                      bd.sh_base    = 0;
                      bd.sh_count   = 0;
                      bd.base       = 0;
                      bd.npriv      = 0;
                      bd.variant    = 0xFFFFu;
                      const int lo_c[3] = {M.lo[0] + bx, M.lo[1] + by, M.lo[2] + bz + z0};
                      const int hi_c[3] = {lo_c[0] + dx - 1, lo_c[1] + dy - 1, lo_c[2] + sz - 1};
                      for (int d = 0; d < 3; ++d)
                        {
                          int nbc[3];
                          if (M.neighbor(lo_c, d, 0, nbc))
                            bd.shared |= (1u << (2 * d));
                          if (M.neighbor(hi_c, d, 1, nbc))
                            bd.shared |= (1u << (2 * d + 1));
                        }
                      if (op->brick_bz == 4 && op->nb.brick_lex[mesh_brick])
                        bd.shared |= BRICK_LEX; // lexicographic box numbering (mesh.h)
                      bricks.push_back(bd);
                      {
                        // on the boundary of this rank's box in a partitioned direction?
                        const int bl[3] = {bx, by, bz + z0}, bh[3] = {bx + dx, by + dy, bz + z0 + sz};
                        bool      onb   = false;
                        for (int d = 0; d < 3; ++d)
                          if (M.p.part[d] > 1 && (bl[d] == 0 || bh[d] == M.nl[d]))
                            onb = true;
                        op->h_brick_boundary.push_back(onb ? 1 : 0);
                      }
                    }
                  first += (size_t)dx * dy * dz;
                }
            ibz_base += n_slabs;
            nbz_total += n_slabs;
          }
        op->n_bricks = (int)bricks.size();
        op->d_bricks = dev_upload(bricks, ctx->stream);
        // list of the DoFs on shared brick faces: every brick contributes the shared entities it owns
        // (lower entities of its cells, or upper entities on a domain-boundary face) lying on a
        // shared lower face
        {
          std::vector<uint32_t> list;
          op->shared_ranges_ok = true;
          for (BrickDesc &bd : bricks)
            {
              const size_t list_begin = list.size();
              const int    ncells = bd.b[0] * bd.b[1] * bd.b[2];
              for (int c = 0; c < ncells; ++c)
                {
                  const int cc[3] = {c % bd.b[0], (c / bd.b[0]) % bd.b[1], c / (bd.b[0] * bd.b[1])};
                  if (cc[0] != 0 && cc[1] != 0 && cc[2] != 0)
                    continue;
                  for (int e = 0; e < 27; ++e)
                    {
                      const int ee[3] = {e % 3, (e / 3) % 3, e / 9};
                      bool      owned = true, shared = false;
                      for (int d = 0; d < 3; ++d)
                        {
                          if (ee[d] == 2 && !(cc[d] == bd.b[d] - 1 && !((bd.shared >> (2 * d + 1)) & 1u)))
                            owned = false;
                          if (ee[d] == 0 && cc[d] == 0 && ((bd.shared >> (2 * d)) & 1u))
                            shared = true;
                        }
                      if (!owned || !shared)
                        continue;
                      const uint32_t *ci = op->nb.cidx.data() + (size_t)(bd.first_cell + c) * 27;
                      if (ci[e] == INVALID_INDEX)
                        continue;
                      // all DoFs of the entity (local coordinates 0, 1..k-1 or k per direction)
                      int lo3[3], hi3[3];
                      for (int d = 0; d < 3; ++d)
                        {
                          lo3[d] = ee[d] == 0 ? 0 : (ee[d] == 2 ? degree : 1);
                          hi3[d] = ee[d] == 1 ? degree - 1 : lo3[d];
                        }
                      for (int z = lo3[2]; z <= hi3[2]; ++z)
                        for (int y = lo3[1]; y <= hi3[1]; ++y)
                          for (int x = lo3[0]; x <= hi3[0]; ++x)
                            list.push_back(expand_start_index(ci, degree, x, y, z));
                    }
                }
              // contiguous range of the brick's own shared DoFs (holds when the kernel brick is a mesh brick)
              bd.sh_count = (uint32_t)(list.size() - list_begin);
              if (bd.sh_count > 0)
                {
                  uint32_t mn = 0xFFFFFFFFu, mx = 0;
                  for (size_t i = list_begin; i < list.size(); ++i)
                    {
                      mn = std::min(mn, list[i]);
                      mx = std::max(mx, list[i]);
                    }
                  bd.sh_base = mn;
                  if (mx - mn + 1 != bd.sh_count && !(bd.shared & BRICK_LEX))
                    op->shared_ranges_ok = false;
                }
            }
          cudaFree(op->d_bricks);
          op->d_bricks = dev_upload(bricks, ctx->stream);
          std::sort(list.begin(), list.end());
          list.erase(std::unique(list.begin(), list.end()), list.end());
          op->n_shared      = (long long)list.size();
          op->d_shared_list = dev_upload(list, ctx->stream);
        }
        // ---- per-variant tile maps for the coalesced gather / store (kernel brick == mesh brick only)
        op->maps.load_tab = nullptr;
        if (op->brick_bz == 4)
          {
            const int k = degree, n = k + 1, TX = 4 * k + 1, TY = 4 * k + 1, NPTS = TX * TY * TX;
            const int CS  = (n * n * n) | 1;
            const int NFP = (NPTS - (4 * k) * (4 * k) * (4 * k) + 3) / 4 * 4;
            struct Variant
            {
              std::vector<uint16_t> load;       // [NPTS]
              std::vector<uint32_t> store;      // sorted by class
              std::vector<uint32_t> store_off;  // [17]
              std::vector<uint32_t> ftab;       // sorted by mask
              std::vector<uint32_t> for_off;    // [9]
              std::vector<uint16_t> sh_tab;     // offsets (in the own range) of the brick's own DoFs on shared lower faces
              uint32_t              flags = 0;
              bool operator==(const Variant &o) const
              {
                return load == o.load && store == o.store && store_off == o.store_off && ftab == o.ftab && for_off == o.for_off &&
                       sh_tab == o.sh_tab && flags == o.flags;
              }
            };
            std::map<uint32_t, uint16_t> variant_of; // signature -> variant
            std::vector<Variant>         variants;
            std::vector<uint32_t>        foreign_gidx((size_t)bricks.size() * NFP, INVALID_INDEX);
            bool                         ok = true;
            auto plain_gidx = [&](const BrickDesc &bd, const std::vector<uint32_t> &tab, int px, int py, int pz) {
              const int cx = std::min(px / k, bd.b[0] - 1), cy = std::min(py / k, bd.b[1] - 1), cz = std::min(pz / k, bd.b[2] - 1);
              return expand_start_index(tab.data() + (size_t)(bd.first_cell + (cz * bd.b[1] + cy) * bd.b[0] + cx) * 27, k, px - cx * k,
                                        py - cy * k, pz - cz * k);
            };
            unsigned verify_counter = 0;
            for (size_t bidx = 0; bidx < bricks.size() && ok; ++bidx)
              {
                BrickDesc &    bd  = bricks[bidx];
                const uint32_t sig = bd.b[0] | (bd.b[1] << 4) | (bd.b[2] << 8) | ((uint32_t)bd.shared << 12);
                const int      e3[3] = {bd.b[0] * k + 1, bd.b[1] * k + 1, bd.b[2] * k + 1};
                auto owned_point = [&](const int pp[3], bool &on_shared_lo) {
                  on_shared_lo = false;
                  for (int d = 0; d < 3; ++d)
                    {
                      if (pp[d] == e3[d] - 1 && ((bd.shared >> (2 * d + 1)) & 1u))
                        return false; // upper face shared: owned by the neighbour brick
                      if (pp[d] == 0 && ((bd.shared >> (2 * d)) & 1u))
                        on_shared_lo = true;
                    }
                  return true;
                };
                // pass 1: ranges of owned DoFs (plain indices: constrained DoFs included)
                uint32_t mn = 0xFFFFFFFFu, mx = 0, cnt = 0, smn = 0xFFFFFFFFu, smx = 0, scnt = 0;
                for (int pz = 0; pz < e3[2]; ++pz)
                  for (int py = 0; py < e3[1]; ++py)
                    for (int px = 0; px < e3[0]; ++px)
                      {
                        const int pp[3] = {px, py, pz};
                        bool      shlo;
                        if (!owned_point(pp, shlo))
                          continue;
                        const uint32_t g = plain_gidx(bd, op->nb.cidx_plain, px, py, pz);
                        mn               = std::min(mn, g);
                        mx               = std::max(mx, g);
                        ++cnt;
                        if (shlo)
                          {
                            smn = std::min(smn, g);
                            smx = std::max(smx, g);
                            ++scnt;
                          }
                      }
                const bool lexb = (bd.shared & BRICK_LEX) != 0;
                if (cnt == 0 || mx - mn + 1 != cnt || (!lexb && scnt > 0 && (smx - smn + 1 != scnt || smx != mx)) || cnt >= 0x1FFFu)
                  {
                    ok = false;
                    break;
                  }
                bd.base     = mn;
                bd.npriv    = (uint16_t)(cnt - scnt);
                bd.sh_base  = scnt > 0 ? smn : mn + cnt;
                bd.sh_count = scnt;
                // pass 2: variant tables for the first brick of a signature (verified on every 97th brick), foreign
                // indices for every brick
                const bool first_of_sig = variant_of.find(sig) == variant_of.end();
                const bool verify       = !first_of_sig && ((++verify_counter) % 97 == 0);
                Variant    var;
                struct E
                {
                  uint32_t cls, key, entry;
                };
                std::vector<E> st_entries, f_entries;
                std::vector<uint32_t> f_gidx_by_p; // parallel to f_entries before sorting
                var.load.assign(NPTS, 0xFFFFu);
                for (int pz = 0; pz < e3[2]; ++pz)
                  for (int py = 0; py < e3[1]; ++py)
                    for (int px = 0; px < e3[0]; ++px)
                      {
                        const int pp[3] = {px, py, pz};
                        int       cc[3], ll[3];
                        unsigned  mask = 0;
                        for (int d = 0; d < 3; ++d)
                          {
                            const int ch = pp[d] / k, l = pp[d] - ch * k;
                            if (ch < bd.b[d])
                              {
                                cc[d] = ch;
                                ll[d] = l;
                                if (l == 0 && ch > 0)
                                  mask |= 1u << d;
                              }
                            else
                              {
                                cc[d] = ch - 1;
                                ll[d] = k;
                              }
                          }
                        const uint32_t o    = (uint32_t)(((cc[2] * bd.b[1] + cc[1]) * bd.b[0] + cc[0]) * CS + (ll[2] * n + ll[1]) * n + ll[0]);
                        const uint32_t plin = (uint32_t)((pz * TY + py) * TX + px);
                        const uint32_t gv   = plain_gidx(bd, op->nb.cidx, px, py, pz);
                        bool           shlo;
                        if (owned_point(pp, shlo))
                          {
                            if (gv != INVALID_INDEX)
                              {
                                const uint32_t i = gv - mn;
                                var.load[i]      = (uint16_t)plin;
                                st_entries.push_back({(shlo ? 8u : 0u) + mask, i, i | (o << 13) | (mask << 26) | (shlo ? STORE_SHARED : 0u)});
                              }
                            else
                              var.flags |= 1u;
                          }
                        else
                          {
                            f_entries.push_back({mask, plin, plin | (o << 13) | (mask << 26)});
                            f_gidx_by_p.push_back(gv);
                            if (gv == INVALID_INDEX)
                              var.flags |= 1u;
                          }
                      }
                // foreign entries sorted by (mask, tile point); the per-brick index list follows the same order
                std::vector<size_t> perm(f_entries.size());
                for (size_t i = 0; i < perm.size(); ++i)
                  perm[i] = i;
                std::stable_sort(perm.begin(), perm.end(), [&](size_t a, size_t b2) { return f_entries[a].cls < f_entries[b2].cls; });
                if ((int)perm.size() > NFP)
                  {
                    ok = false;
                    break;
                  }
                for (size_t j = 0; j < perm.size(); ++j)
                  foreign_gidx[bidx * NFP + j] = f_gidx_by_p[perm[j]];
                if (first_of_sig || verify)
                  {
                    // store table in the order of the own range (entry i describes DoF base + i)
                    var.store_off.assign(17, 0);
                    var.store.assign(cnt, 0xFFFFFFFFu);
                    for (const E &x : st_entries)
                      var.store[x.key] = x.entry;
                    for (uint32_t i = 0; i < cnt; ++i)
                      if (var.store[i] != 0xFFFFFFFFu && (var.store[i] & STORE_SHARED))
                        var.sh_tab.push_back((uint16_t)i);
                    if (lexb)
                      var.flags |= 2u; // the private / shared DoFs are interleaved in the own range
                    var.for_off.assign(9, 0);
                    for (size_t j = 0; j < perm.size(); ++j)
                      {
                        var.ftab.push_back(f_entries[perm[j]].entry);
                        var.for_off[f_entries[perm[j]].cls + 1]++;
                      }
                    for (int c = 0; c < 8; ++c)
                      var.for_off[c + 1] += var.for_off[c];
                    if (first_of_sig)
                      {
                        variant_of[sig] = (uint16_t)variants.size();
                        variants.push_back(var);
                      }
                    else if (!(var == variants[variant_of[sig]]))
                      {
                        ok = false;
                        break;
                      }
                  }
                bd.variant = variant_of[sig];
              }
            if (ok && !variants.empty())
              {
                const int             nv = (int)variants.size();
                std::vector<uint16_t> h_load((size_t)nv * NPTS, 0xFFFFu);
                std::vector<uint32_t> h_store((size_t)nv * NPTS, 0xFFFFFFFFu), h_soff((size_t)nv * 17), h_for((size_t)nv * NFP, 0), h_foff((size_t)nv * 9),
                  h_fl(nv);
                for (int v = 0; v < nv; ++v)
                  {
                    std::copy(variants[v].load.begin(), variants[v].load.end(), h_load.begin() + (size_t)v * NPTS);
                    std::copy(variants[v].store.begin(), variants[v].store.end(), h_store.begin() + (size_t)v * NPTS);
                    std::copy(variants[v].store_off.begin(), variants[v].store_off.end(), h_soff.begin() + (size_t)v * 17);
                    std::copy(variants[v].ftab.begin(), variants[v].ftab.end(), h_for.begin() + (size_t)v * NFP);
                    std::copy(variants[v].for_off.begin(), variants[v].for_off.end(), h_foff.begin() + (size_t)v * 9);
                    h_fl[v] = variants[v].flags;
                  }
                // own shared DoFs per variant (pre-initialisation of the next kernel's destination on lex bricks)
                const int             NSH = NPTS - (4 * k - 1) * (4 * k - 1) * (4 * k - 1);
                std::vector<uint16_t> h_sh((size_t)nv * NSH, 0);
                std::vector<uint32_t> h_shc(nv, 0);
                for (int v = 0; v < nv; ++v)
                  {
                    h_shc[v] = (uint32_t)variants[v].sh_tab.size();
                    if ((int)h_shc[v] > NSH)
                      throw std::runtime_error("internal: shared table overflow");
                    std::copy(variants[v].sh_tab.begin(), variants[v].sh_tab.end(), h_sh.begin() + (size_t)v * NSH);
                  }
                op->d_map_bufs[7] = dev_upload(h_sh, ctx->stream);
                op->d_map_bufs[8] = dev_upload(h_shc, ctx->stream);
                op->maps.sh_tab   = (const uint16_t *)op->d_map_bufs[7];
                op->maps.sh_cnt   = (const uint32_t *)op->d_map_bufs[8];
                op->maps.sh_stride = NSH;
                op->d_map_bufs[0]      = dev_upload(h_load, ctx->stream);
                op->d_map_bufs[1]      = dev_upload(h_store, ctx->stream);
                op->d_map_bufs[2]      = dev_upload(h_soff, ctx->stream);
                op->d_map_bufs[3]      = dev_upload(h_for, ctx->stream);
                op->d_map_bufs[4]      = dev_upload(h_foff, ctx->stream);
                op->d_map_bufs[5]      = dev_upload(h_fl, ctx->stream);
                op->d_map_bufs[6]      = dev_upload(foreign_gidx, ctx->stream);
                op->maps.load_tab      = (const uint16_t *)op->d_map_bufs[0];
                op->maps.store_tab     = (const uint32_t *)op->d_map_bufs[1];
                op->maps.store_off     = (const uint32_t *)op->d_map_bufs[2];
                op->maps.for_tab       = (const uint32_t *)op->d_map_bufs[3];
                op->maps.for_off       = (const uint32_t *)op->d_map_bufs[4];
                op->maps.flags         = (const uint32_t *)op->d_map_bufs[5];
                op->maps.foreign_gidx  = (const uint32_t *)op->d_map_bufs[6];
                op->maps.stride        = NPTS;
                op->maps.nfp           = NFP;
                op->shared_ranges_ok   = true;
                cudaFree(op->d_bricks);
                op->d_bricks = dev_upload(bricks, ctx->stream);
                // ---- warp-specialised kernels: tables of the regular variant (full brick, six shared faces, no
                // constrained DoFs) in the skewed tile layout of kernels_fast.cuh, and the lists of regular / other bricks
                const char *nofast = getenv("DASM_NO_FAST");
                if (degree >= 2 && degree <= 4 && !(nofast && nofast[0] == '1'))
                  {
                    // ---- TMA-fed kernels (kernels_tma.cuh): every lex brick with its 7 upper neighbours
                    if (op->nb.n_lex > 0)
                      {
                        const int nbx = (M.nl[0] + 3) / 4, nby = (M.nl[1] + 3) / 4;
                        const int R   = 4 * k;
                        op->h_tma.assign(bricks.size(), TmaBrick());
                        std::vector<uint32_t> foreign;
                        std::vector<uint32_t> fast_ids, slow_ids;
                        static const int      off7[7][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}};
                        const int             NFORt = 3 * R * R + 3 * R + 1, NFPt = (NFORt + 3) / 4 * 4;
                        for (size_t b = 0; b < bricks.size(); ++b)
                          {
                            const BrickDesc &bd = bricks[b];
                            if (!(bd.shared & BRICK_LEX))
                              {
                                slow_ids.push_back((uint32_t)b);
                                continue;
                              }
                            fast_ids.push_back((uint32_t)b);
                            TmaBrick &t  = op->h_tma[b];
                            t.base       = bd.base;
                            t.flags      = 0;
                            t.list_off   = 0;
                            const auto &o = M.cell_ijk[bd.first_cell];
                            for (int q = 0; q < 7; ++q)
                              {
                                int  c2[3];
                                bool local = true;
                                for (int d = 0; d < 3; ++d)
                                  {
                                    c2[d] = o[d] + 4 * off7[q][d];
                                    if (c2[d] >= M.p.nc[d])
                                      c2[d] -= M.p.nc[d]; // periodic wrap (a lex brick has neighbours across all faces)
                                    if (c2[d] < M.lo[d] || c2[d] >= M.hi[d])
                                      local = false;
                                  }
                                t.nb[q] = 0;
                                if (local)
                                  {
                                    const size_t nbk = ((size_t)((c2[2] - M.lo[2]) / 4) * nby + (c2[1] - M.lo[1]) / 4) * nbx + (c2[0] - M.lo[0]) / 4;
                                    if ((bricks[nbk].shared & BRICK_LEX) && (c2[0] - M.lo[0]) % 4 == 0 && (c2[1] - M.lo[1]) % 4 == 0 &&
                                        (c2[2] - M.lo[2]) % 4 == 0)
                                      t.nb[q] = bricks[nbk].base;
                                    else
                                      local = false;
                                  }
                                if (!local)
                                  t.flags = TMA_MODE1;
                              }
                            if (t.flags & TMA_MODE1)
                              {
                                t.list_off = (uint32_t)foreign.size();
                                for (int j = 0; j < NFPt; ++j)
                                  {
                                    int X = R, Y = R, Z = R;
                                    if (j < R * R)
                                      Y = j % R, Z = j / R;
                                    else if (j < 2 * R * R)
                                      X = (j - R * R) % R, Z = (j - R * R) / R;
                                    else if (j < 3 * R * R)
                                      X = (j - 2 * R * R) % R, Y = (j - 2 * R * R) / R;
                                    else if (j < 3 * R * R + R)
                                      Z = j - 3 * R * R;
                                    else if (j < 3 * R * R + 2 * R)
                                      Y = j - 3 * R * R - R;
                                    else if (j < 3 * R * R + 3 * R)
                                      X = j - 3 * R * R - 2 * R;
                                    uint32_t g = 0;
                                    if (j < NFORt)
                                      {
                                        g = plain_gidx(bd, op->nb.cidx, X, Y, Z);
                                        if (g == INVALID_INDEX)
                                          throw std::runtime_error("internal: constrained DoF in the closure of a lex brick");
                                      }
                                    foreign.push_back(g);
                                  }
                              }
                          }
                        std::stable_partition(fast_ids.begin(), fast_ids.end(), [&](uint32_t b) { return op->h_brick_boundary[b] != 0; });
                        op->n_fast_boundary = 0;
                        for (const uint32_t b : fast_ids)
                          op->n_fast_boundary += op->h_brick_boundary[b] ? 1 : 0;
                        const TmaChunked ch = tma_build_list(op, fast_ids, op->n_fast_boundary);
                        cudaFree(op->d_fast_ids);
                        cudaFree(op->d_slow_ids);
                        op->tma_ok        = true;
                        op->fast_ok       = true;
                        op->tma_any_mode1 = ch.any_mode1;
                        op->tma_any_mode1_interior = ch.any_mode1_interior;
                        op->d_tma_lap     = dev_upload(ch.descs, ctx->stream);
                        op->d_tma_lap_chunks = dev_upload(ch.chunk_start, ctx->stream);
                        op->tma_lap_n_chunks = (int)ch.chunk_start.size() - 1;
                        op->tma_lap_n_chunks_boundary = ch.n_chunks_boundary;
                        op->d_tma_foreign = dev_upload(foreign, ctx->stream);
                        op->d_fast_ids    = dev_upload(fast_ids, ctx->stream);
                        op->d_slow_ids    = dev_upload(slow_ids, ctx->stream);
                        op->n_fast        = (int)fast_ids.size();
                        op->n_slow        = (int)slow_ids.size();
                        op->h_fast_ids    = fast_ids;
                      }
                    // 1-D matrices of the Kronecker form of the Cartesian cell matrix in even-odd form
                    {
                      std::vector<double> A(n * n);
                      bool                eo = eo_pack_centrosymmetric(n, op->basis.M_ref.data(), op->lap_P[0], op->lap_Q[0]);
                      for (int d = 0; d < 3; ++d)
                        {
                          for (int i = 0; i < n * n; ++i)
                            A[i] = op->cart.g[d] * op->basis.K_ref[i];
                          eo = eo_pack_centrosymmetric(n, A.data(), op->lap_P[1 + d], op->lap_Q[1 + d]) && eo;
                        }
                      if (!eo)
                        op->fast_ok = false;
                    }
                  }
              }
            else
              op->use_brick = false; // no consistent tile maps: use the generic kernels
          }
        CUDA_CHECK(cudaMalloc(&op->d_acc, std::max<size_t>(1, (size_t)op->n_vec) * op->esize()));
        CUDA_CHECK(cudaMemset(op->d_acc, 0, std::max<size_t>(1, (size_t)op->n_vec) * op->esize()));
        cudaDeviceProp prop;
        CUDA_CHECK(cudaGetDeviceProperties(&prop, ctx->device));
        op->n_sm = prop.multiProcessorCount;
        op->max_smem = (int)prop.sharedMemPerBlockOptin;
      }
  }
  if (deterministic_env())
    setup_deterministic(op);
  *out = op;
  DASM_API_END
}

// LaplaceOperatorMatrixFree on an unstructured all-hex mesh given by arrays (SURVEY 8(b) dasm_mesh_create_from_arrays; the ball of
// element_centered_preconditioners_01.cc:398-402).  Connectivity, orientation words, compressed indices and geometry: csrc/unstructured.h.
// The generic kernels run it: (k+1)^3 oriented addresses per cell (= read_dof_values of ConstraintInfoReduced with the orientation
// word, vector_access_reduced.h:267-405, expanded once at set-up) and merged coefficients or quadrature points per cell.
extern "C" int
dasm_op_create_unstructured(dasm_ctx *ctx, int degree, int number_type, const char *mapping_type, long long n_vertices, const double *coords,
                            long long n_cells, const uint32_t *cell_vertices, const double *support_points, int dirichlet, dasm_op **out)
{
  DASM_API_BEGIN
  DASM_REQUIRE(degree >= 1 && degree <= MAX_DEGREE, "degree must be in 1..8");
  DASM_REQUIRE(number_type == DASM_F64 || number_type == DASM_F32, "unknown number type");
  DASM_REQUIRE(ctx != nullptr, "unstructured operator: context required");
  DASM_REQUIRE(n_vertices > 0 && n_cells > 0 && coords != nullptr && cell_vertices != nullptr, "unstructured mesh: empty mesh");
  const std::string mt = mapping_type ? mapping_type : "";
  if (mt != "" && mt != "merged" && mt != "construct q" && mt != "quadratic geometry")
    {
      if (mt == "linear geometry")
        throw std::runtime_error("Mapping type <" + mt + "> is built for structured meshes only (brick kernels)");
      throw std::runtime_error("Mapping type <" + mt + "> is not known!"); // operator.h:747-752
    }
  std::unique_ptr<UMesh> U(new UMesh);
  U->n_vertices = n_vertices;
  U->n_cells    = n_cells;
  U->dirichlet  = dirichlet != 0;
  U->coords.assign(coords, coords + 3 * n_vertices);
  U->cells.assign(cell_vertices, cell_vertices + 8 * n_cells);
  if (support_points)
    U->support.assign(support_points, support_points + 81 * n_cells);
  U->build();
  UMesh::Numbering un = U->number_dofs(degree);
  auto             op = new dasm_op(degree);
  op->ctx              = ctx;
  op->mesh             = nullptr;
  op->k                = degree;
  op->ntype            = number_type;
  op->compress_indices = true; // 27 start indices + orientation word per cell are the stored format; the kernels read their expansion
  op->mapping_type     = mt;
  op->n_cells          = n_cells;
  op->n_owned          = un.n_dofs;
  op->n_ghost          = 0;
  op->n_vec            = un.n_dofs;
  op->n_global_dofs    = un.n_dofs;
  op->nb.k             = degree;
  op->nb.n_owned       = (uint32_t)un.n_dofs;
  op->nb.cidx          = un.cidx;
  op->nb.cidx_plain    = un.cidx_plain;
  op->nb.constrained   = un.constrained;
  op->h_plain          = std::move(un.plain);
    {
      CUDA_CHECK(cudaSetDevice(ctx->device));
      op->d_cidx        = dev_upload(op->nb.cidx, ctx->stream);
      op->d_plain       = dev_upload(op->h_plain, ctx->stream);
      op->d_constrained = dev_upload(op->nb.constrained, ctx->stream);
      op->n_constrained = (long long)op->nb.constrained.size();
      op->exchange.init(ctx, op->nb.exchange, op->esize());
      const int n3      = (degree + 1) * (degree + 1) * (degree + 1);
      const int per     = (mt == "construct q") ? 3 * n3 : (mt == "quadratic geometry" ? 81 : 6 * n3);
      op->geom_mode     = (mt == "construct q") ? 3 : (mt == "quadratic geometry" ? 5 : 1);
      std::vector<double> g((size_t)n_cells * per);
      for (long long c = 0; c < n_cells; ++c)
        if (op->geom_mode == 5)
          {
            // "quadratic geometry" (operator.h:1035-1159): 81 numbers per cell, the Jacobian is rebuilt per quadrature point in the
            // kernel; only differences of the support points enter it, so they are stored relative to the first one (single precision)
            const double *X = U->support.data() + (size_t)c * 81;
            for (int i = 0; i < 81; ++i)
              g[(size_t)c * 81 + i] = X[i] - X[i % 3];
          }
        else if (op->geom_mode == 3)
          U->cell_geometry(c, op->basis, g.data() + (size_t)c * per, nullptr, nullptr);
        else
          U->cell_geometry(c, op->basis, nullptr, nullptr, g.data() + (size_t)c * per);
      if (number_type == DASM_F64)
        op->d_geom = dev_upload(g, ctx->stream);
      else
        {
          std::vector<float> gf(g.begin(), g.end());
          op->d_geom = dev_upload(gf, ctx->stream);
        }
      op->cart.g[0] = op->cart.g[1] = op->cart.g[2] = 0;
    }
  op->use_brick = false;
  op->umesh     = std::move(U);
  if (deterministic_env())
    setup_deterministic(op);
  *out          = op;
  DASM_API_END
}

// host-only (no device): connectivity, numbering and patch extents of an unstructured mesh.  sizes = {n_dofs, n_lines, n_quads,
// n_constrained}; every output array may be NULL (first call: sizes only).  cidx [cell][27] (constrained entities invalid),
// orientation [cell], plain [cell][(k+1)^3], constrained [n_constrained], extents [cell][3][3]
extern "C" int
dasm_umesh_host_numbering(int degree, long long n_vertices, const double *coords, long long n_cells, const uint32_t *cell_vertices,
                          const double *support_points, int dirichlet, long long sizes[4], uint32_t *cidx, uint32_t *orientation, uint32_t *plain,
                          uint32_t *constrained, double *extents)
{
  DASM_API_BEGIN
  DASM_REQUIRE(degree >= 1 && degree <= MAX_DEGREE, "degree must be in 1..8");
  UMesh U;
  U.n_vertices = n_vertices;
  U.n_cells    = n_cells;
  U.dirichlet  = dirichlet != 0;
  U.coords.assign(coords, coords + 3 * n_vertices);
  U.cells.assign(cell_vertices, cell_vertices + 8 * n_cells);
  if (support_points)
    U.support.assign(support_points, support_points + 81 * n_cells);
  U.build();
  const UMesh::Numbering un = U.number_dofs(degree);
  sizes[0] = un.n_dofs;
  sizes[1] = U.n_lines;
  sizes[2] = U.n_quads;
  sizes[3] = (long long)un.constrained.size();
  if (cidx)
    memcpy(cidx, un.cidx.data(), un.cidx.size() * sizeof(uint32_t));
  if (orientation)
    memcpy(orientation, U.orientation.data(), U.orientation.size() * sizeof(uint32_t));
  if (plain)
    memcpy(plain, un.plain.data(), un.plain.size() * sizeof(uint32_t));
  if (constrained && !un.constrained.empty())
    memcpy(constrained, un.constrained.data(), un.constrained.size() * sizeof(uint32_t));
  if (extents)
    {
      const Basis1D             b(degree);
      const std::vector<double> e = U.patch_extents(b);
      memcpy(extents, e.data(), e.size() * sizeof(double));
    }
  DASM_API_END
}

// the packed orientation word per cell (12 line bits + 6 x 3 quad bits) and the (k+1)^3 oriented addresses per cell (host copies)
extern "C" int
dasm_op_orientations(const dasm_op *op, uint32_t *out)
{
  DASM_API_BEGIN
  DASM_REQUIRE(op->umesh != nullptr, "orientation words exist for operators on unstructured meshes (structured meshes: all standard)");
  memcpy(out, op->umesh->orientation.data(), op->umesh->orientation.size() * sizeof(uint32_t));
  DASM_API_END
}

extern "C" int
dasm_op_plain_indices(const dasm_op *op, uint32_t *out)
{
  DASM_API_BEGIN
  DASM_REQUIRE(op->umesh != nullptr, "plain index export is provided for operators on unstructured meshes");
  memcpy(out, op->h_plain.data(), op->h_plain.size() * sizeof(uint32_t));
  DASM_API_END
}

extern "C" long long dasm_op_n_cells(const dasm_op *op) { return op->n_cells; }
extern "C" int dasm_op_is_unstructured(const dasm_op *op) { return op->umesh ? 1 : 0; }
extern "C" const uint32_t *dasm_op_device_plain_indices(const dasm_op *op) { return op->d_plain; }

// number of cells that share the entity e of a cell (out[cell*27+e]; entities without DoFs: 1): the weights of the two-level
// transfers on unstructured meshes
extern "C" int
dasm_op_entity_valence(const dasm_op *op, uint8_t *out)
{
  DASM_API_BEGIN
  DASM_REQUIRE(op->umesh != nullptr, "entity valences are exported for operators on unstructured meshes");
  const UMesh &         U = *op->umesh;
  std::vector<uint32_t> nv((size_t)U.n_vertices, 0), nl((size_t)U.n_lines, 0), nq((size_t)U.n_quads, 0);
  for (long long c = 0; c < U.n_cells; ++c)
    for (int e = 0; e < 27; ++e)
      {
        const int dim = UMesh::entity_dim(e);
        auto &    cnt = dim == 0 ? nv : (dim == 1 ? nl : nq);
        if (dim < 3)
          cnt[U.entity[c * 27 + e]]++;
      }
  for (long long c = 0; c < U.n_cells; ++c)
    for (int e = 0; e < 27; ++e)
      {
        const int dim = UMesh::entity_dim(e);
        uint32_t  v   = 1;
        if (dim < 3)
          v = (dim == 0 ? nv : (dim == 1 ? nl : nq))[U.entity[c * 27 + e]];
        DASM_REQUIRE(v <= 255, "entity valence above 255");
        out[c * 27 + e] = (uint8_t)v;
      }
  DASM_API_END
}

// harmonic patch extents [cell][3][3] of an unstructured operator (grid_tools.h:54-138), for inspection
extern "C" int
dasm_op_patch_extents(const dasm_op *op, double *out)
{
  DASM_API_BEGIN
  DASM_REQUIRE(op->umesh != nullptr, "patch extent export is provided for operators on unstructured meshes");
  const std::vector<double> e = op->umesh->patch_extents(op->basis);
  memcpy(out, e.data(), e.size() * sizeof(double));
  DASM_API_END
}

extern "C" int
dasm_op_destroy(dasm_op *op)
{
  DASM_API_BEGIN
  if (!op)
    return 0;
  cudaStreamSynchronize(op->ctx->stream);
  cudaFree(op->d_cidx);
  cudaFree(op->d_constrained);
  cudaFree(op->d_geom);
  cudaFree(op->d_qcoef);
  cudaFree(op->d_bricks);
  cudaFree(op->d_acc);
  cudaFree(op->d_shared_list);
  cudaFree(op->d_plain);
  cudaFree(op->d_color_cells);
  cudaFree(op->d_fast_ids);
  cudaFree(op->d_slow_ids);
  cudaFree(op->d_tma_lap);
  cudaFree(op->d_tma_lap_chunks);
  cudaFree(op->d_tma_foreign);
  for (void *p : op->d_map_bufs)
    cudaFree(p);
  op->exchange.destroy();
  op->exchange_ext.destroy();
  for (void *p : op->scratch)
    cudaFree(p);
  delete op;
  DASM_API_END
}

extern "C" long long dasm_op_n_dofs(const dasm_op *op) { return op->n_owned; }
extern "C" long long dasm_op_n_fast_bricks(const dasm_op *op) { return (op->fast_ok && op->geom_mode == 0) ? op->n_fast : 0; }
extern "C" long long dasm_op_n_ghost(const dasm_op *op) { return op->n_ghost; }
extern "C" long long dasm_op_n_import(const dasm_op *op) { return (long long)op->exchange.n_send; }
extern "C" long long dasm_op_vec_size(const dasm_op *op) { return op->n_vec; }
extern "C" long long dasm_op_n_global_dofs(const dasm_op *op) { return op->n_global_dofs; }
extern "C" int dasm_op_degree(const dasm_op *op) { return op->k; }
extern "C" int dasm_op_number_type(const dasm_op *op) { return op->ntype; }
extern "C" int dasm_op_uses_compressed_indices(const dasm_op *op) { return op->compress_indices ? 1 : 0; }
extern "C" const uint32_t *dasm_op_device_indices(const dasm_op *op) { return op->d_cidx; }
extern "C" dasm_ctx *dasm_op_ctx(const dasm_op *op) { return op->ctx; }
extern "C" dasm_mesh *dasm_op_mesh(const dasm_op *op) { return op->mesh; }

extern "C" int
dasm_mesh_global_size(const dasm_mesh *mesh, int n_cells[3], int periodic[3])
{
  DASM_API_BEGIN
  DASM_REQUIRE(mesh != nullptr, "this operation needs an operator on a structured mesh (the operator was built on an unstructured one)");
  for (int d = 0; d < 3; ++d)
    {
      n_cells[d]  = mesh->mesh->p.nc[d];
      periodic[d] = mesh->mesh->p.periodic[d];
    }
  DASM_API_END
}

// update_ghost_values / compress(VectorOperation::add) of a vector in the operator's layout (matrix_free_internal.h:321-352)
extern "C" int
dasm_op_update_ghost_values(dasm_op *op, void *vec)
{
  DASM_API_BEGIN
  CUDA_CHECK(cudaSetDevice(op->ctx->device));
  DISPATCH_TYPE(op->ntype, op->exchange.run<T>((T *)vec, false));
  DASM_API_END
}

extern "C" int
dasm_op_compress_add(dasm_op *op, void *vec)
{
  DASM_API_BEGIN
  CUDA_CHECK(cudaSetDevice(op->ctx->device));
  DISPATCH_TYPE(op->ntype, op->exchange.run<T>((T *)vec, true));
  if (op->exchange.active() && op->n_ghost > 0)
    CUDA_CHECK(cudaMemsetAsync((char *)vec + (size_t)op->n_owned * op->esize(), 0, (size_t)op->n_ghost * op->esize(), op->ctx->stream));
  DASM_API_END
}

extern "C" int
dasm_op_vmult(dasm_op *op, void *dst, const void *src)
{
  DASM_API_BEGIN
  CUDA_CHECK(cudaSetDevice(op->ctx->device));
  DISPATCH_TYPE(op->ntype, op_vmult<T>(op, (T *)dst, (const T *)src, nullptr, nullptr));
  DASM_API_END
}

extern "C" int
dasm_op_vmult_hooks(dasm_op *op, void *dst, const void *src, const dasm_hook *pre, const dasm_hook *post)
{
  DASM_API_BEGIN
  CUDA_CHECK(cudaSetDevice(op->ctx->device));
  DISPATCH_TYPE(op->ntype, op_vmult<T>(op, (T *)dst, (const T *)src, pre, post));
  DASM_API_END
}

template <typename T>
static void
op_inverse_diagonal(dasm_op *op, T *diag)
{
  dasm_ctx *ctx = op->ctx;
  CUDA_CHECK(cudaMemsetAsync(diag, 0, (size_t)op->n_vec * sizeof(T), ctx->stream));
  DISPATCH_DEGREE(op->k, {
    constexpr int   n3    = (K + 1) * (K + 1) * (K + 1);
    const long long total = op->n_cells * n3;
    if (op->geom_mode == 0)
      laplace_diagonal_kernel<K, T, 0><<<nblocks(total, 128), 128, 0, ctx->stream>>>(diag, op->d_cidx, (const T *)nullptr, op->cart, op->n_cells, op->d_plain);
    else if (op->geom_mode == 3)
      laplace_diagonal_kernel<K, T, 2><<<nblocks(total, 128), 128, 0, ctx->stream>>>(diag, op->d_cidx, (const T *)op->d_geom, op->cart, op->n_cells, op->d_plain);
    else if (op->geom_mode == 5)
      laplace_diagonal_kernel<K, T, 3><<<nblocks(total, 128), 128, 0, ctx->stream>>>(diag, op->d_cidx, (const T *)op->d_geom, op->cart, op->n_cells, op->d_plain);
    else
      laplace_diagonal_kernel<K, T, 1><<<nblocks(total, 128), 128, 0, ctx->stream>>>(diag, op->d_cidx, (const T *)op->d_geom, op->cart, op->n_cells, op->d_plain);
  });
  ctx->launches++;
  op->exchange.run<T>(diag, true);
  if (op->n_constrained > 0)
    {
      vec_set_indexed_kernel<T><<<nblocks(op->n_constrained), 256, 0, ctx->stream>>>(diag, T(1), op->d_constrained, op->n_constrained);
      ctx->launches++;
    }
  vec_invert_diag_kernel<T><<<nblocks(op->n_owned), 256, 0, ctx->stream>>>(diag, op->n_owned);
  ctx->launches++;
  CUDA_CHECK(cudaGetLastError());
}

extern "C" int
dasm_op_inverse_diagonal(dasm_op *op, void *diag)
{
  DASM_API_BEGIN
  CUDA_CHECK(cudaSetDevice(op->ctx->device));
  DISPATCH_TYPE(op->ntype, op_inverse_diagonal<T>(op, (T *)diag));
  DASM_API_END
}

extern "C" int
dasm_op_compressed_indices(const dasm_op *op, int plain, uint32_t *out)
{
  const auto &v = plain ? op->nb.cidx_plain : op->nb.cidx;
  memcpy(out, v.data(), v.size() * sizeof(uint32_t));
  return 0;
}

extern "C" long long
dasm_op_constrained_dofs(const dasm_op *op, uint32_t *out)
{
  if (out)
    memcpy(out, op->nb.constrained.data(), op->nb.constrained.size() * sizeof(uint32_t));
  return (long long)op->nb.constrained.size();
}

extern "C" int
dasm_op_merged_coefficients(const dasm_op *op, long long cell, double *out)
{
  DASM_API_BEGIN
  DASM_REQUIRE(cell >= 0 && cell < op->n_cells, "cell out of range");
  if (op->umesh)
    {
      op->umesh->cell_geometry(cell, op->basis, nullptr, nullptr, out);
      return 0;
    }
  const Mesh &M    = *op->mesh->mesh;
  const int   c[3] = {M.cell_ijk[cell][0], M.cell_ijk[cell][1], M.cell_ijk[cell][2]};
  M.merged_coefficients(c, op->basis, out, op->linear_geometry);
  DASM_API_END
}

// LaplaceOperatorBase::rhs(vec, func) for a constant function (include/operator.h:298-330 with VectorTools::create_right_hand_side:
// b_i = int f phi_i, constrained entries zero): assembled on the host (set-up operation) with the operator's geometry, added up
// over the ranks by the ghost exchange
extern "C" int
dasm_op_rhs_constant(dasm_op *op, void *vec, double value)
{
  DASM_API_BEGIN
  CUDA_CHECK(cudaSetDevice(op->ctx->device));
  const int           k = op->k, n = k + 1, n3 = n * n * n;
  const Basis1D &     b = op->basis;
  std::vector<double> host((size_t)op->n_vec, 0.), jxw(n3), t0(n3), t1(n3);
  for (long long c = 0; c < op->n_cells; ++c)
    {
      if (op->umesh)
        op->umesh->cell_geometry(c, b, nullptr, jxw.data(), nullptr);
      else
        {
          const Mesh &M     = *op->mesh->mesh;
          const int   cc[3] = {M.cell_ijk[c][0], M.cell_ijk[c][1], M.cell_ijk[c][2]};
          M.jxw(cc, b, jxw.data(), op->linear_geometry);
        }
      // local vector = (N^T x N^T x N^T) (f JxW), N[q*n+i]
      for (int qz = 0; qz < n; ++qz)
        for (int qy = 0; qy < n; ++qy)
          for (int i = 0; i < n; ++i)
            {
              double s = 0;
              for (int qx = 0; qx < n; ++qx)
                s += b.N[qx * n + i] * jxw[(qz * n + qy) * n + qx];
              t0[(qz * n + qy) * n + i] = s * value;
            }
      for (int qz = 0; qz < n; ++qz)
        for (int j = 0; j < n; ++j)
          for (int i = 0; i < n; ++i)
            {
              double s = 0;
              for (int qy = 0; qy < n; ++qy)
                s += b.N[qy * n + j] * t0[(qz * n + qy) * n + i];
              t1[(qz * n + j) * n + i] = s;
            }
      const uint32_t *ci = op->nb.cidx.data() + c * 27;
      for (int l = 0; l < n; ++l)
        for (int j = 0; j < n; ++j)
          for (int i = 0; i < n; ++i)
            {
              double s = 0;
              for (int qz = 0; qz < n; ++qz)
                s += b.N[qz * n + l] * t1[(qz * n + j) * n + i];
              const uint32_t g = op->umesh ? op->h_plain[(size_t)c * n3 + (l * n + j) * n + i] : expand_start_index(ci, k, i, j, l);
              if (g != INVALID_INDEX)
                host[g] += s;
            }
    }
  if (op->ntype == DASM_F64)
    {
      CUDA_CHECK(cudaMemcpyAsync(vec, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice, op->ctx->stream));
      CUDA_CHECK(cudaStreamSynchronize(op->ctx->stream));
      op->exchange.run<double>((double *)vec, true);
    }
  else
    {
      std::vector<float> hf(host.begin(), host.end());
      CUDA_CHECK(cudaMemcpyAsync(vec, hf.data(), hf.size() * sizeof(float), cudaMemcpyHostToDevice, op->ctx->stream));
      CUDA_CHECK(cudaStreamSynchronize(op->ctx->stream));
      op->exchange.run<float>((float *)vec, true);
    }
  if (op->n_vec > op->n_owned)
    CUDA_CHECK(cudaMemsetAsync((char *)vec + (size_t)op->n_owned * op->esize(), 0, (size_t)(op->n_vec - op->n_owned) * op->esize(), op->ctx->stream));
  DASM_API_END
}

extern "C" int
dasm_op_vec_alloc(dasm_op *op, void **dev)
{
  DASM_API_BEGIN
  CUDA_CHECK(cudaSetDevice(op->ctx->device));
  CUDA_CHECK(cudaMalloc(dev, std::max<size_t>(1, (size_t)op->n_vec) * op->esize()));
  CUDA_CHECK(cudaMemsetAsync(*dev, 0, (size_t)op->n_vec * op->esize(), op->ctx->stream));
  DASM_API_END
}

extern "C" int
dasm_op_vec_free(dasm_op *op, void *dev)
{
  DASM_API_BEGIN
  CUDA_CHECK(cudaStreamSynchronize(op->ctx->stream));
  CUDA_CHECK(cudaFree(dev));
  DASM_API_END
}

template <typename T>
static void
vec_upload(dasm_op *op, T *dev, const double *host)
{
  std::vector<T> tmp(op->n_vec, T(0));
  for (long long i = 0; i < op->n_owned; ++i)
    tmp[i] = (T)host[i];
  CUDA_CHECK(cudaMemcpyAsync(dev, tmp.data(), (size_t)op->n_vec * sizeof(T), cudaMemcpyHostToDevice, op->ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(op->ctx->stream));
}

template <typename T>
static void
vec_download(dasm_op *op, double *host, const T *dev)
{
  std::vector<T> tmp(op->n_owned);
  CUDA_CHECK(cudaMemcpyAsync(tmp.data(), dev, (size_t)op->n_owned * sizeof(T), cudaMemcpyDeviceToHost, op->ctx->stream));
  CUDA_CHECK(cudaStreamSynchronize(op->ctx->stream));
  for (long long i = 0; i < op->n_owned; ++i)
    host[i] = (double)tmp[i];
}

extern "C" int
dasm_op_vec_upload(dasm_op *op, void *dev, const double *host_owned)
{
  DASM_API_BEGIN
  CUDA_CHECK(cudaSetDevice(op->ctx->device));
  DISPATCH_TYPE(op->ntype, vec_upload<T>(op, (T *)dev, host_owned));
  DASM_API_END
}

extern "C" int
dasm_op_vec_download(dasm_op *op, double *host_owned, const void *dev)
{
  DASM_API_BEGIN
  CUDA_CHECK(cudaSetDevice(op->ctx->device));
  DISPATCH_TYPE(op->ntype, vec_download<T>(op, host_owned, (const T *)dev));
  DASM_API_END
}

// ------------------------------------------------------------------------------------------------
// C ABI: FDM preconditioner
// ------------------------------------------------------------------------------------------------
// [deal.II TensorProductMatrixCreator::create_laplace_tensor_product_matrix, one direction]
static void
laplace_tp_matrix_1d(const Basis1D &b, const double ext[3], const int btype[2], int n_overlap, std::vector<double> &Mo,
                     std::vector<double> &Ko)
{
  const int n = b.n, m = n - 2 + 2 * n_overlap;
  Mo.assign(m * m, 0.);
  Ko.assign(m * m, 0.);
  const int o = n_overlap - 1;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      {
        Mo[(i + o) * m + j + o] = b.M_ref[i * n + j] * ext[1];
        Ko[(i + o) * m + j + o] = b.K_ref[i * n + j] / ext[1];
      }
  auto clear = [&](int r) {
    for (int i = 0; i < m; ++i)
      Mo[r * m + i] = Mo[i * m + r] = Ko[r * m + i] = Ko[i * m + r] = 0;
  };
  if (btype[0] == 0)
    {
      for (int i = 0; i < n_overlap; ++i)
        for (int j = 0; j < n_overlap; ++j)
          {
            Mo[i * m + j] += b.M_ref[(n - n_overlap + i) * n + (n - n_overlap + j)] * ext[0];
            Ko[i * m + j] += b.K_ref[(n - n_overlap + i) * n + (n - n_overlap + j)] / ext[0];
          }
    }
  else if (btype[0] == 1)
    clear(n_overlap - 1);
  if (btype[1] == 0)
    {
      const int s = n_overlap + n - 2;
      for (int i = 0; i < n_overlap; ++i)
        for (int j = 0; j < n_overlap; ++j)
          {
            Mo[(s + i) * m + s + j] += b.M_ref[i * n + j] * ext[2];
            Ko[(s + i) * m + s + j] += b.K_ref[i * n + j] / ext[2];
          }
    }
  else if (btype[1] == 1)
    clear(n_overlap + n - 2);
}

template <typename T>
__global__ void
valence_kernel(T *val, const uint32_t *idx, const long long n)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && idx[i] != DEV_INVALID)
    atomicAdd(val + idx[i], T(1));
}

template <int k>
__global__ void
expand_compressed_kernel(uint32_t *out, const uint32_t *cidx, const long long n_cells)
{
  constexpr int   n = k + 1, n3 = n * n * n;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cells * n3)
    return;
  const long long c = i / n3;
  const int       l = i % n3;
  out[i]            = compressed_index<k>(cidx + c * 27, l % n, (l / n) % n, l / (n * n));
}

template <typename T>
__global__ void
weights_from_valence_kernel(T *w, const int symm, const long long n)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    {
      const double v = (double)w[i];
      w[i]           = (v == 0.0) ? T(0) : (T)(1.0 / (symm ? sqrt(v) : v));
    }
}

template <typename T>
__global__ void
gather_entity_weights_kernel(T *cw, const T *w, const uint32_t *cidx, const long long n)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    cw[i] = (cidx[i] == DEV_INVALID) ? T(0) : w[cidx[i] & ~DEV_LEX_FLAG];
}

template <typename T>
static void
fdm_setup_device(dasm_fdm *f, const std::vector<double> &S, const std::vector<double> &lam, const std::vector<float> &ras_cw)
{
  dasm_op * op  = f->op;
  dasm_ctx *ctx = op->ctx;
  std::vector<T> Sd(S.begin(), S.end()), ld(lam.begin(), lam.end());
  f->d_S   = dev_upload(Sd, ctx->stream);
  f->d_lam = dev_upload(ld, ctx->stream);
  const int       m3      = f->m * f->m * f->m;
  const bool      compressed_idx = (f->d_pidx == nullptr);
  const long long n_entries = compressed_idx ? op->n_cells * (op->k + 1) * (op->k + 1) * (op->k + 1) : op->n_cells * m3;
  f->w_pre  = (f->weight_type == DASM_WEIGHT_PRE || f->weight_type == DASM_WEIGHT_SYMM);
  f->w_post = (f->weight_type == DASM_WEIGHT_POST || f->weight_type == DASM_WEIGHT_SYMM || f->weight_type == DASM_WEIGHT_RAS);
  if (f->weight_type == DASM_WEIGHT_NONE)
    {
      f->wmode = 0;
      return;
    }
  if (f->weight_type == DASM_WEIGHT_RAS)
    {
      // 0/1 ownership weights per (cell, entity) computed on the host
      std::vector<T> cw(ras_cw.begin(), ras_cw.end());
      f->d_cw  = dev_upload(cw, ctx->stream);
      f->wmode = compressed_idx ? 1 : 2;
      return;
    }
  // valence = number of patches containing a DoF (matrix_free.h:674-712)
  T *w = dev_alloc<T>(op->n_vec);
  CUDA_CHECK(cudaMemsetAsync(w, 0, (size_t)op->n_vec * sizeof(T), ctx->stream));
  uint32_t *d_full = nullptr;
  if (compressed_idx)
    {
      d_full = dev_alloc<uint32_t>(n_entries);
      DISPATCH_DEGREE(op->k, expand_compressed_kernel<K><<<nblocks(n_entries), 256, 0, ctx->stream>>>(d_full, op->d_cidx, op->n_cells));
    }
  valence_kernel<T><<<nblocks(n_entries), 256, 0, ctx->stream>>>(w, compressed_idx ? d_full : f->d_pidx, n_entries);
  ctx->launches += 2;
  Exchange &ex = f->use_ext ? op->exchange_ext : op->exchange;
  ex.run<T>(w, true);
  weights_from_valence_kernel<T><<<nblocks(op->n_owned), 256, 0, ctx->stream>>>(w, f->weight_type == DASM_WEIGHT_SYMM ? 1 : 0, op->n_owned);
  ctx->launches++;
  ex.run<T>(w, false);
  f->d_wvec = w;
  {
    std::vector<T> hw(op->n_owned);
    CUDA_CHECK(cudaMemcpyAsync(hw.data(), w, (size_t)op->n_owned * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    f->h_weights.assign(hw.begin(), hw.end());
  }
  if (compressed_idx && f->weight_sequence == DASM_WSEQ_COMPRESSED && op->k >= 2)
    {
      // 27 weights per cell by entity (compute_weights_fe_q_dofs_by_entity, matrix_free.h:756-764)
      T *cw = dev_alloc<T>(op->n_cells * 27);
      gather_entity_weights_kernel<T><<<nblocks(op->n_cells * 27), 256, 0, ctx->stream>>>(cw, w, op->d_cidx, op->n_cells * 27);
      ctx->launches++;
      f->d_cw  = cw;
      f->wmode = 1;
    }
  else if (f->weight_sequence == DASM_WSEQ_DG)
    {
      T *wl = dev_alloc<T>(n_entries);
      gather_entity_weights_kernel<T><<<nblocks(n_entries), 256, 0, ctx->stream>>>(wl, w, compressed_idx ? d_full : f->d_pidx, n_entries);
      ctx->launches++;
      f->d_cw  = wl;
      f->wmode = 2;
    }
  else
    f->wmode = 3; // global / local: gathered from the weight vector
  CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  if (d_full)
    cudaFree(d_full);
}

// ASPoissonPreconditioner on an unstructured mesh (include/matrix_free.h:73-894 with the unstructured pieces: grid_tools.h:54-138 for
// the patch extents; experiments/ball.py generates n overlap = 1 only): one patch per cell in the cell's own frame, the patch indices
// are the cell's (k+1)^3 oriented addresses, 1-D matrices from the cell's extent and those of the face neighbours, Dirichlet rows on
// boundary faces; weights through the explicit-list path (valence per DoF; RAS: the cell with the smallest index keeps a DoF)
static dasm_fdm *
fdm_create_unstructured(dasm_op *op, int n_overlap, int weight_type, int weight_sequence, int element_centric)
{
  const int    k = op->k, n = k + 1, n3 = n * n * n;
  const UMesh &U = *op->umesh;
  if (n_overlap != 1 || !element_centric)
    throw std::runtime_error("unstructured meshes: element-centred patches with n overlap = 1 only (as generated for the ball, experiments/ball.py:71-73)");
  if (k < 2)
    throw std::runtime_error("unstructured meshes: the explicit-list FDM kernel is instantiated for degrees >= 2");
  auto f             = new dasm_fdm;
  f->op              = op;
  f->use_ext         = false;
  f->n_overlap       = 1;
  f->weight_type     = weight_type;
  f->weight_sequence = (weight_sequence == DASM_WSEQ_COMPRESSED) ? DASM_WSEQ_DG : weight_sequence; // (27 weights per cell need the standard orientation)
  f->element_centric = 1;
  f->m               = n;
  const std::vector<double> ext = U.patch_extents(op->basis);
  std::map<std::array<uint64_t, 4>, uint32_t> cache;
  std::vector<double>                         S_all, lam_all;
  f->h_inst.resize((size_t)op->n_cells * 3);
  for (long long c = 0; c < op->n_cells; ++c)
    for (int d = 0; d < 3; ++d)
      {
        const double *e3   = ext.data() + (c * 3 + d) * 3;
        const int     bt[2] = {U.face_at_boundary(c, 2 * d) ? (U.dirichlet ? 1 : 2) : 0, U.face_at_boundary(c, 2 * d + 1) ? (U.dirichlet ? 1 : 2) : 0};
        std::array<uint64_t, 4> key;
        memcpy(&key[0], &e3[0], 8);
        memcpy(&key[1], &e3[1], 8);
        memcpy(&key[2], &e3[2], 8);
        key[3]  = (uint64_t)(bt[0] * 3 + bt[1]);
        auto it = cache.find(key);
        if (it == cache.end())
          {
            std::vector<double> Mm, Km, S, lam;
            laplace_tp_matrix_1d(op->basis, e3, bt, 1, Mm, Km);
            generalized_eig(n, Mm, Km, S, lam);
            it = cache.emplace(key, (uint32_t)cache.size()).first;
            S_all.insert(S_all.end(), S.begin(), S.end());
            lam_all.insert(lam_all.end(), lam.begin(), lam.end());
          }
        f->h_inst[c * 3 + d] = it->second;
      }
  f->n_instances = (long long)cache.size();
  f->h_S         = S_all;
  f->h_lam       = lam_all;
  f->d_inst      = dev_upload(f->h_inst, op->ctx->stream);
  f->d_pidx      = dev_upload(op->h_plain, op->ctx->stream);
  std::vector<float> ras_cw;
  if (weight_type == DASM_WEIGHT_RAS)
    {
      std::vector<uint32_t> winner((size_t)op->n_owned, INVALID_INDEX);
      for (long long c = op->n_cells - 1; c >= 0; --c)
        for (int i = 0; i < n3; ++i)
          {
            const uint32_t g = op->h_plain[(size_t)c * n3 + i];
            if (g != INVALID_INDEX)
              winner[g] = (uint32_t)c;
          }
      ras_cw.assign((size_t)op->n_cells * n3, 0.f);
      for (long long c = 0; c < op->n_cells; ++c)
        for (int i = 0; i < n3; ++i)
          {
            const uint32_t g = op->h_plain[(size_t)c * n3 + i];
            if (g != INVALID_INDEX && winner[g] == (uint32_t)c)
              ras_cw[(size_t)c * n3 + i] = 1.f;
          }
    }
  DISPATCH_TYPE(op->ntype, fdm_setup_device<T>(f, S_all, lam_all, ras_cw));
  return f;
}

extern "C" int
dasm_fdm_create(dasm_op *op, int n_overlap, int sub_mesh_approximation, int weight_type, int weight_sequence, int overlap_pre_post,
                int element_centric, dasm_fdm **out)
{
  DASM_API_BEGIN
  (void)overlap_pre_post;
  (void)sub_mesh_approximation;
  CUDA_CHECK(cudaSetDevice(op->ctx->device));
  DASM_REQUIRE(weight_type >= 0 && weight_type <= 4, "Weighting type is not known!");
  DASM_REQUIRE(weight_sequence >= 0 && weight_sequence <= 3, "weight sequence is not known!");
  const int k = op->k;
  n_overlap   = std::min(std::max(n_overlap, 1), k); // precondition.templates.h:195-196
  if (op->umesh)
    {
      *out = fdm_create_unstructured(op, n_overlap, weight_type, weight_sequence, element_centric);
      return 0;
    }
  const Mesh &M = *op->mesh->mesh;
  if ((n_overlap > 1 || !element_centric) && M.n_ranks() > 1)
    DASM_REQUIRE(!op->halo.exchange.empty() || op->halo.cells.empty(),
                 "n overlap > 1 / vertex patches on several ranks need the enlarged ghost layout (DASM_NO_HALO is set)");
  if (!element_centric)
    {
      DASM_REQUIRE(weight_type != DASM_WEIGHT_RAS, "RAS weighting with vertex patches is not implemented in libdasm yet");
      DASM_REQUIRE(2 * k - 1 <= 11, "vertex patches are instantiated up to degree 6");
    }
  auto f             = new dasm_fdm;
  f->op              = op;
  f->use_ext         = (n_overlap > 1 || !element_centric) && M.n_ranks() > 1;
  f->n_overlap       = n_overlap;
  f->weight_type     = weight_type;
  f->weight_sequence = weight_sequence;
  f->element_centric = element_centric;
  f->m               = element_centric ? (k - 1 + 2 * n_overlap) : (2 * k - 1); // matrix_free.h:90-92
  const int m        = f->m;

  // 1-D instances, deduplicated like TensorProductMatrixSymmetricSumCollection::finalize
  // (matrix_free.h:389-392)
  std::map<std::array<uint64_t, 4>, uint32_t> cache;
  std::vector<double>                         S_all, lam_all;
  f->h_inst.resize((size_t)op->n_cells * 3);
  for (long long c = 0; c < op->n_cells; ++c)
    {
      const int cc[3] = {M.cell_ijk[c][0], M.cell_ijk[c][1], M.cell_ijk[c][2]};
      for (int d = 0; d < 3; ++d)
        {
          double ext[3] = {0, M.harmonic_extent(cc, d, op->basis), 0};
          int    bt[2];
          for (int side = 0; side < 2; ++side)
            {
              int nbc[3];
              if (M.neighbor(cc, d, side, nbc))
                {
                  bt[side]      = 0;
                  ext[2 * side] = M.harmonic_extent(nbc, d, op->basis);
                }
              else
                bt[side] = M.p.dirichlet ? 1 : 2;
            }
          if (!element_centric)
            {
              // vertex patch of the cell's upper corner: own cell + right neighbour, extents replaced by 1 when
              // there is no neighbour (collect_patch_extend, matrix_free.h:1511-1524)
              ext[0] = ext[1] != 0.0 ? ext[1] : 1.0;
              ext[1] = ext[2] != 0.0 ? ext[2] : 1.0;
              ext[2] = -1.0; // marks the vertex-patch instances in the cache key
              bt[0] = bt[1] = 0;
            }
          std::array<uint64_t, 4> key;
          memcpy(&key[0], &ext[0], 8);
          memcpy(&key[1], &ext[1], 8);
          memcpy(&key[2], &ext[2], 8);
          key[3]  = (uint64_t)(bt[0] * 3 + bt[1]);
          auto it = cache.find(key);
          if (it == cache.end())
            {
              std::vector<double> Mm, Km, S, lam;
              if (element_centric)
                laplace_tp_matrix_1d(op->basis, ext, bt, n_overlap, Mm, Km);
              else
                {
                  // include/tensor_product_matrix_creator.h:7-61: two cells glued at the vertex, outer nodes dropped
                  const int n = k + 1;
                  Mm.assign(m * m, 0.);
                  Km.assign(m * m, 0.);
                  for (int i = 0; i < n - 1; ++i)
                    for (int j = 0; j < n - 1; ++j)
                      {
                        Mm[i * m + j] += op->basis.M_ref[(i + 1) * n + j + 1] * ext[0];
                        Km[i * m + j] += op->basis.K_ref[(i + 1) * n + j + 1] / ext[0];
                        Mm[(i + n - 2) * m + j + n - 2] += op->basis.M_ref[i * n + j] * ext[1];
                        Km[(i + n - 2) * m + j + n - 2] += op->basis.K_ref[i * n + j] / ext[1];
                      }
                }
              generalized_eig(m, Mm, Km, S, lam);
              const uint32_t id = (uint32_t)cache.size();
              cache[key]        = id;
              S_all.insert(S_all.end(), S.begin(), S.end());
              lam_all.insert(lam_all.end(), lam.begin(), lam.end());
              f->h_inst[c * 3 + d] = id;
            }
          else
            f->h_inst[c * 3 + d] = it->second;
        }
    }
  f->n_instances = (long long)cache.size();
  f->h_S         = S_all;
  f->h_lam       = lam_all;
  f->d_inst      = dev_upload(f->h_inst, op->ctx->stream);

  // explicit patch indices for n_overlap > 1 (dof_tools.h:78-137) and vertex patches (dof_tools.h:206-300)
  if (n_overlap > 1 || !element_centric)
    {
      const int n = k + 1, n3 = n * n * n, m3 = m * m * m;
      // expand the plain compressed indices of every cell on the host
      // (several ranks: the cells around the rank's box - halo cells - follow the local cells)
      const long long n_halo = f->use_ext ? (long long)op->halo.cells.size() : 0;
      std::vector<uint32_t> full((size_t)(op->n_cells + n_halo) * n3);
      std::map<std::array<int, 3>, long long> cell_of;
      for (long long c = 0; c < op->n_cells + n_halo; ++c)
        {
          const bool      loc = c < op->n_cells;
          const auto &    cc  = loc ? M.cell_ijk[c] : op->halo.cells[c - op->n_cells];
          cell_of[{cc[0], cc[1], cc[2]}] = c;
          const uint32_t *ci = loc ? op->nb.cidx.data() + c * 27 : op->halo.cidx.data() + (c - op->n_cells) * 27;
          for (int z = 0; z < n; ++z)
            for (int y = 0; y < n; ++y)
              for (int x = 0; x < n; ++x)
                full[(size_t)c * n3 + (z * n + y) * n + x] = expand_start_index(ci, k, x, y, z);
        }
      auto translate = [&](int i, int &which, int &l) {
        if (i < n_overlap - 1)
          {
            which = 0;
            l     = k + 1 - n_overlap + i;
          }
        else if (i < k + n_overlap)
          {
            which = 1;
            l     = i - (n_overlap - 1);
          }
        else
          {
            which = 2;
            l     = i - (n_overlap + k - 1);
          }
      };
      std::vector<uint32_t> pidx((size_t)op->n_cells * m3, INVALID_INDEX);
      for (long long c = 0; c < op->n_cells && !element_centric; ++c)
        {
          // the 2x2x2 cells above the cell (collect_cells_for_vertex_patch, matrix_free.h:1490-1509); if one is
          // missing the whole patch is invalid (dof_tools.h:222-227)
          const int cc[3] = {M.cell_ijk[c][0], M.cell_ijk[c][1], M.cell_ijk[c][2]};
          long long cells8[8];
          bool      all = true;
          for (int q = 0; q < 8 && all; ++q)
            {
              int cur[3] = {cc[0], cc[1], cc[2]};
              for (int d = 0; d < 3 && all; ++d)
                if ((q >> d) & 1)
                  {
                    int nbc[3];
                    if (!M.neighbor(cur, d, 1, nbc))
                      all = false;
                    else
                      for (int e = 0; e < 3; ++e)
                        cur[e] = nbc[e];
                  }
              if (all)
                cells8[q] = cell_of[{cur[0], cur[1], cur[2]}];
            }
          if (!all)
            continue;
          // positions 1 .. 2k-1 of the (2k+1)^3 lattice of the 8 cells
          for (int pz = 0; pz < m; ++pz)
            for (int py = 0; py < m; ++py)
              for (int px = 0; px < m; ++px)
                {
                  const int g[3] = {px + 1, py + 1, pz + 1};
                  int       q = 0, l[3];
                  for (int d = 0; d < 3; ++d)
                    {
                      const int half = (g[d] > k) ? 1 : 0; // the shared plane g = k is taken from the lower cell
                      q |= half << d;
                      l[d] = g[d] - half * k;
                    }
                  pidx[(size_t)c * m3 + (pz * m + py) * m + px] = full[(size_t)cells8[q] * n3 + (l[2] * n + l[1]) * n + l[0]];
                }
        }
      for (long long c = 0; c < op->n_cells && element_centric; ++c)
        {
          const int cc[3] = {M.cell_ijk[c][0], M.cell_ijk[c][1], M.cell_ijk[c][2]};
          for (int pz = 0; pz < m; ++pz)
            for (int py = 0; py < m; ++py)
              for (int px = 0; px < m; ++px)
                {
                  const int pp[3] = {px, py, pz};
                  int       cur[3] = {cc[0], cc[1], cc[2]}, li[3];
                  bool      ok = true;
                  for (int d = 0; d < 3 && ok; ++d)
                    {
                      int which;
                      translate(pp[d], which, li[d]);
                      if (which != 1)
                        {
                          int nbc[3];
                          if (!M.neighbor(cur, d, which == 0 ? 0 : 1, nbc))
                            ok = false;
                          else
                            {
                              cur[0] = nbc[0];
                              cur[1] = nbc[1];
                              cur[2] = nbc[2];
                            }
                        }
                    }
                  if (!ok)
                    continue;
                  const long long nc = cell_of[{cur[0], cur[1], cur[2]}];
                  pidx[(size_t)c * m3 + (pz * m + py) * m + px] = full[(size_t)nc * n3 + (li[2] * n + li[1]) * n + li[0]];
                }
        }
      f->d_pidx = dev_upload(pidx, op->ctx->stream);
      if (op->deterministic)
        {
          // overlapping patches reach beyond the cell: they get their own colouring
          std::vector<uint32_t> cells;
          greedy_coloring(pidx.data(), op->n_cells, m * m * m, op->n_vec, f->color_ptr, cells);
          f->d_color_cells = dev_upload(cells, op->ctx->stream);
        }
    }

  if (f->d_pidx == nullptr && op->d_plain != nullptr && f->m >= 3) // (the kernels with explicit lists start at patch size 3)
    {
      // the operator stores plain indices (compress_indices = false): the patch lists are those n^3 indices per cell
      const size_t ne = (size_t)op->n_cells * (k + 1) * (k + 1) * (k + 1);
      f->d_pidx       = dev_alloc<uint32_t>(std::max<size_t>(ne, 1));
      CUDA_CHECK(cudaMemcpyAsync(f->d_pidx, op->d_plain, ne * sizeof(uint32_t), cudaMemcpyDeviceToDevice, op->ctx->stream));
    }

  // RAS ownership: the patch of the touching cell with the smallest global lexicographic id owns
  // the DoF (matrix_free.h:536-673 uses the global cell-batch numbering for the same purpose)
  std::vector<float> ras_cw;
  if (weight_type == DASM_WEIGHT_RAS)
    {
      auto gid = [&](const int c[3]) { return ((long long)c[2] * M.p.nc[1] + c[1]) * M.p.nc[0] + c[0]; };
      auto entity_owner_is = [&](const int cc[3], int e) {
        int s[3];
        M.cell_slot(cc, e, s);
        long long best = -1;
        int       cand[3][2], ncand[3];
        for (int d = 0; d < 3; ++d)
          {
            ncand[d] = 0;
            if (s[d] % 2 == 1)
              cand[d][ncand[d]++] = s[d] / 2;
            else
              {
                int a = s[d] / 2 - 1, b2 = s[d] / 2;
                if (M.p.periodic[d])
                  {
                    a  = (a + M.p.nc[d]) % M.p.nc[d];
                    b2 = b2 % M.p.nc[d];
                  }
                if (a >= 0 && a < M.p.nc[d])
                  cand[d][ncand[d]++] = a;
                if (b2 >= 0 && b2 < M.p.nc[d] && (ncand[d] == 0 || cand[d][0] != b2))
                  cand[d][ncand[d]++] = b2;
              }
          }
        for (int a = 0; a < ncand[0]; ++a)
          for (int b2 = 0; b2 < ncand[1]; ++b2)
            for (int g = 0; g < ncand[2]; ++g)
              {
                const int       tc[3] = {cand[0][a], cand[1][b2], cand[2][g]};
                const long long id    = gid(tc);
                if (best < 0 || id < best)
                  best = id;
              }
        return best == gid(cc);
      };
      if (n_overlap == 1)
        {
          ras_cw.resize((size_t)op->n_cells * 27);
          for (long long c = 0; c < op->n_cells; ++c)
            {
              const int cc[3] = {M.cell_ijk[c][0], M.cell_ijk[c][1], M.cell_ijk[c][2]};
              for (int e = 0; e < 27; ++e)
                ras_cw[c * 27 + e] = entity_owner_is(cc, e) ? 1.f : 0.f;
            }
        }
      else
        {
          const int m3 = m * m * m;
          ras_cw.assign((size_t)op->n_cells * m3, 0.f);
          for (long long c = 0; c < op->n_cells; ++c)
            {
              const int cc[3] = {M.cell_ijk[c][0], M.cell_ijk[c][1], M.cell_ijk[c][2]};
              for (int pz = 0; pz < m; ++pz)
                for (int py = 0; py < m; ++py)
                  for (int px = 0; px < m; ++px)
                    {
                      const int pp[3] = {px, py, pz};
                      int       e = 0, mul = 1;
                      bool      core = true;
                      for (int d = 0; d < 3; ++d, mul *= 3)
                        {
                          const int i = pp[d];
                          if (i < n_overlap - 1 || i >= k + n_overlap)
                            core = false;
                          const int l = i - (n_overlap - 1);
                          e += mul * (l == 0 ? 0 : (l == k ? 2 : 1));
                        }
                      if (core && entity_owner_is(cc, e))
                        ras_cw[(size_t)c * m3 + (pz * m + py) * m + px] = 1.f;
                    }
            }
        }
    }
  DISPATCH_TYPE(op->ntype, fdm_setup_device<T>(f, S_all, lam_all, ras_cw));
  if (op->use_brick && f->d_pidx == nullptr)
    {
      // weight codes for the brick kernel: every compressed weight is a function of the integer patch valence of
      // its entity (1/v, 1/sqrt(v), RAS 0/1), so 1 byte per entity + a 16-entry table replace 27 numbers per cell
      std::vector<uint8_t> codes;
      if (f->wmode == 1)
        {
          std::vector<double> cw((size_t)op->n_cells * 27);
          if (op->ntype == DASM_F64)
            CUDA_CHECK(cudaMemcpy(cw.data(), f->d_cw, cw.size() * sizeof(double), cudaMemcpyDeviceToHost));
          else
            {
              std::vector<float> tmp(cw.size());
              CUDA_CHECK(cudaMemcpy(tmp.data(), f->d_cw, tmp.size() * sizeof(float), cudaMemcpyDeviceToHost));
              cw.assign(tmp.begin(), tmp.end());
            }
          codes.assign((size_t)op->n_cells * 32, 0);
          bool ok = true;
          bool                 seen[16] = {false};
          for (size_t i = 0; i < cw.size() && ok; ++i)
            {
              const double w = cw[i];
              int          code = 0;
              if (w != 0.0)
                {
                  const double v = (weight_type == DASM_WEIGHT_SYMM) ? 1.0 / (w * w) : 1.0 / w;
                  code           = (int)std::lround(v);
                  if (code < 1 || code > 15)
                    ok = false;
                }
              if (ok)
                {
                  if (!seen[code])
                    {
                      seen[code]    = true;
                      f->wtab[code] = w;
                    }
                  else if (f->wtab[code] != w)
                    ok = false;
                  codes[(i / 27) * 32 + (i % 27)] = (uint8_t)code;
                }
            }
          DASM_REQUIRE(ok, "internal: compressed weights are not a function of the patch valence");
          f->d_cwcode = dev_upload(codes, op->ctx->stream);
        }
      // per kernel brick: instance triple of the first cell, and whether all cells of the brick share it
      std::vector<BrickDesc> bricks(op->n_bricks);
      CUDA_CHECK(cudaMemcpy(bricks.data(), op->d_bricks, bricks.size() * sizeof(BrickDesc), cudaMemcpyDeviceToHost));
      std::vector<uint4> tri(op->n_bricks);
      for (int b = 0; b < op->n_bricks; ++b)
        {
          const BrickDesc &bd = bricks[b];
          const uint32_t * i0 = f->h_inst.data() + (size_t)bd.first_cell * 3;
          bool             uniform = true;
          const int        nc = bd.b[0] * bd.b[1] * bd.b[2];
          for (int c = 1; c < nc && uniform; ++c)
            for (int d = 0; d < 3; ++d)
              if (i0[c * 3 + d] != i0[d])
                uniform = false;
          tri[b] = make_uint4(i0[0], i0[1], i0[2], uniform ? 1u : 0u);
        }
      f->d_brick_tri = dev_upload(tri, op->ctx->stream);
      // ---- warp-specialised kernel: the regular bricks with the most frequent instance triple whose cells all have
      // the weight pattern of the first one; the 27 weights must be a tensor product so that they fold into the
      // 1-D matrices (true for 1/valence, 1/sqrt(valence) and the RAS ownership weights away from boundaries)
      if (op->fast_ok && f->m == op->k + 1 && (f->wmode == 0 || f->wmode == 1))
        {
          const int                                     n = op->k + 1;
          std::map<std::array<uint32_t, 3>, int>        count;
          for (const uint32_t b : op->h_fast_ids)
            if (tri[b].w)
              count[{tri[b].x, tri[b].y, tri[b].z}]++;
          std::array<uint32_t, 3> best = {0, 0, 0};
          int                     nbest = 0;
          for (const auto &kv : count)
            if (kv.second > nbest)
              {
                nbest = kv.second;
                best  = kv.first;
              }
          std::vector<uint32_t> fast_ids, slow_ids;
          const uint8_t *       ref_codes = nullptr;
          std::vector<char>     is_fast(op->n_bricks, 0);
          if (f->wmode == 1)
            {
              // reference weight pattern: the most frequent one among the first cells of the candidate bricks
              std::map<std::array<uint8_t, 27>, std::pair<int, const uint8_t *>> patterns;
              for (const uint32_t b : op->h_fast_ids)
                {
                  const uint8_t *         cb = codes.data() + (size_t)bricks[b].first_cell * 32;
                  std::array<uint8_t, 27> key;
                  std::copy(cb, cb + 27, key.begin());
                  auto &e = patterns[key];
                  e.first++;
                  e.second = cb;
                }
              int most = 0;
              for (const auto &kv : patterns)
                if (kv.second.first > most)
                  {
                    most      = kv.second.first;
                    ref_codes = kv.second.second;
                  }
            }
          for (const uint32_t b : op->h_fast_ids)
            {
              if (nbest == 0 || !tri[b].w || tri[b].x != best[0] || tri[b].y != best[1] || tri[b].z != best[2])
                continue;
              bool same = true;
              if (f->wmode == 1)
                {
                  const uint8_t *cb = codes.data() + (size_t)bricks[b].first_cell * 32;
                  for (int c = 0; c < 64 && same; ++c)
                    same = (memcmp(cb + c * 32, ref_codes, 27) == 0);
                }
              is_fast[b] = same;
            }
          // tensor-product factorisation of the weights
          double w1[3][3] = {{1, 1, 1}, {1, 1, 1}, {1, 1, 1}};
          bool   w_ok     = true;
          if (f->wmode == 1 && ref_codes != nullptr)
            {
              double W[27];
              for (int e = 0; e < 27; ++e)
                W[e] = f->wtab[ref_codes[e]];
              const double c = W[13];
              if (c == 0.)
                w_ok = false;
              else
                {
                  for (int e = 0; e < 3; ++e)
                    {
                      w1[0][e] = W[e + 3 + 9] / c;
                      w1[1][e] = W[1 + 3 * e + 9] / c;
                      w1[2][e] = W[1 + 3 + 9 * e];
                    }
                  for (int e = 0; e < 27; ++e)
                    {
                      const double p = w1[0][e % 3] * w1[1][(e / 3) % 3] * w1[2][e / 9];
                      if (std::fabs(p - W[e]) > (op->ntype == DASM_F64 ? 1e-14 : 1e-6) * std::max(1., std::fabs(W[e])))
                        w_ok = false;
                    }
                }
            }
          for (int b = 0; b < op->n_bricks; ++b)
            ((w_ok && is_fast[b]) ? fast_ids : slow_ids).push_back((uint32_t)b);
          if (!fast_ids.empty())
            {
              // even-odd form: the eigenvectors of the symmetric 1-D problem are even or odd under i -> n-1-i; the even
              // ones are ordered first (the eigenvalues follow), the weights must be symmetric
              auto   ent   = [&](int i) { return i == 0 ? 0 : (i == n - 1 ? 2 : 1); };
              const int m_ev = (n + 1) / 2;
              bool   eo_ok = true;
              double lam_p[3][9];
              for (int d = 0; d < 3 && eo_ok; ++d)
                {
                  const double *S = f->h_S.data() + (size_t)best[d] * n * n;
                  const double *l = f->h_lam.data() + (size_t)best[d] * n;
                  if (f->wmode == 1 && std::fabs(w1[d][0] - w1[d][2]) > 1e-14 * std::fabs(w1[d][0]))
                    eo_ok = false;
                  std::vector<int> perm, odd;
                  double           smax = 0;
                  for (int i = 0; i < n * n; ++i)
                    smax = std::max(smax, std::fabs(S[i]));
                  for (int a = 0; a < n; ++a)
                    {
                      double de = 0, dod = 0;
                      for (int i = 0; i < n; ++i)
                        {
                          de  = std::max(de, std::fabs(S[(n - 1 - i) * n + a] - S[i * n + a]));
                          dod = std::max(dod, std::fabs(S[(n - 1 - i) * n + a] + S[i * n + a]));
                        }
                      if (de <= 1e-11 * smax)
                        perm.push_back(a);
                      else if (dod <= 1e-11 * smax)
                        odd.push_back(a);
                      else
                        eo_ok = false;
                    }
                  if ((int)perm.size() != m_ev)
                    eo_ok = false;
                  if (!eo_ok)
                    break;
                  perm.insert(perm.end(), odd.begin(), odd.end());
                  std::vector<double> B(n * n), C(n * n);
                  for (int a = 0; a < n; ++a)
                    {
                      lam_p[d][a] = l[perm[a]];
                      for (int i = 0; i < n; ++i)
                        {
                          // first stage: u_a = sum_i S[i][a] w_i v_i;  second stage: y_o = w_o sum_a S[o][a] u_a
                          B[a * n + i] = S[i * n + perm[a]] * ((f->wmode == 1 && f->w_pre) ? w1[d][ent(i)] : 1.);
                          C[i * n + a] = S[i * n + perm[a]] * ((f->wmode == 1 && f->w_post) ? w1[d][ent(i)] : 1.);
                        }
                    }
                  eo_ok = eo_pack_forward(n, B.data(), f->fast_P[d], f->fast_Q[d]) && eo_pack_backward(n, C.data(), f->fast_P[3 + d], f->fast_Q[3 + d]);
                }
              if (eo_ok)
                for (int z = 0; z < n; ++z)
                  for (int y = 0; y < n; ++y)
                    for (int x = 0; x < n; ++x)
                      f->fast_inv[(z * n + y) * n + x] = 1. / (lam_p[0][x] + lam_p[1][y] + lam_p[2][z]);
              if (!eo_ok)
                {
                  fast_ids.clear();
                  slow_ids.clear();
                }
            }
          if (!fast_ids.empty())
            {
              std::stable_partition(fast_ids.begin(), fast_ids.end(), [&](uint32_t b) { return op->h_brick_boundary[b] != 0; });
              f->n_fast_boundary = 0;
              for (const uint32_t b : fast_ids)
                f->n_fast_boundary += op->h_brick_boundary[b] ? 1 : 0;
              f->fast_ok    = true;
              f->n_fast     = (int)fast_ids.size();
              f->n_slow     = (int)slow_ids.size();
              f->d_fast_ids = dev_upload(fast_ids, op->ctx->stream);
              f->d_slow_ids = dev_upload(slow_ids, op->ctx->stream);
              if (op->tma_only && f->n_slow > 0)
                f->fast_ok = false; // (no brick kernel for the remaining bricks: the generic kernel runs all cells)
              if (op->tma_ok)
                {
                  const TmaChunked ch = tma_build_list(op, fast_ids, f->n_fast_boundary);
                  f->tma_any_mode1    = ch.any_mode1;
                  f->tma_any_mode1_interior = ch.any_mode1_interior;
                  f->d_tma_list       = dev_upload(ch.descs, op->ctx->stream);
                  f->d_tma_chunks     = dev_upload(ch.chunk_start, op->ctx->stream);
                  f->tma_n_chunks     = (int)ch.chunk_start.size() - 1;
                  f->tma_n_chunks_boundary = ch.n_chunks_boundary;
                }
            }
        }
    }
  *out = f;
  DASM_API_END
}

extern "C" int
dasm_fdm_destroy(dasm_fdm *f)
{
  DASM_API_BEGIN
  if (!f)
    return 0;
  cudaStreamSynchronize(f->op->ctx->stream);
  cudaFree(f->d_inst);
  cudaFree(f->d_S);
  cudaFree(f->d_lam);
  cudaFree(f->d_wvec);
  cudaFree(f->d_cw);
  cudaFree(f->d_cwcode);
  cudaFree(f->d_brick_tri);
  cudaFree(f->d_fast_ids);
  cudaFree(f->d_slow_ids);
  cudaFree(f->d_tma_list);
  cudaFree(f->d_tma_chunks);
  cudaFree(f->d_pidx);
  cudaFree(f->d_color_cells);
  delete f;
  DASM_API_END
}

extern "C" int
dasm_fdm_vmult(dasm_fdm *f, void *dst, const void *src)
{
  DASM_API_BEGIN
  CUDA_CHECK(cudaSetDevice(f->op->ctx->device));
  DISPATCH_TYPE(f->op->ntype, fdm_vmult<T>(f, (T *)dst, (const T *)src, nullptr, nullptr));
  DASM_API_END
}

extern "C" int
dasm_fdm_vmult_hooks(dasm_fdm *f, void *dst, const void *src, const dasm_hook *pre, const dasm_hook *post)
{
  DASM_API_BEGIN
  CUDA_CHECK(cudaSetDevice(f->op->ctx->device));
  DISPATCH_TYPE(f->op->ntype, fdm_vmult<T>(f, (T *)dst, (const T *)src, pre, post));
  DASM_API_END
}

extern "C" long long dasm_fdm_n_instances(const dasm_fdm *f) { return f->n_instances; }
extern "C" long long dasm_fdm_n_fast_bricks(const dasm_fdm *f) { return f->fast_ok ? f->n_fast : 0; }
extern "C" long long
dasm_fdm_memory_consumption(const dasm_fdm *f)
{
  return (long long)(f->n_instances * (f->m * f->m + f->m) * f->op->esize() + f->op->n_cells * 3 * sizeof(uint32_t));
}
extern "C" int
dasm_fdm_is_symmetric(const dasm_fdm *f)
{
  return (f->weight_type == DASM_WEIGHT_NONE || f->weight_type == DASM_WEIGHT_SYMM) ? 1 : 0;
}
extern "C" int dasm_fdm_patch_size_1d(const dasm_fdm *f) { return f->m; }

// entity weights [cell][27] -> per-entry weights [cell][n^3]
template <int k, typename T>
__global__ void
expand_entity_weights_kernel(T *out, const T *cw, const long long n_cells)
{
  constexpr int   n = k + 1, n3 = n * n * n;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cells * n3)
    return;
  const long long c = i / n3;
  const int       l = i % n3;
  int             ex, ey, ez, o;
  split_1d<k>(l % n, ex, o);
  split_1d<k>((l / n) % n, ey, o);
  split_1d<k>(l / (n * n), ez, o);
  out[i] = cw[c * 27 + ex + 3 * ey + 9 * ez];
}

template <typename T>
__global__ void
fill_kernel(T *out, const T v, const long long n)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    out[i] = v;
}

// Patch layout of the preconditioner for callers that replace the FDM block inverse (block_asm.cu): explicit DoF indices
// d_idx[cell][m^3] (0xFFFFFFFF: outside the domain / constrained) and the weight d_w[cell][m^3] every entry is multiplied with
// (before and / or after the block solve: *w_pre, *w_post) - the data of Restrictors::ElementCenteredRestrictor
// (include/restrictors.h:48-338).  Both arrays are device buffers of the caller.
extern "C" int
dasm_fdm_export_patches(dasm_fdm *f, uint32_t *d_idx, void *d_w, int *w_pre, int *w_post)
{
  DASM_API_BEGIN
  dasm_op * op  = f->op;
  dasm_ctx *ctx = op->ctx;
  CUDA_CHECK(cudaSetDevice(ctx->device));
  const int       m3 = f->m * f->m * f->m;
  const long long ne = op->n_cells * m3;
  if (f->d_pidx == nullptr)
    {
      DISPATCH_DEGREE(op->k, expand_compressed_kernel<K><<<nblocks(ne), 256, 0, ctx->stream>>>(d_idx, op->d_cidx, op->n_cells));
    }
  else
    CUDA_CHECK(cudaMemcpyAsync(d_idx, f->d_pidx, (size_t)ne * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
  ctx->launches++;
  *w_pre  = f->w_pre ? 1 : 0;
  *w_post = f->w_post ? 1 : 0;
  DISPATCH_TYPE(op->ntype, {
    T *w = (T *)d_w;
    if (f->wmode == 0)
      fill_kernel<T><<<nblocks(ne), 256, 0, ctx->stream>>>(w, T(1), ne);
    else if (f->wmode == 1)
      {
        DISPATCH_DEGREE(op->k, (expand_entity_weights_kernel<K, T><<<nblocks(ne), 256, 0, ctx->stream>>>(w, (const T *)f->d_cw, op->n_cells)));
      }
    else if (f->wmode == 2)
      CUDA_CHECK(cudaMemcpyAsync(w, f->d_cw, (size_t)ne * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
    else
      gather_entity_weights_kernel<T><<<nblocks(ne), 256, 0, ctx->stream>>>(w, (const T *)f->d_wvec, d_idx, ne);
    ctx->launches++;
  });
  DASM_API_END
}
extern "C" dasm_op *dasm_fdm_op(const dasm_fdm *f) { return f->op; }

extern "C" int
dasm_fdm_weights(const dasm_fdm *f, double *out)
{
  DASM_API_BEGIN
  DASM_REQUIRE(!f->h_weights.empty(), "no global weight vector for this weighting type");
  memcpy(out, f->h_weights.data(), f->h_weights.size() * sizeof(double));
  DASM_API_END
}

extern "C" int
dasm_fdm_instance(const dasm_fdm *f, long long cell, int d, double *S, double *lambda)
{
  DASM_API_BEGIN
  DASM_REQUIRE(cell >= 0 && cell < f->op->n_cells && d >= 0 && d < 3, "cell/direction out of range");
  const uint32_t id = f->h_inst[cell * 3 + d];
  memcpy(S, f->h_S.data() + (size_t)id * f->m * f->m, sizeof(double) * f->m * f->m);
  memcpy(lambda, f->h_lam.data() + (size_t)id * f->m, sizeof(double) * f->m);
  DASM_API_END
}

// ------------------------------------------------------------------------------------------------
// C ABI: Chebyshev smoother  [deal.II PreconditionChebyshev restated]
// ------------------------------------------------------------------------------------------------
template <typename T>
static void
precon_apply(dasm_cheb *c, T *dst, const T *src, const dasm_hook *post)
{
  dasm_op * op  = c->op;
  dasm_ctx *ctx = op->ctx;
  if (c->fdm)
    fdm_vmult<T>(c->fdm, dst, src, nullptr, post);
  else
    {
      // DiagonalMatrixPrePost::vmult, preconditioners.h:965-993
      const long long n = op->n_owned;
      vec_mul_kernel<T><<<nblocks(n), 256, 0, ctx->stream>>>(dst, (const T *)c->d_inv_diag, src, n);
      ctx->launches++;
      if (post != nullptr && post->kind != DASM_HOOK_NONE)
        {
          if (post->kind == DASM_HOOK_CHEB_UPDATE)
            vec_cheb_update_kernel<T><<<nblocks(n), 256, 0, ctx->stream>>>(dst, dst, (const T *)post->v0, (const T *)post->v1, (T)post->f1, (T)post->f2, n);
          else if (post->kind == DASM_HOOK_SCALE)
            vec_scale_kernel<T><<<nblocks(n), 256, 0, ctx->stream>>>(dst, dst, (T)post->f2, n);
          else
            throw std::runtime_error("unsupported post hook kind");
          ctx->launches++;
        }
    }
}

static void
cheb_set_ev(dasm_cheb *c, double min_ev, double max_ev)
{
  c->min_ev = min_ev;
  c->max_ev = max_ev;
  const double alpha = c->smoothing_range > 1. ? max_ev / c->smoothing_range : std::min(0.9 * max_ev, min_ev);
  if (c->poly == DASM_POLY_FOURTH_KIND)
    c->delta = c->theta = max_ev;
  else
    {
      c->delta = (max_ev - alpha) * 0.5;
      c->theta = (max_ev + alpha) * 0.5;
    }
  c->ev_ready = true;
}

template <typename T>
__global__ void
initial_guess_kernel(T *v, const long long first, const long long n)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    v[i] = (T)((i + first) % 11);
}

template <typename T>
__global__ void
vec_add_scalar_kernel(T *v, const T a, const long long n)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    v[i] += a;
}

template <typename T>
__global__ void
vec_axpy_kernel(T *y, const T a, const T *x, const long long n)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    y[i] += a * x[i];
}

template <typename T>
__global__ void
vec_xpay_kernel(T *y, const T a, const T *x, const long long n) // y = x + a y
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    y[i] = x[i] + a * y[i];
}

// symmetric tridiagonal eigenvalues (Ritz values of the Lanczos/CG process), ascending
static std::vector<double>
tridiagonal_eigenvalues(std::vector<double> d, std::vector<double> e)
{
  // implicit QL (tql1-style)
  const int n = d.size();
  e.push_back(0.);
  for (int l = 0; l < n; ++l)
    {
      int iter = 0, m;
      do
        {
          for (m = l; m < n - 1; ++m)
            {
              const double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
              if (std::fabs(e[m]) <= 1e-300 + 2.3e-16 * dd)
                break;
            }
          if (m != l)
            {
              if (++iter > 200)
                break;
              double g = (d[l + 1] - d[l]) / (2. * e[l]);
              double r = std::hypot(g, 1.);
              g        = d[m] - d[l] + e[l] / (g + (g >= 0 ? std::fabs(r) : -std::fabs(r)));
              double s = 1., c = 1., p = 0.;
              int    i;
              for (i = m - 1; i >= l; --i)
                {
                  double f = s * e[i], b = c * e[i];
                  r        = std::hypot(f, g);
                  e[i + 1] = r;
                  if (r == 0.)
                    {
                      d[i + 1] -= p;
                      e[m] = 0.;
                      break;
                    }
                  s        = f / r;
                  c        = g / r;
                  g        = d[i + 1] - p;
                  r        = (d[i] - g) * s + 2. * c * b;
                  p        = s * r;
                  d[i + 1] = g + p;
                  g        = c * r - b;
                }
              if (r == 0. && i >= l)
                continue;
              d[l] -= p;
              e[l] = g;
              e[m] = 0.;
            }
        }
      while (m != l);
    }
  std::sort(d.begin(), d.end());
  return d;
}

template <typename T>
static void
cheb_estimate(dasm_cheb *c)
{
  dasm_op *       op  = c->op;
  dasm_ctx *      ctx = op->ctx;
  const long long n   = op->n_owned;
  cudaStream_t    s   = ctx->stream;
  T *v = dev_alloc<T>(op->n_vec), *w = dev_alloc<T>(op->n_vec), *t = dev_alloc<T>(op->n_vec), *pv = dev_alloc<T>(op->n_vec);
  CUDA_CHECK(cudaMemsetAsync(v, 0, op->n_vec * sizeof(T), s));
  // set_initial_guess: v_i = (global index) mod 11, minus mean; constrained DoFs zeroed.  The
  // global index is the local one plus the rank's offset (first_local_range) - with one rank 0.
  long long first = 0, n_global = n;
  if (ctx->n_ranks > 1)
    {
      // offsets via allgather of owned sizes
      std::vector<long long> sizes(ctx->n_ranks);
      long long *            d_sizes = dev_alloc<long long>(ctx->n_ranks);
      CUDA_CHECK(cudaMemcpyAsync(d_sizes + ctx->rank, &n, sizeof(long long), cudaMemcpyHostToDevice, s));
      NCCL_CHECK(ncclAllGather(d_sizes + ctx->rank, d_sizes, sizeof(long long), ncclChar, ctx->comm, s));
      CUDA_CHECK(cudaMemcpyAsync(sizes.data(), d_sizes, sizeof(long long) * ctx->n_ranks, cudaMemcpyDeviceToHost, s));
      CUDA_CHECK(cudaStreamSynchronize(s));
      cudaFree(d_sizes);
      n_global = 0;
      for (int r = 0; r < ctx->n_ranks; ++r)
        {
          if (r < ctx->rank)
            first += sizes[r];
          n_global += sizes[r];
        }
    }
  initial_guess_kernel<T><<<nblocks(n), 256, 0, s>>>(v, first, n);
  ctx->launches++;
  {
    // mean value
    T *ones = w;
    vec_scale_kernel<T><<<nblocks(n), 256, 0, s>>>(ones, v, T(0), n);
    vec_add_scalar_kernel<T><<<nblocks(n), 256, 0, s>>>(ones, T(1), n);
    const double sum = device_dot<T>(ctx, v, ones, n);
    vec_add_scalar_kernel<T><<<nblocks(n), 256, 0, s>>>(v, (T)(-sum / (double)n_global), n);
    ctx->launches += 3;
  }
  if (op->n_constrained > 0)
    {
      vec_set_indexed_kernel<T><<<nblocks(op->n_constrained), 256, 0, s>>>(v, T(0), op->d_constrained, op->n_constrained);
      ctx->launches++;
    }
  const bool symmetric = (c->fdm == nullptr) || dasm_fdm_is_symmetric(c->fdm);
  int        algo      = c->ev_algo;
  if (algo == DASM_EV_DEFAULT)
    algo = symmetric ? DASM_EV_LANCZOS : DASM_EV_POWER_ITERATION;
  if (algo == DASM_EV_POWER_ITERATION)
    {
      double lam = 0;
      double nrm = std::sqrt(device_dot<T>(ctx, v, v, n));
      vec_scale_kernel<T><<<nblocks(n), 256, 0, s>>>(v, v, (T)(1. / nrm), n);
      ctx->launches++;
      for (int it = 0; it < c->n_ev_it; ++it)
        {
          op_vmult<T>(op, t, v, nullptr, nullptr);
          precon_apply<T>(c, w, t, nullptr);
          lam = device_dot<T>(ctx, v, w, n);
          nrm = std::sqrt(device_dot<T>(ctx, w, w, n));
          vec_scale_kernel<T><<<nblocks(n), 256, 0, s>>>(v, w, (T)(1. / nrm), n);
          ctx->launches++;
        }
      cheb_set_ev(c, std::fabs(lam), 1.2 * std::fabs(lam));
    }
  else
    {
      // preconditioned CG on A x = v, Ritz values from the CG coefficients
      T *x = dev_alloc<T>(op->n_vec), *r = v, *z = w, *p = pv, *Ap = t;
      (void)x;
      precon_apply<T>(c, z, r, nullptr);
      CUDA_CHECK(cudaMemcpyAsync(p, z, op->n_vec * sizeof(T), cudaMemcpyDeviceToDevice, s));
      double              rz = device_dot<T>(ctx, r, z, n);
      const double        r0 = std::sqrt(device_dot<T>(ctx, r, r, n));
      std::vector<double> alphas, betas;
      for (int it = 0; it < c->n_ev_it; ++it)
        {
          op_vmult<T>(op, Ap, p, nullptr, nullptr);
          const double pAp = device_dot<T>(ctx, p, Ap, n);
          if (pAp == 0)
            break;
          const double alpha = rz / pAp;
          alphas.push_back(alpha);
          vec_axpy_kernel<T><<<nblocks(n), 256, 0, s>>>(r, (T)(-alpha), Ap, n);
          ctx->launches++;
          const double rn = std::sqrt(device_dot<T>(ctx, r, r, n));
          if (rn < 1e-10 * r0)
            break;
          precon_apply<T>(c, z, r, nullptr);
          const double rz_new = device_dot<T>(ctx, r, z, n);
          const double beta   = rz_new / rz;
          betas.push_back(beta);
          rz = rz_new;
          vec_xpay_kernel<T><<<nblocks(n), 256, 0, s>>>(p, (T)beta, z, n);
          ctx->launches++;
        }
      const int           kk = alphas.size();
      std::vector<double> d(kk), e(std::max(kk - 1, 0));
      for (int i = 0; i < kk; ++i)
        {
          d[i] = 1. / alphas[i] + (i > 0 ? betas[i - 1] / alphas[i - 1] : 0.);
          if (i + 1 < kk)
            e[i] = std::sqrt(betas[i]) / alphas[i];
        }
      const auto ev = tridiagonal_eigenvalues(d, e);
      DASM_REQUIRE(!ev.empty(), "eigenvalue estimation failed");
      cheb_set_ev(c, ev.front(), 1.2 * ev.back());
      cudaFree(x);
    }
  CUDA_CHECK(cudaStreamSynchronize(s));
  cudaFree(v);
  cudaFree(w);
  cudaFree(t);
  cudaFree(pv);
}

extern "C" int
dasm_cheb_create(dasm_op *op, dasm_fdm *fdm, int degree, double smoothing_range, int polynomial_type, int ev_algorithm, int optimize,
                 int eig_cg_n_iterations, dasm_cheb **out)
{
  DASM_API_BEGIN
  CUDA_CHECK(cudaSetDevice(op->ctx->device));
  DASM_REQUIRE(degree >= 1, "Chebyshev degree must be >= 1");
  DASM_REQUIRE(polynomial_type == DASM_POLY_FIRST_KIND || polynomial_type == DASM_POLY_FOURTH_KIND, "Polynomial type is not known!");
  DASM_REQUIRE(ev_algorithm >= 0 && ev_algorithm <= 2, "Eigen-value algorithm is not known!");
  DASM_REQUIRE(optimize >= 0 && optimize <= 3, "optimize level not implemented"); // templates.h:497-527
  DASM_REQUIRE(fdm == nullptr || fdm->op == op, "preconditioner belongs to another operator");
  auto c             = new dasm_cheb;
  c->op              = op;
  c->fdm             = fdm;
  c->degree          = degree;
  c->smoothing_range = smoothing_range > 0 ? smoothing_range : 20.;
  c->poly            = polynomial_type;
  c->ev_algo         = ev_algorithm;
  c->optimize        = optimize;
  c->n_ev_it         = eig_cg_n_iterations > 0 ? eig_cg_n_iterations : 40; // templates.h:109
  const size_t bytes = std::max<size_t>(1, (size_t)op->n_vec) * op->esize();
  for (void **p : {&c->t1, &c->t1b, &c->t2, &c->xold, &c->xin, &c->bin})
    {
      CUDA_CHECK(cudaMalloc(p, bytes));
      CUDA_CHECK(cudaMemsetAsync(*p, 0, bytes, op->ctx->stream));
    }
  if (fdm == nullptr)
    {
      CUDA_CHECK(cudaMalloc(&c->d_inv_diag, bytes));
      DISPATCH_TYPE(op->ntype, op_inverse_diagonal<T>(op, (T *)c->d_inv_diag));
    }
  *out = c;
  DASM_API_END
}

extern "C" dasm_op *dasm_cheb_op(const dasm_cheb *c) { return c->op; }
extern "C" void dasm_set_last_error(const char *msg) { g_last_error = msg; }

extern "C" int
dasm_cheb_destroy(dasm_cheb *c)
{
  DASM_API_BEGIN
  if (!c)
    return 0;
  cudaStreamSynchronize(c->op->ctx->stream);
  for (void *p : {c->t1, c->t1b, c->t2, c->xold, c->xin, c->bin, c->d_inv_diag, c->d_stage})
    cudaFree(p);
  if (c->pipe.ready)
    {
      cudaStreamSynchronize(c->pipe.h2d);
      cudaStreamSynchronize(c->pipe.d2h);
      for (int i = 0; i < 2; ++i)
        {
          cudaFree(c->pipe.x[i]);
          cudaFree(c->pipe.b[i]);
          cudaFree(c->pipe.sx[i]);
          cudaFree(c->pipe.sb[i]);
          cudaEventDestroy(c->pipe.ev_h2d[i]);
          cudaEventDestroy(c->pipe.ev_run[i]);
          cudaEventDestroy(c->pipe.ev_d2h[i]);
        }
      cudaStreamDestroy(c->pipe.h2d);
      cudaStreamDestroy(c->pipe.d2h);
    }
  delete c;
  DASM_API_END
}

extern "C" int
dasm_cheb_estimate_eigenvalues(dasm_cheb *c, double *min_ev, double *max_ev)
{
  DASM_API_BEGIN
  CUDA_CHECK(cudaSetDevice(c->op->ctx->device));
  DISPATCH_TYPE(c->op->ntype, cheb_estimate<T>(c));
  if (min_ev)
    *min_ev = c->min_ev;
  if (max_ev)
    *max_ev = c->max_ev;
  DASM_API_END
}

extern "C" int
dasm_cheb_set_eigenvalues(dasm_cheb *c, double min_ev, double max_ev)
{
  DASM_API_BEGIN
  cheb_set_ev(c, min_ev, max_ev);
  DASM_API_END
}

// x (in/out) = dst.  first_is_step: x1 = x0 + f2 P^-1 (b - A x0), else x1 = f2 P^-1 b.
template <typename T>
static void
cheb_run(dasm_cheb *c, T *x_user, const T *b, bool first_is_step)
{
  dasm_op *       op  = c->op;
  dasm_ctx *      ctx = op->ctx;
  const long long n   = op->n_owned;
  if (!c->ev_ready)
    cheb_estimate<T>(c);
  op->last_compressed = nullptr; // the first sweep waits for the main stream
  T *t1 = (T *)c->t1, *t2 = (T *)c->t2;
  const double f2_0 = (c->poly == DASM_POLY_FOURTH_KIND) ? 4. / (3. * c->theta) : 1. / c->theta;
  const int    n_terms = (c->degree < 2 || std::fabs(c->delta) < 1e-40) ? 1 : c->degree;

  // buffers: cur = current iterate, old = previous iterate.  The new iterate is produced in t2 by
  // the preconditioner sweep + update and the three buffers rotate; the last term writes to x_user.
  T *cur = x_user, *old = (T *)c->xold, *nxt = t2, *spare = nullptr;
  (void)spare;
  dasm_hook res_hook  = {DASM_HOOK_RESIDUAL, 0, 0, b, nullptr};
  bool      have_old  = false;
  double    rho_old   = 0, sigma = 0;
  if (n_terms > 1)
    {
      sigma   = c->theta / c->delta;
      rho_old = 1. / sigma;
    }
  // Fully fused sequence (optimize 2 with the FDM brick preconditioner): DoFs on shared brick faces receive
  // their contributions by red.add directly in the destination, which the PREVIOUS kernel zeroed there (the
  // owning brick adds the part of the epilogue that does not depend on the operator result); no pass over the
  // shared DoFs remains except one at the start of the step.
  const bool direct = op->use_brick && op->shared_ranges_ok && c->fdm != nullptr && fdm_uses_brick(c->fdm) && c->optimize >= 2 &&
                      !(getenv("DASM_NO_DIRECT") && getenv("DASM_NO_DIRECT")[0] == '1');
  T *t1buf[2] = {t1, (T *)c->t1b};
  // the residual buffers alternate; the sequence starts with the one the previous fused call left zeroed on the shared DoFs
  const int t1_off = (direct && c->t1_zero_idx >= 0) ? c->t1_zero_idx : 0;
  const bool t1_ready = direct && c->t1_zero_idx >= 0;
  c->t1_zero_idx = -1;
  for (int term = 0; term < n_terms; ++term)
    {
      double f1 = 0, f2 = f2_0;
      if (term > 0)
        {
          const int j = term - 1;
          if (c->poly == DASM_POLY_FOURTH_KIND)
            {
              f1 = (2 * j + 1.) / (2 * j + 5.);
              f2 = (8 * j + 12.) / (c->theta * (2 * j + 5.));
            }
          else
            {
              const double rho = 1. / (2. * sigma - rho_old);
              f1               = rho * rho_old;
              f2               = 2. * rho / c->delta;
              rho_old          = rho;
            }
        }
      const bool  need_A   = !(term == 0 && !first_is_step);
      const bool  last     = (term == n_terms - 1);
      T *         t1cur    = t1buf[(term + t1_off) & 1];
      const T *   old_used = (have_old && f1 != 0.) ? old : nullptr;
      dasm_hook   upd;
      if (!need_A)
        upd = {DASM_HOOK_SCALE, 0, f2, nullptr, nullptr};
      else
        upd = {DASM_HOOK_CHEB_UPDATE, f1, f2, cur, old_used};
      if (direct)
        {
          // every kernel zeroes the destination of the NEXT kernel on the shared DoFs (the brick that owns a shared DoF adds the
          // base of the epilogue together with its contribution, kernels_brick.cuh SHARED_DIRECT)
          NextInit<T> ni_upd = no_next_init<T>();
          ni_upd.out         = nxt;
          NextInit<T> ni_res = no_next_init<T>();
          ni_res.out         = t1buf[(term + 1 + t1_off) & 1]; // (also after the last term: ready for the next call)
          const T *rhs_for_P = b;
          if (need_A)
            {
              if (term == 0 && !t1_ready)
                {
                  NextInit<T> first = ni_res;
                  first.out         = t1cur;
                  init_shared<T>(op, first);
                }
              op_vmult_brick<T>(op, t1cur, cur, &res_hook, SHARED_DIRECT, ni_upd); // t1 = b - A cur
              rhs_for_P = t1cur;
            }
          else
            init_shared<T>(op, ni_upd);
          fdm_vmult_brick<T>(c->fdm, nxt, rhs_for_P, &upd, SHARED_DIRECT, ni_res);
          if (last)
            c->t1_zero_idx = (term + 1 + t1_off) & 1;
        }
      else
        {
          const T *rhs_for_P;
          if (!need_A)
            rhs_for_P = b; // x0 = 0: residual is b
          else
            {
              op_vmult<T>(op, t1, cur, nullptr, &res_hook); // t1 = b - A cur
              rhs_for_P = t1;
            }
          precon_apply<T>(c, nxt, rhs_for_P, &upd); // nxt = P^-1 rhs, then nxt = (1+f1) cur - f1 old + f2 nxt
        }
      // rotate: old <- cur, cur <- nxt, nxt <- (old buffer)
      T *tmp   = old;
      old      = cur;
      cur      = nxt;
      nxt      = tmp;
      have_old = (term > 0) || first_is_step;
    }
  if (cur != x_user)
    {
      CUDA_CHECK(cudaMemcpyAsync(x_user, cur, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
    }
  CUDA_CHECK(cudaGetLastError());
}

extern "C" int
dasm_cheb_vmult(dasm_cheb *c, void *dst, const void *src)
{
  DASM_API_BEGIN
  CUDA_CHECK(cudaSetDevice(c->op->ctx->device));
  DISPATCH_TYPE(c->op->ntype, cheb_run<T>(c, (T *)dst, (const T *)src, false));
  DASM_API_END
}

extern "C" int
dasm_cheb_step(dasm_cheb *c, void *dst, const void *src)
{
  DASM_API_BEGIN
  CUDA_CHECK(cudaSetDevice(c->op->ctx->device));
  DISPATCH_TYPE(c->op->ntype, cheb_run<T>(c, (T *)dst, (const T *)src, true));
  DASM_API_END
}

template <typename T>
static void
cheb_host(dasm_cheb *c, double *dst, const double *src, bool step)
{
  dasm_op *       op = c->op;
  cudaStream_t    s  = op->ctx->stream;
  const long long n  = op->n_owned;
  // host vectors are double (the reference's outer vectors): copied straight into the device vectors in double precision,
  // through a staging buffer owned by the smoother (allocated on its device) and a conversion kernel in single precision
  T *x = (T *)c->xin, *b = (T *)c->bin;
  if (sizeof(T) == sizeof(double))
    {
      CUDA_CHECK(cudaMemcpyAsync(b, src, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
      if (step)
        CUDA_CHECK(cudaMemcpyAsync(x, dst, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
      cheb_run<T>(c, x, b, step);
      CUDA_CHECK(cudaMemcpyAsync(dst, x, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
  else
    {
      if (c->d_stage == nullptr)
        CUDA_CHECK(cudaMalloc(&c->d_stage, (size_t)n * sizeof(double)));
      double *d_stage = (double *)c->d_stage;
      CUDA_CHECK(cudaMemcpyAsync(d_stage, src, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
      vec_convert_kernel<T, double><<<nblocks(n), 256, 0, s>>>(b, d_stage, n);
      op->ctx->launches++;
      if (step)
        {
          CUDA_CHECK(cudaMemcpyAsync(d_stage, dst, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
          vec_convert_kernel<T, double><<<nblocks(n), 256, 0, s>>>(x, d_stage, n);
          op->ctx->launches++;
        }
      cheb_run<T>(c, x, b, step);
      vec_convert_kernel<double, T><<<nblocks(n), 256, 0, s>>>(d_stage, x, n);
      op->ctx->launches++;
      CUDA_CHECK(cudaMemcpyAsync(dst, d_stage, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
  CUDA_CHECK(cudaStreamSynchronize(s));
}

// n independent problems (x_i, b_i) on one smoother: x_i <- step(x_i, b_i) (or x_i = vmult(b_i)).  The host -> device copies of problem
// i + 1 and the device -> host copy of problem i - 1 run on their own streams next to the kernels of problem i (PCIe is full duplex),
// so that a step costs max(copy in, copy out, kernels) instead of their sum.  Pinned host buffers are needed for the overlap.
template <typename T>
static void
cheb_host_batch(dasm_cheb *c, const int n_steps, double *const *dst, const double *const *src, const bool step)
{
  dasm_op *       op = c->op;
  cudaStream_t    s  = op->ctx->stream;
  const long long n  = op->n_owned;
  const bool      f64 = sizeof(T) == sizeof(double);
  auto &          P   = c->pipe;
  if (!P.ready)
    {
      CUDA_CHECK(cudaStreamCreateWithFlags(&P.h2d, cudaStreamNonBlocking));
      CUDA_CHECK(cudaStreamCreateWithFlags(&P.d2h, cudaStreamNonBlocking));
      for (int i = 0; i < 2; ++i)
        {
          CUDA_CHECK(cudaEventCreateWithFlags(&P.ev_h2d[i], cudaEventDisableTiming));
          CUDA_CHECK(cudaEventCreateWithFlags(&P.ev_run[i], cudaEventDisableTiming));
          CUDA_CHECK(cudaEventCreateWithFlags(&P.ev_d2h[i], cudaEventDisableTiming));
          CUDA_CHECK(cudaMalloc(&P.x[i], std::max<size_t>(1, (size_t)op->n_vec) * sizeof(T)));
          CUDA_CHECK(cudaMalloc(&P.b[i], std::max<size_t>(1, (size_t)op->n_vec) * sizeof(T)));
          CUDA_CHECK(cudaMemset(P.x[i], 0, std::max<size_t>(1, (size_t)op->n_vec) * sizeof(T)));
          CUDA_CHECK(cudaMemset(P.b[i], 0, std::max<size_t>(1, (size_t)op->n_vec) * sizeof(T)));
          if (!f64)
            {
              CUDA_CHECK(cudaMalloc(&P.sx[i], std::max<size_t>(1, (size_t)n) * sizeof(double)));
              CUDA_CHECK(cudaMalloc(&P.sb[i], std::max<size_t>(1, (size_t)n) * sizeof(double)));
            }
        }
      P.ready = true;
    }
  CUDA_CHECK(cudaStreamSynchronize(s)); // (earlier work on the smoother's stream is finished before the copy streams start)
  for (int i = 0; i < n_steps; ++i)
    {
      const int slot = i & 1;
      T *       x = (T *)P.x[slot], *b = (T *)P.b[slot];
      double *  hx = f64 ? (double *)x : P.sx[slot], *hb = f64 ? (double *)b : P.sb[slot];
      if (i >= 2)
        CUDA_CHECK(cudaStreamWaitEvent(P.h2d, P.ev_d2h[slot], 0)); // the slot's previous result has left the device
      CUDA_CHECK(cudaMemcpyAsync(hb, src[i], (size_t)n * sizeof(double), cudaMemcpyHostToDevice, P.h2d));
      if (step)
        CUDA_CHECK(cudaMemcpyAsync(hx, dst[i], (size_t)n * sizeof(double), cudaMemcpyHostToDevice, P.h2d));
      CUDA_CHECK(cudaEventRecord(P.ev_h2d[slot], P.h2d));
      CUDA_CHECK(cudaStreamWaitEvent(s, P.ev_h2d[slot], 0));
      if (!f64)
        {
          vec_convert_kernel<T, double><<<nblocks(n), 256, 0, s>>>(b, hb, n);
          if (step)
            vec_convert_kernel<T, double><<<nblocks(n), 256, 0, s>>>(x, hx, n);
          op->ctx->launches += step ? 2 : 1;
        }
      cheb_run<T>(c, x, b, step);
      if (!f64)
        {
          vec_convert_kernel<double, T><<<nblocks(n), 256, 0, s>>>(hx, x, n);
          op->ctx->launches++;
        }
      CUDA_CHECK(cudaEventRecord(P.ev_run[slot], s));
      CUDA_CHECK(cudaStreamWaitEvent(P.d2h, P.ev_run[slot], 0));
      CUDA_CHECK(cudaMemcpyAsync(dst[i], hx, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, P.d2h));
      CUDA_CHECK(cudaEventRecord(P.ev_d2h[slot], P.d2h));
    }
  CUDA_CHECK(cudaStreamSynchronize(P.d2h));
  CUDA_CHECK(cudaStreamSynchronize(P.h2d));
  CUDA_CHECK(cudaStreamSynchronize(s));
}

extern "C" int
dasm_cheb_step_host_batch(dasm_cheb *c, int n_steps, double *const *dst_owned, const double *const *src_owned)
{
  DASM_API_BEGIN
  DASM_REQUIRE(n_steps >= 0 && (n_steps == 0 || (dst_owned != nullptr && src_owned != nullptr)), "step_host_batch: invalid arguments");
  CUDA_CHECK(cudaSetDevice(c->op->ctx->device));
  DISPATCH_TYPE(c->op->ntype, cheb_host_batch<T>(c, n_steps, dst_owned, src_owned, true));
  DASM_API_END
}

extern "C" int
dasm_cheb_vmult_host_batch(dasm_cheb *c, int n_steps, double *const *dst_owned, const double *const *src_owned)
{
  DASM_API_BEGIN
  DASM_REQUIRE(n_steps >= 0 && (n_steps == 0 || (dst_owned != nullptr && src_owned != nullptr)), "vmult_host_batch: invalid arguments");
  CUDA_CHECK(cudaSetDevice(c->op->ctx->device));
  DISPATCH_TYPE(c->op->ntype, cheb_host_batch<T>(c, n_steps, dst_owned, src_owned, false));
  DASM_API_END
}

extern "C" int
dasm_cheb_step_host(dasm_cheb *c, double *dst_owned, const double *src_owned)
{
  DASM_API_BEGIN
  CUDA_CHECK(cudaSetDevice(c->op->ctx->device));
  DISPATCH_TYPE(c->op->ntype, cheb_host<T>(c, dst_owned, src_owned, true));
  DASM_API_END
}

extern "C" int
dasm_cheb_vmult_host(dasm_cheb *c, double *dst_owned, const double *src_owned)
{
  DASM_API_BEGIN
  CUDA_CHECK(cudaSetDevice(c->op->ctx->device));
  DISPATCH_TYPE(c->op->ntype, cheb_host<T>(c, dst_owned, src_owned, false));
  DASM_API_END
}

// ------------------------------------------------------------------------------------------------
// C ABI: Krylov solvers (SolverCG / SolverGMRES of the reference's solve(), element_centered_preconditioners_01.cc:108-203)
// ------------------------------------------------------------------------------------------------
template <typename T>
static void
krylov_precon(dasm_op *op, const int kind, void *handle, const T *inv_diag, T *z, const T *r)
{
  dasm_ctx *      ctx = op->ctx;
  const long long n   = op->n_owned;
  switch (kind)
    {
      case DASM_PRECON_IDENTITY:
        CUDA_CHECK(cudaMemcpyAsync(z, r, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
        break;
      case DASM_PRECON_DIAGONAL:
        vec_mul_kernel<T><<<nblocks(n), 256, 0, ctx->stream>>>(z, inv_diag, r, n);
        ctx->launches++;
        break;
      case DASM_PRECON_FDM:
        fdm_vmult<T>((dasm_fdm *)handle, z, r, nullptr, nullptr);
        break;
      case DASM_PRECON_CHEBYSHEV:
        cheb_run<T>((dasm_cheb *)handle, z, r, false);
        break;
      case DASM_PRECON_MULTIGRID:
        if (dasm_mg_vmult_outer((dasm_mg *)handle, z, r, op->ntype) != 0)
          throw std::runtime_error(dasm_last_error());
        break;
      case DASM_PRECON_BLOCK_ASM:
        if (dasm_asm_vmult((dasm_asm *)handle, z, r) != 0)
          throw std::runtime_error(dasm_last_error());
        break;
      default:
        throw std::runtime_error("Preconditioner kind is not known!");
    }
}

template <typename T>
static void
krylov_solve(dasm_op *op, const int solver, const int pkind, void *ph, T *x, const T *b, const int max_it, const double abs_tol,
             const double rel_tol, int restart, int *n_it, double *residual)
{
  dasm_ctx *      ctx = op->ctx;
  cudaStream_t    s   = ctx->stream;
  const long long n   = op->n_owned;
  const size_t    nv  = (size_t)op->n_vec;
  // work vectors come from a pool kept by the operator: a solve after the first one allocates nothing (cudaMalloc / cudaFree of
  // up to 35 vectors per call cost more than the iterations of a multigrid-preconditioned solve)
  size_t pool_used = 0;
  auto   alloc     = [&]() {
    if (pool_used == op->krylov_pool.size())
      {
        T *q = dev_alloc<T>(nv);
        op->krylov_pool.push_back(q);
        op->scratch.push_back(q);
      }
    T *p = (T *)op->krylov_pool[pool_used++];
    CUDA_CHECK(cudaMemsetAsync(p, 0, nv * sizeof(T), s));
    return p;
  };
  T *inv_diag = nullptr;
  if (pkind == DASM_PRECON_DIAGONAL)
    {
      inv_diag = alloc();
      op_inverse_diagonal<T>(op, inv_diag);
    }
  CUDA_CHECK(cudaMemsetAsync(x, 0, nv * sizeof(T), s));
  const double r0     = std::sqrt(device_dot<T>(ctx, b, b, n));
  const double target = std::max(abs_tol, rel_tol * r0);
  int          its    = 0;
  double       res    = r0;
  bool         done   = r0 <= target;
  if (solver == DASM_SOLVER_CG && !done)
    {
      T *r = alloc(), *z = alloc(), *p = alloc(), *Ap = alloc();
      CUDA_CHECK(cudaMemcpyAsync(r, b, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, s));
      krylov_precon<T>(op, pkind, ph, inv_diag, z, r);
      CUDA_CHECK(cudaMemcpyAsync(p, z, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, s));
      double rz = device_dot<T>(ctx, r, z, n);
      while (its < max_it)
        {
          op_vmult<T>(op, Ap, p, nullptr, nullptr);
          const double alpha = rz / device_dot<T>(ctx, p, Ap, n);
          vec_axpy_kernel<T><<<nblocks(n), 256, 0, s>>>(x, (T)alpha, p, n);
          vec_axpy_kernel<T><<<nblocks(n), 256, 0, s>>>(r, (T)(-alpha), Ap, n);
          ctx->launches += 2;
          ++its;
          res = std::sqrt(device_dot<T>(ctx, r, r, n));
          if (res <= target)
            {
              done = true;
              break;
            }
          krylov_precon<T>(op, pkind, ph, inv_diag, z, r);
          const double rz_new = device_dot<T>(ctx, r, z, n);
          vec_xpay_kernel<T><<<nblocks(n), 256, 0, s>>>(p, (T)(rz_new / rz), z, n);
          ctx->launches++;
          rz = rz_new;
        }
    }
  else if (!done)
    {
      // GMRES(restart) with right preconditioning; Z_j = M^-1 v_j kept, so that x += Z y needs no further preconditioner call
      restart = std::max(1, restart);
      std::vector<T *> V(restart + 1), Z(restart);
      for (auto &p : V)
        p = alloc();
      for (auto &p : Z)
        p = alloc();
      T *                 w = alloc();
      std::vector<double> H((size_t)(restart + 1) * restart), g(restart + 1), cs(restart), sn(restart);
      while (its < max_it && !done)
        {
          // r = b - A x
          op_vmult<T>(op, w, x, nullptr, nullptr);
          vec_residual_kernel<T><<<nblocks(n), 256, 0, s>>>(w, b, n); // w = b - w
          ctx->launches++;
          const double beta = std::sqrt(device_dot<T>(ctx, w, w, n));
          vec_scale_kernel<T><<<nblocks(n), 256, 0, s>>>(V[0], w, (T)(1. / beta), n);
          ctx->launches++;
          std::fill(H.begin(), H.end(), 0.);
          std::fill(g.begin(), g.end(), 0.);
          g[0]   = beta;
          int kk = 0;
          for (int j = 0; j < restart; ++j)
            {
              krylov_precon<T>(op, pkind, ph, inv_diag, Z[j], V[j]);
              op_vmult<T>(op, w, Z[j], nullptr, nullptr);
              for (int i = 0; i <= j; ++i)
                {
                  const double h         = device_dot<T>(ctx, w, V[i], n);
                  H[(size_t)i * restart + j] = h;
                  vec_axpy_kernel<T><<<nblocks(n), 256, 0, s>>>(w, (T)(-h), V[i], n);
                  ctx->launches++;
                }
              const double hn                  = std::sqrt(device_dot<T>(ctx, w, w, n));
              H[(size_t)(j + 1) * restart + j] = hn;
              vec_scale_kernel<T><<<nblocks(n), 256, 0, s>>>(V[j + 1], w, (T)(hn > 0 ? 1. / hn : 0.), n);
              ctx->launches++;
              for (int i = 0; i < j; ++i)
                {
                  const double t                   = cs[i] * H[(size_t)i * restart + j] + sn[i] * H[(size_t)(i + 1) * restart + j];
                  H[(size_t)(i + 1) * restart + j] = -sn[i] * H[(size_t)i * restart + j] + cs[i] * H[(size_t)(i + 1) * restart + j];
                  H[(size_t)i * restart + j]       = t;
                }
              const double d = std::hypot(H[(size_t)j * restart + j], H[(size_t)(j + 1) * restart + j]);
              cs[j]          = H[(size_t)j * restart + j] / d;
              sn[j]          = H[(size_t)(j + 1) * restart + j] / d;
              H[(size_t)j * restart + j]       = d;
              H[(size_t)(j + 1) * restart + j] = 0;
              g[j + 1]       = -sn[j] * g[j];
              g[j]           = cs[j] * g[j];
              ++its;
              kk  = j + 1;
              res = std::fabs(g[j + 1]);
              if (res <= target || its >= max_it)
                {
                  done = res <= target;
                  break;
                }
            }
          // back substitution and update
          std::vector<double> y(kk);
          for (int i = kk - 1; i >= 0; --i)
            {
              double v = g[i];
              for (int l = i + 1; l < kk; ++l)
                v -= H[(size_t)i * restart + l] * y[l];
              y[i] = v / H[(size_t)i * restart + i];
            }
          for (int i = 0; i < kk; ++i)
            {
              vec_axpy_kernel<T><<<nblocks(n), 256, 0, s>>>(x, (T)y[i], Z[i], n);
              ctx->launches++;
            }
          if (its >= max_it)
            break;
        }
    }
  CUDA_CHECK(cudaStreamSynchronize(s));
  if (n_it)
    *n_it = its;
  if (residual)
    *residual = res;
  if (!done)
    throw std::runtime_error("SolverControl::NoConvergence: the solver did not converge in " + std::to_string(max_it) + " iterations");
}

extern "C" int
dasm_solve(dasm_op *op, int solver, int precon_kind, void *precon, void *x, const void *b, int max_it, double abs_tol, double rel_tol,
           int restart, int *n_it, double *residual)
{
  DASM_API_BEGIN
  DASM_REQUIRE(solver == DASM_SOLVER_CG || solver == DASM_SOLVER_GMRES, "Solver is not known!");
  DASM_REQUIRE((precon_kind == DASM_PRECON_FDM || precon_kind == DASM_PRECON_CHEBYSHEV || precon_kind == DASM_PRECON_MULTIGRID ||
                precon_kind == DASM_PRECON_BLOCK_ASM) == (precon != nullptr),
               "preconditioner handle does not match its kind");
  CUDA_CHECK(cudaSetDevice(op->ctx->device));
  DISPATCH_TYPE(op->ntype, krylov_solve<T>(op, solver, precon_kind, precon, (T *)x, (const T *)b, max_it, abs_tol, rel_tol, restart, n_it,
                                           residual));
  DASM_API_END
}

// ------------------------------------------------------------------------------------------------
// Power kernel (power_kernel_01.likwid.cc:122-308, 479-599): dst_0 = A src (Laplace), dst_1 = M dst_0 (mass operator on the same
// cells), "sequential" = two sweeps over all cells, "power" = the second operator runs on a cell as soon as every cell that
// contributes to its dst_0 entries has been processed (determine_pre_post), so that dst_0 is still on chip.  On the device a wave of
// `cell_granularity` cells is one launch of the generic cell kernel on a cell range, followed by one launch of the second operator on
// the list of cells that became complete (dst_0 of the last waves then sits in the 126 MB L2).  `batch_size` > 1 releases cells in
// batches (the "use matrix-free batches" variant), 1 tracks every cell (the "own batches" variant).
// ------------------------------------------------------------------------------------------------
struct dasm_power
{
  dasm_op *              op;
  long long              granularity;
  std::vector<long long> wave_first;      // cells [wave_first[w], wave_first[w+1])
  std::vector<long long> post_ptr;        // post cells of wave w: post_ids[post_ptr[w] .. post_ptr[w+1])
  uint32_t *             d_post_ids = nullptr;
  double                 cell_volume = 1;
};

extern "C" int
dasm_power_create(dasm_op *op, long long cell_granularity, int batch_size, dasm_power **out)
{
  DASM_API_BEGIN
  DASM_REQUIRE(op->geom_mode == 0 && op->mesh != nullptr, "power kernel: Cartesian structured mesh expected (MappingQ1 on a hyper-cube in the reference)");
  DASM_REQUIRE(op->mesh->mesh->n_ranks() == 1, "power kernel: one rank");
  CUDA_CHECK(cudaSetDevice(op->ctx->device));
  auto p         = new dasm_power;
  p->op          = op;
  const long long nc = op->n_cells;
  p->granularity = (cell_granularity <= 0 || cell_granularity > nc) ? nc : cell_granularity;
  batch_size     = std::max(batch_size, 1);
  const Mesh &M  = *op->mesh->mesh;
  p->cell_volume = M.h(0) * M.h(1) * M.h(2);
  const long long n_waves = (nc + p->granularity - 1) / p->granularity;
  for (long long w = 0; w <= n_waves; ++w)
    p->wave_first.push_back(std::min(w * p->granularity, nc));
  // last wave that touches an entity with DoFs (the 27 start indices identify the entities; constrained ones are not accessed)
  const int                               k = op->k;
  std::unordered_map<uint32_t, long long> last;
  auto has_dofs = [&](const int e) { return k > 1 || ((e % 3 != 1) && ((e / 3) % 3 != 1) && (e / 9 != 1)); };
  for (long long c = 0; c < nc; ++c)
    for (int e = 0; e < 27; ++e)
      {
        const uint32_t s = op->nb.cidx[c * 27 + e];
        if (s != INVALID_INDEX && has_dofs(e))
          last[s] = c / p->granularity; // (cells are visited in increasing wave order)
      }
  std::vector<long long> ready(nc, 0);
  for (long long c = 0; c < nc; ++c)
    {
      long long r = c / p->granularity;
      for (int e = 0; e < 27; ++e)
        {
          const uint32_t s = op->nb.cidx[c * 27 + e];
          if (s != INVALID_INDEX && has_dofs(e))
            r = std::max(r, last[s]);
        }
      ready[c] = r;
    }
  if (batch_size > 1)
    for (long long c0 = 0; c0 < nc; c0 += batch_size)
      {
        long long r = 0;
        for (long long c = c0; c < std::min(nc, c0 + batch_size); ++c)
          r = std::max(r, ready[c]);
        for (long long c = c0; c < std::min(nc, c0 + batch_size); ++c)
          ready[c] = r;
      }
  std::vector<uint32_t> ids(nc);
  p->post_ptr.assign(n_waves + 1, 0);
  for (long long c = 0; c < nc; ++c)
    p->post_ptr[ready[c] + 1]++;
  for (long long w = 0; w < n_waves; ++w)
    p->post_ptr[w + 1] += p->post_ptr[w];
  std::vector<long long> fill(p->post_ptr.begin(), p->post_ptr.end() - 1);
  for (long long c = 0; c < nc; ++c)
    ids[fill[ready[c]]++] = (uint32_t)c;
  p->d_post_ids = dev_upload(ids, op->ctx->stream);
  *out          = p;
  DASM_API_END
}

template <typename T>
static void
power_launch(dasm_op *op, const int geom, const double cell_volume, T *dst, const T *src, const long long first, const long long count,
             const uint32_t *cell_ids)
{
  if (count <= 0)
    return;
  dasm_ctx *ctx = op->ctx;
  DISPATCH_DEGREE(op->k, {
    constexpr int  n = K + 1, CPB = cells_per_block<K>();
    const size_t   smem = (size_t)CPB * 4 * n * n * n * sizeof(T);
    const unsigned grid = (unsigned)((count + CPB - 1) / CPB);
    const uint32_t *ci  = cell_ids ? op->d_cidx : op->d_cidx + first * 27;
    CartesianCoef   cc  = op->cart;
    if (geom == 0)
      {
        auto kern = laplace_generic_kernel<K, T, 0>;
        CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, CPB * n * n, smem, ctx->stream>>>(src, dst, ci, (const T *)nullptr, cc, count, nullptr, cell_ids);
      }
    else if (geom == 4)
      {
        cc.g[0]   = cell_volume;
        auto kern = laplace_generic_kernel<K, T, 4>;
        CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, CPB * n * n, smem, ctx->stream>>>(src, dst, ci, (const T *)nullptr, cc, count, nullptr, cell_ids);
      }
    else
      {
        auto kern = laplace_generic_kernel<K, T, 5>;
        CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, CPB * n * n, smem, ctx->stream>>>(src, dst, ci, (const T *)nullptr, cc, count, nullptr, cell_ids);
      }
  });
  ctx->launches++;
}

// fused != 0: power kernel; 0: sequential.  The results are ADDED to dst_0 / dst_1 (distribute_local_to_global, as in the reference,
// which zeroes the two vectors once before the repetitions, power_kernel_01.likwid.cc:443-445)
extern "C" int
dasm_power_run(dasm_power *p, void *dst_0, void *dst_1, const void *src, int fused, int do_computation)
{
  DASM_API_BEGIN
  dasm_op *op = p->op;
  CUDA_CHECK(cudaSetDevice(op->ctx->device));
  const int g0 = do_computation ? 0 : 5, g1 = do_computation ? 4 : 5;
  DISPATCH_TYPE(op->ntype, {
    if (!fused)
      {
        for (size_t w = 0; w + 1 < p->wave_first.size(); ++w)
          power_launch<T>(op, g0, p->cell_volume, (T *)dst_0, (const T *)src, p->wave_first[w], p->wave_first[w + 1] - p->wave_first[w], nullptr);
        for (size_t w = 0; w + 1 < p->wave_first.size(); ++w)
          power_launch<T>(op, g1, p->cell_volume, (T *)dst_1, (const T *)dst_0, p->wave_first[w], p->wave_first[w + 1] - p->wave_first[w], nullptr);
      }
    else
      for (size_t w = 0; w + 1 < p->wave_first.size(); ++w)
        {
          power_launch<T>(op, g0, p->cell_volume, (T *)dst_0, (const T *)src, p->wave_first[w], p->wave_first[w + 1] - p->wave_first[w], nullptr);
          power_launch<T>(op, g1, p->cell_volume, (T *)dst_1, (const T *)dst_0, 0, p->post_ptr[w + 1] - p->post_ptr[w], p->d_post_ids + p->post_ptr[w]);
        }
  });
  CUDA_CHECK(cudaGetLastError());
  DASM_API_END
}

// number of cells the second operator processes after wave w (the post_indices_ptr of determine_pre_post), for inspection
extern "C" long long
dasm_power_n_waves(const dasm_power *p)
{
  return (long long)p->wave_first.size() - 1;
}

extern "C" long long
dasm_power_post_count(const dasm_power *p, long long wave)
{
  return (wave >= 0 && wave + 1 < (long long)p->post_ptr.size()) ? p->post_ptr[wave + 1] - p->post_ptr[wave] : -1;
}

extern "C" int
dasm_power_destroy(dasm_power *p)
{
  DASM_API_BEGIN
  if (p)
    {
      cudaFree(p->d_post_ids);
      delete p;
    }
  DASM_API_END
}
