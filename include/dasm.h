/* libdasm - C ABI of the B200-native multigrid-smoother hot path of dealii-asm.
 *
 * The reference (peterrum/dealii-asm) has no FFI layer: its boundary is the C++ class surface
 * LaplaceOperatorMatrixFree / ASPoissonPreconditioner / PreconditionChebyshev built by the factories
 * create_fdm_preconditioner / create_system_preconditioner.  Every entry point below cites the
 * reference interface it replaces (file:line into the reference tree); the C++ mirror classes with
 * the reference's names live in include/dasm/ (operator.h, preconditioners.h, precondition.h) and are
 * thin wrappers over these calls.  INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions: plain pointers and sizes only; every call returns 0 on success and non-zero on
 * error with the message available from dasm_last_error() (the C++ wrappers rethrow it, mirroring
 * deal.II's AssertThrow).  Vector arguments named dst/src/vec are DEVICE pointers to
 * dasm_op_vec_size() numbers of the operator's number type (locally owned DoFs first, then ghost
 * DoFs, like LinearAlgebra::distributed::Vector); the *_host variants take HOST pointers of the
 * locally owned size and include the host<->device copies.  All work is enqueued on the context's
 * CUDA stream; dasm_ctx_sync() waits for it.  Objects are not re-entrant (one host thread per
 * context), matching the reference's mutable scratch members (matrix_free.h:1545-1564).
 */
#ifndef DASM_H
#define DASM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

  typedef struct dasm_ctx  dasm_ctx;
  typedef struct dasm_mesh dasm_mesh;
  typedef struct dasm_op   dasm_op;
  typedef struct dasm_fdm  dasm_fdm;
  typedef struct dasm_cheb dasm_cheb;

  enum dasm_number_type
  {
    DASM_F64 = 0,
    DASM_F32 = 1
  };
  /* Restrictors::WeightingType, include/restrictors.h:8-15 */
  enum dasm_weight_type
  {
    DASM_WEIGHT_NONE = 0,
    DASM_WEIGHT_PRE  = 1,
    DASM_WEIGHT_POST = 2,
    DASM_WEIGHT_RAS  = 3,
    DASM_WEIGHT_SYMM = 4
  };
  /* "weight sequence", include/precondition.templates.h:201-203 */
  enum dasm_weight_sequence
  {
    DASM_WSEQ_GLOBAL     = 0,
    DASM_WSEQ_LOCAL      = 1,
    DASM_WSEQ_DG         = 2,
    DASM_WSEQ_COMPRESSED = 3
  };
  enum dasm_map_kind
  {
    DASM_MAP_CARTESIAN = 0, /* use cartesian mesh = true,  matrix_free_loop_08.likwid.cc:185-191 */
    DASM_MAP_SINE      = 1, /* use cartesian mesh = false, matrix_free_loop_08.likwid.cc:193-199 */
    DASM_MAP_KERSHAW   = 2  /* include/kershaw.h:39-80; map_params = {eps_y, eps_z}            */
  };
  enum dasm_polynomial_type
  {
    DASM_POLY_FIRST_KIND  = 0,
    DASM_POLY_FOURTH_KIND = 1
  };
  enum dasm_ev_algorithm
  {
    DASM_EV_LANCZOS         = 0,
    DASM_EV_POWER_ITERATION = 1,
    DASM_EV_DEFAULT         = 2 /* lanczos if A and P symmetric else power iteration (templates.h:113-114) */
  };
  /* The pre/post DoF-range hooks of the reference are host lambdas
   * (std::function<void(unsigned,unsigned)>, operator.h:1367-1373, matrix_free.h:960-986); a kernel
   * cannot call them, so the updates PreconditionChebyshev/PreconditionRelaxation actually pass are
   * enumerated here and fused into the kernels. */
  enum dasm_hook_kind
  {
    DASM_HOOK_NONE        = 0,
    DASM_HOOK_ZERO_DST    = 1, /* pre : dst = 0                                            */
    DASM_HOOK_RESIDUAL    = 2, /* post: dst = v0 - dst                      (t = b - A x)   */
    DASM_HOOK_CHEB_UPDATE = 3, /* post: dst = (1+f1) v0 - f1 v1 + f2 dst    (x+ from P^-1 t) */
    DASM_HOOK_SCALE       = 4  /* post: dst = f2 dst                                        */
  };
  typedef struct dasm_hook
  {
    int         kind;
    double      f1, f2;
    const void *v0, *v1; /* device vectors of the operator's number type */
  } dasm_hook;

  const char *dasm_last_error(void);
  const char *dasm_version(void);

  /* ---- context ------------------------------------------------------------------------------ */
  int dasm_ctx_create(int device, dasm_ctx **out);
  int dasm_ctx_destroy(dasm_ctx *ctx);
  int dasm_ctx_sync(dasm_ctx *ctx);
  /* number of kernels this library has launched on the context since creation */
  long long dasm_ctx_launch_count(const dasm_ctx *ctx);
  /* the cudaStream_t all work of this context is enqueued on (for CUDA-event timing by the caller) */
  void *dasm_ctx_stream(dasm_ctx *ctx);
  /* per-kernel-class device timing with CUDA events on the context's stream (replaces the LIKWID
   * marker regions of matrix_free_loop_08.likwid.cc:365-377).  klass: 0 Laplace cell kernel,
   * 1 FDM cell kernel, 2 vector epilogues, 3 ghost exchange.  enable(1) resets the counters. */
  int dasm_ctx_enable_kernel_timing(dasm_ctx *ctx, int on);
  int dasm_ctx_kernel_time(dasm_ctx *ctx, int klass, double *ms, long long *count);
  /* NCCL communicator for ghost / overlap-layer exchange (replaces the MPI communicator of
   * Utilities::MPI::Partitioner used in matrix_free_internal.h:21-83).  id = 128-byte ncclUniqueId. */
  int dasm_nccl_unique_id(void *id128);
  int dasm_ctx_comm_init(dasm_ctx *ctx, int n_ranks, int rank, const void *id128);

  /* ---- mesh --------------------------------------------------------------------------------- */
  /* GridGenerator::subdivided_hyper_cube_balanced, include/grid_generator.h:107-156:
   * n_subdivisions s -> n_refine, subdivisions[3]; cells per direction = subdivisions * 2^n_refine */
  int dasm_decompose_balanced(int n_subdivisions, int *n_refine, int subdivisions[3]);
  /* structured hex mesh on [0,length]^3; partition = ranks per direction (brick partition, replaces
   * the p4est partition of parallel::distributed::Triangulation, matrix_free_loop_08.likwid.cc:160) */
  int dasm_mesh_create_structured(dasm_ctx *ctx, const int n_cells[3], const int periodic[3], int dirichlet,
                                  const double length[3], int map_kind, const double map_params[4],
                                  const int partition[3], int rank, dasm_mesh **out);
  int       dasm_mesh_destroy(dasm_mesh *mesh);
  long long dasm_mesh_n_cells(const dasm_mesh *mesh);        /* local cells */
  long long dasm_mesh_n_global_cells(const dasm_mesh *mesh);
  /* processing order -> global (i,j,k) of every local cell, out[n_cells*3] (host) */
  int dasm_mesh_cell_coordinates(const dasm_mesh *mesh, int *out);

  /* Host-only view (no device, `ctx` of the mesh may be NULL) of what dasm_op_create sets up for this rank's part of a
   * partitioned mesh: the brick-grouped owner-cell numbering and the ghost-exchange lists (the counterpart of
   * Utilities::MPI::Partitioner, include/matrix_free_internal.h:21-83).  Used by the multi-process CPU tests.
   * sizes = {n_owned, n_ghost, n_local_cells, n_peers, n_send_total, n_recv_total}; every array may be NULL:
   * cidx_plain[n_local_cells*27], peers[n_peers], send_count[n_peers], recv_count[n_peers], send_idx[n_send_total]
   * (owned DoFs sent to the peers for a ghost update, peer after peer), recv_idx[n_recv_total] (ghost DoFs filled). */
  int dasm_mesh_host_numbering(const dasm_mesh *mesh, int degree, long long sizes[6], unsigned int *cidx_plain, int *peers,
                               long long *send_count, long long *recv_count, unsigned int *send_idx, unsigned int *recv_idx);

  /* The same for the enlarged ghost layout a preconditioner with overlapping patches needs on several ranks (include/matrix_free.h:154-213:
   * all DoFs of the cells around the rank's cells are ghosts): sizes = {n_owned, n_ghost (old + new), n_local_cells, n_halo_cells,
   * n_peers, n_send_total, n_recv_total}; halo_cells[n_halo_cells*3] global coordinates; cidx_plain[(n_local_cells + n_halo_cells)*27]. */
  int dasm_mesh_host_halo_numbering(const dasm_mesh *mesh, int degree, long long sizes[7], int *halo_cells, unsigned int *cidx_plain,
                                    int *peers, long long *send_count, long long *recv_count, unsigned int *send_idx,
                                    unsigned int *recv_idx);

  /* Host-only test hook: even-odd blocks (m x m block P on the even parts, h x h block Q on the odd parts, m = ceil(n/2),
   * h = floor(n/2)) of an n x n 1-D matrix as the warp-specialised kernels use them.  kind 0: centrosymmetric matrix
   * (mass / stiffness), 1: forward eigenvector matrix (rows = eigen index, even eigenvectors first), 2: backward.
   * Returns 0, or 1 (with dasm_last_error) when the matrix does not have the symmetry. */
  int dasm_test_eo_pack(int n, int kind, const double *A, double *P, double *Q);

  /* ---- LaplaceOperatorMatrixFree (include/operator.h:266-1628) ------------------------------- */
  /* ctor operator.h:466-482 + setup_mapping_and_indices 490-753.  mapping_type in {"", "merged", "linear geometry",
   * "quadratic geometry", "construct q"}; unknown names return the reference's error (operator.h:747-752). */
  int dasm_op_create(dasm_mesh *mesh, int degree, int number_type, const char *mapping_type, int compress_indices,
                     dasm_op **out);
  int       dasm_op_destroy(dasm_op *op);
  long long dasm_op_n_fast_bricks(const dasm_op *op); /* diagnostics: bricks processed by the warp-specialised kernel */
  long long dasm_op_n_dofs(const dasm_op *op);        /* locally owned (vector_partitioner->locally_owned_size) */
  long long dasm_op_n_ghost(const dasm_op *op);
  long long dasm_op_n_import(const dasm_op *op);      /* owned DoFs other ranks ghost (partitioner->n_import_indices()) */
  long long dasm_op_vec_size(const dasm_op *op);      /* owned + ghost */
  long long dasm_op_n_global_dofs(const dasm_op *op); /* m(), operator.h:1451 */
  int       dasm_op_degree(const dasm_op *op);
  int       dasm_op_number_type(const dasm_op *op);
  int       dasm_op_uses_compressed_indices(const dasm_op *op); /* operator.h:484-488 */
  /* vmult(dst, src), operator.h:1353-1365 (dst zeroed, constrained DoFs stay zero) */
  int dasm_op_vmult(dasm_op *op, void *dst, const void *src);
  /* vmult(dst, src, pre, post), operator.h:1367-1430 (constrained DoFs: dst = src when post given) */
  int dasm_op_vmult_hooks(dasm_op *op, void *dst, const void *src, const dasm_hook *pre, const dasm_hook *post);
  /* LaplaceOperatorBase::rhs(vec, func), operator.h:53-56 / 298-330, for a constant function f = value: vec_i = int f phi_i
   * (VectorTools::create_right_hand_side; constrained entries zero).  get_constraints(): dasm_op_constrained_dofs. */
  int dasm_op_rhs_constant(dasm_op *op, void *vec, double value);
  /* compute_inverse_diagonal, operator.h:1512-1524 */
  int dasm_op_inverse_diagonal(dasm_op *op, void *diag);
  /* 27 compressed start indices per local cell (ConstraintInfoReduced::compressed_dof_indices,
   * vector_access_reduced.h:30-164); plain != 0 keeps the index of constrained entities instead of
   * 0xFFFFFFFF.  out[n_cells*27] host. */
  int dasm_op_compressed_indices(const dasm_op *op, int plain, uint32_t *out);
  /* list of constrained (homogeneous Dirichlet) owned DoFs; returns count, fills out if non-NULL */
  long long dasm_op_constrained_dofs(const dasm_op *op, uint32_t *out);
  /* merged coefficients of one local cell, out[6*(degree+1)^3] host doubles (operator.h:674-711) */
  int dasm_op_merged_coefficients(const dasm_op *op, long long cell, double *out);
  /* host <-> device helpers for vectors of the operator's number type (host side double) */
  int dasm_op_vec_alloc(dasm_op *op, void **dev);
  int dasm_op_vec_free(dasm_op *op, void *dev);
  int dasm_op_vec_upload(dasm_op *op, void *dev, const double *host_owned);
  int dasm_op_vec_download(dasm_op *op, double *host_owned, const void *dev);

  /* ---- ASPoissonPreconditioner (include/matrix_free.h:63-1568) ------------------------------- */
  /* create_fdm_preconditioner, precondition.templates.h:162-247, + ctor matrix_free.h:73-894.  On several ranks n_overlap > 1 and
   * vertex patches use the enlarged ghost layout of the operator (every vector of dasm_op_vec_size() entries already holds it; this
   * replaces set_partitioner, operator.h:780-849) */
  int dasm_fdm_create(dasm_op *op, int n_overlap, int sub_mesh_approximation, int weight_type, int weight_sequence,
                      int overlap_pre_post, int element_centric, dasm_fdm **out);
  int dasm_fdm_destroy(dasm_fdm *fdm);
  /* vmult(dst, src), matrix_free.h:925-955 */
  int dasm_fdm_vmult(dasm_fdm *fdm, void *dst, const void *src);
  /* vmult(dst, src, pre, post), matrix_free.h:960-986 */
  int dasm_fdm_vmult_hooks(dasm_fdm *fdm, void *dst, const void *src, const dasm_hook *pre, const dasm_hook *post);
  long long dasm_fdm_n_fast_bricks(const dasm_fdm *fdm);   /* diagnostics, as dasm_op_n_fast_bricks */
  long long dasm_fdm_n_instances(const dasm_fdm *fdm);      /* n_fdm_instances(), matrix_free.h:1000-1004 */
  long long dasm_fdm_memory_consumption(const dasm_fdm *fdm); /* matrix_free.h:988-992 */
  int       dasm_fdm_is_symmetric(const dasm_fdm *fdm);     /* matrix_free.h:896-904 */
  int       dasm_fdm_patch_size_1d(const dasm_fdm *fdm);    /* matrix_free.h:90-92 */
  /* global weight vector (1/valence, 1/sqrt(valence)), host doubles of owned size; matrix_free.h:674-712 */
  int dasm_fdm_weights(const dasm_fdm *fdm, double *out_owned);
  /* eigenvector matrix S[m*m] (row-major, column = eigenvector) and eigenvalues[m] of cell/direction */
  int dasm_fdm_instance(const dasm_fdm *fdm, long long cell, int direction, double *S, double *lambda);

  /* ---- PreconditionChebyshev (deal.II; configured in precondition.templates.h:89-158, selected in
   *      create_system_preconditioner 439-584; invoked via PreconditionerAdapter, preconditioners.h:897-926) */
  /* fdm == NULL selects the point-Jacobi preconditioner (DiagonalMatrixPrePost, preconditioners.h:951-997) */
  int dasm_cheb_create(dasm_op *op, dasm_fdm *fdm, int degree, double smoothing_range, int polynomial_type,
                       int ev_algorithm, int optimize, int eig_cg_n_iterations, dasm_cheb **out);
  int dasm_cheb_destroy(dasm_cheb *cheb);
  int dasm_cheb_estimate_eigenvalues(dasm_cheb *cheb, double *min_ev, double *max_ev);
  int dasm_cheb_set_eigenvalues(dasm_cheb *cheb, double min_ev, double max_ev);
  int dasm_cheb_vmult(dasm_cheb *cheb, void *dst, const void *src); /* x = p(P^-1 A) P^-1 b, x0 = 0 */
  int dasm_cheb_step(dasm_cheb *cheb, void *dst, const void *src);  /* PreconditionerBase::step, preconditioners.h:725-742 */
  /* the call the reference's user makes with host vectors: copies src (and dst for step) in, result out */
  int dasm_cheb_step_host(dasm_cheb *cheb, double *dst_owned, const double *src_owned);
  int dasm_cheb_vmult_host(dasm_cheb *cheb, double *dst_owned, const double *src_owned);
  /* n independent problems on one smoother, pipelined: x_i <- step(x_i, b_i) with dst_owned[i] = x_i (in / out) and src_owned[i] = b_i.
   * The host -> device copies of problem i + 1 and the device -> host copy of problem i - 1 overlap the kernels of problem i (two sets
   * of device vectors, one copy stream per direction): a step costs max(copy in, copy out, kernels) instead of their sum.  The host
   * buffers should be pinned (cudaHostAlloc / cudaHostRegister); distinct problems must not share buffers. */
  int dasm_cheb_step_host_batch(dasm_cheb *cheb, int n, double *const *dst_owned, const double *const *src_owned);
  int dasm_cheb_vmult_host_batch(dasm_cheb *cheb, int n, double *const *dst_owned, const double *const *src_owned);

  /* ---- Krylov solvers on the device: solve() of element_centered_preconditioners_01.cc:108-203 --------------------
   * solver: DASM_SOLVER_CG (SolverCG) or DASM_SOLVER_GMRES (SolverGMRES, right preconditioning, restart after `restart` vectors;
   * the reference's default max_n_tmp_vectors = 30); stopping rule of ReductionControl(max_it, abs_tol, rel_tol): the residual
   * norm drops below max(abs_tol, rel_tol * ||r0||).  x starts from 0 (`x = 0` in dispatch()).
   * preconditioner: DASM_PRECON_IDENTITY (handle NULL), DASM_PRECON_DIAGONAL (NULL; DiagonalMatrix of the inverse diagonal),
   * DASM_PRECON_FDM (dasm_fdm*), DASM_PRECON_CHEBYSHEV (dasm_cheb*).  n_it = solver_control.last_step(), residual = last
   * residual norm (estimate for GMRES).  Returns an error when max_it is reached (SolverControl::NoConvergence). */
  enum
  {
    DASM_SOLVER_CG    = 0,
    DASM_SOLVER_GMRES = 1
  };
  enum
  {
    DASM_PRECON_IDENTITY  = 0,
    DASM_PRECON_DIAGONAL  = 1,
    DASM_PRECON_FDM       = 2,
    DASM_PRECON_CHEBYSHEV = 3,
    DASM_PRECON_MULTIGRID = 4, /* dasm_mg*  */
    DASM_PRECON_BLOCK_ASM = 5  /* dasm_asm* */
  };
  int dasm_solve(dasm_op *op, int solver, int precon_kind, void *precon, void *x, const void *b, int max_it, double abs_tol,
                 double rel_tol, int restart, int *n_it, double *residual);


  /* ---- layout accessors used by the components built on top of the operator (mg.cu, block_asm.cu) -------------------------- */
  dasm_ctx * dasm_op_ctx(const dasm_op *op);
  dasm_mesh *dasm_op_mesh(const dasm_op *op);
  dasm_op *  dasm_fdm_op(const dasm_fdm *fdm);
  dasm_op *  dasm_cheb_op(const dasm_cheb *cheb);
  /* DEVICE pointer to the 27 compressed start indices per local cell (constrained entities 0xFFFFFFFF; bit 31 = the entity lies in a
   * lexicographic brick box and expands with the strides 1, 4k, 16k^2) */
  const uint32_t *dasm_op_device_indices(const dasm_op *op);
  /* cells per direction of the whole mesh and the periodicity flags */
  int dasm_mesh_global_size(const dasm_mesh *mesh, int n_cells[3], int periodic[3]);
  /* update_ghost_values() / compress(VectorOperation::add) of a device vector in the operator's layout
   * (VectorDataExchange, include/matrix_free_internal.h:3-109, 321-352); compress zeroes the ghost part afterwards */
  int dasm_op_update_ghost_values(dasm_op *op, void *vec);
  int dasm_op_compress_add(dasm_op *op, void *vec);
  /* patch layout of a preconditioner: DoF indices d_idx[n_cells * m^3] (0xFFFFFFFF = not part of the patch) and per-entry weights
   * d_w[n_cells * m^3] (number type of the operator), both DEVICE buffers of the caller; *w_pre / *w_post: the weights are applied
   * before / after the block solve.  This is the data of Restrictors::ElementCenteredRestrictor (include/restrictors.h:48-338). */
  int  dasm_fdm_export_patches(dasm_fdm *fdm, uint32_t *d_idx, void *d_w, int *w_pre, int *w_post);
  /* the same on the host: idx[n_cells * m^3], w[n_cells * m^3] (doubles) */
  int  dasm_fdm_patches_host(dasm_fdm *fdm, uint32_t *idx, double *w, int *w_pre, int *w_post);
  void dasm_set_last_error(const char *msg);

  /* ---- Unstructured all-hex meshes (BASELINE configs[3]: the ball, element_centered_preconditioners_01.cc:398-402; SURVEY 8(b)
   *      dasm_mesh_create_from_arrays) ------------------------------------------------------------------------------------------
   * The mesh comes as arrays: 3 coordinates per vertex, 8 vertices per cell in lexicographic order (x fastest, deal.II's vertex order)
   * and, optionally, the 27 support points per cell of a triquadratic cell map (MappingQCache(2), lexicographic; NULL: trilinear
   * cells from the vertices).  The library derives lines / quads, the packed orientation word per cell and the 3^3 compressed start
   * indices (ConstraintInfoReduced, include/vector_access_reduced.h:30-164, include/reduced_access.h:154-285); DoFs of boundary
   * entities are constrained when `dirichlet` is set.  mapping_type: "" or "merged" (6 coefficients per quadrature point), "construct
   * q".  The operator works with dasm_op_vmult*, dasm_op_inverse_diagonal, dasm_op_rhs_constant, dasm_fdm_create (n overlap = 1,
   * all weighting types incl. ras; weight sequences global / local / dg), dasm_cheb_* and dasm_solve; multigrid transfers and the
   * exact-block ASM need a structured mesh.  One rank. */
  int dasm_op_create_unstructured(dasm_ctx *ctx, int degree, int number_type, const char *mapping_type, long long n_vertices,
                                  const double *coords, long long n_cells, const uint32_t *cell_vertices, const double *support_points,
                                  int dirichlet, dasm_op **out);
  int dasm_op_orientations(const dasm_op *op, uint32_t *out);          /* [n_cells] */
  int dasm_op_plain_indices(const dasm_op *op, uint32_t *out);         /* [n_cells][(k+1)^3] oriented addresses, host copy */
  int dasm_op_patch_extents(const dasm_op *op, double *out);           /* [n_cells][3][3], include/grid_tools.h:54-138 */
  long long dasm_op_n_cells(const dasm_op *op);
  int dasm_op_is_unstructured(const dasm_op *op);
  const uint32_t *dasm_op_device_plain_indices(const dasm_op *op);     /* device copy of the oriented addresses (NULL: compressed storage) */
  int dasm_op_entity_valence(const dasm_op *op, uint8_t *out);         /* [n_cells][27] cells per entity (unstructured meshes) */
  /* host-only (no device needed): sizes = {n_dofs, n_lines, n_quads, n_constrained}; output arrays may be NULL */
  int dasm_umesh_host_numbering(int degree, long long n_vertices, const double *coords, long long n_cells, const uint32_t *cell_vertices,
                                const double *support_points, int dirichlet, long long sizes[4], uint32_t *cidx, uint32_t *orientation,
                                uint32_t *plain, uint32_t *constrained, double *extents);

  /* ---- Power kernel (power_kernel_01.likwid.cc:122-308, 479-599): dst_0 += A src (Laplace), dst_1 += M dst_0 (mass operator), with the
   *      second operator applied to a cell as soon as its dst_0 entries are complete (fused != 0) or in a second sweep (fused == 0).
   * cell_granularity: cells per wave (0: all cells = one wave); batch_size 1: every cell is released on its own ("own batches"),
   * > 1: cells are released in batches of that size ("matrix-free batches").  do_computation == 0: gather / scatter only.
   * Cartesian structured meshes, one rank.  Vectors: device pointers in the operator's layout. */
  typedef struct dasm_power dasm_power;
  int       dasm_power_create(dasm_op *op, long long cell_granularity, int batch_size, dasm_power **out);
  int       dasm_power_run(dasm_power *p, void *dst_0, void *dst_1, const void *src, int fused, int do_computation);
  long long dasm_power_n_waves(const dasm_power *p);
  long long dasm_power_post_count(const dasm_power *p, long long wave);
  int       dasm_power_destroy(dasm_power *p);

  /* ---- Orientation-aware compressed vector access (ConstraintInfoReduced::read_dof_values / distribute_local_to_global,
   *      include/vector_access_reduced.h:267-548, with adjust_for_orientation, include/reduced_access.h:528-702) ------------------
   * d_cidx: 27 start indices per cell (0xFFFFFFFF = constrained), d_orientation: one packed word per cell (12 line bits + 6 x 3 quad
   * bits, the "post" word of compress_orientation; NULL = all standard), local: [n_cells][(degree+1)^3] values in lexicographic
   * order.  All pointers are DEVICE pointers; `stream` is a cudaStream_t.  The structured meshes of this library only produce the
   * standard orientation; these entry points are the general form for unstructured meshes. */
  int dasm_reduced_access_read(int degree, int number_type, const uint32_t *d_cidx, const uint32_t *d_orientation, long long n_cells,
                               const void *src, void *local, void *stream);
  int dasm_reduced_access_distribute(int degree, int number_type, const uint32_t *d_cidx, const uint32_t *d_orientation,
                                     long long n_cells, void *dst, const void *local, void *stream);

  /* ---- Geometric / polynomial multigrid V-cycle (include/multigrid.h:109-537: PreconditionerGMG; deal.II Multigrid +
   *      MGTransferGlobalCoarsening + PreconditionMG as set up in element_centered_preconditioners_01.cc:540-740) ------------------
   * Two-level transfer between two operators on the same context: geometric (the fine mesh has twice the cells of the coarse mesh in
   * every direction, same degree: "mg type" h) or polynomial (same mesh, lower degree: "mg type" p).  Prolongation = embedding,
   * restriction = its transpose; constrained DoFs are read as zero and not written (MGTwoLevelTransfer). */
  typedef struct dasm_transfer dasm_transfer;
  int dasm_transfer_create(dasm_op *fine, dasm_op *coarse, dasm_transfer **out);
  /* two operators on unstructured meshes: parent == NULL: both on the same mesh (polynomial transfer, coarse degree <= fine degree);
   * otherwise parent[fine cell] = coarse cell | child position << 28 (bit 28 / 29 / 30: upper half in x / y / z of the parent's frame):
   * geometric 2:1 transfer between a mesh and its global refinement whose children keep the parent's frame */
  int dasm_transfer_create_unstructured(dasm_op *fine, dasm_op *coarse, const uint32_t *parent, dasm_transfer **out);
  int dasm_transfer_destroy(dasm_transfer *t);
  int dasm_transfer_prolongate_and_add(dasm_transfer *t, void *dst_fine, const void *src_coarse);
  int dasm_transfer_restrict_and_add(dasm_transfer *t, void *dst_coarse, const void *src_fine);
  /* levels[0] = coarsest.  smoothers[l] (l >= 1): pre- and post-smoother of level l (MGSmootherRelaxation with one step: vmult
   * before and step after the coarse-grid correction); smoothers[0]: coarse-grid solver (MGCoarseGridApplyPreconditioner,
   * include/multigrid.h:96-108).  All level operators share one number type (float in the reference's matrix-free set-up,
   * element_centered_preconditioners_01.cc:787-792).  one_sided_v_cycle: no post-smoothing (include/multigrid.h:303-312). */
  typedef struct dasm_mg dasm_mg;
  int dasm_mg_create(int n_levels, dasm_op **level_ops, dasm_cheb **smoothers, int one_sided_v_cycle, dasm_mg **out);
  /* same with caller-built transfers (transfers[l] between the levels l and l - 1, NULL entries are created by the library; the caller
   * keeps their ownership): the geometric levels of an unstructured mesh need dasm_transfer_create_unstructured's parent map */
  int dasm_mg_create_with_transfers(int n_levels, dasm_op **level_ops, dasm_cheb **smoothers, dasm_transfer **transfers, int one_sided_v_cycle,
                                    dasm_mg **out);
  int dasm_mg_destroy(dasm_mg *mg);
  /* PreconditionerGMG::vmult (include/multigrid.h:463-469): dst = V-cycle(src); dst / src are device vectors of the finest level's
   * layout in the number type `outer_number_type` (converted to the level number type and back, PreconditionMG copy_to_mg /
   * copy_from_mg) */
  int dasm_mg_vmult_outer(dasm_mg *mg, void *dst, const void *src, int outer_number_type);
  int dasm_mg_vmult(dasm_mg *mg, void *dst, const void *src); /* vectors of the level number type */

  /* ---- Exact-block additive Schwarz (RestrictedPreconditioner over ElementCenteredRestrictor with RestrictedMatrixView blocks
   *      inverted by gauss_jordan: include/preconditioners.h:528-605, 744-813; include/restrictors.h:17-378) ---------------------
   * The patches (indices, weights, weighting type) are those of `layout`; the block of a patch is the restriction of the operator's
   * matrix to the patch DoFs, obtained by applying the operator to coloured unit vectors, and inverted on the device. */
  typedef struct dasm_asm dasm_asm;
  int       dasm_asm_create(dasm_fdm *layout, dasm_asm **out);
  int       dasm_asm_destroy(dasm_asm *a);
  int       dasm_asm_vmult(dasm_asm *a, void *dst, const void *src);
  long long dasm_asm_memory_consumption(const dasm_asm *a);
  /* inverse block of one local cell, out[m^3 * m^3] host doubles, row-major (rows / columns of entries outside the patch are those
   * of the identity) */
  int dasm_asm_block(const dasm_asm *a, long long cell, double *out);

#ifdef __cplusplus
}
#endif
#endif
