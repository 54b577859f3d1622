"""Python host-side mirror of the reference's operator / preconditioner interface over libdasm's C ABI.

Class and factory names follow the reference (file:line into the reference tree):
  LaplaceOperatorMatrixFree      include/operator.h:266-1628
  ASPoissonPreconditioner        include/matrix_free.h:63-1568
  PreconditionChebyshev          deal.II class configured in include/precondition.templates.h:89-158
  create_fdm_preconditioner      include/precondition.templates.h:162-247
  create_system_preconditioner   include/precondition.templates.h:251-818 (types Chebyshev and FDM)
The parameter dictionaries use the reference's JSON vocabulary ("n overlap", "weighting type",
"weight sequence", "degree", "optimize", "smoothing range", "ev algorithm", "polynomial type", ...).

All compute happens in libdasm.so (CUDA, sm_100a).  There is no CPU fallback: importing works
without a GPU (so symbols can be checked), creating a Context does not.
PyTorch is used only for device memory (tensors whose data_ptr is handed to the C ABI).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdasm.so")

# the Chebyshev step runs as two fused brick kernels per term (vector updates in the kernel epilogues)
FUSED_KERNELS = True

F64, F32 = 0, 1
WEIGHT = {"none": 0, "pre": 1, "post": 2, "ras": 3, "symm": 4}
WSEQ = {"global": 0, "local": 1, "dg": 2, "DG": 2, "compressed": 3}
MAP = {"cartesian": 0, "sine": 1, "kershaw": 2}
HOOK_NONE, HOOK_ZERO_DST, HOOK_RESIDUAL, HOOK_CHEB_UPDATE, HOOK_SCALE = 0, 1, 2, 3, 4


class DasmError(RuntimeError):
    pass


class Hook(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("f1", ctypes.c_double), ("f2", ctypes.c_double),
                ("v0", ctypes.c_void_p), ("v1", ctypes.c_void_p)]


_lib = None


def lib():
    """loads libdasm.so (fails loudly if it was not built: run __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DasmError("libdasm.so not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
        try:
            import torch  # noqa: F401  (loads the CUDA runtime / NCCL the library links against)
        except Exception:
            pass
        _lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        _lib.dasm_last_error.restype = ctypes.c_char_p
        _lib.dasm_version.restype = ctypes.c_char_p
        _lib.dasm_ctx_stream.restype = ctypes.c_void_p
        for name in ("dasm_ctx_launch_count", "dasm_mesh_n_cells", "dasm_mesh_n_global_cells", "dasm_op_n_dofs",
                     "dasm_op_n_ghost", "dasm_op_n_import", "dasm_op_vec_size", "dasm_op_n_global_dofs", "dasm_op_constrained_dofs",
                     "dasm_fdm_n_instances", "dasm_fdm_memory_consumption", "dasm_op_n_fast_bricks", "dasm_fdm_n_fast_bricks", "dasm_op_n_cells", "dasm_power_n_waves", "dasm_power_post_count"):
            getattr(_lib, name).restype = ctypes.c_longlong
    return _lib


def _check(rc):
    if rc != 0:
        raise DasmError(lib().dasm_last_error().decode())


def _ptr(t):
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def decompose_balanced(s):
    nr = ctypes.c_int()
    sub = (ctypes.c_int * 3)()
    _check(lib().dasm_decompose_balanced(int(s), ctypes.byref(nr), sub))
    return nr.value, list(sub)


class Context:
    def __init__(self, device=0):
        self.h = ctypes.c_void_p()
        _check(lib().dasm_ctx_create(int(device), ctypes.byref(self.h)))
        self.device = device

    def sync(self):
        _check(lib().dasm_ctx_sync(self.h))

    def launch_count(self):
        return lib().dasm_ctx_launch_count(self.h)

    def enable_kernel_timing(self, on=True):
        _check(lib().dasm_ctx_enable_kernel_timing(self.h, int(on)))

    def kernel_time(self, klass):
        """(total ms, launches) of kernel class 0 Laplace / 1 FDM / 2 vector / 3 exchange since enable_kernel_timing."""
        ms, n = ctypes.c_double(), ctypes.c_longlong()
        _check(lib().dasm_ctx_kernel_time(self.h, int(klass), ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def stream_ptr(self):
        return lib().dasm_ctx_stream(self.h)

    def comm_init(self, n_ranks, rank, id_bytes):
        buf = (ctypes.c_char * 128).from_buffer_copy(id_bytes)
        _check(lib().dasm_ctx_comm_init(self.h, int(n_ranks), int(rank), buf))

    @staticmethod
    def nccl_unique_id():
        buf = (ctypes.c_char * 128)()
        _check(lib().dasm_nccl_unique_id(buf))
        return bytes(buf)


class Mesh:
    """structured hex mesh (hyper-rectangle, optionally sine-deformed or Kershaw)."""

    def __init__(self, ctx, n_cells, periodic=(0, 0, 0), dirichlet=True, length=(1., 1., 1.), map_kind="cartesian",
                 map_params=(0., 0., 0., 0.), partition=(1, 1, 1), rank=0):
        self.ctx = ctx
        self.h = ctypes.c_void_p()
        self.n_cells_dir = tuple(int(c) for c in n_cells)
        self.periodic = tuple(int(p) for p in periodic)
        self.dirichlet = bool(dirichlet)
        self.length = tuple(float(x) for x in length)
        self.map_kind = map_kind
        self.map_params = tuple(map_params)
        nc = (ctypes.c_int * 3)(*self.n_cells_dir)
        per = (ctypes.c_int * 3)(*self.periodic)
        ln = (ctypes.c_double * 3)(*self.length)
        mp = (ctypes.c_double * 4)(*[float(x) for x in map_params])
        pt = (ctypes.c_int * 3)(*[int(x) for x in partition])
        _check(lib().dasm_mesh_create_structured(ctx.h if ctx is not None else None, nc, per, int(bool(dirichlet)), ln, MAP[map_kind], mp, pt, int(rank),
                                                 ctypes.byref(self.h)))

    @classmethod
    def hyper_cube_balanced(cls, ctx, n_subdivisions, periodic=True, **kw):
        """GridGenerator::subdivided_hyper_cube_balanced + refine_global, matrix_free_loop_08.likwid.cc:160-174."""
        nr, sub = decompose_balanced(n_subdivisions)
        n_cells = [s * 2 ** nr for s in sub]
        return cls(ctx, n_cells, periodic=(int(periodic),) * 3, length=[float(s) for s in sub], **kw)

    @property
    def n_cells(self):
        return lib().dasm_mesh_n_cells(self.h)

    def host_numbering(self, degree):
        """host-only (no device needed): numbering sizes, plain compressed indices and the ghost-exchange lists of this rank"""
        sizes = (ctypes.c_longlong * 6)()
        none = ctypes.c_void_p()
        _check(lib().dasm_mesh_host_numbering(self.h, int(degree), sizes, none, none, none, none, none, none))
        n_owned, n_ghost, n_cells, n_peers, ns, nr = [int(v) for v in sizes]
        cidx = np.zeros((n_cells, 27), dtype=np.uint32)
        peers = np.zeros(max(n_peers, 1), dtype=np.int32)
        sc, rc = np.zeros(max(n_peers, 1), dtype=np.int64), np.zeros(max(n_peers, 1), dtype=np.int64)
        si, ri = np.zeros(max(ns, 1), dtype=np.uint32), np.zeros(max(nr, 1), dtype=np.uint32)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        _check(lib().dasm_mesh_host_numbering(self.h, int(degree), sizes, p(cidx), p(peers), p(sc), p(rc), p(si), p(ri)))
        return dict(n_owned=n_owned, n_ghost=n_ghost, n_cells=n_cells, cidx_plain=cidx, peers=peers[:n_peers], send_count=sc[:n_peers],
                    recv_count=rc[:n_peers], send_idx=si[:ns], recv_idx=ri[:nr])

    def host_halo_numbering(self, degree):
        """host-only: the enlarged ghost layout (halo cells around the rank's box); cidx_plain / coords cover local + halo cells"""
        sizes = (ctypes.c_longlong * 7)()
        none = ctypes.c_void_p()
        _check(lib().dasm_mesh_host_halo_numbering(self.h, int(degree), sizes, none, none, none, none, none, none, none))
        n_owned, n_ghost, n_cells, n_halo, n_peers, ns, nr = [int(v) for v in sizes]
        halo = np.zeros((max(n_halo, 1), 3), dtype=np.int32)
        cidx = np.zeros((n_cells + n_halo, 27), dtype=np.uint32)
        peers = np.zeros(max(n_peers, 1), dtype=np.int32)
        sc, rc = np.zeros(max(n_peers, 1), dtype=np.int64), np.zeros(max(n_peers, 1), dtype=np.int64)
        si, ri = np.zeros(max(ns, 1), dtype=np.uint32), np.zeros(max(nr, 1), dtype=np.uint32)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        _check(lib().dasm_mesh_host_halo_numbering(self.h, int(degree), sizes, p(halo), p(cidx), p(peers), p(sc), p(rc), p(si), p(ri)))
        coords = np.concatenate([self.cell_coordinates(), halo[:n_halo]], axis=0)
        return dict(n_owned=n_owned, n_ghost=n_ghost, n_cells=n_cells, n_halo=n_halo, coords=coords, cidx_plain=cidx, peers=peers[:n_peers],
                    send_count=sc[:n_peers], recv_count=rc[:n_peers], send_idx=si[:ns], recv_idx=ri[:nr])

    def cell_coordinates(self):
        out = np.zeros((self.n_cells, 3), dtype=np.int32)
        lib().dasm_mesh_cell_coordinates(self.h, out.ctypes.data_as(ctypes.c_void_p))
        return out

    def __del__(self):
        try:
            lib().dasm_mesh_destroy(self.h)
        except Exception:
            pass


class LaplaceOperatorMatrixFree:
    """include/operator.h:266-1628."""

    def __init__(self, mesh, degree, number="double", mapping_type="", compress_indices=True):
        import torch
        self.mesh = mesh
        self.ctx = mesh.ctx
        self.degree = degree
        self.number = number
        self.ntype = F64 if number == "double" else F32
        self.torch_dtype = torch.float64 if number == "double" else torch.float32
        self.h = ctypes.c_void_p()
        _check(lib().dasm_op_create(mesh.h, int(degree), self.ntype, mapping_type.encode(), int(compress_indices),
                                    ctypes.byref(self.h)))

    @classmethod
    def from_arrays(cls, ctx, vertices, cells, degree, support=None, dirichlet=True, number="double", mapping_type=""):
        """operator on an unstructured all-hex mesh (the ball of element_centered_preconditioners_01.cc:398-402): vertices [V, 3],
        cells [C, 8] (lexicographic vertex order), support [C, 27, 3] support points of the triquadratic cell map or None
        (dealii-asm_b200/grid.py builds such arrays)."""
        import torch
        self = cls.__new__(cls)
        self.mesh = None
        self.ctx = ctx
        self.degree = degree
        self.number = number
        self.ntype = F64 if number == "double" else F32
        self.torch_dtype = torch.float64 if number == "double" else torch.float32
        self.h = ctypes.c_void_p()
        v = np.ascontiguousarray(vertices, dtype=np.float64)
        c = np.ascontiguousarray(cells, dtype=np.uint32)
        sp = np.ascontiguousarray(support, dtype=np.float64) if support is not None else None
        _check(lib().dasm_op_create_unstructured(ctx.h, int(degree), self.ntype, mapping_type.encode(), ctypes.c_longlong(v.shape[0]),
                                                 v.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(c.shape[0]), c.ctypes.data_as(ctypes.c_void_p),
                                                 sp.ctypes.data_as(ctypes.c_void_p) if sp is not None else None, int(bool(dirichlet)),
                                                 ctypes.byref(self.h)))
        return self

    # -- sizes
    def n_dofs(self):
        return lib().dasm_op_n_dofs(self.h)

    def n_cells(self):
        return lib().dasm_op_n_cells(self.h)

    def orientations(self):
        """packed orientation word per cell (unstructured meshes), include/reduced_access.h:66-152"""
        out = np.zeros(self.n_cells(), dtype=np.uint32)
        _check(lib().dasm_op_orientations(self.h, out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def plain_indices(self):
        out = np.zeros((self.n_cells(), (self.degree + 1) ** 3), dtype=np.uint32)
        _check(lib().dasm_op_plain_indices(self.h, out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def patch_extents(self):
        out = np.zeros((self.n_cells(), 3, 3))
        _check(lib().dasm_op_patch_extents(self.h, out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def n_fast_bricks(self):
        return lib().dasm_op_n_fast_bricks(self.h)

    def vec_size(self):
        return lib().dasm_op_vec_size(self.h)

    def m(self):
        return lib().dasm_op_n_global_dofs(self.h)

    def uses_compressed_indices(self):
        return bool(lib().dasm_op_uses_compressed_indices(self.h))

    @staticmethod
    def is_matrix_free():
        return True

    def is_symmetric(self):
        return True

    def el(self, i, j):
        raise NotImplementedError("ExcNotImplemented")  # operator.h:1457-1463

    def Tvmult(self, dst, src):
        raise NotImplementedError("ExcNotImplemented")  # operator.h:1432-1438

    def initialize_dof_vector(self):
        import torch
        return torch.zeros(self.vec_size(), dtype=self.torch_dtype, device="cuda:%d" % self.ctx.device)

    def to_device(self, host_owned):
        import torch
        v = torch.zeros(self.vec_size(), dtype=self.torch_dtype)
        v[: self.n_dofs()] = torch.as_tensor(np.asarray(host_owned)).to(self.torch_dtype)
        out = v.to("cuda:%d" % self.ctx.device)
        torch.cuda.synchronize()
        return out

    def to_host(self, dev):
        self.ctx.sync()
        return dev[: self.n_dofs()].double().cpu().numpy()

    # -- operator
    def vmult(self, dst, src, pre=None, post=None):
        if pre is None and post is None:
            _check(lib().dasm_op_vmult(self.h, _ptr(dst), _ptr(src)))
        else:
            _check(lib().dasm_op_vmult_hooks(self.h, _ptr(dst), _ptr(src), ctypes.byref(pre) if pre else None,
                                             ctypes.byref(post) if post else None))

    def rhs(self, vec, value=1.0):
        """LaplaceOperatorBase::rhs for a constant right-hand-side function (include/operator.h:53-56, 298-330)."""
        _check(lib().dasm_op_rhs_constant(self.h, _ptr(vec), ctypes.c_double(value)))

    def get_constraints(self):
        """the constrained (homogeneous Dirichlet) DoFs, include/operator.h:46-51"""
        return self.constrained_dofs()

    def compute_inverse_diagonal(self, diag):
        _check(lib().dasm_op_inverse_diagonal(self.h, _ptr(diag)))

    def compressed_indices(self, plain=False):
        out = np.zeros((self.n_cells(), 27), dtype=np.uint32)
        lib().dasm_op_compressed_indices(self.h, int(plain), out.ctypes.data_as(ctypes.c_void_p))
        return out

    def constrained_dofs(self):
        n = lib().dasm_op_constrained_dofs(self.h, None)
        out = np.zeros(n, dtype=np.uint32)
        if n:
            lib().dasm_op_constrained_dofs(self.h, out.ctypes.data_as(ctypes.c_void_p))
        return out

    def merged_coefficients(self, cell):
        n3 = (self.degree + 1) ** 3
        out = np.zeros(6 * n3)
        _check(lib().dasm_op_merged_coefficients(self.h, ctypes.c_longlong(cell), out.ctypes.data_as(ctypes.c_void_p)))
        return out.reshape(6, n3)

    def __del__(self):
        try:
            lib().dasm_op_destroy(self.h)
        except Exception:
            pass


class ASPoissonPreconditioner:
    """include/matrix_free.h:63-1568 (element-centred FDM additive Schwarz)."""

    def __init__(self, op, n_overlap=1, sub_mesh_approximation=3, weight_type="post", weight_local_global="global",
                 overlap_pre_post=True, element_centric=True):
        self.op = op
        self.h = ctypes.c_void_p()
        if weight_type not in WEIGHT:
            raise DasmError("Weighting type <%s> is not known!" % weight_type)  # precondition.templates.h:24-28
        if weight_local_global not in WSEQ:
            raise DasmError("weight sequence <%s> is not known!" % weight_local_global)
        _check(lib().dasm_fdm_create(op.h, int(n_overlap), int(sub_mesh_approximation), WEIGHT[weight_type],
                                     WSEQ[weight_local_global], int(overlap_pre_post), int(element_centric),
                                     ctypes.byref(self.h)))

    def vmult(self, dst, src, pre=None, post=None):
        if pre is None and post is None:
            _check(lib().dasm_fdm_vmult(self.h, _ptr(dst), _ptr(src)))
        else:
            _check(lib().dasm_fdm_vmult_hooks(self.h, _ptr(dst), _ptr(src), ctypes.byref(pre) if pre else None,
                                              ctypes.byref(post) if post else None))

    def is_symmetric(self):
        return bool(lib().dasm_fdm_is_symmetric(self.h))

    def n_fdm_instances(self):
        return lib().dasm_fdm_n_instances(self.h)

    def n_fast_bricks(self):
        return lib().dasm_fdm_n_fast_bricks(self.h)

    def memory_consumption(self):
        return lib().dasm_fdm_memory_consumption(self.h)

    def patch_size_1d(self):
        return lib().dasm_fdm_patch_size_1d(self.h)

    def weights(self):
        out = np.zeros(self.op.n_dofs())
        _check(lib().dasm_fdm_weights(self.h, out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def instance(self, cell, direction):
        m = self.patch_size_1d()
        S = np.zeros((m, m))
        lam = np.zeros(m)
        _check(lib().dasm_fdm_instance(self.h, ctypes.c_longlong(cell), int(direction), S.ctypes.data_as(ctypes.c_void_p),
                                       lam.ctypes.data_as(ctypes.c_void_p)))
        return S, lam

    def step(self, dst, src):
        raise NotImplementedError("ExcNotImplemented")  # PreconditionerBase::step default, preconditioners.h:735-741

    def __del__(self):
        try:
            lib().dasm_fdm_destroy(self.h)
        except Exception:
            pass


class PreconditionChebyshev:
    """deal.II PreconditionChebyshev as configured by create_chebyshev_preconditioner
    (include/precondition.templates.h:89-158) around an FDM or point-Jacobi (fdm=None) preconditioner."""

    POLY = {"1st kind": 0, "4th kind": 1}
    EV = {"lanczos": 0, "power iteration": 1, None: 2}

    def __init__(self, op, fdm=None, degree=3, smoothing_range=20., polynomial_type="1st kind", ev_algorithm=None, optimize=2,
                 eig_cg_n_iterations=40):
        self.op = op
        self.fdm = fdm
        self.degree = degree
        if polynomial_type not in self.POLY:
            raise DasmError("Polynomial type <%s> is not known!" % polynomial_type)
        if ev_algorithm not in self.EV:
            raise DasmError("Eigen-value algorithm <%s> is not known!" % ev_algorithm)
        self.h = ctypes.c_void_p()
        _check(lib().dasm_cheb_create(op.h, fdm.h if fdm is not None else None, int(degree), ctypes.c_double(smoothing_range),
                                      self.POLY[polynomial_type], self.EV[ev_algorithm], int(optimize), int(eig_cg_n_iterations),
                                      ctypes.byref(self.h)))

    def estimate_eigenvalues(self):
        mn, mx = ctypes.c_double(), ctypes.c_double()
        _check(lib().dasm_cheb_estimate_eigenvalues(self.h, ctypes.byref(mn), ctypes.byref(mx)))
        return mn.value, mx.value

    def set_eigenvalues(self, min_ev, max_ev):
        _check(lib().dasm_cheb_set_eigenvalues(self.h, ctypes.c_double(min_ev), ctypes.c_double(max_ev)))

    def vmult(self, dst, src):
        _check(lib().dasm_cheb_vmult(self.h, _ptr(dst), _ptr(src)))

    def step(self, dst, src):
        _check(lib().dasm_cheb_step(self.h, _ptr(dst), _ptr(src)))

    def step_host(self, dst_np, src_np):
        """host-buffer call (float64 numpy arrays of the owned size; dst is updated in place)."""
        _check(lib().dasm_cheb_step_host(self.h, dst_np.ctypes.data_as(ctypes.c_void_p), src_np.ctypes.data_as(ctypes.c_void_p)))

    def step_host_batch(self, dst_list, src_list):
        """pipelined host-buffer call for independent problems: dst_list[i] (x_i, updated in place) and src_list[i] (b_i) are float64
        numpy arrays of the owned size, ideally views of pinned memory; copies of neighbouring problems overlap the kernels."""
        n = len(dst_list)
        assert n == len(src_list)
        d = (ctypes.c_void_p * n)(*[a.ctypes.data for a in dst_list])
        s_ = (ctypes.c_void_p * n)(*[a.ctypes.data for a in src_list])
        _check(lib().dasm_cheb_step_host_batch(self.h, n, d, s_))

    def vmult_host_batch(self, dst_list, src_list):
        n = len(dst_list)
        d = (ctypes.c_void_p * n)(*[a.ctypes.data for a in dst_list])
        s_ = (ctypes.c_void_p * n)(*[a.ctypes.data for a in src_list])
        _check(lib().dasm_cheb_vmult_host_batch(self.h, n, d, s_))

    def vmult_host(self, dst_np, src_np):
        _check(lib().dasm_cheb_vmult_host(self.h, dst_np.ctypes.data_as(ctypes.c_void_p), src_np.ctypes.data_as(ctypes.c_void_p)))

    def __del__(self):
        try:
            lib().dasm_cheb_destroy(self.h)
        except Exception:
            pass


class MGTwoLevelTransfer:
    """Two-level transfer of MGTransferGlobalCoarsening between two operators (geometric 2:1 or polynomial), the transfer
    operators PreconditionerGMG sets up in include/multigrid.h:338-349."""

    def __init__(self, fine, coarse, parent=None):
        """parent (unstructured meshes, geometric transfer): parent[fine cell] = coarse cell | child position << 28
        (dealii-asm_b200/grid.py ball_parents)"""
        self.fine, self.coarse = fine, coarse
        self.h = ctypes.c_void_p()
        if parent is None:
            _check(lib().dasm_transfer_create(fine.h, coarse.h, ctypes.byref(self.h)))
        else:
            par = np.ascontiguousarray(parent, dtype=np.uint32)
            _check(lib().dasm_transfer_create_unstructured(fine.h, coarse.h, par.ctypes.data_as(ctypes.c_void_p), ctypes.byref(self.h)))

    def prolongate_and_add(self, dst_fine, src_coarse):
        _check(lib().dasm_transfer_prolongate_and_add(self.h, _ptr(dst_fine), _ptr(src_coarse)))

    def restrict_and_add(self, dst_coarse, src_fine):
        _check(lib().dasm_transfer_restrict_and_add(self.h, _ptr(dst_coarse), _ptr(src_fine)))

    def __del__(self):
        try:
            lib().dasm_transfer_destroy(self.h)
        except Exception:
            pass


class PreconditionerGMG:
    """include/multigrid.h:109-537: V-cycle over level operators (coarsest first) with one smoother per level
    (smoothers[0] = coarse-grid solver).  vmult takes vectors of the OUTER number type (that of `outer_op`, double in the
    reference) and converts to the level number type and back."""

    def __init__(self, level_ops, smoothers, outer_op=None, use_one_sided_v_cycle=False, transfers=None):
        """transfers: optional list (one entry per level, entry l between the levels l and l - 1, None = built by the library)"""
        assert len(level_ops) == len(smoothers) and len(level_ops) >= 1
        self.level_ops, self.smoothers = list(level_ops), list(smoothers)
        self.transfers = list(transfers) if transfers is not None else None
        self.outer_ntype = (outer_op or level_ops[-1]).ntype
        n = len(level_ops)
        ops = (ctypes.c_void_p * n)(*[o.h for o in level_ops])
        sms = (ctypes.c_void_p * n)(*[s.h for s in smoothers])
        self.h = ctypes.c_void_p()
        if transfers is None:
            _check(lib().dasm_mg_create(n, ops, sms, 1 if use_one_sided_v_cycle else 0, ctypes.byref(self.h)))
        else:
            assert len(transfers) == n
            trs = (ctypes.c_void_p * n)(*[(t.h if t is not None else None) for t in transfers])
            _check(lib().dasm_mg_create_with_transfers(n, ops, sms, trs, 1 if use_one_sided_v_cycle else 0, ctypes.byref(self.h)))

    def vmult(self, dst, src):
        _check(lib().dasm_mg_vmult_outer(self.h, _ptr(dst), _ptr(src), int(self.outer_ntype)))

    def __del__(self):
        try:
            lib().dasm_mg_destroy(self.h)
        except Exception:
            pass


class RestrictedPreconditioner:
    """Exact-block additive Schwarz (include/preconditioners.h:744-813 over Restrictors::ElementCenteredRestrictor,
    include/restrictors.h:17-378): the patches and weights of `layout` (an ASPoissonPreconditioner) with the exact inverse of the
    restricted operator matrix per patch."""

    def __init__(self, layout):
        self.layout = layout
        self.op = layout.op
        self.h = ctypes.c_void_p()
        _check(lib().dasm_asm_create(layout.h, ctypes.byref(self.h)))

    def vmult(self, dst, src):
        _check(lib().dasm_asm_vmult(self.h, _ptr(dst), _ptr(src)))

    def memory_consumption(self):
        lib().dasm_asm_memory_consumption.restype = ctypes.c_longlong
        return lib().dasm_asm_memory_consumption(self.h)

    def block_inverse(self, cell):
        m3 = self.layout.patch_size_1d() ** 3
        out = np.zeros((m3, m3))
        _check(lib().dasm_asm_block(self.h, ctypes.c_longlong(cell), out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def __del__(self):
        try:
            lib().dasm_asm_destroy(self.h)
        except Exception:
            pass


class PowerKernel:
    """power_kernel_01.likwid.cc:479-599: dst_0 += A src, dst_1 += M dst_0 (mass operator); fused (power kernel: the second operator
    runs on a cell once its dst_0 entries are complete, wave by wave) or sequential (two sweeps)."""

    def __init__(self, op, cell_granularity=0, batch_size=1):
        self.op = op
        self.h = ctypes.c_void_p()
        _check(lib().dasm_power_create(op.h, ctypes.c_longlong(cell_granularity), int(batch_size), ctypes.byref(self.h)))

    def run(self, dst_0, dst_1, src, fused=True, do_computation=True):
        _check(lib().dasm_power_run(self.h, _ptr(dst_0), _ptr(dst_1), _ptr(src), int(fused), int(do_computation)))

    def n_waves(self):
        return lib().dasm_power_n_waves(self.h)

    def post_counts(self):
        return [lib().dasm_power_post_count(self.h, ctypes.c_longlong(w)) for w in range(self.n_waves())]

    def __del__(self):
        try:
            lib().dasm_power_destroy(self.h)
        except Exception:
            pass


def umesh_host_numbering(degree, vertices, cells, support=None, dirichlet=True):
    """host-only (no device): numbering of an unstructured mesh as the library builds it: dict(n_dofs, n_lines, n_quads, cidx [C, 27],
    orientation [C], plain [C, (k+1)^3], constrained, extents [C, 3, 3])"""
    v = np.ascontiguousarray(vertices, dtype=np.float64)
    c = np.ascontiguousarray(cells, dtype=np.uint32)
    sp = np.ascontiguousarray(support, dtype=np.float64) if support is not None else None
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else None
    sizes = (ctypes.c_longlong * 4)()
    args = (int(degree), ctypes.c_longlong(v.shape[0]), p(v), ctypes.c_longlong(c.shape[0]), p(c), p(sp), int(bool(dirichlet)), sizes)
    _check(lib().dasm_umesh_host_numbering(*args, None, None, None, None, None))
    n_dofs, n_lines, n_quads, n_con = [int(x) for x in sizes]
    C = c.shape[0]
    cidx = np.zeros((C, 27), dtype=np.uint32)
    ori = np.zeros(C, dtype=np.uint32)
    plain = np.zeros((C, (degree + 1) ** 3), dtype=np.uint32)
    con = np.zeros(max(n_con, 1), dtype=np.uint32)
    ext = np.zeros((C, 3, 3))
    _check(lib().dasm_umesh_host_numbering(*args, p(cidx), p(ori), p(plain), p(con), p(ext)))
    return dict(n_dofs=n_dofs, n_lines=n_lines, n_quads=n_quads, cidx=cidx, orientation=ori, plain=plain, constrained=con[:n_con], extents=ext)


def reduced_access_read(degree, cidx, orientation, src):
    """ConstraintInfoReduced::read_dof_values with the packed orientation word per cell (include/vector_access_reduced.h:267-405,
    include/reduced_access.h:528-702): cidx [cells, 27] uint32 and orientation [cells] uint32 torch CUDA tensors (orientation may be
    None), src a CUDA vector; returns local values [cells, (degree+1)^3]."""
    import torch
    n_cells = cidx.shape[0]
    local = torch.zeros((n_cells, (degree + 1) ** 3), dtype=src.dtype, device=src.device)
    nt = F64 if src.dtype == torch.float64 else F32
    _check(lib().dasm_reduced_access_read(int(degree), nt, _ptr(cidx), _ptr(orientation), ctypes.c_longlong(n_cells), _ptr(src), _ptr(local),
                                          ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return local


def reduced_access_distribute(degree, cidx, orientation, dst, local):
    """ConstraintInfoReduced::distribute_local_to_global (include/vector_access_reduced.h:407-548): the transpose of reduced_access_read."""
    import torch
    nt = F64 if dst.dtype == torch.float64 else F32
    _check(lib().dasm_reduced_access_distribute(int(degree), nt, _ptr(cidx), _ptr(orientation), ctypes.c_longlong(cidx.shape[0]), _ptr(dst),
                                                _ptr(local), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))


def solve(op, x, b, preconditioner=None, params=None):
    """solve() of element_centered_preconditioners_01.cc:108-203 on the device: `params` is the reference's "solver" JSON block
    (type CG | GMRES, max iterations 1000, abs tolerance 1e-10, rel tolerance 1e-2, max n tmp vectors 30);
    preconditioner: None (Identity), "Diagonal", an ASPoissonPreconditioner or a PreconditionChebyshev.  x is overwritten
    (the reference starts from x = 0).  Returns (n_iterations, last residual norm)."""
    params = params or {}
    types = {"CG": 0, "GMRES": 1}
    t = params.get("type", "")
    if t not in types:
        raise DasmError("Solver <" + t + "> is not known!")
    if preconditioner is None:
        kind, h = 0, None
    elif isinstance(preconditioner, str) and preconditioner == "Diagonal":
        kind, h = 1, None
    elif isinstance(preconditioner, ASPoissonPreconditioner):
        kind, h = 2, preconditioner.h
    elif isinstance(preconditioner, PreconditionChebyshev):
        kind, h = 3, preconditioner.h
    elif isinstance(preconditioner, PreconditionerGMG):
        kind, h = 4, preconditioner.h
    elif isinstance(preconditioner, RestrictedPreconditioner):
        kind, h = 5, preconditioner.h
    else:
        raise DasmError("Preconditioner <%r> is not known!" % (preconditioner,))
    n_it, res = ctypes.c_int(), ctypes.c_double()
    _check(lib().dasm_solve(op.h, types[t], kind, h, _ptr(x), _ptr(b), int(params.get("max iterations", 1000)),
                            ctypes.c_double(float(params.get("abs tolerance", 1e-10))), ctypes.c_double(float(params.get("rel tolerance", 1e-2))),
                            int(params.get("max n tmp vectors", 30)), ctypes.byref(n_it), ctypes.byref(res)))
    return n_it.value, res.value


# ---- factories (JSON vocabulary of the reference) -----------------------------------------------
def get_weighting_type(params):
    """include/precondition.templates.h:10-29."""
    t = params.get("weighting type", "symm")
    if t not in WEIGHT:
        raise DasmError("Weighting type <" + t + "> is not known!")
    return t


def _as_bool(v):
    if isinstance(v, str):
        return v.lower() == "true"
    return bool(v)


def create_fdm_preconditioner(op, params):
    """include/precondition.templates.h:162-247."""
    n_overlap = min(int(params.get("n overlap", 1)), op.degree)
    weight_type = get_weighting_type(params)
    sub_mesh = int(params.get("sub mesh approximation", 3))
    seq = params.get("weight sequence", "global" if n_overlap > 1 else "compressed")
    return ASPoissonPreconditioner(op, n_overlap, sub_mesh, weight_type, seq, _as_bool(params.get("overlap pre post", True)),
                                   _as_bool(params.get("element centric", True)))


def create_system_preconditioner(op, params):
    """include/precondition.templates.h:251-818; the types on the hot path: Chebyshev (around FDM or
    Diagonal) and FDM.  Other types (AMG, AdditiveSchwarzPreconditioner, ...) are out of scope."""
    t = params.get("type", "")
    if t == "Chebyshev":
        pp = params.get("preconditioner", {})
        pt = pp.get("type", "")
        if pt == "":
            raise DasmError("ExcNotImplemented")
        if pt == "Diagonal":
            fdm = None
            optimize = int(params.get("optimize", 3))
        elif pt == "FDM":
            fdm = create_fdm_preconditioner(op, pp)
            optimize = int(params.get("optimize", 2 if int(pp.get("n overlap", 1)) == 1 else 1))
        else:
            raise DasmError("Preconditioner <" + pt + "> is not known!")
        cheb = PreconditionChebyshev(op, fdm, int(params.get("degree", 3)), float(params.get("smoothing range", 20.)),
                                     params.get("polynomial type", "1st kind"), params.get("ev algorithm", None), optimize)
        cheb.estimate_eigenvalues()
        return cheb
    if t == "FDM":
        return create_fdm_preconditioner(op, params)
    raise DasmError("Preconditioner <" + t + "> is not known!")
