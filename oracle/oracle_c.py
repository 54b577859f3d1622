"""ctypes front end of the C restatement (oracle/dasm_oracle_c.c): a multi-threaded Chebyshev + FDM-ASM smoother
step on the CPU, used as cross-check of the numpy oracle and as the CPU baseline of bench.py.
TEST / BASELINE INFRASTRUCTURE ONLY."""
import ctypes
import os

import numpy as np

import dasm_oracle as o

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(os.path.join(_HERE, "liboracle_c.so"))
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def max_threads():
    return lib().oracle_c_max_threads()


class CSmoother:
    """Chebyshev(degree) around FDM-ASM (n_overlap = 1) on the data of a numpy-oracle problem."""

    def __init__(self, mesh, oop, oP, degree, max_ev, min_ev=None, polynomial_type="1st kind"):
        self.n = oop.n
        self.n_dofs = oop.n_dofs
        n3 = self.n ** 3
        self.N = np.ascontiguousarray(oop.basis.N, dtype=np.float64)
        self.Dq = np.ascontiguousarray(oop.basis.Dq, dtype=np.float64)
        idx = oop.cell_dofs.astype(np.int64).copy()
        valid = (idx != int(o.INVALID))
        valid &= ~oop.constrained[np.where(valid, idx, 0)]
        self.idx = np.ascontiguousarray(np.where(valid, idx, int(o.INVALID)).astype(np.uint32))
        # G[c, q, d, e] -> [c][6][n3]
        G = oop.G
        comp = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]
        self.G = np.ascontiguousarray(np.stack([G[:, :, a, b] for a, b in comp], axis=1), dtype=np.float64)
        self.S = np.ascontiguousarray(oP.S, dtype=np.float64)
        self.lam = np.ascontiguousarray(oP.lam, dtype=np.float64)
        self.wl = np.ascontiguousarray(oP.w_local, dtype=np.float64) if oP.w_local is not None else None
        self.w_pre = int(oP.weight_type in ("pre", "symm"))
        self.w_post = int(oP.weight_type in ("post", "symm", "ras"))
        # colouring by cell parity (cells of one colour share no DoF); needs even cell counts in periodic directions
        cols = {}
        for c in range(mesh.C):
            ijk = mesh.cell_ijk(c)
            cols.setdefault(tuple(i % 2 for i in ijk), []).append(c)
        for d in range(3):
            if mesh.periodic[d]:
                assert mesh.n_cells[d] % 2 == 0 and mesh.n_cells[d] >= 4, "colouring needs even periodic cell counts >= 4"
        self.colours = [np.ascontiguousarray(np.array(v, dtype=np.int64)) for v in cols.values()]
        self.ch = o.Chebyshev(oop, oP, degree=degree, polynomial_type=polynomial_type)
        self.ch.set_eigenvalues(max_ev, min_ev)
        self.coef = self.ch.coefficients()

    def vmult_A(self, x, y):
        L = lib()
        L.oracle_c_zero(ctypes.c_int64(self.n_dofs), _p(y))
        for col in self.colours:
            L.oracle_c_vmult_cells(self.n, _p(self.N), _p(self.Dq), _p(self.idx), _p(self.G), _p(col), ctypes.c_int64(len(col)),
                                   _p(x), _p(y))

    def vmult_P(self, r, z):
        L = lib()
        L.oracle_c_zero(ctypes.c_int64(self.n_dofs), _p(z))
        for col in self.colours:
            L.oracle_c_fdm_cells(self.n, _p(self.idx), _p(self.S), _p(self.lam), _p(self.wl), self.w_pre, self.w_post, _p(col),
                                 ctypes.c_int64(len(col)), _p(r), _p(z))

    def step(self, x, b):
        """x <- Chebyshev step (in place on a copy); returns the new iterate."""
        L = lib()
        n = ctypes.c_int64(self.n_dofs)
        cur = np.ascontiguousarray(x, dtype=np.float64).copy()
        old = None
        t = np.empty(self.n_dofs)
        z = np.empty(self.n_dofs)
        for term, (f1, f2) in enumerate(self.coef):
            self.vmult_A(cur, t)
            L.oracle_c_residual(n, _p(b), _p(t))
            self.vmult_P(t, z)
            new = np.empty(self.n_dofs)
            L.oracle_c_cheb_update(n, ctypes.c_double(f1), ctypes.c_double(f2), _p(cur), _p(old) if (old is not None and f1 != 0) else None,
                                   _p(z), _p(new))
            old, cur = cur, new
        return cur


# ---------------------------------------------------------------------------------------------------------------
# optimised CPU baseline (oracle/cpu_baseline.c): periodic Cartesian mesh, lexicographic DoF numbering
# ---------------------------------------------------------------------------------------------------------------
_blib = None


def baseline_lib():
    global _blib
    if _blib is None:
        _blib = ctypes.CDLL(os.path.join(_HERE, "libcpu_baseline.so"))
    return _blib


def baseline_max_threads():
    return baseline_lib().cpu_baseline_max_threads()


def baseline_set_threads(n):
    """torch.distributed.run exports OMP_NUM_THREADS=1; the CPU baseline is given all host cores explicitly"""
    baseline_lib().cpu_baseline_set_threads(int(n))


def lexicographic_cell_dofs(nc, k):
    """[C, n^3] DoF indices of the periodic (k nx) x (k ny) x (k nz) lattice, x fastest; cells lexicographic."""
    n = k + 1
    nd = [k * c for c in nc]
    cz, cy, cx = np.meshgrid(np.arange(nc[2]), np.arange(nc[1]), np.arange(nc[0]), indexing="ij")
    l, j, i = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    X = (k * cx.reshape(-1, 1) + i.reshape(1, -1)) % nd[0]
    Y = (k * cy.reshape(-1, 1) + j.reshape(1, -1)) % nd[1]
    Z = (k * cz.reshape(-1, 1) + l.reshape(1, -1)) % nd[2]
    return ((Z * nd[1] + Y) * nd[0] + X).astype(np.uint32), int(np.prod(nd))


class CartesianBaseline:
    """Chebyshev(degree) + FDM-ASM (n_overlap = 1) smoother step of oracle/cpu_baseline.c on a periodic Cartesian mesh."""

    def __init__(self, nc, lengths, k, degree, weight_type="symm", max_ev=2.4, min_ev=1.0, polynomial_type="1st kind"):
        self.nc = tuple(int(c) for c in nc)
        self.k = k
        n = k + 1
        self.h = np.array([lengths[d] / nc[d] for d in range(3)], dtype=np.float64)
        b = o.Basis1D(k)
        M, K = b.reference_mass_stiffness()
        self.M = np.ascontiguousarray(M, dtype=np.float64)
        self.K = np.ascontiguousarray(K, dtype=np.float64)
        S, lam = [], []
        for d in range(3):
            ext = (self.h[d], self.h[d], self.h[d])
            Md, Kd = o.laplace_tensor_product_matrix_1d(M, K, ext, (o.INTERNAL, o.INTERNAL), 1)
            s, l = o.generalized_eig(Md, Kd)
            S.append(s)
            lam.append(l)
        self.S = np.ascontiguousarray(np.stack(S), dtype=np.float64)
        self.lam = np.ascontiguousarray(np.stack(lam), dtype=np.float64)
        # weights of the gathered cell vector: valence = 2^(number of coordinates on the cell boundary)
        self.w = None
        self.w_pre = int(weight_type in ("pre", "symm"))
        self.w_post = int(weight_type in ("post", "symm"))
        if weight_type != "none":
            e = np.array([1 if (i == 0 or i == k) else 0 for i in range(n)])
            val = 2.0 ** (e[:, None, None] + e[None, :, None] + e[None, None, :])
            self.w = np.ascontiguousarray((1.0 / (np.sqrt(val) if weight_type == "symm" else val)).reshape(-1), dtype=np.float64)

        # coefficient recurrences of the oracle's Chebyshev (deal.II PreconditionChebyshev)
        helper = o.Chebyshev.__new__(o.Chebyshev)
        helper.degree, helper.poly, helper.smoothing_range = degree, polynomial_type, 20.0
        helper.set_eigenvalues(max_ev, min_ev)
        co = helper.coefficients()
        self.f1 = np.ascontiguousarray([c[0] for c in co], dtype=np.float64)
        self.f2 = np.ascontiguousarray([c[1] for c in co], dtype=np.float64)
        self.degree = degree
        self.n_dofs = int(np.prod([k * c for c in self.nc]))
        self.work = np.empty(4 * self.n_dofs)

    def step(self, x, b):
        """in place on x (float64, lexicographic numbering)"""
        nc = (ctypes.c_int * 3)(*self.nc)
        rc = baseline_lib().cpu_baseline_cheb_step(self.k, nc, _p(self.h), _p(self.M), _p(self.K), _p(self.S), _p(self.lam), _p(self.w),
                                                   self.w_pre, self.w_post, len(self.f1), _p(self.f1), _p(self.f2), _p(x), _p(b),
                                                   _p(self.work))
        if rc != 0:
            raise RuntimeError("cpu_baseline_cheb_step: unsupported size")
        return x
