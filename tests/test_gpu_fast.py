"""GPU parity of the warp-specialised kernels (kernels_fast.cuh): meshes with regular 4x4x4 bricks (all faces shared,
no constrained DoFs) go through laplace_fast_kernel / fdm_fast_kernel, the remaining bricks through the brick kernels.
Checked against the CPU oracle at oracle-sized meshes and against the brick kernels alone (DASM_NO_FAST=1) at larger
ones.  Tolerances as in test_gpu_parity.py."""
import os

import numpy as np
import pytest

import dasm_oracle as o
from __graft_entry__ import load_package
from parity_util import oracle_problem

pytestmark = pytest.mark.gpu

TOL = {"double": 1e-12, "float": 1e-5}
NPDT = {"double": np.float64, "float": np.float32}


@pytest.fixture(scope="module")
def pkg():
    return load_package()


@pytest.fixture(scope="module")
def ctx(pkg):
    return pkg.Context(0)


def relerr(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


FAST_MESHES = {
    "periodic_2bricks": dict(n_cells=(8, 4, 4), periodic=(1, 1, 1)),
    "periodic_aniso": dict(n_cells=(4, 8, 4), periodic=(1, 1, 1), length=(1.0, 3.0, 0.5)),
    "dirichlet_interior": dict(n_cells=(12, 12, 12), periodic=(0, 0, 0), dirichlet=True),
    "mixed": dict(n_cells=(8, 12, 4), periodic=(1, 0, 1), dirichlet=True),
    "ragged": dict(n_cells=(9, 4, 6), periodic=(1, 1, 1)),
}


class no_fast:
    """operators created inside use the brick kernels only"""

    def __enter__(self):
        os.environ["DASM_NO_FAST"] = "1"

    def __exit__(self, *a):
        os.environ.pop("DASM_NO_FAST", None)


@pytest.mark.parametrize("k", [2, 3, 4, 5, 6])
@pytest.mark.parametrize("number", ["double", "float"])
@pytest.mark.parametrize("name", ["periodic_2bricks", "periodic_aniso"])
def test_fast_vmult_vs_oracle(pkg, ctx, name, k, number):
    mesh = pkg.Mesh(ctx, **FAST_MESHES[name])
    op = pkg.LaplaceOperatorMatrixFree(mesh, k, number)
    assert op.n_fast_bricks() == 2
    oop, _ = oracle_problem(pkg, mesh, op, with_fdm=False)
    x = np.random.default_rng(k).uniform(-1, 1, op.n_dofs())
    yd = op.initialize_dof_vector()
    op.vmult(yd, op.to_device(x))
    assert relerr(op.to_host(yd), oop.vmult(x)) < TOL[number]


@pytest.mark.parametrize("k", [2, 3, 4, 5, 6])
@pytest.mark.parametrize("wt", ["none", "pre", "post", "symm", "ras"])
def test_fast_fdm_vs_oracle(pkg, ctx, k, wt):
    mesh = pkg.Mesh(ctx, **FAST_MESHES["periodic_aniso"])
    op = pkg.LaplaceOperatorMatrixFree(mesh, k, "double")
    fdm = pkg.create_fdm_preconditioner(op, {"weighting type": wt})
    assert fdm.n_fast_bricks() == (0 if wt == "ras" else 2)  # RAS: per-entry weights, generic kernel
    oop, oP = oracle_problem(pkg, mesh, op, 1, wt)
    x = np.random.default_rng(k).uniform(-1, 1, op.n_dofs())
    zd = op.initialize_dof_vector()
    fdm.vmult(zd, op.to_device(x))
    assert relerr(op.to_host(zd), oP.vmult(x)) < 1e-12


@pytest.mark.parametrize("k,number,wt,degree,poly,is_step", [(4, "double", "symm", 3, "1st kind", True), (3, "float", "post", 2, "4th kind", True),
                                                             (2, "double", "pre", 4, "1st kind", False), (4, "float", "symm", 3, "1st kind", False),
                                                             (5, "double", "symm", 3, "1st kind", True), (5, "float", "post", 3, "1st kind", False),
                                                             (6, "double", "post", 2, "4th kind", True), (6, "float", "symm", 3, "1st kind", True)])
def test_fast_chebyshev_vs_oracle(pkg, ctx, k, number, wt, degree, poly, is_step):
    mesh = pkg.Mesh(ctx, **FAST_MESHES["periodic_2bricks"])
    op = pkg.LaplaceOperatorMatrixFree(mesh, k, number)
    dt = NPDT[number]
    fdm = pkg.create_fdm_preconditioner(op, {"weighting type": wt})
    assert fdm.n_fast_bricks() == 2
    oop, oP = oracle_problem(pkg, mesh, op, 1, wt, dtype=dt)
    cheb = pkg.PreconditionChebyshev(op, fdm, degree=degree, polynomial_type=poly)
    cheb.set_eigenvalues(0.9, 2.2)
    och = o.Chebyshev(oop, oP, degree=degree, polynomial_type=poly)
    och.set_eigenvalues(2.2, 0.9)
    rng = np.random.default_rng(13)
    b = rng.uniform(-1, 1, op.n_dofs())
    x0 = rng.uniform(-1, 1, op.n_dofs())
    xd, bd = op.to_device(x0), op.to_device(b)
    if is_step:
        cheb.step(xd, bd)
        ref = och.step(x0.astype(dt), b.astype(dt))
    else:
        cheb.vmult(xd, bd)
        ref = och.vmult(b.astype(dt))
    assert relerr(op.to_host(xd), ref.astype(np.float64)) < TOL[number]


@pytest.mark.parametrize("name", ["dirichlet_interior", "mixed", "ragged"])
@pytest.mark.parametrize("k,number", [(4, "double"), (3, "double"), (4, "float"), (2, "double")])
def test_fast_and_brick_kernels_agree(pkg, ctx, name, k, number):
    """meshes with regular AND irregular bricks: vmult, FDM and a fused Chebyshev step with the fast kernels on the
    regular bricks equal the brick kernels on all bricks."""
    kw = FAST_MESHES[name]
    rng = np.random.default_rng(5)
    res = []
    for fast in (True, False):
        if fast:
            mesh = pkg.Mesh(ctx, **kw)
            op = pkg.LaplaceOperatorMatrixFree(mesh, k, number)
            fdm = pkg.create_fdm_preconditioner(op, {"weighting type": "symm"})
            assert op.n_fast_bricks() > 0 and fdm.n_fast_bricks() > 0
        else:
            with no_fast():
                mesh = pkg.Mesh(ctx, **kw)
                op = pkg.LaplaceOperatorMatrixFree(mesh, k, number)
                fdm = pkg.create_fdm_preconditioner(op, {"weighting type": "symm"})
            assert op.n_fast_bricks() == 0 and fdm.n_fast_bricks() == 0
        if not res:
            x = rng.uniform(-1, 1, op.n_dofs())
            b = rng.uniform(-1, 1, op.n_dofs())
            con = op.constrained_dofs()
            x[con] = 0
            b[con] = 0
        yd, zd = op.initialize_dof_vector(), op.initialize_dof_vector()
        op.vmult(yd, op.to_device(x))
        fdm.vmult(zd, op.to_device(x))
        cheb = pkg.PreconditionChebyshev(op, fdm, degree=3)
        cheb.set_eigenvalues(0.8, 2.5)
        xd = op.to_device(x)
        cheb.step(xd, op.to_device(b))
        res.append((op.to_host(yd).copy(), op.to_host(zd).copy(), op.to_host(xd).copy()))
    tol = 1e-13 if number == "double" else 2e-6
    for a, b_ in zip(res[0], res[1]):
        assert relerr(a, b_) < tol


def test_fast_large_properties(pkg, ctx):
    """BASELINE-sized building block (k = 4, 32^3 cells): linearity, symmetry of A and of the symmetric FDM-ASM,
    null space of the periodic Laplacian."""
    mesh = pkg.Mesh(ctx, (32, 32, 32), periodic=(1, 1, 1))
    op = pkg.LaplaceOperatorMatrixFree(mesh, 4, "double")
    fdm = pkg.create_fdm_preconditioner(op, {"weighting type": "symm"})
    assert op.n_fast_bricks() == 512 and fdm.n_fast_bricks() == 512
    rng = np.random.default_rng(7)
    n = op.n_dofs()
    x, y = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    out = op.initialize_dof_vector()

    def A(v):
        op.vmult(out, op.to_device(v))
        return op.to_host(out).copy()

    def P(v):
        fdm.vmult(out, op.to_device(v))
        return op.to_host(out).copy()

    Ax, Ay = A(x), A(y)
    assert relerr(A(2.5 * x - 0.5 * y), 2.5 * Ax - 0.5 * Ay) < 1e-12
    assert abs(x @ Ay - y @ Ax) < 1e-10 * abs(x @ Ay)
    assert np.linalg.norm(A(np.ones(n))) < 1e-9 * np.linalg.norm(Ax)
    Px, Py = P(x), P(y)
    assert abs(x @ Py - y @ Px) < 1e-10 * abs(x @ Py)
    assert relerr(P(2.5 * x - 0.5 * y), 2.5 * Px - 0.5 * Py) < 1e-12
