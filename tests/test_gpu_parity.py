"""GPU parity tests: the CUDA path through the C ABI against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): 1e-12 relative in double, 1e-5 relative in float, measured in the
l2 norm relative to the result; DoF indexing and constrained-DoF lists bit-exact."""
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


MESHES = {
    "periodic": dict(n_cells=(4, 3, 5), periodic=(1, 1, 1)),
    "dirichlet": dict(n_cells=(3, 4, 2), periodic=(0, 0, 0), dirichlet=True),
    "mixed_aniso": dict(n_cells=(5, 2, 3), periodic=(1, 0, 0), dirichlet=True, length=(2.0, 1.0, 3.0)),
    "sine": dict(n_cells=(3, 3, 3), periodic=(1, 1, 1), map_kind="sine"),
    "kershaw": dict(n_cells=(6, 2, 2), periodic=(0, 0, 0), dirichlet=True, map_kind="kershaw", map_params=(0.3, 0.3, 0, 0)),
}


def run_vmult(pkg, ctx, meshkw, k, number):
    mesh = pkg.Mesh(ctx, **meshkw)
    op = pkg.LaplaceOperatorMatrixFree(mesh, k, number)
    oop, _ = oracle_problem(pkg, mesh, op, with_fdm=False)
    rng = np.random.default_rng(k)
    x = rng.uniform(-1, 1, op.n_dofs())
    xd = op.to_device(x)
    yd = op.initialize_dof_vector()
    op.vmult(yd, xd)
    y = op.to_host(yd)
    ref = oop.vmult(x)
    return relerr(y, ref)


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("number", ["double", "float"])
def test_vmult_degrees_periodic(pkg, ctx, k, number):
    kw = dict(MESHES["periodic"])
    if k >= 6:
        kw["n_cells"] = (2, 2, 3)
    assert run_vmult(pkg, ctx, kw, k, number) < TOL[number]


@pytest.mark.parametrize("name", ["dirichlet", "mixed_aniso", "sine", "kershaw"])
@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6])
def test_vmult_meshes(pkg, ctx, name, k):
    assert run_vmult(pkg, ctx, MESHES[name], k, "double") < TOL["double"]


@pytest.mark.parametrize("name", ["sine", "kershaw", "mixed_aniso"])
@pytest.mark.parametrize("k,number", [(2, "double"), (4, "double"), (3, "float"), (6, "double")])
def test_vmult_quadratic_geometry(pkg, ctx, name, k, number):
    """mapping type "quadratic geometry" (operator.h:1035-1159): Jacobian rebuilt from 27 coefficients per cell."""
    mesh = pkg.Mesh(ctx, **MESHES[name])
    op = pkg.LaplaceOperatorMatrixFree(mesh, k, number, mapping_type="quadratic geometry")
    oop, _ = oracle_problem(pkg, mesh, op, with_fdm=False)
    x = np.random.default_rng(k).uniform(-1, 1, op.n_dofs())
    yd = op.initialize_dof_vector()
    op.vmult(yd, op.to_device(x))
    assert relerr(op.to_host(yd), oop.vmult(x)) < (1e-12 if number == "double" else 2e-5)


@pytest.mark.parametrize("name", ["sine", "kershaw"])
@pytest.mark.parametrize("k", [2, 4])
def test_vmult_linear_geometry(pkg, ctx, name, k):
    """mapping type "linear geometry" (operator.h:916-1033): trilinear map through the 8 cell vertices."""
    mesh = pkg.Mesh(ctx, **MESHES[name])
    op = pkg.LaplaceOperatorMatrixFree(mesh, k, "double", mapping_type="linear geometry")
    oop, _ = oracle_problem(pkg, mesh, op, with_fdm=False, mapping_degree=1)
    x = np.random.default_rng(k).uniform(-1, 1, op.n_dofs())
    yd = op.initialize_dof_vector()
    op.vmult(yd, op.to_device(x))
    assert relerr(op.to_host(yd), oop.vmult(x)) < 1e-12
    d = op.initialize_dof_vector()
    op.compute_inverse_diagonal(d)
    assert relerr(op.to_host(d), oop.inverse_diagonal()) < 1e-12


@pytest.mark.parametrize("name", ["sine", "kershaw", "mixed_aniso"])
@pytest.mark.parametrize("k,number", [(1, "double"), (2, "double"), (4, "double"), (3, "float"), (6, "double")])
def test_vmult_construct_q(pkg, ctx, name, k, number):
    """mapping type "construct q" (operator.h:712-746, 1221-1333): quadrature-point coordinates per cell, Jacobians by collocation
    differentiation in the kernel; vmult and the inverse diagonal against the oracle's restatement (for k = 1 the two Gauss points
    cannot represent the Q2 geometry, so this differs from "merged" - in the reference too)."""
    from parity_util import oracle_mesh
    mesh = pkg.Mesh(ctx, **MESHES[name])
    op = pkg.LaplaceOperatorMatrixFree(mesh, k, number, mapping_type="construct q")
    oop, _ = oracle_problem(pkg, mesh, op, with_fdm=False)
    G = o.construct_q_coefficients(oracle_mesh(mesh), o.Basis1D(k))
    if k >= 2:
        assert np.allclose(G, oop.G, rtol=1e-9, atol=1e-12 * np.abs(oop.G).max())
    oop.G = G.astype(oop.dtype)
    x = np.random.default_rng(k).uniform(-1, 1, op.n_dofs())
    yd = op.initialize_dof_vector()
    op.vmult(yd, op.to_device(x))
    assert relerr(op.to_host(yd), oop.vmult(x)) < (1e-12 if number == "double" else 5e-5)
    if number == "double":
        d = op.initialize_dof_vector()
        op.compute_inverse_diagonal(d)
        assert relerr(op.to_host(d), oop.inverse_diagonal()) < 1e-12


@pytest.mark.parametrize("name", ["periodic", "dirichlet", "kershaw"])
@pytest.mark.parametrize("k,number", [(2, "double"), (4, "double"), (3, "float")])
def test_plain_indices(pkg, ctx, name, k, number):
    """AdditionalData::compress_indices = false (operator.h:285-295, plain branch of do_cell_integral_global 1343-1350): (k+1)^3
    indices per cell instead of the 27 compressed start indices; vmult, FDM and a Chebyshev step equal the oracle."""
    mesh = pkg.Mesh(ctx, **MESHES[name])
    mt = "quadratic geometry" if name == "kershaw" else ""
    op = pkg.LaplaceOperatorMatrixFree(mesh, k, number, mapping_type="merged" if name == "kershaw" else mt, compress_indices=False)
    assert not op.uses_compressed_indices()
    dt = NPDT[number]
    oop, oP = oracle_problem(pkg, mesh, op, 1, "symm", dtype=dt)
    rng = np.random.default_rng(k)
    x, b = rng.uniform(-1, 1, op.n_dofs()), rng.uniform(-1, 1, op.n_dofs())
    con = op.constrained_dofs()
    x[con] = 0
    b[con] = 0
    yd = op.initialize_dof_vector()
    op.vmult(yd, op.to_device(x))
    assert relerr(op.to_host(yd), oop.vmult(x)) < TOL[number]
    fdm = pkg.create_fdm_preconditioner(op, {"weighting type": "symm"})
    fdm.vmult(yd, op.to_device(b))
    assert relerr(op.to_host(yd), oP.vmult(b)) < TOL[number]
    cheb = pkg.PreconditionChebyshev(op, fdm, degree=3)
    cheb.set_eigenvalues(0.9, 2.3)
    och = o.Chebyshev(oop, oP, degree=3)
    och.set_eigenvalues(2.3, 0.9)
    xd = op.to_device(x)
    cheb.step(xd, op.to_device(b))
    assert relerr(op.to_host(xd), och.step(x.astype(dt), b.astype(dt)).astype(np.float64)) < TOL[number]


@pytest.mark.parametrize("name", ["dirichlet", "mixed_aniso", "kershaw", "periodic"])
def test_rhs_and_constraints(pkg, ctx, name):
    """LaplaceOperatorBase::rhs (f = 1: b_i = int phi_i, constrained entries zero, operator.h:298-330) and get_constraints:
    against the oracle's assembly; the entries sum to the volume of the domain when nothing is constrained."""
    from parity_util import oracle_mesh
    k = 3
    mesh = pkg.Mesh(ctx, **MESHES[name])
    op = pkg.LaplaceOperatorMatrixFree(mesh, k, "double", mapping_type="" if name in ("periodic", "dirichlet", "mixed_aniso") else "merged")
    oop, _ = oracle_problem(pkg, mesh, op, with_fdm=False)
    om = oracle_mesh(mesh)
    bas = o.Basis1D(k)
    J = om.jacobians(bas)
    w = np.einsum("c,b,a->cba", bas.qw, bas.qw, bas.qw).reshape(-1)
    jxw = np.linalg.det(J) * w[None, :]
    N = bas.N
    loc = np.einsum("qk,rj,si,cqrs->ckji", N, N, N, jxw.reshape(-1, k + 1, k + 1, k + 1)).reshape(om.C, -1)
    ref = np.zeros(oop.n_dofs)
    np.add.at(ref, oop.idx.reshape(-1), (loc * oop.mask).reshape(-1))
    bd = op.initialize_dof_vector()
    op.rhs(bd, 1.0)
    got = op.to_host(bd)
    assert relerr(got, ref) < 1e-12
    assert np.array_equal(np.sort(op.get_constraints()), np.nonzero(oop.constrained)[0].astype(np.uint32))
    if name == "periodic":
        assert abs(got.sum() - np.prod(mesh.length)) < 1e-12 * np.prod(mesh.length)


def test_vmult_merged_on_cartesian_equals_default(pkg, ctx):
    mesh = pkg.Mesh(ctx, (3, 3, 3), periodic=(1, 1, 1))
    a = pkg.LaplaceOperatorMatrixFree(mesh, 3, "double", mapping_type="")
    b = pkg.LaplaceOperatorMatrixFree(mesh, 3, "double", mapping_type="merged")
    x = np.random.default_rng(1).uniform(-1, 1, a.n_dofs())
    ya, yb = a.initialize_dof_vector(), b.initialize_dof_vector()
    a.vmult(ya, a.to_device(x))
    b.vmult(yb, b.to_device(x))
    assert relerr(a.to_host(ya), b.to_host(yb)) < 1e-13


def test_unknown_mapping_type_raises(pkg, ctx):
    mesh = pkg.Mesh(ctx, (2, 2, 2), periodic=(1, 1, 1))
    with pytest.raises(pkg.DasmError, match="is not known"):
        pkg.LaplaceOperatorMatrixFree(mesh, 2, "double", mapping_type="cubic geometry")


def test_vmult_linearity_and_symmetry_large(pkg, ctx):
    """size-independent properties on a mesh too large for the oracle: linearity, symmetry x'Ay = y'Ax,
    constant in the null space of the periodic Laplacian."""
    mesh = pkg.Mesh(ctx, (16, 16, 16), periodic=(1, 1, 1))
    op = pkg.LaplaceOperatorMatrixFree(mesh, 4, "double")
    rng = np.random.default_rng(7)
    n = op.n_dofs()
    x, y = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    out = op.initialize_dof_vector()

    def A(v):
        op.vmult(out, op.to_device(v))
        return op.to_host(out).copy()

    Ax, Ay = A(x), A(y)
    assert relerr(A(2.5 * x - 0.5 * y), 2.5 * Ax - 0.5 * Ay) < 1e-12
    assert abs(x @ Ay - y @ Ax) < 1e-10 * abs(x @ Ay)
    assert np.linalg.norm(A(np.ones(n))) < 1e-9 * np.linalg.norm(Ax)


@pytest.mark.parametrize("name,k", [("periodic", 4), ("dirichlet", 3), ("sine", 2), ("kershaw", 2), ("mixed_aniso", 3)])
def test_inverse_diagonal(pkg, ctx, name, k):
    mesh = pkg.Mesh(ctx, **MESHES[name])
    op = pkg.LaplaceOperatorMatrixFree(mesh, k, "double")
    oop, _ = oracle_problem(pkg, mesh, op, with_fdm=False)
    d = op.initialize_dof_vector()
    op.compute_inverse_diagonal(d)
    assert relerr(op.to_host(d), oop.inverse_diagonal()) < 1e-12


@pytest.mark.parametrize("wt", ["none", "pre", "post", "symm", "ras"])
@pytest.mark.parametrize("seq", ["compressed", "global", "local", "dg"])
def test_fdm_weightings(pkg, ctx, wt, seq):
    mesh = pkg.Mesh(ctx, **MESHES["mixed_aniso"])
    op = pkg.LaplaceOperatorMatrixFree(mesh, 3, "double")
    fdm = pkg.create_fdm_preconditioner(op, {"n overlap": 1, "weighting type": wt, "weight sequence": seq})
    oop, oP = oracle_problem(pkg, mesh, op, 1, wt)
    x = np.random.default_rng(3).uniform(-1, 1, op.n_dofs())
    zd = op.initialize_dof_vector()
    fdm.vmult(zd, op.to_device(x))
    assert relerr(op.to_host(zd), oP.vmult(x)) < 1e-12
    assert fdm.is_symmetric() == oP.is_symmetric()
    if wt in ("pre", "post", "symm"):
        assert relerr(fdm.weights(), oP.weights) < 1e-14


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("number", ["double", "float"])
def test_fdm_degrees(pkg, ctx, k, number):
    kw = dict(MESHES["periodic"])
    if k >= 6:
        kw["n_cells"] = (2, 2, 3)
    mesh = pkg.Mesh(ctx, **kw)
    op = pkg.LaplaceOperatorMatrixFree(mesh, k, number)
    fdm = pkg.create_fdm_preconditioner(op, {"weighting type": "symm"})
    oop, oP = oracle_problem(pkg, mesh, op, 1, "symm")
    x = np.random.default_rng(k).uniform(-1, 1, op.n_dofs())
    x -= x.mean()
    zd = op.initialize_dof_vector()
    fdm.vmult(zd, op.to_device(x))
    assert relerr(op.to_host(zd), oP.vmult(x)) < (1e-11 if number == "double" else 2e-4)


@pytest.mark.parametrize("name", ["dirichlet", "sine", "kershaw"])
def test_fdm_meshes(pkg, ctx, name):
    mesh = pkg.Mesh(ctx, **MESHES[name])
    op = pkg.LaplaceOperatorMatrixFree(mesh, 3, "double")
    fdm = pkg.create_fdm_preconditioner(op, {"weighting type": "post"})
    oop, oP = oracle_problem(pkg, mesh, op, 1, "post")
    x = np.random.default_rng(5).uniform(-1, 1, op.n_dofs())
    zd = op.initialize_dof_vector()
    fdm.vmult(zd, op.to_device(x))
    assert relerr(op.to_host(zd), oP.vmult(x)) < 1e-12
    # eigen-decomposition of one cell/direction: S^T M S = I is invariant to sign/order; compare S diag(1/l) S^T
    S, lam = fdm.instance(0, 0)
    So, lo = oP.S[oP.mesh.cell_order[0], 0], oP.lam[oP.mesh.cell_order[0], 0]
    assert np.allclose(S @ np.diag(1 / lam) @ S.T, So @ np.diag(1 / lo) @ So.T, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("n_overlap,wt", [(2, "post"), (2, "symm"), (2, "none"), (3, "post"), (2, "ras")])
def test_fdm_overlap(pkg, ctx, n_overlap, wt):
    mesh = pkg.Mesh(ctx, (3, 4, 3), periodic=(1, 0, 0), dirichlet=True)
    op = pkg.LaplaceOperatorMatrixFree(mesh, 3, "double")
    fdm = pkg.create_fdm_preconditioner(op, {"n overlap": n_overlap, "weighting type": wt})
    assert fdm.patch_size_1d() == 3 - 1 + 2 * n_overlap
    oop, oP = oracle_problem(pkg, mesh, op, n_overlap, wt)
    x = np.random.default_rng(11).uniform(-1, 1, op.n_dofs())
    zd = op.initialize_dof_vector()
    fdm.vmult(zd, op.to_device(x))
    assert relerr(op.to_host(zd), oP.vmult(x)) < 1e-11


@pytest.mark.parametrize("name,k,wt", [("periodic", 3, "symm"), ("dirichlet", 2, "post"), ("mixed_aniso", 4, "pre"), ("kershaw", 3, "none"),
                                       ("periodic", 4, "post")])
def test_fdm_vertex_patch(pkg, ctx, name, k, wt):
    """fdmv: element centric = false, patches of (2k-1)^3 DoFs around vertices (matrix_free.h:90-92, dof_tools.h:206-300,
    tensor_product_matrix_creator.h:7-61)."""
    mesh = pkg.Mesh(ctx, **MESHES[name])
    op = pkg.LaplaceOperatorMatrixFree(mesh, k, "double")
    fdm = pkg.create_fdm_preconditioner(op, {"element centric": False, "weighting type": wt})
    assert fdm.patch_size_1d() == 2 * k - 1
    oop, oP = oracle_problem(pkg, mesh, op, 1, wt, element_centric=False)
    x = np.random.default_rng(17).uniform(-1, 1, op.n_dofs())
    zd = op.initialize_dof_vector()
    fdm.vmult(zd, op.to_device(x))
    assert relerr(op.to_host(zd), oP.vmult(x)) < 1e-11
    # as Chebyshev preconditioner (label cheby-2-1-<wt>-v-c of matrix_free_loop_08)
    cheb = pkg.PreconditionChebyshev(op, fdm, degree=2, optimize=1)
    cheb.set_eigenvalues(0.5, 3.0)
    och = o.Chebyshev(oop, oP, degree=2)
    och.set_eigenvalues(3.0, 0.5)
    b = np.random.default_rng(18).uniform(-1, 1, op.n_dofs())
    xd = op.to_device(x)
    cheb.step(xd, op.to_device(b))
    assert relerr(op.to_host(xd), och.step(x, b)) < 1e-11


def test_fdm_n_instances_cartesian(pkg, ctx):
    mesh = pkg.Mesh(ctx, (4, 4, 4), periodic=(1, 1, 1))
    op = pkg.LaplaceOperatorMatrixFree(mesh, 4, "double")
    fdm = pkg.create_fdm_preconditioner(op, {})
    assert fdm.n_fdm_instances() == 1  # O(1) instances on a Cartesian mesh (matrix_free.h:1000-1004)
    mesh2 = pkg.Mesh(ctx, (4, 4, 4), periodic=(0, 0, 0))
    op2 = pkg.LaplaceOperatorMatrixFree(mesh2, 4, "double")
    assert pkg.create_fdm_preconditioner(op2, {}).n_fdm_instances() == 3  # left boundary / interior / right boundary


CHEB_CASES = [
    ("periodic", 4, "double", "symm", 3, "1st kind", True),
    ("periodic", 4, "double", "post", 1, "1st kind", True),
    ("periodic", 3, "float", "symm", 3, "1st kind", True),
    ("dirichlet", 3, "double", "post", 3, "4th kind", True),
    ("kershaw", 2, "double", "symm", 2, "1st kind", False),
    ("sine", 3, "double", "pre", 4, "4th kind", False),
    ("mixed_aniso", 2, "double", "diag", 3, "1st kind", True),
    ("dirichlet", 4, "float", "diag", 2, "4th kind", False),
]


@pytest.mark.parametrize("name,k,number,wt,degree,poly,is_step", CHEB_CASES)
def test_chebyshev_step_and_vmult(pkg, ctx, name, k, number, wt, degree, poly, is_step):
    mesh = pkg.Mesh(ctx, **MESHES[name])
    op = pkg.LaplaceOperatorMatrixFree(mesh, k, number)
    dt = NPDT[number]
    if wt == "diag":
        fdm = None
        oop, _ = oracle_problem(pkg, mesh, op, with_fdm=False, dtype=dt)
        oP = o.JacobiPreconditioner(oop)
    else:
        fdm = pkg.create_fdm_preconditioner(op, {"weighting type": wt})
        oop, oP = oracle_problem(pkg, mesh, op, 1, wt, dtype=dt)
    cheb = pkg.PreconditionChebyshev(op, fdm, degree=degree, polynomial_type=poly)
    cheb.set_eigenvalues(0.9, 2.2)
    och = o.Chebyshev(oop, oP, degree=degree, polynomial_type=poly)
    och.set_eigenvalues(2.2, 0.9)
    rng = np.random.default_rng(13)
    b = rng.uniform(-1, 1, op.n_dofs())
    x0 = rng.uniform(-1, 1, op.n_dofs())
    xd = op.to_device(x0)
    bd = op.to_device(b)
    if is_step:
        cheb.step(xd, bd)
        ref = och.step(x0.astype(dt), b.astype(dt))
    else:
        cheb.vmult(xd, bd)
        ref = och.vmult(b.astype(dt))
    assert relerr(op.to_host(xd), ref.astype(np.float64)) < (1e-12 if number == "double" else 1e-5)
    # host-buffer entry point (the e2e path of bench.py)
    xh = x0.copy()
    if is_step:
        cheb.step_host(xh, b)
    else:
        cheb.vmult_host(xh, b)
    assert relerr(xh, ref.astype(np.float64)) < (1e-12 if number == "double" else 1e-5)


@pytest.mark.parametrize("wt,alg", [("post", "power iteration"), ("symm", "lanczos"), ("diag", "lanczos"), ("none", "power iteration")])
def test_eigenvalue_estimates(pkg, ctx, wt, alg):
    mesh = pkg.Mesh(ctx, (3, 3, 2), periodic=(0, 0, 0), dirichlet=True)
    op = pkg.LaplaceOperatorMatrixFree(mesh, 3, "double")
    if wt == "diag":
        fdm = None
        oop, _ = oracle_problem(pkg, mesh, op, with_fdm=False)
        oP = o.JacobiPreconditioner(oop)
    else:
        fdm = pkg.create_fdm_preconditioner(op, {"weighting type": wt})
        oop, oP = oracle_problem(pkg, mesh, op, 1, wt)
    cheb = pkg.PreconditionChebyshev(op, fdm, degree=3, ev_algorithm=alg, eig_cg_n_iterations=20)
    mn, mx = cheb.estimate_eigenvalues()
    och = o.Chebyshev(oop, oP, degree=3, ev_algorithm=alg, eig_cg_n_iterations=20)
    omn, omx = och.estimate_eigenvalues()
    assert mx == pytest.approx(omx, rel=1e-8)
    assert mn == pytest.approx(omn, rel=1e-6)


def test_factory_system_preconditioner(pkg, ctx):
    mesh = pkg.Mesh(ctx, (3, 3, 3), periodic=(1, 1, 1))
    op = pkg.LaplaceOperatorMatrixFree(mesh, 3, "double")
    pc = pkg.create_system_preconditioner(op, {"type": "Chebyshev", "degree": 2, "preconditioner": {"type": "FDM", "n overlap": 1,
                                                                                                  "weighting type": "post"}})
    assert isinstance(pc, pkg.PreconditionChebyshev)
    with pytest.raises(pkg.DasmError, match="is not known"):
        pkg.create_system_preconditioner(op, {"type": "Bogus"})
    with pytest.raises(pkg.DasmError, match="is not known"):
        pkg.create_fdm_preconditioner(op, {"weighting type": "bogus"})


def test_smoother_reduces_residual_large(pkg, ctx):
    """property at a larger size: a Chebyshev(3)+FDM step reduces the high-frequency residual."""
    mesh = pkg.Mesh(ctx, (12, 12, 12), periodic=(0, 0, 0), dirichlet=True)
    op = pkg.LaplaceOperatorMatrixFree(mesh, 4, "double")
    fdm = pkg.create_fdm_preconditioner(op, {"weighting type": "symm"})
    cheb = pkg.PreconditionChebyshev(op, fdm, degree=3)
    mn, mx = cheb.estimate_eigenvalues()
    assert 1.0 < mx < 4.0
    rng = np.random.default_rng(0)
    n = op.n_dofs()
    con = op.constrained_dofs()
    xt = rng.uniform(-1, 1, n)
    xt[con] = 0
    xtd = op.to_device(xt)
    bd = op.initialize_dof_vector()
    op.vmult(bd, xtd)
    xd = op.initialize_dof_vector()
    e0 = np.linalg.norm(xt)
    for _ in range(3):
        cheb.step(xd, bd)
    e1 = np.linalg.norm(op.to_host(xd) - xt)
    assert e1 < 0.5 * e0


# ---- Krylov solvers on the device (dasm_solve) against the oracle's restatement of the reference's solve() ---------------------
@pytest.mark.parametrize("solver,precon", [("CG", None), ("CG", "Diagonal"), ("CG", "fdm-symm"), ("GMRES", "Diagonal"),
                                           ("GMRES", "fdm-post"), ("GMRES", "cheb-fdm-post"), ("CG", "cheb-diag")])
def test_krylov_iteration_counts(pkg, ctx, solver, precon):
    """element_centered_preconditioners_01.cc:108-203: CG / GMRES (right preconditioning) with ReductionControl(1000, 1e-10, 1e-2);
    same iteration count as the oracle and the same solution to 1e-10 (the oracle's solvers are pinned to the reference's
    iteration counts 24 / 23 / 9 in tests/test_oracle_golden.py)."""
    mesh = pkg.Mesh(ctx, (6, 5, 4), periodic=(0, 0, 0), dirichlet=True)
    k = 3
    op = pkg.LaplaceOperatorMatrixFree(mesh, k, "double")
    wt = "symm" if precon in (None, "Diagonal", "fdm-symm", "cheb-diag") else "post"
    oop, oP = oracle_problem(pkg, mesh, op, 1, wt)
    n = op.n_dofs()
    rng = np.random.default_rng(11)
    b = rng.uniform(-1, 1, n)
    b[op.constrained_dofs()] = 0.0
    A = lambda v: oop.vmult(v, copy_constrained=True)
    if precon is None:
        P, dev_P = (lambda v: v.copy()), None
    elif precon == "Diagonal":
        P, dev_P = o.JacobiPreconditioner(oop).vmult, "Diagonal"
    elif precon.startswith("fdm"):
        dev_P = pkg.create_fdm_preconditioner(op, {"weighting type": wt})
        P = oP.vmult
    elif precon == "cheb-diag":
        dev_P = pkg.PreconditionChebyshev(op, None, degree=3)
        dev_P.set_eigenvalues(0.5, 2.0)
        och = o.Chebyshev(oop, o.JacobiPreconditioner(oop), degree=3)
        och.set_eigenvalues(2.0, 0.5)
        P = och.vmult
    else:
        fdm = pkg.create_fdm_preconditioner(op, {"weighting type": wt})
        dev_P = pkg.PreconditionChebyshev(op, fdm, degree=2)
        dev_P.set_eigenvalues(1.0, 2.2)
        och = o.Chebyshev(oop, oP, degree=2)
        och.set_eigenvalues(2.2, 1.0)
        P = och.vmult
    if solver == "CG":
        x_ref, its_ref = o.solve_cg(A, P, b)
    else:
        x_ref, its_ref = o.solve_gmres(A, P, b)
    xd = op.initialize_dof_vector()
    its, res = pkg.solve(op, xd, op.to_device(b), dev_P, {"type": solver})
    assert its == its_ref
    assert relerr(op.to_host(xd), x_ref) < 1e-10
    assert res <= 1e-2 * np.linalg.norm(b) * (1 + 1e-8)


# ---- orientation-aware compressed access on the device against the reference's golden and the oracle -------------------------------
def test_reduced_access_orientation(pkg, ctx):
    """ConstraintInfoReduced::read_dof_values / distribute_local_to_global with the packed orientation word
    (include/vector_access_reduced.h:267-548, include/reduced_access.h:528-702): all 19 post-variant cases of the reference's
    reduced_access_02.result bit-exactly, random orientation words against the oracle's adjust_for_orientation, and
    distribute = read^T."""
    import json
    import os
    import torch
    from test_oracle_golden import _reduced_access_setup
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reduced_access_02.json")))
    dev = torch.device("cuda", 0)
    n_checked = 0
    for case in gold:
        degree, do_post, orientations = case["args"][0], case["args"][1], case["args"][2:]
        dofs, g = _reduced_access_setup(3, degree)
        word = o.compress_orientation(orientations, True)
        cidx = torch.tensor(np.array(dofs, dtype=np.int64).reshape(1, 27), dtype=torch.int64, device=dev).to(torch.int32)
        ori = torch.tensor([word], dtype=torch.int64, device=dev).to(torch.int32)
        src = torch.tensor(g, dtype=torch.float64, device=dev)
        loc = pkg.reduced_access_read(degree, cidx, ori, src)
        assert [int(v) for v in loc[0].cpu().numpy()] == case["local"]  # (the pre and the post variant print the same vector)
        n_checked += 1
    assert n_checked == 38
    # random words (line bits and quad flags 0 / 1, the pinned rows of the orientation table), several cells, degrees 2..5
    rng = np.random.default_rng(3)
    for degree in (2, 3, 5):
        dofs, g = _reduced_access_setup(3, degree)
        n_cells = 7
        words, ref = [], []
        table = o.orientation_table(degree - 1)
        for c in range(n_cells):
            orientations = list(rng.integers(0, 2, 18))
            w = o.compress_orientation([int(v) for v in orientations], True)
            words.append(w)
            ref.append(o.gather_post([float(v) for v in g], 3, degree, dofs, w, table))
        cidx = torch.tensor(np.tile(np.array(dofs, dtype=np.int64), (n_cells, 1)), device=dev).to(torch.int32)
        ori = torch.tensor(np.array(words, dtype=np.int64), device=dev).to(torch.int32)
        src = torch.tensor(g, dtype=torch.float64, device=dev)
        loc = pkg.reduced_access_read(degree, cidx, ori, src)
        assert np.array_equal(loc.cpu().numpy(), np.array(ref))
        # transpose: <read(x), y> = <x, distribute(y)>
        x = torch.rand(len(g), dtype=torch.float64, device=dev)
        y = torch.rand((n_cells, (degree + 1) ** 3), dtype=torch.float64, device=dev)
        rx = pkg.reduced_access_read(degree, cidx, ori, x)
        dy = torch.zeros(len(g), dtype=torch.float64, device=dev)
        pkg.reduced_access_distribute(degree, cidx, ori, dy, y)
        assert abs(float((rx * y).sum()) - float((x * dy).sum())) < 1e-10


@pytest.mark.parametrize("k,number,n_cells", [(4, "double", (8, 8, 8)), (3, "float", (4, 4, 8)), (2, "double", (3, 4, 2))])
def test_step_host_batch(pkg, ctx, k, number, n_cells):
    """the pipelined host entry point (dasm_cheb_step_host_batch): five chained problems on two host buffer pairs give the
    results of five blocking dasm_cheb_step_host calls on the same data"""
    import torch
    periodic = (1, 1, 1) if n_cells[0] % 4 == 0 else (0, 0, 0)
    mesh = pkg.Mesh(ctx, n_cells, periodic=periodic, dirichlet=True)
    op = pkg.LaplaceOperatorMatrixFree(mesh, k, number)
    fdm = pkg.create_fdm_preconditioner(op, {"weighting type": "symm"})
    cheb = pkg.PreconditionChebyshev(op, fdm, degree=3)
    cheb.set_eigenvalues(0.9, 2.2)
    rng = np.random.default_rng(1)
    n = op.n_dofs()
    x0 = [rng.uniform(-1, 1, n) for _ in range(2)]
    b0 = [rng.uniform(-1, 1, n) for _ in range(2)]
    for v in x0 + b0:
        v[op.constrained_dofs()] = 0
    # sequential: problem i uses pair i % 2, x updated in place
    xs = [v.copy() for v in x0]
    for i in range(5):
        cheb.step_host(xs[i % 2], b0[i % 2])
    hx = [torch.tensor(v).pin_memory() for v in x0]
    hb = [torch.tensor(v).pin_memory() for v in b0]
    cheb.step_host_batch([hx[i % 2].numpy() for i in range(5)], [hb[i % 2].numpy() for i in range(5)])
    for i in range(2):
        assert relerr(hx[i].numpy(), xs[i]) < (1e-13 if number == "double" else 1e-6)
    # vmult variant
    ys = [np.zeros(n), np.zeros(n)]
    for i in range(2):
        cheb.vmult_host(ys[i], b0[i])
    hy = [torch.zeros(n, dtype=torch.float64).pin_memory() for _ in range(2)]
    cheb.vmult_host_batch([hy[0].numpy(), hy[1].numpy()], [hb[0].numpy(), hb[1].numpy()])
    for i in range(2):
        assert relerr(hy[i].numpy(), ys[i]) < (1e-13 if number == "double" else 1e-6)


def test_deterministic_mode(pkg, ctx, monkeypatch):
    """DASM_DETERMINISTIC=1: the cells are coloured so that no two cells of a colour share a vector entry and the generic kernels run
    colour by colour: repeated applications are bitwise identical (the default kernels add shared-face contributions with red.add in
    varying order, reproducible to 1e-15 only) and agree with the oracle and with the default path."""
    import importlib
    k = 4
    rng = np.random.default_rng(5)

    def run(op, fdm, cheb, x, b):
        xd, bd, yd = op.to_device(x), op.to_device(b), op.initialize_dof_vector()
        op.vmult(yd, xd)
        y = op.to_host(yd).copy()
        fdm.vmult(yd, bd)
        z = op.to_host(yd).copy()
        cheb.step(xd, bd)
        return y, z, op.to_host(xd).copy()

    def build(number="double", n_overlap=1):
        mesh = pkg.Mesh(ctx, (8, 8, 8), periodic=(1, 1, 1))
        op = pkg.LaplaceOperatorMatrixFree(mesh, k, number)
        fdm = pkg.create_fdm_preconditioner(op, {"weighting type": "symm", "n overlap": n_overlap})
        cheb = pkg.PreconditionChebyshev(op, fdm, degree=3)
        cheb.set_eigenvalues(0.9, 2.2)
        return mesh, op, fdm, cheb

    mesh0, op0, fdm0, cheb0 = build()                       # default path (TMA-fed kernels)
    assert op0.n_fast_bricks() > 0
    x, b = rng.uniform(-1, 1, op0.n_dofs()), rng.uniform(-1, 1, op0.n_dofs())
    ref = run(op0, fdm0, cheb0, x, b)
    monkeypatch.setenv("DASM_DETERMINISTIC", "1")
    mesh1, op1, fdm1, cheb1 = build()
    assert op1.n_fast_bricks() == 0
    first = run(op1, fdm1, cheb1, x, b)
    for _ in range(3):
        again = run(op1, fdm1, cheb1, x, b)
        for u, v in zip(first, again):
            assert np.array_equal(u, v)                      # bitwise
    for u, v in zip(first, ref):
        assert relerr(u, v) < 1e-13
    oop, oP = oracle_problem(pkg, mesh1, op1, 1, "symm")
    assert relerr(first[0], oop.vmult(x)) < 1e-12 and relerr(first[1], oP.vmult(b)) < 1e-12
    # overlapping patches (own colouring of the patches) and an unstructured mesh
    mesh2, op2, fdm2, cheb2 = build(n_overlap=2)
    a1, a2 = run(op2, fdm2, cheb2, x, b), run(op2, fdm2, cheb2, x, b)
    assert all(np.array_equal(u, v) for u, v in zip(a1, a2))
    _, oP2 = oracle_problem(pkg, mesh2, op2, 2, "symm")
    assert relerr(a1[1], oP2.vmult(b)) < 1e-12
    grid = importlib.import_module("dealii-asm_b200.grid")
    g = grid.hyper_ball(1)
    opb = pkg.LaplaceOperatorMatrixFree.from_arrays(ctx, g["vertices"], g["cells"], 3, g["support"])
    fdmb = pkg.create_fdm_preconditioner(opb, {"weighting type": "symm"})
    chebb = pkg.PreconditionChebyshev(opb, fdmb, degree=3)
    chebb.set_eigenvalues(0.9, 2.2)
    xb, bb = rng.uniform(-1, 1, opb.n_dofs()), rng.uniform(-1, 1, opb.n_dofs())
    b1, b2 = run(opb, fdmb, chebb, xb, bb), run(opb, fdmb, chebb, xb, bb)
    assert all(np.array_equal(u, v) for u, v in zip(b1, b2))
    # Krylov solve: identical iteration count and bitwise identical solution in repeated runs
    rhs = op1.to_device(b - b.mean())                        # (periodic mesh: the right-hand side must be orthogonal to the constants)
    sols = []
    for _ in range(2):
        sol = op1.initialize_dof_vector()
        its, res = pkg.solve(op1, sol, rhs, cheb1, {"type": "GMRES", "rel tolerance": 1e-6})
        sols.append((its, res, op1.to_host(sol).copy()))
    assert sols[0][0] == sols[1][0] and sols[0][1] == sols[1][1] and np.array_equal(sols[0][2], sols[1][2])
