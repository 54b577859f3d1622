"""GPU parity of the rows SURVEY.md section 8 marks "next": multigrid V-cycle (two-level transfers, cycle, outer Krylov solver) and
exact-block additive Schwarz, against the oracle (whose multigrid is pinned to the reference's dummy_mg_chebyshev_fdm_* iteration
counts in tests/test_oracle_golden.py).  Tolerances: 1e-12 (double) / 1e-5 (float) relative l2 for single operations, 1e-10 for a
whole V-cycle in double; iteration counts identical."""
import numpy as np
import pytest

import dasm_oracle as o
from __graft_entry__ import load_package
from parity_util import oracle_problem

pytestmark = pytest.mark.gpu

NPDT = {"double": np.float64, "float": np.float32}


@pytest.fixture(scope="module")
def pkg():
    return load_package()


@pytest.fixture(scope="module")
def ctx(pkg):
    return pkg.Context(0)


def relerr(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def make_level(pkg, ctx, n_cells, k, number, periodic, wt="post", n_overlap=1, with_fdm=True):
    mesh = pkg.Mesh(ctx, n_cells, periodic=periodic, dirichlet=True)
    op = pkg.LaplaceOperatorMatrixFree(mesh, k, number)
    oop, oP = oracle_problem(pkg, mesh, op, n_overlap, wt, dtype=NPDT[number], with_fdm=with_fdm)
    from parity_util import oracle_mesh
    return dict(mesh=mesh, op=op, oop=oop, oP=oP, omesh=oracle_mesh(mesh))


@pytest.mark.parametrize("case", [
    dict(fine=(4, 4, 4), coarse=(2, 2, 2), kf=2, kc=2, periodic=(0, 0, 0)),
    dict(fine=(4, 2, 6), coarse=(2, 1, 3), kf=3, kc=3, periodic=(0, 0, 0)),
    dict(fine=(8, 4, 4), coarse=(4, 2, 2), kf=2, kc=2, periodic=(1, 1, 1)),
    dict(fine=(8, 8, 8), coarse=(4, 4, 4), kf=4, kc=4, periodic=(1, 0, 1)),  # lex bricks on the fine level
    dict(fine=(3, 2, 2), coarse=(3, 2, 2), kf=4, kc=2, periodic=(0, 0, 0)),  # polynomial transfer
    dict(fine=(4, 4, 4), coarse=(4, 4, 4), kf=3, kc=1, periodic=(1, 1, 0)),
])
@pytest.mark.parametrize("number", ["double", "float"])
def test_two_level_transfer(pkg, ctx, case, number):
    """MGTwoLevelTransfer::prolongate_and_add / restrict_and_add (geometric and polynomial) against the oracle."""
    lf = make_level(pkg, ctx, case["fine"], case["kf"], number, case["periodic"], with_fdm=False)
    lc = make_level(pkg, ctx, case["coarse"], case["kc"], number, case["periodic"], with_fdm=False)
    tr = pkg.MGTwoLevelTransfer(lf["op"], lc["op"])
    otr = o.TwoLevelTransfer(lf["omesh"], lf["oop"], lc["omesh"], lc["oop"])
    rng = np.random.default_rng(3)
    uc, uf = rng.uniform(-1, 1, lc["op"].n_dofs()), rng.uniform(-1, 1, lf["op"].n_dofs())
    base_f, base_c = rng.uniform(-1, 1, lf["op"].n_dofs()), rng.uniform(-1, 1, lc["op"].n_dofs())
    base_f[lf["oop"].constrained] = 0
    base_c[lc["oop"].constrained] = 0
    tol = 1e-12 if number == "double" else 2e-6
    d = lf["op"].to_device(base_f)
    tr.prolongate_and_add(d, lc["op"].to_device(uc))
    ref = otr.prolongate_and_add(base_f.astype(NPDT[number]), uc.astype(NPDT[number]))
    assert relerr(lf["op"].to_host(d), ref.astype(np.float64)) < tol
    d = lc["op"].to_device(base_c)
    tr.restrict_and_add(d, lf["op"].to_device(uf))
    ref = otr.restrict_and_add(base_c.astype(NPDT[number]), uf.astype(NPDT[number]))
    assert relerr(lc["op"].to_host(d), ref.astype(np.float64)) < tol


def build_mg(pkg, ctx, sizes, k, number, wt, periodic=(0, 0, 0), degree=1, evs=(1.0, 2.2)):
    levels = [make_level(pkg, ctx, s, k, number, periodic, wt) for s in sizes]
    sm, osm = [], []
    for lv in levels:
        fdm = pkg.create_fdm_preconditioner(lv["op"], {"weighting type": wt})
        ch = pkg.PreconditionChebyshev(lv["op"], fdm, degree=degree)
        ch.set_eigenvalues(*evs)
        och = o.Chebyshev(lv["oop"], lv["oP"], degree=degree)
        och.set_eigenvalues(evs[1], evs[0])
        lv["fdm"] = fdm
        sm.append(ch)
        osm.append(och)
    otr = [None] + [o.TwoLevelTransfer(levels[l]["omesh"], levels[l]["oop"], levels[l - 1]["omesh"], levels[l - 1]["oop"])
                    for l in range(1, len(levels))]
    omg = o.Multigrid([lv["oop"] for lv in levels], osm, otr)
    return levels, sm, omg


@pytest.mark.parametrize("wt,degree", [("post", 1), ("symm", 2)])
def test_v_cycle_double(pkg, ctx, wt, degree):
    """PreconditionerGMG::vmult (include/multigrid.h:463-469): one V-cycle over 1 / 8 / 64 / 512 cells, Q3, Chebyshev + FDM smoothers
    on every level, against the oracle's restatement of Multigrid::level_v_step."""
    sizes = [(1, 1, 1), (2, 2, 2), (4, 4, 4), (8, 8, 8)]
    levels, sm, omg = build_mg(pkg, ctx, sizes, 3, "double", wt, degree=degree)
    mg = pkg.PreconditionerGMG([lv["op"] for lv in levels], sm)
    top = levels[-1]
    rng = np.random.default_rng(5)
    r = rng.uniform(-1, 1, top["op"].n_dofs())
    r[top["oop"].constrained] = 0
    d = top["op"].initialize_dof_vector()
    mg.vmult(d, top["op"].to_device(r))
    assert relerr(top["op"].to_host(d), omg.vmult(r)) < 1e-10


@pytest.mark.parametrize("solver,wt", [("GMRES", "post"), ("CG", "symm")])
def test_multigrid_krylov_iteration_counts(pkg, ctx, solver, wt):
    """configs[2]-style solve: double outer Krylov solver, float multigrid levels (element_centered_preconditioners_01.cc:787-792);
    iteration count identical to the oracle's, solution equal to float level accuracy."""
    sizes = [(1, 1, 1), (2, 2, 2), (4, 4, 4), (8, 8, 8)]
    k = 3
    levels, sm, omg = build_mg(pkg, ctx, sizes, k, "float", wt, degree=2 if solver == "CG" else 1)
    mesh = levels[-1]["mesh"]
    op = pkg.LaplaceOperatorMatrixFree(mesh, k, "double")
    oop, _ = oracle_problem(pkg, mesh, op, with_fdm=False)
    mg = pkg.PreconditionerGMG([lv["op"] for lv in levels], sm, outer_op=op)
    rng = np.random.default_rng(7)
    b = rng.uniform(-1, 1, op.n_dofs())
    b[oop.constrained] = 0
    A = lambda v: oop.vmult(v, copy_constrained=True)
    params = {"type": solver, "rel tolerance": 1e-6}
    if solver == "CG":
        x_ref, its_ref = o.solve_cg(A, omg.vmult, b, rel_tol=1e-6)
    else:
        x_ref, its_ref = o.solve_gmres(A, omg.vmult, b, rel_tol=1e-6)
    xd = op.initialize_dof_vector()
    its, res = pkg.solve(op, xd, op.to_device(b), mg, params)
    assert its == its_ref
    assert 2 <= its <= 12  # mesh-independent multigrid convergence
    assert relerr(op.to_host(xd), x_ref) < 1e-4
    assert np.linalg.norm(b - A(op.to_host(xd))) <= 1e-6 * np.linalg.norm(b) * 1.01


def test_multigrid_kershaw(pkg, ctx):
    """BASELINE configs[2] in small: Kershaw mesh (eps = 0.3), "quadratic geometry", Q3, CG + multigrid with Chebyshev(FDM symm)
    smoothers: converges, same iteration count as the oracle."""
    k, number = 3, "double"
    sizes = [(3, 3, 3), (6, 6, 6)]
    levels, sm, osm = [], [], []
    for s in sizes:
        mesh = pkg.Mesh(ctx, s, periodic=(0, 0, 0), dirichlet=True, map_kind="kershaw", map_params=(0.3, 0.3, 0, 0))
        op = pkg.LaplaceOperatorMatrixFree(mesh, k, number, mapping_type="quadratic geometry")
        oop, oP = oracle_problem(pkg, mesh, op, 1, "symm")
        from parity_util import oracle_mesh
        fdm = pkg.create_fdm_preconditioner(op, {"weighting type": "symm"})
        ch = pkg.PreconditionChebyshev(op, fdm, degree=3)
        mn, mx = ch.estimate_eigenvalues()
        och = o.Chebyshev(oop, oP, degree=3)
        och.set_eigenvalues(mx, mn)
        levels.append(dict(mesh=mesh, op=op, oop=oop, oP=oP, omesh=oracle_mesh(mesh), fdm=fdm))
        sm.append(ch)
        osm.append(och)
    otr = [None, o.TwoLevelTransfer(levels[1]["omesh"], levels[1]["oop"], levels[0]["omesh"], levels[0]["oop"])]
    omg = o.Multigrid([lv["oop"] for lv in levels], osm, otr)
    mg = pkg.PreconditionerGMG([lv["op"] for lv in levels], sm)
    op, oop = levels[1]["op"], levels[1]["oop"]
    rng = np.random.default_rng(9)
    b = rng.uniform(-1, 1, op.n_dofs())
    b[oop.constrained] = 0
    A = lambda v: oop.vmult(v, copy_constrained=True)
    x_ref, its_ref = o.solve_cg(A, omg.vmult, b, rel_tol=1e-8)
    xd = op.initialize_dof_vector()
    its, res = pkg.solve(op, xd, op.to_device(b), mg, {"type": "CG", "rel tolerance": 1e-8})
    assert its == its_ref
    assert relerr(op.to_host(xd), x_ref) < 1e-8


@pytest.mark.parametrize("wt,n_overlap", [("none", 1), ("symm", 1), ("post", 2), ("ras", 1), ("pre", 2)])
@pytest.mark.parametrize("meshkw", [dict(n_cells=(3, 4, 3), periodic=(0, 0, 0)), dict(n_cells=(4, 3, 6), periodic=(1, 0, 1))])
def test_exact_block_asm(pkg, ctx, wt, n_overlap, meshkw):
    """RestrictedPreconditioner::vmult with gauss_jordan(R A R^T) blocks (include/preconditioners.h:528-605, 775-808) against the
    oracle's dense restatement; on this Cartesian mesh the FDM block with the same overlap is the same matrix (fdm_01.cc:148-177)."""
    k = 2
    mesh = pkg.Mesh(ctx, dirichlet=True, **meshkw)
    op = pkg.LaplaceOperatorMatrixFree(mesh, k, "double")
    oop, oP = oracle_problem(pkg, mesh, op, n_overlap, wt)
    from parity_util import oracle_mesh
    fdm = pkg.create_fdm_preconditioner(op, {"n overlap": n_overlap, "weighting type": wt})
    asm = pkg.RestrictedPreconditioner(fdm)
    A = oop.dense()
    lex_rank = np.arange(oP.mesh.C)
    oasm = o.ExactBlockASM(A, oracle_mesh(mesh), k, oop.cell_dofs, oop.n_dofs, oop.constrained, n_overlap, wt, cell_rank=lex_rank)
    # blocks: library cell c (processing order) = oracle cell cell_order[c]
    order = oasm.mesh.cell_order
    for c in (0, len(order) // 2, len(order) - 1):
        assert np.allclose(asm.block_inverse(c), oasm.block_inv[order[c]], rtol=1e-9, atol=1e-11)
    rng = np.random.default_rng(1)
    r = rng.uniform(-1, 1, op.n_dofs())
    d = op.initialize_dof_vector()
    asm.vmult(d, op.to_device(r))
    assert relerr(op.to_host(d), oasm.vmult(r)) < 1e-11
    # FDM == exact ASM on a Cartesian mesh
    d2 = op.initialize_dof_vector()
    fdm.vmult(d2, op.to_device(r))
    assert relerr(op.to_host(d2), op.to_host(d)) < 1e-10
    # as a preconditioner of the device CG
    if wt in ("none", "symm"):
        b = r.copy()
        b[oop.constrained] = 0
        x_ref, its_ref = o.solve_cg(lambda v: oop.vmult(v, copy_constrained=True), oasm.vmult, b)
        xd = op.initialize_dof_vector()
        its, _ = pkg.solve(op, xd, op.to_device(b), asm, {"type": "CG"})
        assert its == its_ref
