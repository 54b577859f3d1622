"""CPU-only: the C restatement (CPU baseline of bench.py) agrees with the numpy oracle."""
import numpy as np
import pytest

import dasm_oracle as o


@pytest.mark.parametrize("k,wt,mapfun", [(3, "symm", o.sine_map), (4, "post", None), (2, "none", None)])
def test_c_port_matches_numpy_oracle(k, wt, mapfun):
    from __graft_entry__ import build_oracle
    build_oracle()
    import oracle_c
    nc = (4, 4, 4)
    mesh = o.StructuredMesh(3, nc, (True, True, True), mapfun=mapfun)
    mesh.cell_order = o.brick_major_order(nc)
    cd, nd, con, comp = o.number_dofs_owner_cell(mesh, k)
    b = o.Basis1D(k)
    G = o.merged_coefficients(mesh.jacobians(b), b, 3)
    oop = o.LaplaceOperator(3, k, cd, nd, con, G)
    oP = o.FDMPreconditioner(mesh, k, cd, nd, con, 1, wt)
    sm = oracle_c.CSmoother(mesh, oop, oP, 3, 2.4, 1.0)
    rng = np.random.default_rng(0)
    x, bb = rng.uniform(-1, 1, nd), rng.uniform(-1, 1, nd)
    ch = o.Chebyshev(oop, oP, degree=3)
    ch.set_eigenvalues(2.4, 1.0)
    r1, r2 = sm.step(x, bb), ch.step(x, bb)
    assert np.linalg.norm(r1 - r2) / np.linalg.norm(r2) < 1e-13


@pytest.mark.parametrize("k,nc,wt", [(4, (8, 4, 6), "symm"), (2, (9, 4, 4), "post"), (3, (4, 5, 7), "none"), (5, (8, 4, 4), "symm")])
def test_cartesian_baseline_step(k, nc, wt):
    """oracle/cpu_baseline.c (SIMD across cells, Kronecker form, fused pre / post operations, slab threading) against the numpy
    oracle on the same periodic Cartesian mesh in the lexicographic numbering."""
    import oracle_c
    L = (1.0, 0.75, 1.5)
    mesh = o.StructuredMesh(3, nc, (True, True, True), lengths=L)
    cd, nd = oracle_c.lexicographic_cell_dofs(nc, k)
    con = np.zeros(nd, dtype=bool)
    b = o.Basis1D(k)
    G = o.merged_coefficients(mesh.jacobians(b), b, 3)
    oop = o.LaplaceOperator(3, k, cd, nd, con, G)
    oP = o.FDMPreconditioner(mesh, k, cd, nd, con, 1, wt)
    och = o.Chebyshev(oop, oP, degree=3)
    och.set_eigenvalues(2.4, 1.0)
    rng = np.random.default_rng(5)
    x0, rhs = rng.uniform(-1, 1, nd), rng.uniform(-1, 1, nd)
    ref = och.step(x0, rhs)
    base = oracle_c.CartesianBaseline(nc, L, k, 3, wt, max_ev=2.4, min_ev=1.0)
    x = base.step(x0.copy(), rhs)
    assert np.linalg.norm(x - ref) / np.linalg.norm(ref) < 1e-12
