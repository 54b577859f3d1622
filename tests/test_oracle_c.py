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
