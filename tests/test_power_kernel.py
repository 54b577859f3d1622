"""Power kernel (power_kernel_01.likwid.cc): dst_0 = A src, dst_1 = M dst_0, fused per wave of cells or in two sweeps; the device
result against the oracle, the wave schedule against the restatement of determine_pre_post."""
import numpy as np
import pytest

import dasm_oracle as o
from __graft_entry__ import load_package
from parity_util import oracle_mesh, oracle_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    return load_package()


@pytest.fixture(scope="module")
def ctx(pkg):
    return pkg.Context(0)


def relerr(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("k,number", [(2, "double"), (3, "double"), (4, "double"), (4, "float"), (6, "double")])
@pytest.mark.parametrize("granularity,batch", [(0, 1), (16, 1), (24, 4), (7, 1)])
def test_power_kernel(pkg, ctx, k, number, granularity, batch):
    # hyper-cube without constraints (power_kernel_01.likwid.cc:322-333: empty AffineConstraints), Cartesian cells
    n_cells = (4, 4, 5) if k <= 4 else (3, 2, 3)
    mesh = pkg.Mesh(ctx, n_cells, periodic=(0, 0, 0), dirichlet=False, length=(1.0, 1.0, 1.25))
    op = pkg.LaplaceOperatorMatrixFree(mesh, k, number)
    dt = np.float64 if number == "double" else np.float32
    oop, _ = oracle_problem(pkg, mesh, op, with_fdm=False, dtype=dt)
    omesh = oracle_mesh(mesh)
    b = o.Basis1D(k)
    w = np.einsum("c,b,a->cba", b.qw, b.qw, b.qw).reshape(-1)
    JxW = np.linalg.det(omesh.jacobians(b)) * w[None, :]
    rng = np.random.default_rng(k)
    x = rng.uniform(-1, 1, op.n_dofs())
    d0 = oop.vmult(x)
    d1 = o.mass_vmult(oop, d0, JxW)
    tol = 1e-12 if number == "double" else 2e-5
    pk = pkg.PowerKernel(op, granularity, batch)
    xd = op.to_device(x)
    for fused in (True, False):
        a0, a1 = op.initialize_dof_vector(), op.initialize_dof_vector()
        pk.run(a0, a1, xd, fused=fused)
        assert relerr(op.to_host(a0), d0.astype(np.float64)) < tol
        assert relerr(op.to_host(a1), d1.astype(np.float64)) < tol
        # results are added (distribute_local_to_global): a second run doubles dst_0
        pk.run(a0, a1, xd, fused=fused)
        assert relerr(op.to_host(a0), 2 * d0.astype(np.float64)) < tol
    # do_computation = false: gather / scatter only: dst_0 = valence * src
    a0, a1 = op.initialize_dof_vector(), op.initialize_dof_vector()
    pk.run(a0, a1, xd, fused=True, do_computation=False)
    val = np.zeros(op.n_dofs())
    np.add.at(val, oop.idx.reshape(-1), 1.0)
    assert relerr(op.to_host(a0), val * x) < tol
    assert relerr(op.to_host(a1), val * val * x) < tol
    # wave schedule = determine_pre_post on the cells in processing order (vertices = the corner DoFs)
    n = k + 1
    corners = [i + n * (j + n * l) for l in (0, k) for j in (0, k) for i in (0, k)]
    cv = oop.cell_dofs[omesh.cell_order][:, corners]
    _, _, post, ptr = o.determine_pre_post(cv, granularity, batch, track_individual_cell=(batch == 1))
    if batch == 1:
        assert [int(c) for c in np.diff(ptr)] == pk.post_counts()
    else:
        # batches: the reference counts batches, the library counts their cells
        nc = op.n_cells()
        cells_per_wave = [int(sum(min(batch, nc - e * batch) for e in post[ptr[w]:ptr[w + 1]])) for w in range(len(ptr) - 1)]
        assert cells_per_wave == pk.post_counts()
