#!/usr/bin/env python
"""BASELINE.json configs[2]: Kershaw-deformed mesh (eps_y = eps_z = 0.3), degree 4, "quadratic geometry" matrix-free operator,
multigrid-preconditioned Krylov solver with Chebyshev(3) + FDM(n_overlap = 1) smoothers on every level, float levels, double outer
solver (element_centered_preconditioners_01.cc:352-413, 540-740, 787-792; experiments/kershaw.sh).  The reference's instance L = 5
is subdivided_hyper_cube(6) + 3 refinements = 48^3 = 110,592 cells, 7,189,057 DoFs; its coarse solver is Trilinos AMG, which does
not exist here: the coarse level (6^3 cells) is solved by a Chebyshev(FDM) sweep of higher degree instead (stated in the output).

  python tools/solve_kershaw.py out.json [n_refinements=3] [degree=4]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402


def main():
    out = sys.argv[1]
    n_ref = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    pkg = load_package()
    ctx = pkg.Context(0)
    dev = torch.device("cuda", 0)
    results = []
    for solver, wt, cheb_degree in (("CG", "symm", 3), ("GMRES", "post", 3), ("CG", "symm", 2), ("CG", "symm", 5)):
        t0 = time.perf_counter()
        levels, smoothers, ev, keepalive = [], [], [], []
        for r in range(n_ref + 1):
            c = 6 * 2 ** r
            mesh = pkg.Mesh(ctx, (c, c, c), periodic=(0, 0, 0), dirichlet=True, map_kind="kershaw", map_params=(0.3, 0.3, 0, 0))
            op = pkg.LaplaceOperatorMatrixFree(mesh, k, "float", mapping_type="quadratic geometry")
            fdm = pkg.create_fdm_preconditioner(op, {"n overlap": 1, "weighting type": wt})
            ch = pkg.PreconditionChebyshev(op, fdm, degree=(8 if r == 0 else cheb_degree), optimize=2)
            ev.append(ch.estimate_eigenvalues())
            levels.append(op)
            smoothers.append(ch)
            keepalive.append((mesh, fdm))  # (not stored on the operator: operator -> fdm -> operator would be a reference cycle, whose
            # members the garbage collector finalises in arbitrary order - the FDM object must be destroyed before its operator)
        mesh = levels[-1].mesh
        op = pkg.LaplaceOperatorMatrixFree(mesh, k, "double", mapping_type="quadratic geometry")
        mg = pkg.PreconditionerGMG(levels, smoothers, outer_op=op)
        n = op.n_dofs()
        # right-hand side: f = 1 tested with the shape functions ~ row sums of the mass matrix; a smooth positive vector serves the
        # same purpose here (A x = b with x = 0 start): b = A * (random smooth field) is avoided to keep the problem generic
        g = torch.Generator(device=dev)
        g.manual_seed(0)
        b = torch.zeros(op.vec_size(), dtype=torch.float64, device=dev)
        b[:n] = 1.0
        con = torch.from_numpy(op.constrained_dofs().astype(np.int64)).to(dev)
        b[con] = 0
        x = op.initialize_dof_vector()
        ctx.sync()
        t_setup = time.perf_counter() - t0
        params = {"type": solver, "rel tolerance": 1e-8, "abs tolerance": 1e-20, "max iterations": 300}  # experiments: reduction 1e-8
        try:
            its, res = pkg.solve(op, x, b, mg, params)  # warm-up (the reference also solves twice, :221-236)
        except pkg.DasmError as e:
            results.append(dict(solver=solver, weighting=wt, smoother="Chebyshev(%d) + FDM n=1" % cheb_degree, failed=str(e)))
            json.dump(results, open(out, "w"), indent=1)
            continue
        ctx.sync()
        t1 = time.perf_counter()
        its, res = pkg.solve(op, x, b, mg, params)
        ctx.sync()
        dt = time.perf_counter() - t1
        y = op.initialize_dof_vector()
        op.vmult(y, x)
        x_con = float(torch.linalg.norm(x[con]))  # constrained entries of the solution (zero: homogeneous Dirichlet values)
        y[con] = x[con]
        true_res = float(torch.linalg.norm(b[:n] - y[:n])) / float(torch.linalg.norm(b[:n]))
        row = dict(config="kershaw eps=0.3", degree=k, cells=int(mesh.n_cells), n_dofs=int(n), solver=solver, weighting=wt,
                   smoother="Chebyshev(%d) + FDM n=1" % cheb_degree, coarse="Chebyshev(8) + FDM n=1 on 6^3 cells (no AMG here)",
                   level_number="float", outer_number="double", n_levels=n_ref + 1, iterations=its, time_to_solution_s=dt,
                   setup_s=t_setup, true_relative_residual=true_res, reported_residual=res, norm_x_constrained=x_con, max_ev=[e[1] for e in ev],
                   dofs_per_s_per_iteration=n * its / dt)
        print(row, flush=True)
        results.append(row)
        json.dump(results, open(out, "w"), indent=1)
        del mg, smoothers, levels, op
        torch.cuda.empty_cache()
    json.dump(results, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
