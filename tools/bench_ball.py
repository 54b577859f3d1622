#!/usr/bin/env python
"""BASELINE configs[3]: the ball (hyper_ball, degree 5, triquadratic cells, n overlap = 1) through the unstructured-mesh path
(dasm_op_create_unstructured, generic kernels).  One JSON line per measurement: vmult, FDM additive Schwarz per weighting,
Chebyshev(3) + FDM smoother step, and the Krylov solves the reference's experiments/ball.py generates (CG + symm, GMRES + post;
Chebyshev(3)-FDM as the preconditioner instead of the p-multigrid the reference wraps around it).

  python tools/bench_ball.py out.jsonl [n_refinements=4] [degree=5] [numbers=double,float] [mapping_type=]"""
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402


def timed(fn, stream, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def main():
    out = open(sys.argv[1], "a")
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    numbers = sys.argv[4].split(",") if len(sys.argv) > 4 else ["double", "float"]
    mapping_type = sys.argv[5] if len(sys.argv) > 5 else ""
    pkg = load_package()
    grid = importlib.import_module("dealii-asm_b200.grid")
    ctx = pkg.Context(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.ExternalStream(ctx.stream_ptr(), device=dev)
    t0 = time.perf_counter()
    g = grid.hyper_ball(L)
    t_grid = time.perf_counter() - t0
    for number in numbers:
        t0 = time.perf_counter()
        op = pkg.LaplaceOperatorMatrixFree.from_arrays(ctx, g["vertices"], g["cells"], k, g["support"], number=number, mapping_type=mapping_type)
        t_op = time.perf_counter() - t0
        n = op.n_dofs()
        base = {"mesh": "hyper_ball", "n_refinements": L, "degree": k, "number": number, "mapping_type": mapping_type or "merged", "n_cells": int(op.n_cells()),
                "n_dofs": int(n)}
        gen = torch.Generator(device=dev)
        gen.manual_seed(1)
        x = torch.rand(n, generator=gen, dtype=op.torch_dtype, device=dev) * 2 - 1
        b = torch.rand(n, generator=gen, dtype=op.torch_dtype, device=dev) * 2 - 1
        con = torch.tensor(op.constrained_dofs().astype(np.int64), device=dev)
        x[con] = 0
        b[con] = 0
        y = torch.zeros_like(x)
        dt = timed(lambda: op.vmult(y, x), stream, 10)
        out.write(json.dumps(dict(base, variant="vmult", value=n / dt, unit="DoFs/s per vmult", grid_s=t_grid, op_setup_s=t_op)) + "\n")
        out.flush()
        for wt in ("none", "pre", "post", "symm", "ras"):
            t0 = time.perf_counter()
            fdm = pkg.create_fdm_preconditioner(op, {"weighting type": wt, "weight sequence": "dg"})
            t_fdm = time.perf_counter() - t0
            dp = timed(lambda: fdm.vmult(y, b), stream, 10)
            row = dict(base, variant="fdm-%s-1" % wt, value=n / dp, unit="DoFs/s per preconditioner application", fdm_setup_s=t_fdm,
                       n_fdm_instances=int(fdm.n_fdm_instances()))
            if wt in ("symm", "post"):
                cheb = pkg.PreconditionChebyshev(op, fdm, degree=3)
                t0 = time.perf_counter()
                mn, mx = cheb.estimate_eigenvalues()
                row.update(max_ev=mx, ev_s=time.perf_counter() - t0)
                xs = x.clone()
                dc = timed(lambda: cheb.step(xs, b), stream, 5)
                row.update(cheby_step_value=n * 3 / dc, cheby_step_unit="DoFs/s per Chebyshev term", cheby_step_ms=dc * 1e3)
                if number == "double":  # (a relative residual of 1e-8 is below single precision)
                    rhs = op.initialize_dof_vector()
                    op.rhs(rhs, 1.0)
                    sol = op.initialize_dof_vector()
                    solver = "CG" if wt == "symm" else "GMRES"
                    ctx.sync()
                    times = []
                    for _ in range(2):
                        t0 = time.perf_counter()
                        its, res = pkg.solve(op, sol, rhs, cheb, {"type": solver, "rel tolerance": 1e-8})
                        ctx.sync()
                        times.append(time.perf_counter() - t0)
                    row.update(solver=solver, iterations=its, solve_s=min(times), residual=res, max_u=float(op.to_host(sol).max()),
                               max_u_exact=1.0 / 6.0)
                del cheb
            out.write(json.dumps(row) + "\n")
            out.flush()
            del fdm
        del op, x, b, y
        torch.cuda.empty_cache()
    # the reference's ball experiment (experiments/ball.py): CG (symm) / GMRES (post) with hp-multigrid, Chebyshev(3) + FDM smoothers,
    # float levels under a double outer solver; p sequence "bisect" (k -> k/2 -> ... -> 1), then global coarsening to the 32-cell ball
    for solver, wt in (("CG", "symm"), ("GMRES", "post")):
        t0 = time.perf_counter()
        spec = [(l, 1) for l in range(L + 1)]
        kk, ps = k, []
        while kk > 1:
            ps.append(kk)
            kk = max(1, kk // 2)
        spec += [(L, q) for q in reversed(ps)]
        grids = {l: (g if l == L else grid.hyper_ball(l)) for l in range(L + 1)}
        ops, sms, keep = [], [], []
        for (l, q) in spec:
            gl = grids[l]
            lop = pkg.LaplaceOperatorMatrixFree.from_arrays(ctx, gl["vertices"], gl["cells"], q, gl["support"], number="float", mapping_type=mapping_type)
            fdm = pkg.create_fdm_preconditioner(lop, {"weighting type": wt, "weight sequence": "dg"}) if q > 1 else None
            ch = pkg.PreconditionChebyshev(lop, fdm, degree=3 if (l, q) != spec[0] else 8)
            ch.estimate_eigenvalues()
            ops.append(lop)
            sms.append(ch)
            keep.append(fdm)
        trs = [None]
        for i in range(1, len(spec)):
            par = grid.ball_parents(spec[i][0]) if spec[i][0] != spec[i - 1][0] else None
            trs.append(pkg.MGTwoLevelTransfer(ops[i], ops[i - 1], par))
        op = pkg.LaplaceOperatorMatrixFree.from_arrays(ctx, g["vertices"], g["cells"], k, g["support"], number="double", mapping_type=mapping_type)
        mg = pkg.PreconditionerGMG(ops, sms, outer_op=op, transfers=trs)
        rhs, sol = op.initialize_dof_vector(), op.initialize_dof_vector()
        op.rhs(rhs, 1.0)
        ctx.sync()
        t_setup = time.perf_counter() - t0
        pkg.solve(op, sol, rhs, mg, {"type": solver, "rel tolerance": 1e-8})   # warm-up
        ctx.sync()
        times = []
        for _ in range(3):   # (the solver allocates its Krylov vectors per call: the wall time varies, the minimum is reported)
            t0 = time.perf_counter()
            its, res = pkg.solve(op, sol, rhs, mg, {"type": solver, "rel tolerance": 1e-8})
            ctx.sync()
            times.append(time.perf_counter() - t0)
        t_solve = min(times)
        x = op.to_host(sol)
        out.write(json.dumps({"mesh": "hyper_ball", "n_refinements": L, "degree": k, "n_cells": int(op.n_cells()), "n_dofs": int(op.n_dofs()),
                              "variant": "%s + hp-multigrid, Chebyshev(3) + FDM %s smoothers" % (solver, wt), "mapping_type": mapping_type or "merged", "levels": spec, "iterations": its,
                              "time_to_solution_s": t_solve, "time_to_solution_all_s": times, "setup_s": t_setup, "residual": res, "rel_tolerance": 1e-8,
                              "max_u": float(x.max()), "max_u_exact": 1.0 / 6.0}) + "\n")
        out.flush()
        del mg, trs, sms, keep, ops, op
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
