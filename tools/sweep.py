#!/usr/bin/env python
"""Degree / precision / variant sweep of BASELINE.json configs[1]: one JSON line per (degree, number type, variant).

variants (labels of matrix_free_loop_08):   vmult | diag (cheby-3-3-diag) | fdm1 (cheby-3-2-symm-1-c) | fdm2 (cheby-3-1-symm-2-g)
                                            | fdmv (cheby-3-1-post-v-c, vertex patches, degree <= 6)
value = DoFs/s (vmult: per operator application; smoothers: per Chebyshev term = n_dofs * 3 * steps / time).
vmult / fdm1 run on ~1e8 DoFs, the generic-kernel variants on ~3e7 DoFs (host set-up of the explicit patch lists).

  python tools/sweep.py out.jsonl [degrees] [numbers]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

CELLS_BIG = {1: 464, 2: 232, 3: 156, 4: 116, 5: 92, 6: 76, 7: 64, 8: 56}
CELLS_SMALL = {1: 312, 2: 156, 3: 104, 4: 76, 5: 64, 6: 52, 7: 44, 8: 40}


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
    degrees = [int(d) for d in sys.argv[2].split(",")] if len(sys.argv) > 2 else list(range(1, 9))
    numbers = sys.argv[3].split(",") if len(sys.argv) > 3 else ["double", "float"]
    pkg = load_package()
    ctx = pkg.Context(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.ExternalStream(ctx.stream_ptr(), device=dev)
    for number in numbers:
        for k in degrees:
            for size, variants in (("big", ["vmult", "diag", "fdm1"]), ("small", ["fdm2", "fdmv"])):
                c = (CELLS_BIG if size == "big" else CELLS_SMALL)[k]
                try:
                    mesh = pkg.Mesh(ctx, (c, c, c), periodic=(1, 1, 1), length=(c / 64.0,) * 3)
                    op = pkg.LaplaceOperatorMatrixFree(mesh, k, number)
                    n = op.n_dofs()
                    g = torch.Generator(device=dev)
                    g.manual_seed(1)
                    x = torch.zeros(op.vec_size(), dtype=op.torch_dtype, device=dev)
                    b = torch.zeros_like(x)
                    x[:n] = torch.rand(n, generator=g, dtype=op.torch_dtype, device=dev) * 2 - 1
                    b[:n] = torch.rand(n, generator=g, dtype=op.torch_dtype, device=dev) * 2 - 1
                    y = torch.zeros_like(x)
                except Exception as e:
                    out.write(json.dumps({"degree": k, "number": number, "size": size, "failed": repr(e)}) + "\n")
                    out.flush()
                    continue
                for v in variants:
                    row = {"degree": k, "number": number, "variant": v, "cells": c, "n_dofs": int(n)}
                    try:
                        t0 = time.perf_counter()
                        if v == "vmult":
                            dt = timed(lambda: op.vmult(y, x), stream, 10)
                            row.update(value=n / dt, unit="DoFs/s per vmult", fast_bricks=int(op.n_fast_bricks()))
                        else:
                            if v == "diag":
                                fdm = None
                            elif v == "fdm1":
                                fdm = pkg.create_fdm_preconditioner(op, {"n overlap": 1, "weighting type": "symm", "weight sequence": "compressed"})
                            elif v == "fdm2":
                                if k < 2:
                                    raise RuntimeError("n overlap is clamped to the degree")
                                fdm = pkg.create_fdm_preconditioner(op, {"n overlap": 2, "weighting type": "symm", "weight sequence": "global"})
                            else:
                                if k > 6:
                                    raise RuntimeError("vertex patches are instantiated up to degree 6")
                                fdm = pkg.create_fdm_preconditioner(op, {"weighting type": "post", "element centric": False})
                            cheb = pkg.PreconditionChebyshev(op, fdm, degree=3, optimize=2 if v == "fdm1" else (3 if v == "diag" else 1))
                            cheb.set_eigenvalues(1.0, 2.4)
                            dt = timed(lambda: cheb.step(x, b), stream, 5)
                            row.update(value=n * 3 / dt, unit="DoFs/s per Chebyshev term")
                            if fdm is not None:
                                dp = timed(lambda: fdm.vmult(y, b), stream, 5)
                                row.update(precon_alone=n / dp, fast_bricks=int(fdm.n_fast_bricks()))
                        row["setup_and_run_s"] = time.perf_counter() - t0
                    except Exception as e:
                        row["failed"] = repr(e)[:200]
                    out.write(json.dumps(row) + "\n")
                    out.flush()
                    x[:n] = torch.rand(n, generator=g, dtype=op.torch_dtype, device=dev) * 2 - 1  # keep the iterate bounded
                del op, mesh, x, b, y
                torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
