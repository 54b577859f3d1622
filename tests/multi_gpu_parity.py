"""Multi-rank parity (run under torchrun with 2, 4 or 8 ranks, one GPU each):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_parity.py

Every rank builds its brick of the partitioned mesh AND (on its own GPU) the same mesh on one rank; inputs are a
smooth periodic function of the DoF position, so both discretisations hold the same function; the results of
vmult, the FDM preconditioner and a Chebyshev step are compared DoF by DoF through (global cell, local index).
Tolerance 1e-12 relative (double)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from __graft_entry__ import load_package  # noqa: E402
import dasm_oracle as o  # noqa: E402  (only the Gauss-Lobatto points and index expansion helpers)

PART = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def field(pos, L):
    return np.sin(2 * np.pi * pos[..., 0] / L[0]) * np.cos(2 * np.pi * pos[..., 1] / L[1]) + 0.3 * np.sin(4 * np.pi * pos[..., 2] / L[2]) \
        + 0.1 * np.cos(2 * np.pi * (pos[..., 0] / L[0] + pos[..., 2] / L[2]))


def cell_local_values(pkg, mesh, op, k, L, nc):
    """(cell coords [C,3], expanded plain indices [C,n^3], values of `field` at the DoF positions [C,n^3])"""
    n = k + 1
    coords = mesh.cell_coordinates()
    comp = op.compressed_indices(plain=True)
    idx = o.expand_compressed(comp, k, 3).astype(np.int64)
    gll = o.gauss_lobatto_points(n)
    h = [L[d] / nc[d] for d in range(3)]
    zz, yy, xx = np.meshgrid(gll, gll, gll, indexing="ij")
    ref = np.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], axis=-1)  # [n^3, 3], x fastest
    pos = (coords[:, None, :] + ref[None, :, :]) * np.array(h)[None, None, :]
    return coords, idx, field(pos, L)


def main():
    pkg = load_package()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = pkg.Context(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        ident = [pkg.Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        ctx.comm_init(world, rank, ident[0])
    part = PART[world]
    ok = True
    # the (24, 8, 8) mesh has interior bricks on every rank of a 2 x 1 x 1 partition: the halo exchange overlaps with them
    # (16, 16, 16) / (24, 8, 8): lex bricks on every rank, also next to the partition boundary (index-list mode of the TMA-fed
    # kernels); "ras": the patch that keeps an interface DoF lives on the neighbouring rank (compress must run)
    for k, number, nc, wt in ((4, "double", (8, 8, 8), "symm"), (3, "double", (8, 6, 4), "post"), (2, "double", (8, 8, 8), "none"),
                              (4, "double", (24, 8, 8), "symm"), (3, "float", (24, 8, 8), "post"), (4, "double", (16, 16, 16), "symm"),
                              (3, "double", (16, 16, 16), "ras"), (2, "double", (8, 8, 8), "ras")) + \
            (((4, "double", (32, 32, 32), "symm"), (3, "float", (32, 32, 16), "post")) if world >= 4 else ()) + \
            ((3, "double", (8, 8, 8), "symm-2"), (2, "double", (8, 4, 6), "post-2"), (3, "double", (8, 4, 4), "post-3"),
             (2, "double", (8, 8, 4), "post-v"), (3, "float", (8, 6, 4), "symm-v")):
        # "<weighting>-<n overlap>" / "<weighting>-v" (vertex patches): overlapping patches reach into the cells of the neighbour
        # ranks (enlarged ghost layout, include/matrix_free.h:154-213; labels *-2-g-p-n / *-v-c of matrix_free_loop_08)
        fdm_params = {"weighting type": wt.split("-")[0]}
        if wt.endswith("-v"):
            fdm_params["element centric"] = False
        elif "-" in wt:
            fdm_params["n overlap"] = int(wt.split("-")[1])
        L = tuple(float(c) / 4 for c in nc)
        vsize = None
        results = {}
        for tag, prt, rk, c in (("multi", part, rank, ctx), ("single", (1, 1, 1), 0, pkg.Context(local_rank))):
            mesh = pkg.Mesh(c, nc, periodic=(1, 1, 1), length=L, partition=prt, rank=rk)
            op = pkg.LaplaceOperatorMatrixFree(mesh, k, number)
            fdm = pkg.create_fdm_preconditioner(op, fdm_params)
            cheb = pkg.PreconditionChebyshev(op, fdm, degree=3, optimize=2 if len(fdm_params) == 1 else 1)
            cheb.set_eigenvalues(1.0, 2.4)
            coords, idx, vals = cell_local_values(pkg, mesh, op, k, L, nc)
            nvec = op.vec_size()
            xh = np.zeros(nvec)
            xh[idx.reshape(-1)] = vals.reshape(-1)
            bh = np.zeros(nvec)
            bh[idx.reshape(-1)] = (vals ** 2).reshape(-1) - 0.4
            x = torch.zeros(nvec, dtype=op.torch_dtype, device=dev)
            b = torch.zeros(nvec, dtype=op.torch_dtype, device=dev)
            nown = op.n_dofs()
            x[:nown] = torch.as_tensor(xh[:nown]).to(dev)
            b[:nown] = torch.as_tensor(bh[:nown]).to(dev)
            torch.cuda.synchronize()
            y = torch.zeros_like(x)
            z = torch.zeros_like(x)
            op.vmult(y, x)
            fdm.vmult(z, b)
            cheb.step(x, b)
            c.sync()
            out = {}
            for name, v in (("vmult", y), ("fdm", z), ("cheb", x)):
                vh = v.double().cpu().numpy()
                # values of owned DoFs per (cell, local index); ghost entries are not defined after the operation
                loc = np.where(idx < nown, vh[np.minimum(idx, nvec - 1)], np.nan)
                out[name] = {tuple(cc): loc[i] for i, cc in enumerate(coords)}
            results[tag] = out
        for name in ("vmult", "fdm", "cheb"):
            num = den = 0.0
            for cc, v in results["multi"][name].items():
                r = results["single"][name][cc]
                m = ~np.isnan(v) & ~np.isnan(r)
                num += np.sum((v[m] - r[m]) ** 2)
                den += np.sum(r[m] ** 2)
            err = np.sqrt(num / max(den, 1e-300))
            print("rank %d/%d k=%d %s %-6s rel.err multi vs single = %.3e" % (rank, world, k, wt, name, err), flush=True)
            ok &= bool(err < (1e-12 if number == "double" else 2e-5))
    if world > 1:
        t = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(t.item() > 0.5)
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_PARITY", "PASS" if ok else "FAIL", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
