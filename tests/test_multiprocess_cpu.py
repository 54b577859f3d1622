"""World-size-2 CPU test (gloo) of the multi-rank host logic: brick partition, owner-cell numbering with ghosts and the
ghost-exchange lists that both sides compute independently from the structured topology (the counterpart of
Utilities::MPI::Partitioner / VectorDataExchange, include/matrix_free_internal.h:3-83).  No device is involved: the
lists come from dasm_mesh_host_numbering and the exchange itself runs over torch.distributed / gloo.

  update_ghost_values: every DoF a rank sees through its cells (owned or ghost) holds the value of a function of the
                       global DoF position after the exchange
  compress(add):       ghost contributions sent back and added at the owner; the sum over all DoF copies is conserved
                       and every shared DoF receives exactly one contribution per ghosting rank
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _field(pos, L):
    return np.sin(2 * np.pi * pos[..., 0] / L[0]) + 0.5 * np.cos(2 * np.pi * pos[..., 1] / L[1]) * (1 + pos[..., 2] / L[2])


def _worker(rank, world, port, nc, periodic, k, result, part=(2, 1, 1), halo=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dasm_oracle as o  # only index expansion and Gauss-Lobatto points
    from __graft_entry__ import load_package
    pkg = load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = tuple(float(c) for c in nc)
        mesh = pkg.Mesh(None, nc, periodic=periodic, dirichlet=False, length=L, partition=part, rank=rank)
        nb = mesh.host_halo_numbering(k) if halo else mesh.host_numbering(k)
        n_owned, n_ghost = nb["n_owned"], nb["n_ghost"]
        n = k + 1
        idx = o.expand_compressed(nb["cidx_plain"], k, 3).astype(np.int64)  # [cells, n^3]
        coords = nb["coords"] if halo else mesh.cell_coordinates()
        gll = o.gauss_lobatto_points(n)
        zz, yy, xx = np.meshgrid(gll, gll, gll, indexing="ij")
        ref = np.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], axis=-1)
        pos = coords[:, None, :] + ref[None, :, :]  # unit cells
        if any(periodic):
            pos = np.where(np.array(periodic, dtype=bool)[None, None, :], np.mod(pos, np.array(nc)[None, None, :]), pos)
        vals = _field(pos, L)
        assert idx.max() < n_owned + n_ghost
        # ---- update_ghost_values
        x = np.full(n_owned + n_ghost, np.nan)
        own = idx < n_owned
        x[idx[own]] = vals[own]
        assert not np.isnan(x[:n_owned]).any()  # every owned DoF is touched by a local cell
        assert idx.max() == n_owned + n_ghost - 1 or n_ghost == 0  # (the cells reach the last ghost)
        so = ro = 0
        for p, ns, nr in zip(nb["peers"], nb["send_count"], nb["recv_count"]):
            send = torch.from_numpy(np.ascontiguousarray(x[nb["send_idx"][so:so + ns]]))
            recv = torch.empty(int(nr), dtype=torch.float64)
            reqs = ([dist.isend(send, int(p))] if ns > 0 else []) + ([dist.irecv(recv, int(p))] if nr > 0 else [])
            for r in reqs:
                r.wait()
            x[nb["recv_idx"][ro:ro + nr]] = recv.numpy()
            so += ns
            ro += nr
        err = np.nanmax(np.abs(x[idx] - vals))
        assert not np.isnan(x[idx]).any()
        # ---- compress(add): ghosts hold 1, owned 0; after the exchange an owned DoF holds its number of ghost copies
        y = np.zeros(n_owned + n_ghost)
        y[n_owned:] = 1.0
        so = ro = 0
        for p, ns, nr in zip(nb["peers"], nb["send_count"], nb["recv_count"]):
            send = torch.from_numpy(np.ascontiguousarray(y[nb["recv_idx"][ro:ro + nr]]))  # ghost values go back
            recv = torch.empty(int(ns), dtype=torch.float64)
            reqs = ([dist.isend(send, int(p))] if nr > 0 else []) + ([dist.irecv(recv, int(p))] if ns > 0 else [])
            for r in reqs:
                r.wait()
            np.add.at(y, nb["send_idx"][so:so + ns], recv.numpy())
            so += ns
            ro += nr
        t = torch.tensor([float(y[:n_owned].sum()), float(n_ghost), float(n_owned)], dtype=torch.float64)
        dist.all_reduce(t)
        result[rank] = (float(err), float(t[0]), float(t[1]), float(t[2]), float(y[:n_owned].max()) if n_owned else 0.0)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nc,periodic,k,part", [((8, 4, 4), (1, 1, 1), 3, (2, 1, 1)), ((6, 3, 2), (0, 0, 0), 2, (2, 1, 1)),
                                                ((8, 4, 5), (1, 0, 1), 4, (2, 1, 1)), ((8, 8, 4), (1, 1, 1), 2, (2, 2, 1)),
                                                ((4, 6, 4), (0, 1, 0), 3, (1, 2, 2))])
def test_ghost_exchange_lists(nc, periodic, k, part):
    world = part[0] * part[1] * part[2]
    port = 29600 + (os.getpid() + 7 * k) % 300
    mgr = mp.get_context("spawn").Manager()
    result = mgr.dict()
    mp.spawn(_worker, args=(world, port, nc, periodic, k, result, part), nprocs=world, join=True)
    assert len(result) == world
    n_dofs_expected = 1
    for d in range(3):
        n_dofs_expected *= nc[d] * k + (0 if periodic[d] else 1)
    for rank in range(world):
        err, added, n_ghost_total, n_owned_total, mx = result[rank]
        assert err < 1e-14                       # ghost values equal the owner's values
        assert added == n_ghost_total            # every ghost copy arrives exactly once at an owner
        assert n_owned_total == n_dofs_expected  # the owned ranges partition the global DoFs
        assert mx <= world - 1                   # at most one ghost copy per other rank


@pytest.mark.parametrize("nc,periodic,k,part", [((8, 4, 4), (1, 1, 1), 3, (2, 1, 1)), ((6, 3, 2), (0, 0, 0), 2, (2, 1, 1)),
                                                ((8, 8, 4), (1, 0, 1), 2, (2, 2, 1)), ((4, 6, 4), (0, 1, 0), 3, (1, 2, 2))])
def test_enlarged_ghost_layout(nc, periodic, k, part):
    """the partitioner of a preconditioner with overlapping patches (include/matrix_free.h:154-213): all DoFs of the cells around a
    rank's cells are ghosts; values seen through local AND halo cells are the owner's after the exchange, compress conserves."""
    world = part[0] * part[1] * part[2]
    port = 29300 + (os.getpid() + 11 * k) % 250
    mgr = mp.get_context("spawn").Manager()
    result = mgr.dict()
    mp.spawn(_worker, args=(world, port, nc, periodic, k, result, part, True), nprocs=world, join=True)
    assert len(result) == world
    n_dofs_expected = 1
    for d in range(3):
        n_dofs_expected *= nc[d] * k + (0 if periodic[d] else 1)
    for rank in range(world):
        err, added, n_ghost_total, n_owned_total, mx = result[rank]
        assert err < 1e-14
        assert added == n_ghost_total
        assert n_owned_total == n_dofs_expected
        assert mx <= world - 1
