#!/usr/bin/env python
"""Benchmark of the smoother hot path (driver contract: one JSON line on stdout from rank 0).

A "step" is one Chebyshev(3) smoother step around the element-centred FDM additive-Schwarz preconditioner
(label `cheby-3-2-symm-1-c` of the reference's matrix_free_loop_08) on a periodic Cartesian hyper-rectangle,
FE_Q degree 4, double precision, ~1e8 DoFs per GPU (192x128x64 cells per GPU = `n subdivisions` 41).
Throughput follows the reference's definition (matrix_free_loop_08.likwid.cc:390-395):
    DoFs/s = n_dofs * chebyshev_degree * steps / time        ("DoFs/s per Chebyshev term").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

# exactly one JSON line on stdout: native libraries (NCCL's version banner) write to file descriptor 1 directly, so the
# real stdout is kept aside for the JSON line and descriptor 1 points to stderr for everything else
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PARTITION = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
METRIC = "DoFs/s per Chebyshev term of the Chebyshev(3)+FDM-ASM smoother step (n_dofs*degree*steps/time)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--degree", type=int, default=4)
    ap.add_argument("--number", default="double", choices=["double", "float"])
    ap.add_argument("--cheb-degree", type=int, default=3)
    ap.add_argument("--weighting", default="symm")
    ap.add_argument("--cells", default="192,128,64", help="cells per GPU (x,y,z)")
    ap.add_argument("--cpu-cells", default="", help="mesh of the cpu_baseline leg (default: --cells, the GPU workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--extra", action="store_true", help="also time vmult / FDM alone (printed to stderr)")
    ap.add_argument("--map", default="cartesian", choices=["cartesian", "kershaw", "sine"],
                    help="secondary workloads: kershaw = BASELINE configs[2] (eps 0.3, Dirichlet), not the headline")
    ap.add_argument("--mapping-type", default="", help='operator mapping type ("", "merged", "quadratic geometry", "linear geometry")')
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: an in-process NVML polling thread (2 ms period; the timed
    region is tens of milliseconds, too short for `nvidia-smi -lms`), falling back to one nvidia-smi query."""

    def __init__(self, index=0):
        import threading
        self.index = index
        self.samples = []      # (sm_mhz, reasons bitmask)
        self.max_mhz = None
        self.stop_flag = False
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((float(mhz), int(rs)))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            nv = self.nv
            names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            if self.samples:
                sm = [s[0] for s in self.samples]
                out["sm_mhz"] = float(np.median(sm))
                out["samples"] = len(sm)
                mask = 0
                for s in self.samples:
                    mask |= s[1]
                out["reasons"] = sorted(n for n, bit in names.items() if mask & bit)
            return out
        try:
            r = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=10).stdout.strip().split(", ")
            out["sm_mhz"], out["sm_max_mhz"] = float(r[0]), float(r[1])
            out["note"] = "single nvidia-smi query after the timed region (NVML python binding unavailable)"
        except Exception:
            pass
        return out


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the optimised C restatement of the reference algorithm (oracle/cpu_baseline.c: compile-time degree, SIMD across
# 8 cells, Cartesian Kronecker form, fused pre / post operations, one cell range per thread), all host cores
# ---------------------------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_smoother(args, cells):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_c
    oracle_c.baseline_set_threads(host_cores())  # (torch.distributed.run exports OMP_NUM_THREADS=1)
    nc = tuple(int(c) for c in cells.split(","))
    sm = oracle_c.CartesianBaseline(nc, tuple(c / 64.0 for c in nc), args.degree, args.cheb_degree, args.weighting, max_ev=2.4, min_ev=1.0)
    return sm, sm.n_dofs, oracle_c.baseline_max_threads(), nc


def time_cpu(args, steps, warmup, cells, target_s=None):
    sm, nd, threads, nc = cpu_smoother(args, cells)
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, nd)
    b = rng.uniform(-1, 1, nd)
    t_one = 1.0
    for _ in range(max(warmup, 1)):
        t0 = time.perf_counter()
        sm.step(x, b)
        t_one = time.perf_counter() - t0
    if target_s is not None:
        # bounded sample of about target_s seconds of CPU work
        steps = int(max(1, min(400, target_s / max(t_one, 1e-6))))
    t0 = time.perf_counter()
    for _ in range(steps):
        sm.step(x, b)
    dt = time.perf_counter() - t0
    value = nd * args.cheb_degree * steps / dt
    sample = "Chebyshev(%d)+FDM(%s,n=1) step, degree %d, double, %dx%dx%d periodic Cartesian cells (%d DoFs), %d steps, %.2f s" % (
        args.cheb_degree, args.weighting, args.degree, nc[0], nc[1], nc[2], nd, steps, dt)
    return value, threads, sample, dt / steps * 1e3, steps


def run_reference(args):
    """the reference's CPU path (restated) on the host cores, on the GPU arm's workload: the same mesh as
    one GPU of the GPU arm processes (--cells), every step a full smoother step; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    value, threads, sample, ms, steps = time_cpu(args, steps, max(1, min(args.warmup, 2)), args.cells)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "DoFs/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "matrix_free_loop_08 label cheby-%d-2-%s-1-c: FE_Q(%d), periodic Cartesian hyper-rectangle, %s cells "
                               "(the mesh of ONE GPU of the GPU arm); CPU restatement of the reference algorithm "
                               "(oracle/cpu_baseline.c; deal.II itself cannot be built here)" % (
                                   args.cheb_degree, args.weighting, args.degree, args.cells.replace(",", "x")),
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "DoFs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "DoFs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    pkg = load_package()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("WORLD_SIZE %d != --gpus %d" % (world, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = pkg.Context(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        ident = [pkg.Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        ctx.comm_init(world, rank, ident[0])
    if world not in PARTITION:
        raise SystemExit("--gpus must be 1, 2, 4 or 8")
    part = PARTITION[world]
    cpg = tuple(int(c) for c in args.cells.split(","))
    nc = tuple(cpg[d] * part[d] for d in range(3))
    length = tuple(c / 64.0 for c in nc)  # unit cells of the n_refine = 6 hyper-rectangle

    def barrier():
        torch.cuda.synchronize()
        ctx.sync()
        if world > 1:
            dist.barrier()

    k = args.degree
    S = 8 if args.number == "double" else 4
    t_setup = time.perf_counter()
    if args.map == "cartesian":
        mesh = pkg.Mesh(ctx, nc, periodic=(1, 1, 1), length=length, partition=part, rank=rank)
    else:
        mp_ = (0.3, 0.3, 0., 0.) if args.map == "kershaw" else (0., 0., 0., 0.)
        mesh = pkg.Mesh(ctx, nc, periodic=(0, 0, 0) if args.map == "kershaw" else (1, 1, 1), dirichlet=True, length=(1., 1., 1.),
                        map_kind=args.map, map_params=mp_, partition=part, rank=rank)
    op = pkg.LaplaceOperatorMatrixFree(mesh, k, args.number, mapping_type=args.mapping_type)
    fdm = pkg.create_fdm_preconditioner(op, {"n overlap": 1, "weighting type": args.weighting, "weight sequence": "compressed"})
    cheb = pkg.PreconditionChebyshev(op, fdm, degree=args.cheb_degree, optimize=2)
    cheb.set_eigenvalues(1.0, 2.4)  # fixed so that every rank count does identical arithmetic
    ctx.sync()
    t_setup = time.perf_counter() - t_setup
    n_own = op.n_dofs()
    n_glob = op.m()
    n_cells_local = mesh.n_cells

    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    x = torch.zeros(op.vec_size(), dtype=op.torch_dtype, device=dev)
    b = torch.zeros(op.vec_size(), dtype=op.torch_dtype, device=dev)
    x[:n_own] = torch.rand(n_own, generator=g, dtype=op.torch_dtype, device=dev) * 2 - 1
    b[:n_own] = torch.rand(n_own, generator=g, dtype=op.torch_dtype, device=dev) * 2 - 1
    torch.cuda.synchronize()

    stream = torch.cuda.ExternalStream(ctx.stream_ptr(), device=dev)
    for _ in range(max(args.warmup, 3)):
        cheb.step(x, b)
    barrier()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    ctx.enable_kernel_timing(True)
    l0 = ctx.launch_count()
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        cheb.step(x, b)
    e1.record(stream)
    barrier()
    launches = ctx.launch_count() - l0
    ms = e0.elapsed_time(e1)
    ktimes = [ctx.kernel_time(c) for c in range(4)]
    ctx.enable_kernel_timing(False)
    if sampler is not None:
        n_timed = len(sampler.samples)
        if n_timed < 8 and world == 1:  # (several ranks: every step is collective, no rank-local extension)
            # the timed region lasted only a few sampling periods: keep the same load running (untimed) for more samples
            for _ in range(max(args.steps, 10)):
                cheb.step(x, b)
            barrier()
    clocks = sampler.stop() if sampler else None
    if clocks is not None and sampler is not None:
        clocks["samples_in_timed_region"] = n_timed
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = n_glob * args.cheb_degree * args.steps / (ms * 1e-3)

    # ---- end to end through the host-buffer entry point (pinned host memory, H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        # dasm_cheb_step_host_batch: independent problems streamed through the smoother; the host -> device copies of problem i + 1 and
        # the device -> host copy of problem i - 1 overlap the kernels of problem i.  Every step copies its x and b from pinned host
        # memory and its result back (two host buffer pairs used in turn); the one-call-per-step entry point is timed next to it.
        ke = max(2, min(args.steps, 10))
        hx = [torch.empty(n_own, dtype=torch.float64).pin_memory() for _ in range(2)]
        hb = [torch.empty(n_own, dtype=torch.float64).pin_memory() for _ in range(2)]
        for t_ in hx:
            t_.copy_(x[:n_own].double().cpu())
        for t_ in hb:
            t_.copy_(b[:n_own].double().cpu())
        xn, bn = [t_.numpy() for t_ in hx], [t_.numpy() for t_ in hb]
        cheb.step_host_batch([xn[i % 2] for i in range(3)], [bn[i % 2] for i in range(3)])  # warm-up (allocates the pipeline buffers)
        barrier()
        t0 = time.perf_counter()
        cheb.step_host_batch([xn[i % 2] for i in range(ke)], [bn[i % 2] for i in range(ke)])
        barrier()
        dt = time.perf_counter() - t0
        ks = 3
        cheb.step_host(xn[0], bn[0])
        barrier()
        t0 = time.perf_counter()
        for _ in range(ks):
            cheb.step_host(xn[0], bn[0])
        barrier()
        dts = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt, dts], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt, dts = float(t[0].item()), float(t[1].item())
        e2e = {"value": n_glob * args.cheb_degree * ke / dt, "unit": "DoFs/s", "h2d_bytes_per_step": int(2 * n_own * 8 * world),
               "d2h_bytes_per_step": int(n_own * 8 * world), "steps": ke, "ms_per_step": dt / ke * 1e3,
               "sequential_value": n_glob * args.cheb_degree * ks / dts, "sequential_ms_per_step": dts / ks * 1e3,
               "note": "dasm_cheb_step_host_batch: per step x and b copied from pinned host memory and the result copied back; copies of "
                       "neighbouring steps overlap the kernels (two device buffer sets, one copy stream per direction); sequential_value: "
                       "dasm_cheb_step_host, one blocking call per step"}

    extra = {}
    if args.extra:
        y = torch.zeros_like(x)
        for name, fn in (("vmult", lambda: op.vmult(y, x)), ("fdm", lambda: fdm.vmult(y, x))):
            for _ in range(3):
                fn()
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            for _ in range(10):
                fn()
            a1.record(stream)
            barrier()
            extra[name + "_dofs_per_s"] = n_glob * 10 / (a0.elapsed_time(a1) * 1e-3)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (by accumulated device time inside the timed region)
    peak, peak_src = measured_peak()
    names = ["laplace cell kernel (A-sweep)", "FDM cell kernel (P-sweep)", "vector epilogues", "ghost exchange"]
    dom = int(np.argmax([t[0] for t in ktimes[:2]]))
    dom_ms, dom_n = ktimes[dom]
    idx_bytes = 27 * 4
    d_c = k ** 3
    fused = getattr(pkg, "FUSED_KERNELS", False) and os.environ.get("DASM_FORCE_GENERIC", "0") != "1" and k <= 5
    if fused:
        # A-sweep: read x, read b, write t1 (3 S) + indices; P-sweep: read t1, x, x_old, write x+ (4 S) + indices + 27 weights + 3 ids
        per_dof = [3 * S + idx_bytes / d_c, 4 * S + (idx_bytes + 27 * S + 12) / d_c][dom]
    else:
        # unfused cell kernels: read src, write dst (2 S) + per-cell metadata (SURVEY.md 8(d) `vmult` / `FDM-ASM vmult` rows)
        per_dof = [2 * S + idx_bytes / d_c, 2 * S + (idx_bytes + 27 * S + 12) / d_c][dom]
    cd = args.cheb_degree
    step_bytes_per_term = ((6 + 7 * (cd - 1)) * S + cd * (2 * idx_bytes + 27 * S + 12) / d_c) / cd
    bytes_per_launch = per_dof * n_own
    achieved = bytes_per_launch / (dom_ms / max(dom_n, 1) * 1e-3) / 1e9 if dom_n else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(["laplace", "fdm"][dom])
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_dof": per_dof, "avg_launch_ms": dom_ms / max(dom_n, 1), "launches_timed": dom_n,
                "share_of_step": dom_ms / ms,
                "kernel_time_ms": {names[i]: ktimes[i][0] for i in range(4)},
                # whole step against the model of SURVEY.md 8(d): the first term of `step` has no x_old: (6 + 7 (d - 1)) S per DoF,
                # plus the per-cell metadata of every term
                "step_algorithmic_bytes_per_dof_per_term": step_bytes_per_term,
                "step_frac_of_hbm_roofline": value * step_bytes_per_term / world / (peak * 1e9)}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            v, threads, sample, _, _ = time_cpu(args, 3, 1, args.cpu_cells or args.cells, target_s=12.0)
            cpu = {"value": v, "unit": "DoFs/s", "cores": threads, "kind": "port", "sample": sample}
        except Exception as e:  # the baseline must never take the GPU number down
            cpu = {"value": None, "unit": "DoFs/s", "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}

    line = {
        "metric": METRIC, "value": value, "unit": "DoFs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64" if args.number == "double" else "f32", "data": "synthetic",
        "config": {"workload": "matrix_free_loop_08 label cheby-%d-2-%s-1-c: FE_Q(%d), %s, "
                               "%dx%dx%d cells per GPU (n subdivisions 41), brick partition %dx%dx%d" % (
                                   args.cheb_degree, args.weighting, k,
                                   "periodic Cartesian hyper-rectangle" if args.map == "cartesian" else
                                   "%s-deformed mesh, mapping type '%s' (secondary workload; the roofline bytes omit the geometry data)" % (args.map, args.mapping_type),
                                   cpg[0], cpg[1], cpg[2], part[0], part[1], part[2]),
                   "n_dofs": int(n_glob), "n_dofs_per_gpu": int(n_own), "n_cells_per_gpu": int(n_cells_local),
                   "regular_bricks_laplace": int(op.n_fast_bricks()), "regular_bricks_fdm": int(fdm.n_fast_bricks()),
                   "chebyshev_degree": args.cheb_degree, "l2": "inputs larger than L2 (each vector %.0f MB)" % (n_own * S / 1e6),
                   "setup_s": t_setup},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
    }
    if extra:
        line["extra"] = extra
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
