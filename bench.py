#!/usr/bin/env python
"""bench.py — grid-cells/s for Grid.diff + Grid.interp on a C-grid field (BASELINE.json metric).

A *step* = one pass of the hot path over one synthetic (75, 2400, 3600) fp32 field
(BASELINE configs[2], "C3"): ``Grid.diff`` and ``Grid.interp`` along X (periodic),
Y (fill) and Z (extend) — six fused ``xg_stencil2`` launches, 6 x 648 M output cells.

  value   whole-job cells/s with the field already resident in HBM (device-backed DataArray),
          timed with CUDA events, barrier + synchronize on both sides, max over ranks.
  e2e     the same six Grid calls on a HOST (page-locked numpy) DataArray: every call streams
          H2D -> kernel -> D2H inside the C-ABI (xg_stencil2_host); copies are in the timed region.
  roofline  algorithmic bytes (8 B/cell fp32) / mean per-launch duration measured live
          with CUDA events on the launching stream, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the oracle (numpy restatement of the reference's calls: np.pad + ufunc on a
          moveaxis view) timed single-threaded on the host, rank 0, N=1, on a bounded sample.

``--impl reference`` times the reference's own CPU path (the oracle port; xarray/dask are not
installable) with a thread pool over all host cores on broadcast-dim chunks —
the dask="parallelized" analogue (xgcm/grid.py:786-789).

Multi-GPU (torchrun): every rank owns the time steps of its shard (one field per step, no
data-path collective) => weak scaling.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPE = (75, 2400, 3600)  # (Z, Y, X), C3
DTYPE = np.float32
SEED = 0xC0FFEE
# (axis name, padding, fill) — SURVEY 8d: periodic X, fill Y, extend Z
AXES = (("X", "periodic", None), ("Y", "fill", 0.0), ("Z", "extend", None))
OPS = ("diff", "interp")
METRIC = "grid-cells/s for diff+interp on C-grid field; achieved HBM GB/s vs peak"
WORKLOAD = ("C3 (BASELINE configs[2]): Grid.diff + Grid.interp along X(periodic), Y(fill), Z(extend) of one "
            "75x2400x3600 fp32 C-grid field per step")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", type=int, nargs=3, default=list(SHAPE))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def make_dataset(shape, field):
    import xgcm_b200 as xg

    nz, ny, nx = shape
    coords = {
        "Z": np.arange(nz) + 0.5, "Zl": np.arange(nz) + 0.0,
        "YC": np.arange(ny) + 0.5, "YG": np.arange(ny) + 0.0,
        "XC": np.arange(nx) + 0.5, "XG": np.arange(nx) + 0.0,
    }
    ds = xg.Dataset(coords=coords)
    grid = xg.Grid(
        ds,
        coords={"X": {"center": "XC", "left": "XG"}, "Y": {"center": "YC", "left": "YG"},
                "Z": {"center": "Z", "left": "Zl"}},
        padding={"X": "periodic", "Y": "fill", "Z": "extend"},
        autoparse_metadata=False,
    )
    da = xg.DataArray(field, dims=("Z", "YC", "XC"), name="theta")
    return grid, da


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = f"/tmp/xgcm_b200_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.out = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=self.out, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()  # exact PID we started
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.out.close()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1]))
                    smax.append(float(parts[2]))
                except ValueError:
                    continue
                for name, val in zip(names, parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.remove(self.path)
        except Exception:
            pass
        busy = [s for s in sm if s > 0]
        return {
            "sm_mhz": statistics.median(busy) if busy else None,
            "sm_max_mhz": max(smax) if smax else None,
            "reasons": sorted(reasons),
            "samples": len(sm),
        }


# --------------------------------------------------------------------------- the oracle legs (CPU)
def oracle_step(a, pool=None, nchunks=1):
    """The reference's numpy calls for the six ops (np.pad copy + kernel on a moveaxis view)."""
    from oracle import stencil as oracle

    def one(op, axis_i, bc, fill):
        lo, hi = 1, 0  # center -> left
        if pool is None:
            return oracle.stencil2(op, a, axis_i, lo, hi, bc, 0.0 if fill is None else fill)
        # dask="parallelized" analogue: chunk a broadcast (non-operated) dim, one task per chunk
        chunk_axis = 0 if axis_i != 0 else 1
        nch = max(1, min(nchunks, a.shape[chunk_axis]))
        bounds = np.linspace(0, a.shape[chunk_axis], nch + 1).astype(int)
        out = np.empty(a.shape, a.dtype)

        def task(i):
            sl = [slice(None)] * a.ndim
            sl[chunk_axis] = slice(bounds[i], bounds[i + 1])
            out[tuple(sl)] = oracle.stencil2(op, a[tuple(sl)], axis_i, lo, hi, bc, 0.0 if fill is None else fill)

        list(pool.map(task, range(nch)))
        return out

    axis_index = {"Z": 0, "Y": 1, "X": 2}
    cells = 0
    for ax, bc, fill in AXES:
        for op in OPS:
            r = one(op, axis_index[ax], bc, fill)
            cells += r.size
            del r
    return cells


def run_reference(args):
    """--impl reference: the reference's CPU path on all host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor

    from xgcm_b200 import ops

    cores = os.cpu_count() or 1
    shape = tuple(args.shape)
    a = np.empty(shape, DTYPE)
    ops.fill_uniform_host(a.reshape(-1), SEED)
    nchunks = cores
    budget_s = 150.0  # the whole --steps K --warmup W run should end within a few minutes
    with ThreadPoolExecutor(max_workers=cores) as pool:
        # calibrate on one full step, then bound the per-step sample to a leading block of Z levels
        t0 = time.perf_counter()
        oracle_step(a, pool, nchunks)
        t_full = time.perf_counter() - t0
        steps_total = args.steps + max(args.warmup - 1, 0)
        nz = shape[0]
        nz_s = nz if t_full * steps_total <= budget_s else max(2, int(nz * budget_s / (t_full * steps_total)))
        sub = a if nz_s == nz else np.ascontiguousarray(a[:nz_s])
        for _ in range(max(args.warmup - 1, 0)):
            oracle_step(sub, pool, nchunks)
        t0 = time.perf_counter()
        cells = 0
        for _ in range(args.steps):
            cells += oracle_step(sub, pool, nchunks)
        dt = time.perf_counter() - t0
    value = cells / dt
    sample = (f"{'full workload' if nz_s == nz else f'first {nz_s} of {nz} Z levels of the field'} per step, "
              f"{args.steps} steps; thread pool of {cores} over broadcast-dim chunks (dask='parallelized' analogue); "
              f"one full step took {t_full:.2f} s")
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "shape": list(shape), "cells_per_step": cells // args.steps},
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "numpy": np.__version__,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from xgcm_b200 import _capi, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    dev = torch.device("cuda", torch.cuda.current_device())
    lib = _capi.load()
    shape = tuple(args.shape)
    cells_field = int(np.prod(shape))
    es = np.dtype(DTYPE).itemsize

    # this rank's field: time step `rank` of the synthetic series (counter-based RNG => any shard anywhere)
    x = torch.empty(shape, dtype=torch.float32, device=dev)
    ops.fill_uniform(x, SEED, offset=rank * cells_field)
    grid, da_dev = make_dataset(shape, x)

    def step_device(timer=None):
        n = 0
        for ax, bc, fill in AXES:
            for op in OPS:
                if timer is not None:
                    timer.append((ax, op, torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
                    timer[-1][2].record()
                r = getattr(grid, op)(da_dev, ax)
                if timer is not None:
                    timer[-1][3].record()
                n += r.size
                del r
        return n

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM ------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        sampler.start()
    launches0 = lib.xg_launch_count()
    timers = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    cells = 0
    for _ in range(args.steps):
        cells += step_device(timers)
    e1.record()
    barrier()
    launches = lib.xg_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = cells * world / (ms_total * 1e-3)

    per_kernel = {}
    for ax, op, s, e in timers:
        per_kernel.setdefault(f"{op}_{ax}", []).append(s.elapsed_time(e))
    launch_ms = [t for v in per_kernel.values() for t in v]
    mean_launch_ms = sum(launch_ms) / len(launch_ms)
    peak, peak_src = measured_peak()
    alg_bytes = 2 * es * cells_field  # read n + write n (8 B/cell fp32)
    achieved = alg_bytes / (mean_launch_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "peak_source": peak_src,
        "kernel": "xg_stencil2 (k_stencil_row_vec for X, k_stencil_strided for Y/Z)",
        "algorithmic_bytes_per_launch": alg_bytes,
        "mean_launch_ms": mean_launch_ms,
        "per_op_GBps": {k: alg_bytes / (statistics.median(v) * 1e-3) / 1e9 for k, v in per_kernel.items()},
    }

    # ---- e2e: host buffers through the public Grid API -----------------------------------------
    e2e = None
    if not args.no_e2e:
        host = ops.pinned_empty(shape, DTYPE)
        torch.from_numpy(host).copy_(x)
        torch.cuda.synchronize()
        _, da_host = make_dataset(shape, host)

        def step_host():
            n = 0
            for ax, bc, fill in AXES:
                for op in OPS:
                    r = getattr(grid, op)(da_host, ax)  # numpy in -> numpy out, copies inside
                    n += r.size
                    del r
            return n

        for _ in range(2):
            step_host()
        barrier()
        k = max(1, min(args.steps, 5))
        t0 = time.perf_counter()
        n = 0
        for _ in range(k):
            n += step_host()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        nbytes = 6 * cells_field * es
        e2e = {"value": n * world / float(dt.item()), "unit": "cells/s", "h2d_bytes_per_step": nbytes,
               "d2h_bytes_per_step": nbytes, "steps": k, "ms_per_step": float(dt.item()) / k * 1e3,
               "path": "Grid.diff/interp(numpy DataArray) -> xg_stencil2_host (3-stream slab pipeline)"}
        del host, da_host

    # ---- cpu_baseline (rank 0, N=1 only): the oracle, single thread, one full step ---------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        a = x.cpu().numpy()
        t0 = time.perf_counter()
        n = oracle_step(a)  # numpy elementwise ufuncs / np.pad are single-threaded
        dt = time.perf_counter() - t0
        cpu = {"value": n / dt, "unit": "cells/s", "cores": 1, "kind": "port",
               "sample": f"1 full step (6 ops x {cells_field} cells) of the same workload, numpy {np.__version__}, "
                         f"host has {os.cpu_count()} cores; --impl reference uses all of them",
               "seconds": dt}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "shape": list(shape), "cells_per_step_per_gpu": cells // args.steps,
                       "l2": "each field is 2.59 GB in + 2.59 GB out per launch, >> 126 MB L2: no flush needed",
                       "parallelism": f"time-shards x{world} (one field per rank per step, no collective on the data path)"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
