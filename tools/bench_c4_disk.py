"""BASELINE configs[3] (C4) driven from DISK: time steps stored one per .npy file stream through
xgcm_b200.ingest.ChunkStream (reader thread -> page-locked ring -> copy stream -> device) while Grid.diff + Grid.interp
run on the previous step (X periodic, Y fill, Z extend: the six ops of bench.py's step).

    python tools/bench_c4_disk.py [--steps 6] [--dir /tmp/xgcm_b200_c4] [--keep]

Prints one JSON line: steps/s, cells/s, GB/s read from the files (just written, so mostly page cache: the figure is the
ingest path's ceiling, not the storage device's), ms per step of the kernels alone for comparison.
"""
import argparse, json, os, shutil, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xgcm_b200 as xg
from xgcm_b200 import ingest, ops

SHAPE = (75, 2400, 3600)

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--dir", default="/tmp/xgcm_b200_c4")
    ap.add_argument("--keep", action="store_true")
    args = ap.parse_args()
    nz, ny, nx = SHAPE
    cells = nz * ny * nx
    os.makedirs(args.dir, exist_ok=True)
    host = np.empty(SHAPE, np.float32)
    paths = []
    t0 = time.perf_counter()
    for t in range(args.steps):
        ops.fill_uniform_host(host.reshape(-1)[: cells // 64], 0xC0FFEE, t * cells)  # a slice is enough to make steps differ
        host.reshape(-1)[cells // 64:] = np.float32(t)
        p = os.path.join(args.dir, f"step_{t:05d}.npy")
        np.save(p, host)
        paths.append(p)
    t_write = time.perf_counter() - t0
    ds = xg.Dataset(coords={"Z": np.arange(nz) + 0.5, "Zl": np.arange(nz) + 0.0, "YC": np.arange(ny) + 0.5,
                            "YG": np.arange(ny) + 0.0, "XC": np.arange(nx) + 0.5, "XG": np.arange(nx) + 0.0})
    grid = xg.Grid(ds, coords={"X": {"center": "XC", "left": "XG"}, "Y": {"center": "YC", "left": "YG"},
                               "Z": {"center": "Z", "left": "Zl"}},
                   padding={"X": "periodic", "Y": "fill", "Z": "extend"}, autoparse_metadata=False)
    def six_ops(x):
        da = xg.DataArray(x, dims=("Z", "YC", "XC"))
        for ax in ("X", "Y", "Z"):
            for op in ("diff", "interp"):
                r = getattr(grid, op)(da, ax)
                del r
    stream = ingest.ChunkStream(paths, depth=3)
    # kernels alone
    x = torch.empty(SHAPE, dtype=torch.float32, device="cuda")
    for _ in range(2): six_ops(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); six_ops(x); e1.record(); torch.cuda.synchronize()
    ms_kernels = e0.elapsed_time(e1)
    del x
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 0
    for k, x in stream:
        six_ops(x)
        n += 1
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"workload": "C4 from disk: one 75x2400x3600 fp32 step per .npy file -> ChunkStream -> 6 ops per step",
                      "steps": n, "s_per_step": dt / n, "cells_per_s": 6 * cells * n / dt,
                      "file_GBps": stream.bytes_read / dt / 1e9, "ms_per_step_kernels_alone": ms_kernels,
                      "write_GBps": cells * 4 * args.steps / t_write / 1e9,
                      "note": "files were written just before being read: served largely from the page cache"}))
    if not args.keep:
        shutil.rmtree(args.dir, ignore_errors=True)

if __name__ == "__main__":
    main()
