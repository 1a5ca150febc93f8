"""Per-kernel HBM throughput for every kernel of the library at C3 scale (fp32).

Usage (GPU box): python tools/kernel_bench_all.py [--small]
Prints GB/s of ALGORITHMIC bytes and the fraction of the measured copy peak.
"""

import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xgcm_b200 import ops  # noqa: E402


def timeit(fn, iters=8, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--dtype", default="f32")
    args = ap.parse_args()
    shape = (25, 1200, 1800) if args.small else (75, 2400, 3600)
    dt = torch.float32 if args.dtype == "f32" else torch.float64
    es = 4 if args.dtype == "f32" else 8
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    x = torch.empty(shape, dtype=dt, device="cuda")
    ops.fill_uniform(x, 1)
    cells = x.numel()
    names = "ZYX"
    rows = []

    def report(name, ms, nbytes):
        gbs = nbytes / ms / 1e6
        rows.append((name, ms, gbs, gbs / peak))
        print(f"{name:44s} {ms:8.3f} ms {gbs:8.1f} GB/s  {gbs/peak:5.2f} of measured peak", flush=True)

    y = torch.empty_like(x)
    report("torch copy_ (reference point)", timeit(lambda: y.copy_(x)), 2 * cells * es)
    del y
    for ax in range(3):
        out = torch.empty_like(x)
        report(f"stencil2 diff {names[ax]} c->l periodic", timeit(lambda: ops.stencil2(x, ax, "diff", 1, 0, "periodic", out=out)), 2 * cells * es)
        del out
    report("stencil_multi interp X then Y (fused, 1 pass)", timeit(lambda: ops.stencil_multi(x, [(2, "interp", 1, 0, "periodic", 0.0), (1, "interp", 1, 0, "fill", 0.0)])), 2 * cells * es)
    report("stencil_multi interp Y then Z (fused, 1 pass)", timeit(lambda: ops.stencil_multi(x, [(1, "interp", 1, 0, "fill", 0.0), (0, "interp", 1, 0, "extend", 0.0)])), 2 * cells * es)
    report("stencil_multi interp X,Y,Z (fused, 1 pass)", timeit(lambda: ops.stencil_multi(x, [(2, "interp", 1, 0, "periodic", 0.0), (1, "interp", 1, 0, "fill", 0.0), (0, "interp", 1, 0, "extend", 0.0)])), 2 * cells * es)
    # metric-weighted derivative along X with 2-D dx (Y, X) and along Z with 1-D dz
    dx = (1 + torch.rand((1, shape[1], shape[2]), device="cuda", dtype=dt))
    dz = (1 + torch.rand((shape[0], 1, 1), device="cuda", dtype=dt))
    hfac = (0.2 + torch.rand(shape, device="cuda", dtype=dt))
    out = torch.empty_like(x)
    report("stencil2 derivative X (post = dx(Y,X))", timeit(lambda: ops.stencil2(x, 2, "diff", 1, 0, "periodic", post=dx, out=out)), (2 * cells + dx.numel()) * es)
    report("stencil2 interp Z metric_weighted dz(Z)", timeit(lambda: ops.stencil2(x, 0, "interp", 1, 0, "extend", pre=dz, post=dz, out=out)), 2 * cells * es)
    report("stencil2 diff Y pre=hFac(Z,Y,X) post=dx", timeit(lambda: ops.stencil2(x, 1, "diff", 1, 0, "fill", pre=hfac, post=dx, out=out)), 3 * cells * es)
    del out
    for ax in range(3):
        report(f"cumscan {names[ax]} c->r", timeit(lambda: ops.cumscan(x, ax)), 2 * cells * es)
        report(f"cumscan {names[ax]} c->l fill (drop_last, pad)", timeit(lambda: ops.cumscan(x, ax, False, "drop_last", 1, 0, "fill")), 2 * cells * es)
    report("cumscan Z reverse c->l", timeit(lambda: ops.cumscan(x, 0, True)), 2 * cells * es)
    for ax in range(3):
        n = shape[ax]
        report(f"wreduce sum {names[ax]} (integrate, 1-D weight)", timeit(lambda: ops.wreduce(x, ax, (dz if ax == 0 else None), "sum")), (cells + cells // n) * es)
    report("wreduce sum Z weight hFac(Z,Y,X)", timeit(lambda: ops.wreduce(x, 0, hfac, "sum")), (2 * cells + cells // shape[0]) * es)
    report("wreduce mean Z (average)", timeit(lambda: ops.wreduce(x, 0, dz, "mean")), (cells + cells // shape[0]) * es)
    del hfac
    # vertical transform Z -> 100 levels (BASELINE configs[4])
    m = 100
    depth = torch.cumsum(10 * 1.05 ** torch.arange(shape[0], device="cuda", dtype=dt), 0).reshape(-1, 1, 1)
    target = torch.linspace(float(depth[0]) - 5, float(depth[-1]) + 5, m, device="cuda", dtype=dt)
    cols = shape[1] * shape[2]
    report("vinterp_linear Z->100, shared 1-D theta", timeit(lambda: ops.vinterp_linear(x, depth, target, 0, True), iters=4), cols * (shape[0] + m) * es)
    theta3 = torch.cumsum(0.5 + torch.rand(shape, device="cuda", dtype=dt), 0)
    tg2 = torch.linspace(0, float(theta3.max()), m, device="cuda", dtype=dt)
    report("vinterp_linear Z->100, theta field", timeit(lambda: ops.vinterp_linear(x, theta3, tg2, 0, True), iters=4), cols * (2 * shape[0] + m) * es)
    del theta3
    report("pad X periodic (1,1)", timeit(lambda: ops.pad(x, 2, 1, 1, "periodic")), 2 * cells * es)
    report("binary mul x * dx(Y,X)", timeit(lambda: ops.binary("mul", x, dx)), 2 * cells * es)
    # ---- face connections: cubed sphere, 6 faces x 50 levels x 1020^2 ----------
    import xgcm_b200 as xg

    del x
    torch.cuda.empty_cache()
    nz, nf, n = (10, 6, 510) if args.small else (50, 6, 1020)
    cs = {
        "face": {
            0: {"X": ((3, "X", False), (1, "X", False)), "Y": ((4, "Y", False), (5, "Y", False))},
            1: {"X": ((0, "X", False), (2, "X", False)), "Y": ((4, "X", False), (5, "X", True))},
            2: {"X": ((1, "X", False), (3, "X", False)), "Y": ((4, "Y", True), (5, "Y", True))},
            3: {"X": ((2, "X", False), (0, "X", False)), "Y": ((4, "X", True), (5, "X", False))},
            4: {"X": ((3, "Y", True), (1, "Y", False)), "Y": ((2, "Y", True), (0, "Y", False))},
            5: {"X": ((3, "Y", False), (1, "Y", True)), "Y": ((0, "Y", False), (2, "Y", True))},
        }
    }
    f = torch.empty((nz, nf, n, n), dtype=dt, device="cuda")
    ops.fill_uniform(f, 3)
    fcells = f.numel()
    ds = xg.Dataset(coords={"z": np.arange(nz), "face": np.arange(nf), "y": np.arange(n), "yl": np.arange(n) - 0.5,
                            "x": np.arange(n), "xl": np.arange(n) - 0.5})
    grid = xg.Grid(ds, coords={"X": {"center": "x", "left": "xl"}, "Y": {"center": "y", "left": "yl"}},
                   face_connections=cs)
    da = xg.DataArray(f, dims=("z", "face", "y", "x"))
    report("cubed sphere Grid.diff X (halo planes + fused stencil)", timeit(lambda: grid.diff(da, "X"), iters=6), 2 * fcells * es)
    report("cubed sphere Grid.interp Y (halo planes + fused stencil)", timeit(lambda: grid.interp(da, "Y"), iters=6), 2 * fcells * es)
    from xgcm_b200.padding import pad as xpad

    report("cubed sphere pad X,Y (1,1) materialised", timeit(lambda: xpad(da, grid, {"X": (1, 1), "Y": (1, 1)}), iters=4), 2 * fcells * es)
    out_path = os.path.join("gpurun_out", "kernel_bench_all.json")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump([dict(name=r[0], ms=r[1], GBps=r[2], frac_of_measured_peak=r[3]) for r in rows], open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
