"""Host (Python) time per Grid call with the kernel launch stubbed out — the part of the
BASELINE configs[0] / configs[1] latency that no kernel can fix.  CPU only; run here.

    python tools/prof_host_overhead.py [--profile]
"""
import argparse
import cProfile
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from xgcm_b200 import Grid, device, ops  # noqa: E402
from xgcm_b200.interop import DataArray, Dataset  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--n", type=int, default=2000)
    a = ap.parse_args()

    cpu = torch.device("cpu")
    device.default_device = lambda: cpu
    device.as_device_tensor = lambda data, dev=None: (data, False)
    device.result_like = lambda t, was_host: t
    res = {}

    def stencil2(x, axis, op, lo, hi, padding, fill_value=0.0, pre=None, post=None, halo_lo=None,
                 halo_hi=None, out=None):
        return res.setdefault((x.shape, axis, lo, hi), x)

    ops.stencil2 = stencil2
    nx = 1000
    x = torch.zeros(nx, dtype=torch.float64)
    ds = Dataset(
        {"f": DataArray(x, dims=("XC",), coords={"XC": np.arange(nx) + 0.5})},
        coords={"XC": np.arange(nx) + 0.5, "XG": np.arange(nx) * 1.0},
    )
    grid = Grid(ds, coords={"X": {"center": "XC", "left": "XG"}}, padding="periodic", autoparse_metadata=False)
    f = ds["f"]

    def body():
        for _ in range(a.n):
            grid.diff(f, "X")

    body()
    if a.profile:
        pr = cProfile.Profile()
        pr.enable()
        body()
        pr.disable()
        pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
    t0 = time.perf_counter()
    body()
    dt = time.perf_counter() - t0
    print(f"Grid.diff host time: {dt / a.n * 1e6:.1f} us/call")


if __name__ == "__main__":
    main()
