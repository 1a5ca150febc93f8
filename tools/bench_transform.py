"""BASELINE configs[4]: linear vertical regrid of a 75x2400x3600 fp32 field to 100 target levels
through Grid.transform, on 1 GPU or (torchrun) N GPUs sharding Y (never the operated Z axis).

    python tools/bench_transform.py
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_transform.py

Prints one JSON line: output cells/s, ms per call, achieved algorithmic GB/s ((n + m) * 4 B per column).
"""

import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xgcm_b200 as xg  # noqa: E402
from xgcm_b200 import ops, parallel  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nz, ny, nx, m = 75, 2400, 3600, 100
    y0, y1 = parallel.shard_bounds(ny, world, rank)
    dz = 10 * 1.05 ** np.arange(nz)
    depth = (np.cumsum(dz) - dz / 2).astype(np.float32)
    x = torch.empty((nz, y1 - y0, nx), dtype=torch.float32, device="cuda")
    ops.fill_uniform(x, 0xC0FFEE, offset=y0 * nx)
    ds = xg.Dataset(coords={"Z": depth})
    grid = xg.Grid(ds, coords={"Z": {"center": "Z"}}, autoparse_metadata=False)
    da = xg.DataArray(x, dims=("Z", "Y", "X"), name="theta")
    levels = np.linspace(depth[0] - 5, depth[-1] + 5, m).astype(np.float32)
    for _ in range(3):
        out = grid.transform(da, "Z", levels)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    k = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        out = grid.transform(da, "Z", levels)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / k], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    assert out.dims == ("Y", "X", "Z") and out.shape == (y1 - y0, nx, m)
    if rank == 0:
        cols = ny * nx
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
        gbs = cols / world * (nz + m) * 4 / (float(ms) * 1e-3) / 1e9
        print(json.dumps({"workload": "C5 transform Z->100 levels (linear, mask_edges), Y sharded", "n_gpus": world,
                          "ms_per_call": float(ms), "out_cells_per_s": cols * m / (float(ms) * 1e-3),
                          "per_gpu_algorithmic_GBps": gbs, "frac_of_measured_peak": gbs / peak}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
