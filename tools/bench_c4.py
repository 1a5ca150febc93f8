"""BASELINE configs[3] (C4): the 3600x2400x75 fp32 field x 365 time steps, sharded on TIME across the
GPUs of one box (contiguous blocks of steps per rank, no collective on the data path).

    python tools/bench_c4.py [--steps 365]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_c4.py

946 GB do not fit anywhere, so every step's field is generated ON THE DEVICE (counter-based
generator keyed by the global step index: any rank — or the host — can regenerate any step), then
Grid.diff + Grid.interp run along X (periodic), Y (fill), Z (extend) like bench.py's step.  Timed:
the six Grid calls of every step (CUDA events; generation excluded), max over ranks.  Parity of
sampled global steps against the oracle is the job of tests/test_fullsize_gpu.py::test_config4_sampled_steps
(same generator, same Grid); this tool reports, per rank, a checksum of three sampled steps' results
so runs at different N can be compared with each other.

Prints one JSON line: aggregate cells/s over all ranks, per-rank ms per step, sampled-step checksums.
"""

import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xgcm_b200 as xg  # noqa: E402
from xgcm_b200 import ops, parallel  # noqa: E402

SHAPE = (75, 2400, 3600)
CASES = (("X", 2, "periodic"), ("Y", 1, "fill"), ("Z", 0, "extend"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=365)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nz, ny, nx = SHAPE
    ds = xg.Dataset(coords={"Z": np.arange(nz) + 0.5, "Zl": np.arange(nz) + 0.0, "YC": np.arange(ny) + 0.5,
                            "YG": np.arange(ny) + 0.0, "XC": np.arange(nx) + 0.5, "XG": np.arange(nx) + 0.0})
    grid = xg.Grid(ds, coords={"X": {"center": "XC", "left": "XG"}, "Y": {"center": "YC", "left": "YG"},
                               "Z": {"center": "Z", "left": "Zl"}},
                   padding={"X": "periodic", "Y": "fill", "Z": "extend"}, autoparse_metadata=False)
    t0, t1 = parallel.shard_bounds(args.steps, world, rank)
    x = torch.empty(SHAPE, dtype=torch.float32, device="cuda")
    da = xg.DataArray(x, dims=("Z", "YC", "XC"))
    cells = x.numel()
    sampled = sorted({t0, (t0 + t1) // 2, t1 - 1}) if t1 > t0 else []
    checks = []

    def step_field(t):
        ops.fill_uniform(x, 0xC0FFEE, offset=t * cells)  # global step index keys the values

    for _ in range(2):  # warm-up
        step_field(t0 if t1 > t0 else 0)
        for ax, _, _ in CASES:
            grid.diff(da, ax), grid.interp(da, ax)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for t in range(t0, t1):
        step_field(t)
        e0.record()
        outs = [(op, k, bc, getattr(grid, op)(da, ax)) for ax, k, bc in CASES for op in ("diff", "interp")]
        e1.record()
        torch.cuda.synchronize()
        total_ms += e0.elapsed_time(e1)
        if t in sampled:
            checks.append([t] + [float(out.data.double().sum()) for _, _, _, out in outs])
        del outs
    stats = torch.tensor([total_ms, float(t1 - t0)], device="cuda", dtype=torch.float64)
    worst = stats.clone()
    if world > 1:
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        gathered = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(gathered, stats)
        all_checks = [None] * world
        dist.all_gather_object(all_checks, checks)
    else:
        gathered, all_checks = [stats], [checks]
    if rank == 0:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
        max_ms = float(worst[0])
        total_cells = args.steps * 6 * cells
        print(json.dumps({
            "config": "C4: 75x2400x3600 fp32 x %d time steps, time-sharded over %d GPU(s), Grid.diff + Grid.interp on X/Y/Z" % (args.steps, world),
            "metric": "grid-cells/s", "value": total_cells / (max_ms / 1e3), "n_gpus": world,
            "steps_per_rank": [int(g[1]) for g in gathered],
            "ms_per_step_per_rank": [float(g[0] / max(g[1], 1)) for g in gathered],
            "max_rank_ms": max_ms,
            "hbm_frac_per_gpu": (float(gathered[0][1]) * 6 * cells * 8 / (float(gathered[0][0]) / 1e3)) / 1e9 / peak,
            "sampled_step_checksums": {str(int(c[0])): c[1:] for cs in all_checks for c in cs},
            "data": "synthetic, generated on the device per step (fill_uniform keyed by global step)",
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
