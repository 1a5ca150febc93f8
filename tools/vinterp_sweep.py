"""vinterp_linear (shared theta) throughput vs field size: is the C5 kernel bound by address translation?
Usage (GPU box): python tools/vinterp_sweep.py"""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xgcm_b200 import ops

def timeit(fn, iters=6, warmup=2):
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); return ts[len(ts) // 2]

nz, m = 75, 100
depth = torch.cumsum(10 * 1.05 ** torch.arange(nz, device="cuda", dtype=torch.float32), 0).reshape(-1, 1, 1)
target = torch.linspace(float(depth[0]) - 5, float(depth[-1]) + 5, m, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for ny, nx in [(300, 450), (480, 720), (600, 900), (848, 1272), (1200, 1800), (1696, 2544), (2400, 3600)]:
    x = torch.empty((nz, ny, nx), dtype=torch.float32, device="cuda")
    ops.fill_uniform(x, 1)
    def run():
        flush.zero_()  # evict L2 between calls so small cases are not cache-resident
        return ops.vinterp_linear(x, depth, target, 0, True)
    t_all = timeit(run)
    t_flush = timeit(lambda: flush.zero_())
    ms = t_all - t_flush
    nbytes = ny * nx * (nz + m) * 4
    print(f"{nz}x{ny}x{nx}: in {x.numel()*4/2**20:7.0f} MiB out {ny*nx*m*4/2**20:7.0f} MiB  {ms:7.3f} ms  {nbytes/ms/1e6:7.0f} GB/s", flush=True)
    del x
