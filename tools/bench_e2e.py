"""Host-buffer (e2e) throughput of one streamed stencil call: python tools/bench_e2e.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xgcm_b200 import ops

shape = (75, 2400, 3600)
x = ops.pinned_empty(shape, np.float32)
ops.fill_uniform_host(x.reshape(-1), 1)
out = ops.pinned_empty(shape, np.float32)
for axis in (2, 0):
    for _ in range(2):
        ops.stencil2_host(x, axis, "diff", 1, 0, "periodic", out=out)
    t0 = time.perf_counter()
    k = 5
    for _ in range(k):
        ops.stencil2_host(x, axis, "diff", 1, 0, "periodic", out=out)
    dt = (time.perf_counter() - t0) / k
    print(f"slab={os.environ.get('XG_HOST_SLAB_MB','128')}MB axis={axis}: {dt*1e3:.1f} ms  {x.nbytes/dt/1e9:.1f} GB/s each way  {x.size/dt/1e9:.2f} Gcell/s")
# plain copies for reference
d = torch.empty(shape, dtype=torch.float32, device="cuda")
h = torch.from_numpy(x)
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: torch.from_numpy(out).copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): fn()
    torch.cuda.synchronize()
    print(f"plain {name}: {x.nbytes*3/(time.perf_counter()-t0)/1e9:.1f} GB/s")
