"""xg_stencil_pair at C3 size: plain and metric-fused, against the explicit chain.  python tools/bench_pair.py"""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xgcm_b200 import ops
x = torch.empty((75, 2400, 3600), device="cuda"); v = torch.empty_like(x)
ops.fill_uniform(x, 1); ops.fill_uniform(v, 2)
dx = (1 + torch.rand((1, 2400, 3600), device="cuda"))
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
def t(fn, n=8):
    for _ in range(3): fn()
    torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n
sa, sb = ("diff", 0, 1, "periodic", 0.0), (1, "diff", 0, 1, "periodic", 0.0)
nb = 12 * x.numel()
ms = t(lambda: ops.stencil_pair(x, v, sa, sb, 0)); print(f"pair plain    {ms:.3f} ms  frac {nb/ms/1e6/peak:.3f}")
ms = t(lambda: ops.stencil_pair(x, v, sa, sb, 0, pre_a=dx, pre_b=dx, post=dx)); print(f"pair metrics  {ms:.3f} ms  frac {nb/ms/1e6/peak:.3f}")
def chain():
    a = ops.stencil2(ops.binary("mul", x, dx), 2, "diff", 0, 1, "periodic")
    b = ops.stencil2(ops.binary("mul", v, dx), 1, "diff", 0, 1, "periodic")
    return ops.binary("div", ops.binary("add", a, b), dx)
ms = t(chain, 4); print(f"explicit chain (device ops) {ms:.3f} ms")
