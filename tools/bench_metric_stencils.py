"""Metric-fused stencils at C3 size (derivative X with dx(Y,X), metric-weighted variants).  python tools/bench_metric_stencils.py"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xgcm_b200 import ops, _capi
x = torch.empty((75, 2400, 3600), device="cuda"); ops.fill_uniform(x, 1)
dx = (1 + torch.rand((1, 2400, 3600), device="cuda")); dz = (1 + torch.rand((75, 1, 1), device="cuda"))
hfac = (0.2 + torch.rand((75, 2400, 3600), device="cuda"))
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
def t(fn, n=8):
    for _ in range(3): fn()
    torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n
cases = {
 "derivative X  (/ dx(Y,X))": (lambda: ops.stencil2(x, 2, "diff", 1, 0, "periodic", post=dx), 8),
 "interp X metric_weighted (x dx, / dx)": (lambda: ops.stencil2(x, 2, "interp", 1, 0, "periodic", pre=dx, post=dx), 8),
 "diff X x hFac(Z,Y,X) / dx(Y,X)": (lambda: ops.stencil2(x, 2, "diff", 1, 0, "periodic", pre=hfac, post=dx), 12),
 "interp Z metric_weighted (x dz, / dz)": (lambda: ops.stencil2(x, 0, "interp", 1, 0, "extend", pre=dz, post=dz), 8),
 "derivative Y (/ dx(Y,X))": (lambda: ops.stencil2(x, 1, "diff", 1, 0, "fill", post=dx), 8),
 "diff Y x hFac / dx(Y,X)": (lambda: ops.stencil2(x, 1, "diff", 1, 0, "fill", pre=hfac, post=dx), 12),
 "plain diff X": (lambda: ops.stencil2(x, 2, "diff", 1, 0, "periodic"), 8),
}
for name, (fn, bpc) in cases.items():
    ms = t(fn); print(f"{name:42s} {ms:7.3f} ms  frac {bpc*x.numel()/ms/1e6/peak:5.3f}  [{_capi.last_launch()}]", flush=True)
