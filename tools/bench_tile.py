"""The tile-kernel cases at C3 size (row-axis metric stencils, divergence), one line per run — for env-knob sweeps.
python tools/bench_tile.py [label]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xgcm_b200 import ops, _capi
x = torch.empty((75, 2400, 3600), device="cuda"); v = torch.empty_like(x)
ops.fill_uniform(x, 1); ops.fill_uniform(v, 2)
dx = (1 + torch.rand((1, 2400, 3600), device="cuda")); dy = (1 + torch.rand((1, 2400, 3600), device="cuda"))
ra = dx * dy
hfac = (0.2 + torch.rand((75, 2400, 3600), device="cuda"))
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
def t(fn, n=8):
    for _ in range(3): fn()
    torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n
sa, sb = ("diff", 0, 1, "periodic", 0.0), (1, "diff", 0, 1, "periodic", 0.0)
cases = [
 ("derivY", lambda: ops.stencil2(x, 1, "diff", 1, 0, "fill", post=dx), 8),
 ("diffY_hfac", lambda: ops.stencil2(x, 1, "diff", 1, 0, "fill", pre=hfac, post=dx), 12),
 ("div3m", lambda: ops.stencil_pair(x, v, sa, sb, 0, pre_a=dy, pre_b=dx, post=ra), 12),
 ("vort", lambda: ops.stencil_pair(v, x, ("diff", 1, 0, "periodic", 0.0), (1, "diff", 1, 0, "extend", 0.0), 1, pre_a=dy, pre_b=dx, post=ra), 12),
]
out = []
for name, fn, bpc in cases:
    ms = t(fn); out.append(f"{name} {ms:6.3f} ms {bpc*x.numel()/ms/1e6/peak:5.3f} [{_capi.last_launch().split('(')[-1][:-1]}]")
print((sys.argv[1] if len(sys.argv) > 1 else "default").ljust(30), " | ".join(out), flush=True)
