"""The three row (X) metric stencils at C3 size, one line per run — for env-knob sweeps of the TMA-staged kernel.
python tools/bench_row_tma.py [label]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xgcm_b200 import ops, _capi
x = torch.empty((75, 2400, 3600), device="cuda"); ops.fill_uniform(x, 1)
dx = (1 + torch.rand((1, 2400, 3600), device="cuda"))
hfac = (0.2 + torch.rand((75, 2400, 3600), device="cuda"))
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n
cases = [
 ("derivX", lambda: ops.stencil2(x, 2, "diff", 1, 0, "periodic", post=dx), 8),
 ("interpX_w", lambda: ops.stencil2(x, 2, "interp", 1, 0, "periodic", pre=dx, post=dx), 8),
 ("diffX_hfac", lambda: ops.stencil2(x, 2, "diff", 1, 0, "periodic", pre=hfac, post=dx), 12),
]
out = []
for name, fn, bpc in cases:
    ms = t(fn); out.append(f"{name} {ms:6.3f} ms {bpc*x.numel()/ms/1e6/peak:5.3f} [{_capi.last_launch().split('(')[1][:-1]}]")
print((sys.argv[1] if len(sys.argv) > 1 else "default").ljust(44), " | ".join(out), flush=True)
