"""Fused multi-axis interp at C3 size, one line per run — for env-knob sweeps.  python tools/bench_multi.py [label]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xgcm_b200 import ops, _capi
x = torch.empty((75, 2400, 3600), device="cuda"); ops.fill_uniform(x, 1)
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
def t(fn, n=8):
    for _ in range(3): fn()
    torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n
X, Y, Z = (2, "interp", 1, 0, "periodic", 0.0), (1, "interp", 1, 0, "fill", 0.0), (0, "interp", 1, 0, "extend", 0.0)
cases = [("XY", [X, Y]), ("YZ", [Y, Z]), ("XYZ", [X, Y, Z]), ("XYZ diff", [(2, "diff", 0, 1, "periodic", 0.0), (1, "diff", 0, 1, "periodic", 0.0), (0, "diff", 0, 1, "fill", 0.0)])]
out = []
for name, specs in cases:
    ms = t(lambda: ops.stencil_multi(x, specs)); out.append(f"{name} {ms:6.3f} ms {8*x.numel()/ms/1e6/peak:5.3f} [{_capi.last_launch().split('(')[-1][:-1]}]")
print((sys.argv[1] if len(sys.argv) > 1 else "default").ljust(28), " | ".join(out), flush=True)
