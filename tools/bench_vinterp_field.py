"""C5-sized transform with a per-column theta FIELD (k_vinterp_columns[_tma]).  python tools/bench_vinterp_field.py [nt:extra:wt ...]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xgcm_b200 import _capi, ops
def timeit(fn, iters=6, warmup=2):
    for _ in range(warmup): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); return ts[len(ts) // 2]
nz, ny, nx, m = 75, 2400, 3600, 100
x = torch.empty((nz, ny, nx), dtype=torch.float32, device="cuda"); ops.fill_uniform(x, 1)
th = torch.cumsum(x + 0.5, 0)
levels = torch.linspace(0.0, float(th.max()) + 1, m, device="cuda")
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
nbytes = ny * nx * (2 * nz + m) * 4
for v in (sys.argv[1:] or ["plain", "0:2:0", "5:2:6", "6:1:5", "6:0:5", "4:2:8", "5:1:6", "5:2:4"]):
    for k in ("XG_VINTERP_TMA", "XG_VINTERP_W", "XG_VINTERP_EXTRA", "XG_VINTERP_WT"): os.environ.pop(k, None)
    if v == "plain": os.environ["XG_VINTERP_TMA"] = "0"
    else:
        nt, extra, wt = v.split(":"); os.environ.update(XG_VINTERP_W=nt, XG_VINTERP_EXTRA=extra, XG_VINTERP_WT=wt)
    out = ops.vinterp_linear(x, th, levels, 0, True); path = _capi.last_launch()
    ms = timeit(lambda: ops.vinterp_linear(x, th, levels, 0, True))
    print(f"{v:8s} {path:36s} {ms:7.3f} ms  {nbytes/ms/1e6:7.0f} GB/s  frac {nbytes/ms/1e6/peak:5.3f}", flush=True)
    del out
