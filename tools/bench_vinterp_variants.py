"""C5 (75x2400x3600 fp32 -> 100 levels, shared theta) through ops.vinterp_linear for the tuning knobs of the
TMA kernel (XG_VINTERP_CPL / _W / _EXTRA) and the plain kernel (XG_VINTERP_TMA=0).
Usage (GPU box): python tools/bench_vinterp_variants.py [cpl:w:extra ...]"""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xgcm_b200 import _capi, ops

def timeit(fn, iters=8, warmup=3):
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); return ts[len(ts) // 2]

nz, ny, nx, m = 75, 2400, 3600, 100
dz = 10 * 1.05 ** np.arange(nz)
depth_np = (np.cumsum(dz) - dz / 2).astype(np.float32)
depth = torch.from_numpy(depth_np).cuda().reshape(-1, 1, 1)
target = torch.from_numpy(np.linspace(depth_np[0] - 5, depth_np[-1] + 5, m).astype(np.float32)).cuda()
x = torch.empty((nz, ny, nx), dtype=torch.float32, device="cuda")
ops.fill_uniform(x, 0xC0FFEE)
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
nbytes = ny * nx * (nz + m) * 4
variants = sys.argv[1:] or ["plain", "1:0:2:0", "1:9:2:3", "1:8:2:4", "1:8:2:2", "1:8:2:3", "1:10:0:3", "1:7:2:4", "1:6:2:5", "2:4:2:4", "2:4:2:8", "2:5:0:6"]
ref = None
for v in variants:
    for k in ("XG_VINTERP_TMA", "XG_VINTERP_CPL", "XG_VINTERP_W", "XG_VINTERP_EXTRA", "XG_VINTERP_WT"): os.environ.pop(k, None)
    if v == "plain":
        os.environ["XG_VINTERP_TMA"] = "0"
    else:
        cpl, w, extra, wt = v.split(":")  # columns per lane : teams (0 = auto) : lookahead buffers : warps per team (0 = auto)
        os.environ.update(XG_VINTERP_CPL=cpl, XG_VINTERP_W=w, XG_VINTERP_EXTRA=extra, XG_VINTERP_WT=wt)
    out = ops.vinterp_linear(x, depth, target, 0, True)
    path = _capi.last_launch()
    chk = out[::97, ::89].double().nan_to_num(nan=-7.0).sum().item()
    if ref is None: ref = chk
    ms = timeit(lambda: ops.vinterp_linear(x, depth, target, 0, True))
    print(f"{v:10s} {path:36s} {ms:7.3f} ms  {nbytes/ms/1e6:7.0f} GB/s  frac {nbytes/ms/1e6/peak:5.3f}  checksum_ok={chk == ref}", flush=True)
    del out
