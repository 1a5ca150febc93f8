"""Run one kernel of the library a few times (for ncu): python tools/prof_one.py <name>"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xgcm_b200 import ops

name = sys.argv[1]
shape = (75, 2400, 3600)
x = torch.empty(shape, dtype=torch.float32, device="cuda")
ops.fill_uniform(x, 1)
x2 = torch.empty_like(x) if name == "divergence" else None
if x2 is not None:
    ops.fill_uniform(x2, 2)
dz = (1 + torch.rand((shape[0], 1, 1), device="cuda"))
dx = (1 + torch.rand((1, shape[1], shape[2]), device="cuda"))
depth = torch.cumsum(10 * 1.05 ** torch.arange(shape[0], device="cuda", dtype=torch.float32), 0).reshape(-1, 1, 1)
target = torch.linspace(float(depth[0]) - 5, float(depth[-1]) + 5, 100, device="cuda")
fns = {
    "vinterp_shared": lambda: ops.vinterp_linear(x, depth, target, 0, True),
    "vinterp_field": lambda: ops.vinterp_linear(x, torch.cumsum(x + 0.5, 0), target, 0, True),
    "cumscan_x": lambda: ops.cumscan(x, 2),
    "cumscan_y": lambda: ops.cumscan(x, 1),
    "derivative_x": lambda: ops.stencil2(x, 2, "diff", 1, 0, "periodic", post=dx),
    "derivative_y": lambda: ops.stencil2(x, 1, "diff", 1, 0, "fill", post=dx),
    "divergence": lambda: ops.stencil_pair(x, x2, ("diff", 0, 1, "periodic", 0.0), (1, "diff", 0, 1, "periodic", 0.0), 0,
                                           pre_a=dx, pre_b=dx, post=dx),
    "interp_z_metric": lambda: ops.stencil2(x, 0, "interp", 1, 0, "extend", pre=dz, post=dz),
    "wreduce_z": lambda: ops.wreduce(x, 0, dz, "sum"),
    "wreduce_x": lambda: ops.wreduce(x, 2, None, "sum"),
    "multi_xyz": lambda: ops.stencil_multi(x, [(2, "interp", 1, 0, "periodic", 0.0), (1, "interp", 1, 0, "fill", 0.0), (0, "interp", 1, 0, "extend", 0.0)]),
    "multi_yz": lambda: ops.stencil_multi(x, [(1, "interp", 1, 0, "fill", 0.0), (0, "interp", 1, 0, "extend", 0.0)]),
}
for _ in range(3):
    fns[name]()
torch.cuda.synchronize()
