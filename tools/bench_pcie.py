"""The host's DMA ceiling: every rank copies page-locked buffers H2D and D2H concurrently (full duplex, the mix of
bench.py's e2e step: 1 part up, 6 parts down), nothing else.  Under torchrun this shows what N GPUs sharing the
sockets' memory controllers / root complexes can move at best — the bound of the e2e numbers at N > 1.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_pcie.py
"""
import json, os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xgcm_b200 import device as xg_device

def main():
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    numa = xg_device.bind_to_gpu_numa(local)
    n = 648_000_000  # one C3 field, fp32
    up_h = torch.empty(n, dtype=torch.float32, pin_memory=True); up_h.zero_()
    dn_h = torch.empty(n, dtype=torch.float32, pin_memory=True); dn_h.zero_()
    up_d = torch.empty(n, dtype=torch.float32, device="cuda"); dn_d = torch.zeros(n, dtype=torch.float32, device="cuda")
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    def step():
        with torch.cuda.stream(s_up):
            up_d.copy_(up_h, non_blocking=True)
        with torch.cuda.stream(s_dn):
            for _ in range(6):
                dn_h.copy_(dn_d, non_blocking=True)
    for _ in range(2): step()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    k = 4
    for _ in range(k): step()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        dt = float(t.item())
        print(json.dumps({"n_gpus": world, "s_per_step": dt / k, "d2h_GBps_per_gpu": 6 * n * 4 * k / dt / 1e9,
                          "h2d_GBps_per_gpu": n * 4 * k / dt / 1e9, "numa_rank0": numa,
                          "note": "pure cudaMemcpyAsync of pinned buffers, 1 field up + 6 fields down per step and rank, all ranks at once"}))
    if world > 1: dist.destroy_process_group()
if __name__ == "__main__":
    main()
