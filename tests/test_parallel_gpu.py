"""Sharded operated axis on real GPUs: NCCL halo plane + fused kernel == single-GPU result.
Needs >= 2 GPUs (gpurun --gpus 2); skipped otherwise."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import stencil as oracle

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from xgcm_b200 import parallel

        rng = np.random.default_rng(99)
        shape = (8, 12, 64)
        glob = rng.random(shape).astype(np.float32)
        metric = (1 + rng.random(shape)).astype(np.float32)
        ok = True
        for axis in range(3):
            start, stop = parallel.shard_bounds(shape[axis], world, rank)
            sl = [slice(None)] * 3
            sl[axis] = slice(start, stop)
            loc = torch.from_numpy(np.ascontiguousarray(glob[tuple(sl)])).cuda()
            mloc = torch.from_numpy(np.ascontiguousarray(metric[tuple(sl)])).cuda()
            for (lo, hi), bc, op in [((1, 0), "periodic", "diff"), ((0, 1), "periodic", "interp"),
                                     ((1, 0), "fill", "interp"), ((0, 1), "extend", "diff")]:
                got = parallel.sharded_stencil2(loc, axis, op, lo, hi, bc, 1.5).cpu().numpy()
                want = oracle.stencil2(op, glob, axis, lo, hi, bc, 1.5)[tuple(sl)]
                ok = ok and np.array_equal(got, want)
                got = parallel.sharded_stencil2(loc, axis, op, lo, hi, bc, 1.5, pre=mloc, post=mloc).cpu().numpy()
                want = oracle.stencil2(op, glob, axis, lo, hi, bc, 1.5, metric, metric)[tuple(sl)]
                ok = ok and np.array_equal(got, want)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_sharded_axis_matches_single_gpu():
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
