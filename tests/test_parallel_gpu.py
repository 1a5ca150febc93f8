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
        comm = parallel.Communicator()  # the library's own NCCL communicator (xg_comm_init)
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
                # the fused C-ABI call: pack kernel + one NCCL group on a side stream + edge fix-up
                got = parallel.sharded_stencil2(loc, axis, op, lo, hi, bc, 1.5, comm=comm).cpu().numpy()
                want = oracle.stencil2(op, glob, axis, lo, hi, bc, 1.5)[tuple(sl)]
                ok = ok and np.array_equal(got, want)
                got = parallel.sharded_stencil2(loc, axis, op, lo, hi, bc, 1.5, pre=mloc, post=mloc, comm=comm).cpu().numpy()
                want = oracle.stencil2(op, glob, axis, lo, hi, bc, 1.5, metric, metric)[tuple(sl)]
                ok = ok and np.array_equal(got, want)
        # the bare ring step (xg_halo_exchange): my first plane down, my last plane up
        from xgcm_b200 import _capi
        lib = _capi.load()
        mine = torch.full((2, 1000), float(rank), device="cuda")
        got_lo = torch.full((1000,), -1.0, device="cuda")
        got_hi = torch.full((1000,), -1.0, device="cuda")
        rc = lib.xg_halo_exchange(comm._handle, mine[0].data_ptr(), mine[1].data_ptr(), got_lo.data_ptr(),
                                  got_hi.data_ptr(), 4000, 1, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ok = ok and rc == 0 and bool((got_lo == (rank - 1) % world).all()) and bool((got_hi == (rank + 1) % world).all())
        comm.close()
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


def _faces_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import sys

        sys.path.insert(0, os.path.dirname(__file__))
        import xgcm_b200 as xg
        from oracle import faces as oracle_faces
        from test_faces_gpu import AXES, COORDS, CUBED_SPHERE
        from xgcm_b200 import parallel

        n, nz = 48, 4
        rng = np.random.default_rng(99)
        glob = {k: rng.random((nz, 6, n, n)).astype(np.float32) for k in ("c", "u", "v")}
        gdims = {"c": ("z", "face", "y", "x"), "u": ("z", "face", "xl", "y"), "v": ("z", "face", "x", "yl")}
        ds = xg.Dataset(coords={"z": np.arange(nz), "face": np.arange(6), "y": np.arange(n) + 0.0,
                                "yl": np.arange(n) - 0.5, "x": np.arange(n) + 0.0, "xl": np.arange(n) - 0.5})
        grid = xg.Grid(ds, coords=COORDS, face_connections=CUBED_SPHERE, autoparse_metadata=False,
                       device=f"cuda:{rank}")
        start, stop = parallel.shard_bounds(6, world, rank)

        def local(name):
            t = torch.from_numpy(np.ascontiguousarray(glob[name][:, start:stop])).cuda()
            return xg.DataArray(t, dims=gdims[name])

        ok = True
        cases = [("c", None, None, "X", 1, 0, "diff"), ("c", None, None, "Y", 0, 1, "interp"),
                 ("u", "X", "v", "X", 0, 1, "interp"), ("v", "Y", "u", "Y", 0, 1, "diff"),
                 ("u", "X", "v", "Y", 1, 0, "max")]
        for name, vax, pname, ax, lo, hi, op in cases:
            if vax is None:
                out = parallel.sharded_connected_stencil2(grid, local(name), ax, op, lo, hi)
            else:
                out = parallel.sharded_connected_stencil2(
                    grid, {vax: local(name)}, ax, op, lo, hi,
                    other_component_local={("Y" if vax == "X" else "X"): local(pname)})
            padded = oracle_faces.pad_face_connections(
                glob[name], gdims[name], AXES, "face", CUBED_SPHERE["face"], {ax: (lo, hi)},
                {"X": None, "Y": None}, {"X": 0.0, "Y": 0.0}, vector_axis=vax,
                partner=None if pname is None else glob[pname],
                partner_dims=None if pname is None else gdims[pname])
            k = gdims[name].index([d for d in AXES[ax] if d in gdims[name]][0])
            want = np.moveaxis(oracle.KERNELS[op](np.moveaxis(padded, k, -1)), -1, k)[:, start:stop]
            ok = ok and np.array_equal(out.cpu().numpy(), want.astype(np.float32))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_face_sharded_cubed_sphere_matches_oracle():
    """Cubed-sphere faces split over 2 GPUs: rims crossing the split travel over NCCL, already
    rotated by their owner; scalars and vector components, bit-exact against the global oracle."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_faces_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
