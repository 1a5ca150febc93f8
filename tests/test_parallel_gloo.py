"""N>1 host logic on CPU: world_size-2 gloo groups exercise the shard plan and the halo exchange
(the plumbing around the kernels).  Values the kernel would produce are checked with the oracle."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import stencil as oracle


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, axis, return_dict):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from xgcm_b200 import parallel

        rng = np.random.default_rng(1234)  # same global field on every rank
        shape = [6, 8, 10]
        glob = rng.random(shape)
        metric = 1.0 + rng.random(shape)
        start, stop = parallel.shard_bounds(shape[axis], world, rank)
        sl = [slice(None)] * 3
        sl[axis] = slice(start, stop)
        local = torch.from_numpy(np.ascontiguousarray(glob[tuple(sl)]))
        mloc = torch.from_numpy(np.ascontiguousarray(metric[tuple(sl)]))
        ok = True
        for (lo, hi), bc in [((1, 0), "periodic"), ((0, 1), "periodic"), ((1, 0), "fill"), ((0, 1), "extend")]:
            for weighted in (False, True):
                kw = {}
                if weighted:
                    kw = dict(edge_scale_lo=mloc.select(axis, 0), edge_scale_hi=mloc.select(axis, mloc.shape[axis] - 1))
                hl, hh = parallel.exchange_halo(local, axis, lo, hi, bc == "periodic", **kw)
                src = glob * metric if weighted else glob
                # what the reference computes globally (pad -> kernel), then this rank's block of it
                want = oracle.stencil2("diff", src, axis, lo, hi, bc, 2.5)[tuple(sl)]
                # what the kernel computes locally: the received plane is the halo, the exterior
                # boundary (edge ranks, non-periodic) is synthesised from `bc`
                loc = (local * mloc).numpy() if weighted else local.numpy()
                parts = []
                if lo:
                    if hl is not None:
                        parts.append(np.expand_dims(hl.numpy(), axis))
                    else:
                        assert rank == 0 and bc != "periodic"
                        parts.append(np.take(oracle.pad_axis(loc, axis, 1, 0, bc, 2.5), [0], axis=axis))
                parts.append(loc)
                if hi:
                    if hh is not None:
                        parts.append(np.expand_dims(hh.numpy(), axis))
                    else:
                        assert rank == world - 1 and bc != "periodic"
                        parts.append(np.take(oracle.pad_axis(loc, axis, 0, 1, bc, 2.5), [-1], axis=axis))
                padded = np.concatenate(parts, axis=axis)
                got = np.moveaxis(oracle.diff_forward(np.moveaxis(padded, axis, -1)), -1, axis)
                ok = ok and np.array_equal(got, want)
        # time-shard plan: blocks are contiguous, disjoint and cover everything
        gathered = [None] * world
        dist.all_gather_object(gathered, parallel.shard_bounds(365, world, rank))
        flat = [i for a, b in gathered for i in range(a, b)]
        ok = ok and flat == list(range(365))
        return_dict[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_halo_exchange_world2_gloo(axis):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), axis, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_shard_bounds_and_errors():
    from xgcm_b200 import parallel

    assert [parallel.shard_bounds(365, 8, r) for r in range(8)] == [
        (0, 46), (46, 92), (92, 138), (138, 184), (184, 230), (230, 276), (276, 322), (322, 365)]
    assert parallel.shard_bounds(3, 8, 7) == (3, 3)  # more ranks than items: empty tail blocks
    with pytest.raises(ValueError):
        parallel.shard_bounds(10, 2, 2)
    with pytest.raises(NotImplementedError):
        parallel.sharded_stencil2(torch.zeros(4, 4), 0, "diff", 1, 1, "fill")  # grid_ufunc.py:1136-1159


# ---------------------------------------------------------------------------------------------
# faces of a connected grid sharded across ranks (cubed sphere: 6 faces over 2 or 3 ranks)
def _faces_worker(rank, world, port, return_dict):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys

        sys.path.insert(0, os.path.dirname(__file__))
        import xgcm_b200 as xg
        from _mock_backend import install
        from oracle import faces as oracle_faces
        from test_faces_gpu import AXES, COORDS, CUBED_SPHERE, N
        from xgcm_b200 import parallel

        class _Patch:  # the two lines of pytest's monkeypatch the mock needs
            def setattr(self, obj, name, value):
                setattr(obj, name, value)

        install(_Patch())
        rng = np.random.default_rng(99)  # same global fields on every rank
        nz = 3
        glob = {"c": rng.random((nz, 6, N, N)), "u": rng.random((nz, 6, N, N)), "v": rng.random((nz, 6, N, N))}
        gdims = {"c": ("z", "face", "y", "x"), "u": ("z", "face", "xl", "y"), "v": ("z", "face", "x", "yl")}
        ds = xg.Dataset(coords={"z": np.arange(nz), "face": np.arange(6), "y": np.arange(N) + 0.0,
                                "yl": np.arange(N) - 0.5, "x": np.arange(N) + 0.0, "xl": np.arange(N) - 0.5})
        grid = xg.Grid(ds, coords=COORDS, face_connections=CUBED_SPHERE, autoparse_metadata=False)
        start, stop = parallel.shard_bounds(6, world, rank)

        def local(name):
            return xg.DataArray(np.ascontiguousarray(glob[name][:, start:stop]), dims=gdims[name])

        ok = True
        cases = [("c", None, None, "X", 1, 0), ("c", None, None, "Y", 1, 0), ("c", None, None, "Y", 0, 1),
                 ("u", "X", "v", "X", 0, 1), ("v", "Y", "u", "Y", 0, 1), ("u", "X", "v", "Y", 1, 0)]
        for name, vax, pname, ax, lo, hi in cases:
            if vax is None:
                out = parallel.sharded_connected_stencil2(grid, local(name), ax, "diff", lo, hi)
            else:
                out = parallel.sharded_connected_stencil2(
                    grid, {vax: local(name)}, ax, "diff", lo, hi,
                    other_component_local={("Y" if vax == "X" else "X"): local(pname)})
            padded = oracle_faces.pad_face_connections(
                glob[name], gdims[name], AXES, "face", CUBED_SPHERE["face"], {ax: (lo, hi)},
                {"X": None, "Y": None}, {"X": 0.0, "Y": 0.0}, vector_axis=vax,
                partner=None if pname is None else glob[pname],
                partner_dims=None if pname is None else gdims[pname])
            k = gdims[name].index([d for d in AXES[ax] if d in gdims[name]][0])
            want = np.moveaxis(oracle.diff_forward(np.moveaxis(padded, k, -1)), -1, k)[:, start:stop]
            ok = ok and np.array_equal(out.numpy(), want)
        return_dict[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_face_sharded_connected_stencil_gloo(world):
    """Every rank holds a block of cubed-sphere faces; rims that cross a block boundary are built
    by the owner of the neighbour face and exchanged; the result equals the global oracle."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_faces_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {r: True for r in range(world)}
