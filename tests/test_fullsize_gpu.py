"""BASELINE.json's full sizes (config 3: 75 x 2400 x 3600 fp32, 648 M cells) and a > 2^31-element
field: the oracle cannot hold these in seconds, so parity is checked through size-independent
properties — any sub-block of the full-size result must equal the oracle applied to the matching
input sub-block (plus its halo), bit for bit."""

import numpy as np
import pytest
import torch

from oracle import stencil as oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _field(shape, seed):
    from xgcm_b200 import ops

    x = torch.empty(shape, dtype=torch.float32, device=DEV)
    return ops.fill_uniform(x, seed)


def _check_axis_blocks(x, out, axis, op, lo, hi, bc, fill, rng, nblocks=12):
    """Compare random full lines along `axis` (all cells of the operated axis, a small window of the
    other dims) against the oracle."""
    shape = x.shape
    for _ in range(nblocks):
        sl = []
        for d, n in enumerate(shape):
            if d == axis:
                sl.append(slice(None))
            else:
                w = min(n, 5)
                s0 = int(rng.integers(0, n - w + 1))
                sl.append(slice(s0, s0 + w))
        a = x[tuple(sl)].cpu().numpy()
        want = oracle.stencil2(op, a, axis, lo, hi, bc if (lo or hi) else None, fill)
        np.testing.assert_array_equal(out[tuple(sl)].cpu().numpy(), want)


def test_config3_full_size_stencils():
    from xgcm_b200 import ops

    shape = (75, 2400, 3600)
    x = _field(shape, 0xC0FFEE)
    rng = np.random.default_rng(0)
    for axis, (lo, hi), bc, fill in [(2, (1, 0), "periodic", 0.0), (1, (0, 1), "fill", 1.5), (0, (1, 0), "extend", 0.0),
                                     (0, (1, 1), "fill", 0.0), (2, (0, 0), None, 0.0), (1, (1, 1), "periodic", 0.0)]:
        for op in ("diff", "interp"):
            out = ops.stencil2(x, axis, op, lo, hi, bc, fill)
            _check_axis_blocks(x, out, axis, op, lo, hi, bc, fill, rng)
            del out


def test_config3_full_size_integrate_cumsum_transform():
    from xgcm_b200 import ops

    shape = (75, 2400, 3600)
    x = _field(shape, 7)
    rng = np.random.default_rng(1)
    dz = torch.from_numpy((10 * 1.05 ** np.arange(75)).astype(np.float32)).to(DEV).reshape(75, 1, 1)
    got = ops.wreduce(x, 0, dz, "sum")
    for _ in range(8):
        j, i = int(rng.integers(0, 2396)), int(rng.integers(0, 3596))
        a = x[:, j:j + 4, i:i + 4].cpu().numpy()
        np.testing.assert_array_equal(got[j:j + 4, i:i + 4].cpu().numpy(), oracle.wreduce(a, dz.cpu().numpy(), 0))
    for axis in (0, 1, 2):
        c = ops.cumscan(x, axis, False, "drop_last", 1, 0, "fill", 0.0)
        for _ in range(6):
            sl = [slice(int(s0), int(s0) + 3) for s0 in (rng.integers(0, n - 3) for n in shape)]
            sl[axis] = slice(None)
            a = x[tuple(sl)].cpu().numpy()
            np.testing.assert_array_equal(c[tuple(sl)].cpu().numpy(), oracle.cumscan(a, axis, False, "drop_last", 1, 0, "fill", 0.0))
        del c
    depth = torch.cumsum(dz.reshape(-1), 0)
    levels = torch.linspace(float(depth[0]) - 5, float(depth[-1]) + 5, 100, device=DEV)
    t = ops.vinterp_linear(x, depth.reshape(-1, 1, 1), levels, 0, True)
    assert t.shape == (2400, 3600, 100)
    for _ in range(6):
        j, i = int(rng.integers(0, 2396)), int(rng.integers(0, 3596))
        a = x[:, j:j + 4, i:i + 4].cpu().numpy()
        want = oracle.vinterp_linear(a, depth.cpu().numpy().reshape(-1, 1, 1) * np.ones((1, 4, 4), np.float32), levels.cpu().numpy(), 0, True)
        np.testing.assert_array_equal(t[j:j + 4, i:i + 4].cpu().numpy(), want)


def test_more_than_2_31_elements():
    """64-bit indexing: 2.42 G cells (9.7 GB) per field."""
    from xgcm_b200 import ops

    shape = (9, 16384, 16400)
    assert np.prod(shape) > 2**31
    x = _field(shape, 3)
    rng = np.random.default_rng(2)
    for axis, (lo, hi), bc in [(2, (1, 0), "periodic"), (1, (0, 1), "extend"), (0, (1, 0), "fill")]:
        out = ops.stencil2(x, axis, "diff", lo, hi, bc, 2.0)
        _check_axis_blocks(x, out, axis, "diff", lo, hi, bc, 2.0, rng, nblocks=6)
        # the far corner of the array (flat offsets beyond 2^31)
        sl = [slice(n - 3, n) for n in shape]
        sl[axis] = slice(None)
        a = x[tuple(sl)].cpu().numpy()
        np.testing.assert_array_equal(out[tuple(sl)].cpu().numpy(), oracle.stencil2("diff", a, axis, lo, hi, bc, 2.0))
        del out
    s = ops.wreduce(x, 0, None, "sum")
    np.testing.assert_array_equal(s[-2:, -5:].cpu().numpy(), x[:, -2:, -5:].cpu().numpy().sum(axis=0))
    del s
    c = ops.cumscan(x, 2)
    np.testing.assert_array_equal(c[-1, -1, :].cpu().numpy(), np.cumsum(x[-1, -1, :].cpu().numpy()))


def test_cubed_sphere_full_size_operators():
    """A cubed sphere at production size (50 levels x 6 faces x 1020^2 fp32, 1.25 GB, device
    resident): every operator result on sampled levels must equal the oracle's face-connection
    padding + pairwise kernel on those levels — seams, rotations and all — bit for bit."""
    import xgcm_b200 as xg
    from oracle import faces as oracle_faces
    from test_faces_gpu import AXES, COORDS, CUBED_SPHERE

    nz, nf, n = 50, 6, 1020
    f = _field((nz, nf, n, n), 11)
    ds = xg.Dataset(coords={"z": np.arange(nz), "face": np.arange(nf), "y": np.arange(n) + 0.0,
                            "yl": np.arange(n) - 0.5, "x": np.arange(n) + 0.0, "xl": np.arange(n) - 0.5})
    grid = xg.Grid(ds, coords=COORDS, face_connections=CUBED_SPHERE, autoparse_metadata=False)
    da = xg.DataArray(f, dims=("z", "face", "y", "x"))
    levels = [0, 17, nz - 1]
    host = f[levels].cpu().numpy()
    for op, ax in (("diff", "X"), ("interp", "Y"), ("max", "X")):
        out = getattr(grid, op)(da, ax)
        assert out.is_device and out.shape == (nz, nf, n, n)
        padded = oracle_faces.pad_face_connections(
            host, ("z", "face", "y", "x"), AXES, "face", CUBED_SPHERE["face"], {ax: (1, 0)},
            {"X": None, "Y": None}, {"X": 0.0, "Y": 0.0})
        k = 3 if ax == "X" else 2
        want = np.moveaxis(oracle.KERNELS[op](np.moveaxis(padded, k, -1)), -1, k)
        np.testing.assert_array_equal(out.data[levels].cpu().numpy(), want.astype(np.float32))


def test_config4_sampled_steps():
    """BASELINE configs[3] (x365 time steps, time-sharded): the field of ANY global step is
    regenerated on the device from the step index (tools/bench_c4.py uses the same generator), so
    parity is checked on sampled steps — first, middle, last of the year — exactly as a rank that
    owns them would compute them: Grid.diff / Grid.interp on X (periodic), Y (fill), Z (extend),
    random full lines of every result against the oracle, bit for bit."""
    import xgcm_b200 as xg
    from xgcm_b200 import ops

    shape = (75, 2400, 3600)
    nz, ny, nx = shape
    ds = xg.Dataset(coords={"Z": np.arange(nz) + 0.5, "Zl": np.arange(nz) + 0.0, "YC": np.arange(ny) + 0.5,
                            "YG": np.arange(ny) + 0.0, "XC": np.arange(nx) + 0.5, "XG": np.arange(nx) + 0.0})
    grid = xg.Grid(ds, coords={"X": {"center": "XC", "left": "XG"}, "Y": {"center": "YC", "left": "YG"},
                               "Z": {"center": "Z", "left": "Zl"}},
                   padding={"X": "periodic", "Y": "fill", "Z": "extend"}, autoparse_metadata=False)
    x = torch.empty(shape, dtype=torch.float32, device=DEV)
    da = xg.DataArray(x, dims=("Z", "YC", "XC"))
    rng = np.random.default_rng(4)
    for t in (0, 182, 364):
        ops.fill_uniform(x, 0xC0FFEE, offset=t * x.numel())
        for ax, k, bc in (("X", 2, "periodic"), ("Y", 1, "fill"), ("Z", 0, "extend")):
            for op in ("diff", "interp"):
                out = getattr(grid, op)(da, ax)
                _check_axis_blocks(x, out.data, k, op, 1, 0, bc, 0.0, rng, nblocks=4)
