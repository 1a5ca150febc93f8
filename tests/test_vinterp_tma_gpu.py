"""The bulk-async (TMA) staged shared-theta transform kernel (csrc/xg_vinterp_tma.cu) against the oracle.

Every case asserts that the TMA kernel — not its fallback — served the call (xg_last_launch), and that the
result is bit-identical to the reference port (oracle.vinterp_linear <- xgcm/transform.py:15-41).
"""

import numpy as np
import pytest
import torch

from oracle import stencil as oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TMA = "xg_vinterp_linear(shared, tma)"


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _run(phi, theta, target, axis, mask=True, bypass=False, log=False, expect_tma=True):
    from xgcm_b200 import _capi, ops

    got = ops.vinterp_linear(_t(phi), _t(theta), _t(target), axis, mask, bypass, log).cpu().numpy()
    if expect_tma:
        assert _capi.last_launch() == TMA, _capi.last_launch()
    else:
        assert _capi.last_launch() != TMA
    return got


def _shared_theta(n, dtype, rng, shape, axis):
    th = np.cumsum(0.1 + rng.random(n)).astype(dtype)
    b = [1] * len(shape)
    b[axis] = n
    return th.reshape(b)


@pytest.mark.parametrize("cpl", ["1", "2"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape,axis", [((20, 64), 0), ((75, 7, 44), 0), ((3, 25, 40), 1), ((2, 31, 5, 36), 1),
                                        ((9, 3, 1028), 0), ((300, 4, 32), 0)])
def test_tma_matches_oracle(monkeypatch, cpl, dtype, shape, axis):
    monkeypatch.setenv("XG_VINTERP_CPL", cpl)
    rng = np.random.default_rng(5)
    n = shape[axis]
    phi = rng.standard_normal(shape).astype(dtype)
    phi[rng.random(shape) < 0.03] = np.nan
    phi[rng.random(shape) < 0.01] = np.inf
    phi[rng.random(shape) < 0.01] = -np.inf
    phi[rng.random(shape) < 0.02] = -0.0
    phi[rng.random(shape) < 0.02] = 0.0
    theta = _shared_theta(n, dtype, rng, shape, axis)
    lo, hi = float(theta.min()), float(theta.max())
    for m in (1, 4, 7, 33, 100):
        target = np.linspace(lo - 0.3, hi + 0.3, m).astype(dtype)
        if m >= 7:
            target[2] = np.nan
            target[3] = theta.reshape(-1)[1]       # exact hit on a node
            target[4] = theta.reshape(-1)[n - 1]   # exact hit on the last node
        for th in (theta, np.flip(theta, axis=axis).copy()):  # increasing and decreasing theta (flip)
            for mask, bypass in ((True, False), (False, False)):
                want = oracle.vinterp_linear(phi, np.broadcast_to(th, shape), target, axis, mask, bypass)
                got = _run(phi, th, target, axis, mask, bypass)
                assert got.shape == want.shape and got.dtype == want.dtype
                np.testing.assert_array_equal(got, want, err_msg=f"m={m} mask={mask}")
                # the sign of zeros too
                np.testing.assert_array_equal(np.signbit(got), np.signbit(want))
        # targets in descending and in shuffled order
        for tg in (target[::-1].copy(), rng.permutation(target)):
            want = oracle.vinterp_linear(phi, np.broadcast_to(theta, shape), tg, axis, True, False)
            np.testing.assert_array_equal(_run(phi, theta, tg, axis), want)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_tma_unsorted_and_nan_theta_replays_numpy_search(dtype):
    """theta with NaNs / out of order: the plan is the literal binary_search_with_guess replay and every
    slope takes the true division."""
    rng = np.random.default_rng(8)
    shape = (24, 6, 64)
    phi = rng.standard_normal(shape).astype(dtype)
    theta = np.cumsum(0.1 + rng.random(24)).astype(dtype)
    theta[[5, 6]] = theta[[6, 5]]
    theta[11] = np.nan
    theta[17] = theta[16]  # zero-width interval
    theta = theta.reshape(24, 1, 1)
    target = np.linspace(0.0, float(np.nanmax(theta)) + 0.2, 40).astype(dtype)
    for mask in (True, False):
        want = oracle.vinterp_linear(phi, np.broadcast_to(theta, shape), target, 0, mask, False)
        np.testing.assert_array_equal(_run(phi, theta, target, 0, mask), want)
    want = oracle.vinterp_linear(phi, np.broadcast_to(theta, shape), target, 0, True, True)
    np.testing.assert_array_equal(_run(phi, theta, target, 0, True, True), want)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_tma_slope_division_is_correctly_rounded(dtype):
    rng = np.random.default_rng(77)
    ncol, n = 200_000, 5
    phi = (rng.standard_normal((n, ncol)) * 10.0 ** rng.integers(-12, 12, size=(n, ncol))).astype(dtype)
    phi[:, :50] = 0.0
    phi[2, 50:80] = np.nan
    phi[1, 100:200] = phi[2, 100:200]  # dy == 0
    phi[1, 200:300] = -0.0
    phi[2, 200:300] = 0.0
    if dtype == np.float64:
        phi[3, 300:400] = 1e300
        phi[4, 300:400] = -1e300  # dy overflows
        phi[3, 400:500] = 5e-324
    theta = np.array([0.1, 0.7, 1.9, 3.0000001, 7.3], dtype=dtype).reshape(n, 1)
    target = np.array([0.05, 0.1, 0.33, 0.7000001, 1.0, 2.5, 3.0, 3.5, 7.0, 7.3, 9.0, 0.2], dtype=dtype)
    want = oracle.vinterp_linear(phi, np.broadcast_to(theta, phi.shape), target, 0, True)
    got = _run(phi, theta, target, 0)
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(np.signbit(got), np.signbit(want))


def test_tma_log_method_and_fallbacks():
    rng = np.random.default_rng(9)
    shape = (30, 5, 64)
    phi = rng.random(shape).astype(np.float64)
    theta = np.cumsum(1.0 + rng.random(30)).reshape(30, 1, 1)
    target = np.linspace(0.5, float(theta.max()) + 1, 16)
    want = oracle.vinterp_linear(phi, np.broadcast_to(theta, shape), target, 0, True, False, True)
    np.testing.assert_allclose(_run(phi, theta, target, 0, log=True), want, rtol=1e-12, atol=1e-12, equal_nan=True)
    # layouts the TMA descriptor cannot express go to the plain kernel (and still match)
    phi32 = rng.random((20, 3, 50)).astype(np.float32)  # 50 * 4 B is not a multiple of 16
    th32 = np.cumsum(0.1 + rng.random(20)).astype(np.float32).reshape(20, 1, 1)
    tg32 = np.linspace(0, float(th32.max()), 9).astype(np.float32)
    want = oracle.vinterp_linear(phi32, np.broadcast_to(th32, phi32.shape), tg32, 0, True)
    np.testing.assert_array_equal(_run(phi32, th32, tg32, 0, expect_tma=False), want)


def test_tma_full_size_c5_samples():
    """BASELINE configs[4]: 75 x 2400 x 3600 fp32 -> 100 levels; sampled columns against the oracle."""
    from xgcm_b200 import _capi, ops

    nz, ny, nx, m = 75, 2400, 3600, 100
    x = torch.empty((nz, ny, nx), dtype=torch.float32, device=DEV)
    ops.fill_uniform(x, 0xC0FFEE)
    dz = 10 * 1.05 ** np.arange(nz)
    depth = (np.cumsum(dz) - dz / 2).astype(np.float32)
    levels = np.linspace(depth[0] - 5, depth[-1] + 5, m).astype(np.float32)
    out = ops.vinterp_linear(x, _t(depth.reshape(-1, 1, 1)), _t(levels), 0, True)
    assert _capi.last_launch() == TMA
    assert out.shape == (ny, nx, m)
    rng = np.random.default_rng(3)
    for _ in range(6):
        j0, i0 = int(rng.integers(0, ny - 2)), int(rng.integers(0, nx - 40))
        a = x[:, j0:j0 + 2, i0:i0 + 40].cpu().numpy()
        want = oracle.vinterp_linear(a, depth.reshape(-1, 1, 1) * np.ones((1, 2, 40), np.float32), levels, 0, True)
        np.testing.assert_array_equal(out[j0:j0 + 2, i0:i0 + 40].cpu().numpy(), want)
    # the very last tile of the field
    a = x[:, -1:, -64:].cpu().numpy()
    want = oracle.vinterp_linear(a, depth.reshape(-1, 1, 1) * np.ones((1, 1, 64), np.float32), levels, 0, True)
    np.testing.assert_array_equal(out[-1:, -64:].cpu().numpy(), want)


# ----------------------------------------------------------------------------- theta FIELD (columns kernel)
COLS_TMA = "xg_vinterp_linear(columns, tma)"


def _theta_field(shape, axis, dtype, rng, decreasing_frac=0.3, nan_frac=0.0, unsorted_frac=0.0):
    inc = np.cumsum(0.1 + rng.random(shape), axis=axis).astype(dtype)
    flip_shape = [s if d != axis else 1 for d, s in enumerate(shape)]
    th = np.where(rng.random(flip_shape) < decreasing_frac, np.flip(inc, axis=axis), inc).astype(dtype)
    if unsorted_frac:
        sw = rng.random(flip_shape) < unsorted_frac
        a, b = np.take(th, [3], axis=axis), np.take(th, [5], axis=axis)
        idx3 = [slice(None)] * len(shape)
        idx5 = [slice(None)] * len(shape)
        idx3[axis], idx5[axis] = slice(3, 4), slice(5, 6)
        th[tuple(idx3)] = np.where(sw, b, a)
        th[tuple(idx5)] = np.where(sw, a, b)
    if nan_frac:
        th[rng.random(shape) < nan_frac] = np.nan
    return th


@pytest.mark.parametrize("wt", ["0", "1", "3"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape,axis", [((20, 64), 0), ((75, 5, 44), 0), ((3, 25, 40), 1), ((2, 31, 5, 36), 1), ((9, 2, 1028), 0)])
def test_columns_tma_matches_oracle(monkeypatch, wt, dtype, shape, axis):
    from xgcm_b200 import _capi, ops

    monkeypatch.setenv("XG_VINTERP_WT", wt)
    rng = np.random.default_rng(17)
    phi = rng.standard_normal(shape).astype(dtype)
    phi[rng.random(shape) < 0.02] = np.nan
    for kind in ("sorted", "nan", "unsorted", "allnan_cols"):
        theta = _theta_field(shape, axis, dtype, rng, nan_frac=0.04 if kind == "nan" else 0.0,
                             unsorted_frac=0.3 if kind == "unsorted" else 0.0)
        if kind == "allnan_cols":
            col_shape = [s if d != axis else 1 for d, s in enumerate(shape)]
            theta = np.where(rng.random(col_shape) < 0.2, np.nan, theta).astype(dtype)
        lo, hi = float(np.nanmin(theta)), float(np.nanmax(theta))
        for m in (1, 6, 33, 100):
            target = np.linspace(lo - 0.3, hi + 0.3, m).astype(dtype)
            if m >= 6:
                target[2] = np.nan
                target[3] = theta.reshape(-1)[0] if not np.isnan(theta.reshape(-1)[0]) else target[3]
            for tg in (target, target[::-1].copy()):
                for mask, bypass in ((True, False), (False, False), (True, True)):
                    with np.errstate(invalid="ignore"):
                        want = oracle.vinterp_linear(phi, theta, tg, axis, mask, bypass)
                    got = ops.vinterp_linear(_t(phi), _t(theta), _t(tg), axis, mask, bypass).cpu().numpy()
                    assert _capi.last_launch() == COLS_TMA, _capi.last_launch()
                    np.testing.assert_array_equal(got, want, err_msg=f"{kind} m={m} mask={mask} bypass={bypass}")


def test_columns_tma_log_and_full_size_samples():
    from xgcm_b200 import _capi, ops

    rng = np.random.default_rng(23)
    shape = (30, 5, 64)
    phi = rng.random(shape)
    theta = np.cumsum(1.0 + rng.random(shape), axis=0)
    target = np.linspace(0.5, float(theta.max()) + 1, 16)
    want = oracle.vinterp_linear(phi, theta, target, 0, True, False, True)
    got = ops.vinterp_linear(_t(phi), _t(theta), _t(target), 0, True, False, True).cpu().numpy()
    assert _capi.last_launch() == COLS_TMA
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12, equal_nan=True)
    # C5-sized field with a per-column theta
    nz, ny, nx, m = 75, 2400, 3600, 100
    x = torch.empty((nz, ny, nx), dtype=torch.float32, device=DEV)
    ops.fill_uniform(x, 0xC0FFEE)
    th = torch.cumsum(x + 0.5, 0)
    levels = torch.linspace(0.0, float(th.max()) + 1, m, device=DEV)
    out = ops.vinterp_linear(x, th, levels, 0, True)
    assert _capi.last_launch() == COLS_TMA
    for _ in range(4):
        j0, i0 = int(rng.integers(0, ny - 2)), int(rng.integers(0, nx - 40))
        a = x[:, j0:j0 + 2, i0:i0 + 40].cpu().numpy()
        t = th[:, j0:j0 + 2, i0:i0 + 40].cpu().numpy()
        want = oracle.vinterp_linear(a, t, levels.cpu().numpy(), 0, True)
        np.testing.assert_array_equal(out[j0:j0 + 2, i0:i0 + 40].cpu().numpy(), want)
