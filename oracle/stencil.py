"""CPU oracle: numpy restatement of xgcm's grid-ufunc stencil hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``xgcm_b200/`` may import this; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs do.

Each function restates, with the very numpy calls xarray would make, what the
reference does around / inside its kernels (paths relative to the xgcm tree,
snapshot 052b033a).  Pinning: ``tests/test_oracle_golden.py`` checks this file
against (i) the known answers transcribed from the reference's own tests and
docs and (ii) ``tests/golden/*.npz`` produced by running the reference's
unmodified ``xgcm.gridops`` / ``xgcm.transform`` kernels (``oracle/make_golden.py``,
via ``oracle/ref_loader.py``).

Unpinned (no reference test, xarray source not in the container): NaN inputs to
cumsum / integrate (xarray ``skipna`` default, restated here as nancumsum /
nansum) and the ``extrapolate`` boundary (absent from the reference snapshot).
"""

from __future__ import annotations

import numpy as np

# xgcm/padding.py:15-19
_PAD_MODE = {"periodic": "wrap", "fill": "constant", "extend": "edge"}


# --- kernels: xgcm/gridops.py ------------------------------------------------
def diff_forward(a):
    """gridops.py:23-24"""
    return a[..., 1:] - a[..., :-1]


def interp_forward(a):
    """gridops.py:76-77"""
    return (a[..., :-1] + a[..., 1:]) / 2.0


def pairwise_forward_min(a):
    """gridops.py:123-126"""
    left, right = a[..., :-1], a[..., 1:]
    return np.min(np.stack([left, right], axis=-1), axis=-1)


def pairwise_forward_max(a):
    """gridops.py:172-175"""
    left, right = a[..., :-1], a[..., 1:]
    return np.max(np.stack([left, right], axis=-1), axis=-1)


KERNELS = {
    "diff": diff_forward,
    "interp": interp_forward,
    "min": pairwise_forward_min,
    "max": pairwise_forward_max,
}

# position pair -> (lo, hi) halo widths, gridops.py:27-65 (same table for all ops)
PADDING_WIDTH = {
    ("center", "left"): (1, 0),
    ("left", "center"): (0, 1),
    ("center", "right"): (0, 1),
    ("right", "center"): (1, 0),
    ("center", "outer"): (1, 1),
    ("outer", "center"): (0, 0),
    ("center", "inner"): (0, 0),
    ("inner", "center"): (1, 1),
}


# --- padding: xgcm/padding.py:575-616 ------------------------------------------
def pad_axis(a, axis, lo, hi, padding, fill_value=0.0):
    """One axis of ``_pad_basic``: ``DataArray.pad`` -> ``np.pad``."""
    if lo == 0 and hi == 0:  # padding.py:592-593
        return a
    if padding is None:  # padding.py:601-608
        raise ValueError("No boundary condition was specified")
    if padding == "extrapolate":
        # NOT in the reference snapshot (removed upstream): pre-0.6 meaning,
        # linear extrapolation of the edge.  Parity unpinned.
        a = np.asarray(a)
        m = np.moveaxis(a, axis, -1)
        parts = []
        if lo:
            nxt = m[..., 1:2] if m.shape[-1] > 1 else m[..., 0:1]
            parts.append(2 * m[..., 0:1] - nxt)
        parts.append(m)
        if hi:
            prv = m[..., -2:-1] if m.shape[-1] > 1 else m[..., -1:]
            parts.append(2 * m[..., -1:] - prv)
        return np.moveaxis(np.concatenate(parts, axis=-1), -1, axis)
    widths = [(0, 0)] * a.ndim
    widths[axis] = (lo, hi)
    mode = _PAD_MODE[padding]
    if mode == "constant":
        return np.pad(a, widths, mode="constant", constant_values=fill_value)
    return np.pad(a, widths, mode=mode)


# --- the fused site: pad -> kernel (+ metric weighting) ----------------------------
def stencil2(op, a, axis, lo, hi, padding, fill_value=0.0, pre=None, post=None):
    """grid.py:806-832 around grid_ufunc.py:905-984.

    ``pre`` / ``post`` are metric arrays already shaped to broadcast (numpy
    rules) against the input / output.
    """
    a = np.asarray(a)
    if pre is not None:
        a = a * pre  # grid.py:806-808
    p = pad_axis(a, axis, lo, hi, padding, fill_value)  # padding.py:615
    moved = np.moveaxis(p, axis, -1)  # xr.apply_ufunc puts the core dim last
    r = KERNELS[op](moved)
    r = np.moveaxis(r, -1, axis)  # grid_ufunc.py:56-103
    if post is not None:
        r = r / post  # grid.py:830-832 / :1578
    return r


# --- cumsum: xgcm/grid.py:1306-1414 ---------------------------------------------------
# (from, to) -> (trim, (pad_lo, pad_hi)) ; grid.py:1326-1383
CUMSUM_TABLE_FWD = {
    ("center", "right"): ("none", (0, 0)),
    ("left", "center"): ("none", (0, 0)),
    ("center", "left"): ("drop_last", (1, 0)),
    ("right", "center"): ("drop_last", (1, 0)),
    ("center", "inner"): ("drop_last", (0, 0)),
    ("outer", "center"): ("drop_last", (0, 0)),
    ("center", "outer"): ("none", (1, 0)),
    ("inner", "center"): ("none", (1, 0)),
}
CUMSUM_TABLE_REV = {
    ("center", "left"): ("none", (0, 0)),
    ("right", "center"): ("none", (0, 0)),
    ("center", "right"): ("drop_first", (0, 1)),
    ("left", "center"): ("drop_first", (0, 1)),
    ("center", "inner"): ("drop_first", (0, 0)),
    ("outer", "center"): ("drop_first", (0, 0)),
    ("center", "outer"): ("none", (0, 1)),
    ("inner", "center"): ("none", (0, 1)),
}


def cumscan(a, axis, reverse=False, trim="none", pad_lo=0, pad_hi=0, padding=None,
            fill_value=0.0, pre=None, post=None, skipna=True):
    a = np.asarray(a)
    if pre is not None:
        a = a * pre  # grid.py:1306-1308
    if reverse:
        a = np.flip(a, axis)  # grid.py:1314-1315
    # DataArray.cumsum: skipna default for floats -> nancumsum (sequential order)
    c = np.nancumsum(a, axis=axis) if skipna else np.cumsum(a, axis=axis)
    if c.dtype != a.dtype:
        c = c.astype(a.dtype)
    if reverse:
        c = np.flip(c, axis)
    sl = [slice(None)] * a.ndim
    if trim == "drop_last":
        sl[axis] = slice(0, -1)
    elif trim == "drop_first":
        sl[axis] = slice(1, None)
    c = c[tuple(sl)]
    c = pad_axis(c, axis, pad_lo, pad_hi, padding, fill_value)  # grid.py:1385-1391
    if post is not None:
        c = c / post  # grid.py:1411-1414
    return c


# --- integrate / average: xgcm/grid.py:1598-1605, :1680-1685 --------------------------------
def wreduce(a, w, axis, mode="sum", skipna=True):
    a = np.asarray(a)
    if mode == "sum":
        prod = a if w is None else a * w  # grid.py:1599
        return (np.nansum if skipna else np.sum)(prod, axis=axis)
    # xarray Weighted.mean: sum(da*w over valid) / sum(w over valid), NaN where that is 0; with
    # skipna=False nothing is masked and a NaN cell makes its line NaN (xgcm/grid.py:1680-1685 forwards
    # **kwargs to da.weighted(w).mean)
    wb = np.ones_like(a) if w is None else np.broadcast_to(w, a.shape).astype(a.dtype)
    valid = ~np.isnan(a) if skipna else np.ones(a.shape, bool)
    num = np.sum(np.where(valid, a, 0) * wb, axis=axis)
    den = np.sum(np.where(valid, wb, 0), axis=axis)
    if mode == "wvalid":
        return den.astype(a.dtype)
    with np.errstate(invalid="ignore", divide="ignore"):
        out = num / den
    return np.where(den != 0, out, np.nan).astype(a.dtype)


# --- vertical transform: xgcm/transform.py:15-85 ----------------------------------------------
def _interp_column(phi, theta, target, mask_edges, bypass_checks):
    """transform.py:23-41 for one column; np.interp works in float64 like numba's."""
    if not bypass_checks:
        t = theta[~np.isnan(theta)]
        # an all-NaN column indexes an empty array in the reference (numba reads out of
        # bounds: undefined); the restatement leaves such a column unflipped
        if t.size and t[-1] < t[0]:
            theta = theta[::-1]
            phi = phi[::-1]
    out = np.interp(target, theta, phi)
    if mask_edges:
        tmax = np.nanmax(theta)
        tmin = np.nanmin(theta)
        out = np.where((target < tmin) | (target > tmax), np.nan, out)
    return out


def vinterp_linear(phi, theta, target, axis=-1, mask_edges=False, bypass_checks=False,
                   logarithmic=False):
    """interp_1d_linear over columns along ``axis``; new dim appended LAST
    (transform.py:233-249).  ``theta`` broadcasts against ``phi``."""
    phi = np.asarray(phi)
    target = np.asarray(target)
    # numba gufunc loop resolution (transform.py:15-22): the float32 loop only when
    # all three operands are float32, otherwise everything is cast to float64.
    if not (phi.dtype == np.float32 and np.asarray(theta).dtype == np.float32
            and target.dtype == np.float32):
        phi = phi.astype(np.float64)
        theta = np.asarray(theta).astype(np.float64)
        target = target.astype(np.float64)
    dtype = phi.dtype
    theta = np.broadcast_to(np.asarray(theta), phi.shape) if np.ndim(theta) == phi.ndim else theta
    if np.ndim(theta) == 1:
        shape = [1] * phi.ndim
        shape[axis] = -1
        theta = np.broadcast_to(np.asarray(theta).reshape(shape), phi.shape)
    if logarithmic:  # transform.py:82-84
        with np.errstate(invalid="ignore", divide="ignore"):
            theta = np.log(theta)
            target = np.log(target)
    pm = np.moveaxis(phi, axis, -1)
    tm = np.moveaxis(theta, axis, -1)
    cols = pm.reshape(-1, pm.shape[-1])
    tcols = tm.reshape(-1, tm.shape[-1])
    out = np.empty((cols.shape[0], target.shape[0]), dtype=dtype)
    with np.errstate(invalid="ignore"):
        for c in range(cols.shape[0]):
            out[c] = _interp_column(cols[c], tcols[c].astype(dtype), target.astype(dtype),
                                    mask_edges, bypass_checks)
    return out.reshape(pm.shape[:-1] + (target.shape[0],))


# --- conservative vertical transform: xgcm/transform.py:88-191 ---------------------------------
def _conservative_column(phi, theta_1, theta_2, hat_1, hat_2):
    """transform.py:98-143, one column, arithmetic in the array dtype (like the numba loop)."""
    dt = phi.dtype.type
    out = np.full(hat_1.shape[0], np.nan, dtype=phi.dtype)
    for i in range(theta_1.shape[0]):
        t1, t2 = theta_1[i], theta_2[i]
        if np.isnan(t1) and np.isnan(t2):
            continue
        elif np.isnan(t1):
            tmin = tmax = t2
        elif np.isnan(t2):
            tmin = tmax = t1
        elif t1 < t2:
            tmin, tmax = t1, t2
        else:
            tmin, tmax = t2, t1
        if np.isnan(phi[i]):
            continue
        for j in range(hat_1.shape[0]):
            if hat_1[j] > tmax or hat_2[j] < tmin:
                continue
            if tmax == tmin:
                add = phi[i]
            else:
                hmin = max(tmin, hat_1[j])
                hmax = min(tmax, hat_2[j])
                alpha = dt(dt(hmax - hmin) / dt(tmax - tmin))
                add = dt(alpha * phi[i])
            out[j] = add if np.isnan(out[j]) else dt(out[j] + add)
    return out


def vinterp_conservative(phi, theta, target_bins, axis=-1):
    """interp_1d_conservative (transform.py:145-191) over columns along ``axis``; theta has one
    more point than phi along ``axis``; new dim (m - 1 bins) appended LAST."""
    phi = np.asarray(phi)
    theta = np.asarray(theta)
    target_bins = np.asarray(target_bins)
    if not (phi.dtype == np.float32 and theta.dtype == np.float32 and target_bins.dtype == np.float32):
        phi, theta, target_bins = phi.astype(np.float64), theta.astype(np.float64), target_bins.astype(np.float64)
    d = np.diff(target_bins)
    if np.all(d < 0):
        flip, target_bins = True, target_bins[::-1]
    elif np.all(d > 0):
        flip = False
    else:
        raise ValueError("Target values are not monotonic")
    pm = np.moveaxis(phi, axis, -1)
    tshape = list(phi.shape)
    tshape[axis] += 1
    tm = np.moveaxis(np.broadcast_to(theta, tshape), axis, -1)
    cols = pm.reshape(-1, pm.shape[-1])
    tcols = tm.reshape(-1, tm.shape[-1])
    out = np.empty((cols.shape[0], target_bins.shape[0] - 1), dtype=phi.dtype)
    for c in range(cols.shape[0]):
        out[c] = _conservative_column(cols[c], tcols[c, :-1], tcols[c, 1:], target_bins[:-1], target_bins[1:])
    if flip:
        # NB the reference does `out[::-1]`, which reverses the FIRST axis; for the 1-D columns its
        # tests use that is the bin axis.  We reverse the bin axis for any rank.
        out = out[:, ::-1]
    return out.reshape(pm.shape[:-1] + (target_bins.shape[0] - 1,))


def stencil_pair(op_a, a, axis_a, lo_a, hi_a, bc_a, fill_a, pre_a, op_b, b, axis_b, lo_b, hi_b, bc_b, fill_b, pre_b,
                 subtract=False, post=None):
    """The chain a user writes with the reference (docs/ufunc_examples.md:105-153; xgcm/grid.py:796-832 per term):
    (op_a(a * pre_a, axis_a) +- op_b(b * pre_b, axis_b)) / post, every step a separate numpy pass."""
    ta = stencil2(op_a, a, axis_a, lo_a, hi_a, bc_a, fill_a, pre_a, None)
    tb = stencil2(op_b, b, axis_b, lo_b, hi_b, bc_b, fill_b, pre_b, None)
    r = ta + tb if not subtract else (ta - tb if int(subtract) == 1 else tb - ta)
    if post is not None:
        with np.errstate(invalid="ignore", divide="ignore"):
            r = r / post
    return r.astype(np.asarray(a).dtype)


def log32_port(x):
    """numpy restatement of the float32 log the device evaluates for `method="log"` (csrc/xg_vinterp.cuh
    xg_log<float>): numpy's own AVX2 / AVX512F algorithm (a rational minimax approximation after range reduction to
    (1/sqrt2, sqrt2]), FMAs emulated through float64.  Test infrastructure: pins the port against np.log."""
    x = np.asarray(x, dtype=np.float32)
    P = [np.float32(v) for v in (0.0, 9.999999999999998702752e-01, 2.112677543073053063722e+00,
                                 1.480000633576506585156e+00, 3.808837741388407920751e-01, 2.589979117907922693523e-02)]
    Q = [np.float32(v) for v in (1.0, 2.612677543073109236779e+00, 2.453006071784736363091e+00,
                                 9.864942958519418960339e-01, 1.546476374983906719538e-01, 5.875095403124574342950e-03)]

    def fma(a, b, c):
        return (a.astype(np.float64) * np.float64(b) + np.float64(c)).astype(np.float32) if np.isscalar(b) or b.ndim == 0 \
            else (a.astype(np.float64) * b.astype(np.float64) + np.asarray(c, np.float64)).astype(np.float32)

    with np.errstate(all="ignore"):
        m, e = np.frexp(x)
        m = m.astype(np.float32)
        e = e.astype(np.float32)
        low = m <= np.float32(0.70710678118654752440)
        m = np.where(low, m * np.float32(2), m)
        e = np.where(low, e - 1, e)
        t = m - np.float32(1)
        num = np.full_like(t, P[5])
        for c in (P[4], P[3], P[2], P[1], P[0]):
            num = fma(num, t, np.full_like(t, c))
        den = np.full_like(t, Q[5])
        for c in (Q[4], Q[3], Q[2], Q[1], Q[0]):
            den = fma(den, t, np.full_like(t, c))
        out = fma(e, np.full_like(t, np.float32(0.693147180559945309417232121458176568)), (num / den).astype(np.float32))
        out = np.where(x == 0, np.float32(-np.inf), out)
        out = np.where(x < 0, np.float32(np.nan), out)
        out = np.where(np.isinf(x) & (x > 0), np.float32(np.inf), out)
        out = np.where(np.isnan(x), np.float32(np.nan), out)
    return out.astype(np.float32)
