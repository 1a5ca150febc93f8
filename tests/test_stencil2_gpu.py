"""Parity of the fused stencil kernel (through the C-ABI) against the oracle: bit-exact."""

import itertools

import numpy as np
import pytest
import torch

from oracle import stencil as oracle

pytestmark = pytest.mark.gpu

SHIFTS = sorted(set(oracle.PADDING_WIDTH.values()))  # (0,0) (0,1) (1,0) (1,1)
OPS = ["diff", "interp", "min", "max"]
BCS = [("periodic", 0.0), ("fill", 0.0), ("fill", 1.5), ("fill", float("nan")), ("extend", 0.0)]


def _field(shape, dtype, seed=0, nan_frac=0.0):
    rng = np.random.default_rng(seed)
    a = rng.random(shape).astype(dtype)
    if nan_frac:
        a[rng.random(shape) < nan_frac] = np.nan
    return a


def _run(a, axis, op, lo, hi, bc, fill, pre=None, post=None):
    from xgcm_b200 import ops

    dev = torch.device("cuda:0")
    x = torch.from_numpy(a).to(dev)
    p = None if pre is None else torch.from_numpy(np.ascontiguousarray(pre)).to(dev)
    q = None if post is None else torch.from_numpy(np.ascontiguousarray(post)).to(dev)
    out = ops.stencil2(x, axis, op, lo, hi, bc, fill, pre=p, post=q)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def _check(a, axis, op, lo, hi, bc, fill, pre=None, post=None):
    if a.shape[axis] + lo + hi - 1 <= 0:
        return
    want = oracle.stencil2(op, a, axis, lo, hi, bc if (lo or hi) else None, fill, pre, post)
    got = _run(a, axis, op, lo, hi, bc, fill, pre, post)
    assert got.dtype == want.dtype
    assert got.shape == want.shape
    np.testing.assert_array_equal(got, want)  # NaN == NaN positions, bit-exact values


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("op", OPS)
def test_config2_all_axes_shifts_bcs(dtype, op):
    """MITgcm-like (50, 240, 360) field, every axis x shift x boundary (BASELINE configs[1])."""
    a = _field((50, 240, 360), dtype, seed=1)
    for axis, (lo, hi), (bc, fill) in itertools.product(range(3), SHIFTS, BCS):
        _check(a, axis, op, lo, hi, bc, fill)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize(
    "shape",
    [(1,), (2,), (5,), (127,), (128,), (129,), (4097,), (1000,), (3, 5), (7, 13, 5), (2, 3, 4, 6),
     (5, 1, 7), (1, 9), (9, 1), (3, 130), (2, 3, 256), (2, 3, 258), (4, 1026)],
)
def test_ragged_shapes(dtype, shape):
    a = _field(shape, dtype, seed=2, nan_frac=0.05)
    for axis in range(len(shape)):
        for op, (lo, hi), (bc, fill) in itertools.product(OPS, SHIFTS, BCS):
            _check(a, axis, op, lo, hi, bc, fill)


def test_config1_periodic_1d_fp64():
    """BASELINE configs[0]: 1e6 fp64 cells, periodic, center->left diff and interp."""
    a = _field((1_000_000,), np.float64, seed=3)
    for op in ("diff", "interp"):
        _check(a, 0, op, 1, 0, "periodic", 0.0)
        _check(a, 0, op, 0, 1, "periodic", 0.0)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_metric_weighting_bit_exact(dtype):
    """op(a*m)/m_new must round exactly like the reference's three separate passes
    (xgcm/test/test_metrics_ops.py:35-95 uses .equals)."""
    rng = np.random.default_rng(4)
    shape = (6, 20, 36)
    a = _field(shape, dtype, seed=5)
    full = (1.0 + rng.random(shape)).astype(dtype)
    dx2 = (1.0 + rng.random((1,) + shape[1:])).astype(dtype)  # (Y, X) metric
    dz1 = (1.0 + rng.random((shape[0], 1, 1))).astype(dtype)  # (Z,) metric
    dy1 = (1.0 + rng.random((1, shape[1], 1))).astype(dtype)  # (Y,) metric
    dx1 = (1.0 + rng.random((1, 1, shape[2]))).astype(dtype)  # (X,) metric
    for axis in range(3):
        for (lo, hi), (bc, fill) in itertools.product(SHIFTS, BCS[:3] + BCS[4:]):
            n_out = shape[axis] + lo + hi - 1
            if n_out <= 0:
                continue
            oshape = list(shape)
            oshape[axis] = n_out
            for pre in (None, full, dx2, dz1, dy1, dx1):
                for post_kind in (None, "full", "dx2", "dz1", "dy1", "dx1"):
                    if pre is None and post_kind is None:
                        continue
                    post = None
                    if post_kind is not None:
                        src = {"full": full, "dx2": dx2, "dz1": dz1, "dy1": dy1, "dx1": dx1}[post_kind]
                        pshape = [oshape[d] if src.shape[d] != 1 else 1 for d in range(3)]
                        post = (1.0 + np.random.default_rng(7).random(pshape)).astype(dtype)
                    _check(a, axis, "diff", lo, hi, bc, fill, pre, post)
                    _check(a, axis, "interp", lo, hi, bc, fill, pre, post)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(7, 5, 256), (5, 3, 132), (2, 2, 128), (9, 1, 260), (3, 2, 4, 136), (6, 520),
                                   (7, 5, 1000), (5, 9, 452), (3, 2, 3, 676), (6, 904), (2, 3, 1792)])
def test_metric_shared_divisor_rows_bit_exact(dtype, shape):
    """The z-batched row kernels (register-staged below 2 tiles per row, TMA-staged above; divisor
    dx(Y, X) or dx(X) shared by every level, inverted once):
    same bits as the reference's separate multiply / op / divide passes, every boundary, halo side
    and pre-metric layout, ragged level counts (Zn % 4 != 0) and ragged last chunks."""
    rng = np.random.default_rng(50)
    a = _field(shape, dtype, seed=51, nan_frac=0.01)
    nd = len(shape)
    axis = nd - 1

    def metric(dims):
        shp = [shape[d] if d in dims else 1 for d in range(nd)]
        return (0.5 + rng.random(shp)).astype(dtype)

    posts = [metric((nd - 2, nd - 1)), metric((nd - 1,))]
    pres = [None, metric(range(nd)), metric((nd - 2, nd - 1)), metric((0,)), metric((nd - 1,)), metric((nd - 2,))]
    for (lo, hi) in ((1, 0), (0, 1)):
        for (bc, fill) in BCS + [("extrapolate", 0.0)]:
            for post in posts:
                for pre in pres:
                    for op in ("diff", "interp"):
                        _check(a, axis, op, lo, hi, bc, fill, pre, post)
            _check(a, axis, "max", lo, hi, bc, fill, None, posts[0])
            _check(a, axis, "min", lo, hi, bc, fill, pres[2], posts[0])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("X,kernel", [(256, "row_zb"), (1024, "row_tma")])
def test_shared_divisor_division_is_ieee_on_arbitrary_bit_patterns(dtype, X, kernel):
    """x / b through the shared reciprocal against numpy's division on raw bit patterns: zeros,
    subnormals, infinities, NaNs, quotients that overflow, underflow or land in the subnormal range.
    The numerator handed to the divide is the field value itself (see below), so numerator and divisor
    are both fully controlled."""
    from xgcm_b200 import ops

    rng = np.random.default_rng(52)
    Z, Y = 16, 256 * (1024 // X)
    bits = np.uint32 if dtype == np.float32 else np.uint64
    hi_bit = 32 if dtype == np.float32 else 64

    def patterns(shape):
        raw = rng.integers(0, 2 ** hi_bit, size=shape, dtype=np.uint64).astype(bits)
        return raw.view(dtype)

    specials = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, np.finfo(dtype).max, -np.finfo(dtype).max,
                         np.finfo(dtype).tiny, np.finfo(dtype).smallest_subnormal, 1.0, -1.0, 3.0, 1 / 3],
                        dtype=dtype)
    a = patterns((Z, Y, X))
    a[:, :, :32] = rng.choice(specials, size=(Z, Y, 32))
    b = patterns((1, Y, X))
    b[:, :, 16:48] = rng.choice(specials, size=(1, Y, 32))
    # near-1 quotients: numerator and divisor share the exponent, mantissas random
    a[:, :, 64:128] = (1.0 + rng.random((Z, Y, 64))).astype(dtype)
    b[:, :, 64:128] = (1.0 + rng.random((1, Y, 64))).astype(dtype)
    dev = torch.device("cuda:0")
    # diff = A[x] - A[x-1] on a field whose even columns are +0 is the field itself at the odd ones
    # (x - (+0) == x for every x, -0 and NaN included)
    a2 = a.copy()
    a2[:, :, 0::2] = 0.0
    with np.errstate(all="ignore"):
        want = (a2 / b)[:, :, 1::2]
    m = torch.from_numpy(b).to(dev)
    ones = torch.ones((1, Y, X), dtype=m.dtype, device=dev)  # a level-shared pre-metric: x * 1 is exact
    got = ops.stencil2(torch.from_numpy(a2).to(dev), 2, "diff", 1, 0, "fill", 0.0, pre=ones, post=m).cpu().numpy()
    assert _capi_last_launch() == f"xg_stencil2({kernel})"
    got = got[:, :, 1::2]
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    np.testing.assert_array_equal(got.view(bits)[ok], want.view(bits)[ok])  # signed zeros included


def _capi_last_launch():
    from xgcm_b200 import _capi

    return _capi.last_launch()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(5, 9, 452), (3, 6, 904), (2, 3, 5, 676), (7, 2, 1000), (2, 1, 448)])
def test_metric_tile_kernel_second_to_last_axis(dtype, shape):
    """Metric-fused stencils along the dim next to x with level-shared metrics (derivative('Y') on a
    (Z, Y, X) field): the TMA-staged tile kernel against the reference's separate passes, every shift
    (n_out = n - 1, n, n + 1), boundary incl. extrapolate, metric layout and operator; ragged tiles."""
    from xgcm_b200 import _capi

    rng = np.random.default_rng(60)
    a = _field(shape, dtype, seed=61, nan_frac=0.01)
    nd = len(shape)
    axis = nd - 2

    def metric(dims, n_axis):
        shp = [shape[d] if d in dims else 1 for d in range(nd)]
        if axis in dims:
            shp[axis] = n_axis
        return (0.5 + rng.random(shp)).astype(dtype)

    for (lo, hi) in SHIFTS:
        n_out = shape[axis] + lo + hi - 1
        if n_out <= 0:
            continue
        posts = [metric((nd - 2, nd - 1), n_out), metric((nd - 1,), n_out), metric((nd - 2,), n_out), metric((0,), n_out)]
        pres = [None, metric(range(nd), shape[axis]), metric((nd - 2, nd - 1), shape[axis]), metric((0,), shape[axis]),
                metric((nd - 1,), shape[axis]), metric((nd - 2,), shape[axis])]
        for (bc, fill) in BCS + [("extrapolate", 0.0)]:
            for post in posts:
                for pre in pres:
                    _check(a, axis, "diff", lo, hi, bc, fill, pre, post)
                _check(a, axis, "interp", lo, hi, bc, fill, pres[2], post)
            _check(a, axis, "max", lo, hi, bc, fill, pres[1], posts[0])
            if shape[-1] >= (448 if dtype == np.float32 else 480) and shape[axis] > 1:  # two tiles per row
                assert _capi.last_launch() == "xg_stencil2(tile_tma)"
            _check(a, axis, "min", lo, hi, bc, fill, None, posts[0])
            # a full (Z, Y, X) divisor with a shared pre-metric still stages the pre tile once per level batch
            _check(a, axis, "diff", lo, hi, bc, fill, pres[2], metric(range(nd), n_out))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(9, 5, 452), (6, 3, 904), (5, 2, 3, 676), (3, 7, 1000)])
def test_metric_tile_kernel_outermost_axis(dtype, shape):
    """Stencils along the OUTERMOST dim of a (Z, Y, X) field with dz(Z)-like metrics (interp / derivative
    along Z, metric-weighted): the tile kernel runs them as rows = Z, levels = Y (x Y'), sharing the per-row
    scalar divisor.  Against the reference's separate passes, every shift, boundary and metric layout."""
    from xgcm_b200 import _capi

    rng = np.random.default_rng(80)
    a = _field(shape, dtype, seed=81, nan_frac=0.01)
    nd = len(shape)

    def metric(dims, n_axis):
        shp = [shape[d] if d in dims else 1 for d in range(nd)]
        if 0 in dims:
            shp[0] = n_axis
        return (0.5 + rng.random(shp)).astype(dtype)

    for (lo, hi) in SHIFTS:
        n_out = shape[0] + lo + hi - 1
        posts = [metric((0,), n_out), metric(range(nd), n_out), metric((nd - 2, nd - 1), n_out)]
        pres = [None, metric((0,), shape[0]), metric(range(nd), shape[0]), metric((nd - 1,), shape[0])]
        for (bc, fill) in BCS + [("extrapolate", 0.0)]:
            for post in posts:
                for pre in pres:
                    _check(a, 0, "interp", lo, hi, bc, fill, pre, post)
            _check(a, 0, "diff", lo, hi, bc, fill, pres[1], posts[0])
            if shape[-1] >= (448 if dtype == np.float32 else 480):
                assert _capi.last_launch() == "xg_stencil2(tile_tma)"
            _check(a, 0, "max", lo, hi, bc, fill, pres[2], posts[0])
            _check(a, 0, "min", lo, hi, bc, fill, None, posts[0])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape,axis", [((5, 6, 904), 2), ((5, 6, 904), 1), ((4, 3, 256), 2), ((3, 4, 130), 1), ((6, 5, 904), 0)])
def test_halo_planes_with_metrics(dtype, shape, axis):
    """Explicit halo planes (the neighbour GPU's boundary plane in the sharded path, xgcm_b200.h) replace the
    boundary rule on their side — through every kernel a metric-fused call can reach (row / tile TMA kernels,
    the level-batched row kernel, the strided kernel): out = OP(concat(halo_lo, A x pre, halo_hi)) / post."""
    from xgcm_b200 import ops

    rng = np.random.default_rng(70)
    a = _field(shape, dtype, seed=71)
    nd = len(shape)
    plane = [s for d, s in enumerate(shape) if d != axis]
    hl, hh = rng.random(plane).astype(dtype), rng.random(plane).astype(dtype)
    dev = torch.device("cuda:0")

    def metric(dims, n_axis):
        shp = [shape[d] if d in dims else 1 for d in range(nd)]
        if axis in dims:
            shp[axis] = n_axis
        return (0.5 + rng.random(shp)).astype(dtype)

    for (lo, hi), (bc, fill), op in itertools.product(((1, 0), (0, 1), (1, 1)), (("extend", 0.0), ("fill", 2.5)), ("diff", "interp")):
        n_out = shape[axis] + lo + hi - 1
        for pre_dims, post_dims in (((1, 2), (1, 2)), (None, (1, 2)), (range(nd), (1, 2)), ((0,), (2,)), ((0,), (0,)),
                                    (range(nd), (0,)), (None, (1,))):
            pre = None if pre_dims is None else metric(pre_dims, shape[axis])
            post = metric(post_dims, n_out)
            ap = a if pre is None else a * pre
            parts = ([np.expand_dims(hl, axis)] if lo else []) + [ap] + ([np.expand_dims(hh, axis)] if hi else [])
            padded = np.concatenate(parts, axis=axis)
            want = (np.moveaxis(oracle.KERNELS[op](np.moveaxis(padded, axis, -1)), -1, axis) / post).astype(dtype)
            got = ops.stencil2(torch.from_numpy(a).to(dev), axis, op, lo, hi, bc, fill,
                               pre=None if pre is None else torch.from_numpy(pre).to(dev), post=torch.from_numpy(post).to(dev),
                               halo_lo=torch.from_numpy(hl).to(dev) if lo else None,
                               halo_hi=torch.from_numpy(hh).to(dev) if hi else None).cpu().numpy()
            np.testing.assert_array_equal(got, want, err_msg=f"{op} lo={lo} hi={hi} {bc} pre={pre_dims} post={post_dims}")


def test_metric_4d_outer_broadcast():
    """(T, Z, Y, X) field with dx(Y, X) along X and dz(Z) along Y: multi-group outer index."""
    shape = (3, 4, 10, 16)
    a = _field(shape, np.float32, seed=8)
    rng = np.random.default_rng(9)
    dx = (1 + rng.random((1, 1, 10, 16))).astype(np.float32)
    dz = (1 + rng.random((1, 4, 1, 1))).astype(np.float32)
    dt = (1 + rng.random((3, 1, 1, 1))).astype(np.float32)
    for axis in range(4):
        for pre, post in ((dx, None), (None, dz), (dz, dx), (dt, dz), (dx, dt)):
            lo, hi = 1, 0
            _check(a, axis, "diff", lo, hi, "periodic", 0.0, pre, post)
            _check(a, axis, "interp", 0, 1, "extend", 0.0, pre, post)


def test_errors():
    from xgcm_b200 import ops

    x = torch.zeros((4, 4), device="cuda")
    with pytest.raises(ValueError):
        ops.stencil2(x, 0, "diff", 1, 0, None)  # padding.py:601-608
    with pytest.raises(ValueError):
        ops.stencil2(x, 0, "diff", 1, 0, "bogus")
    with pytest.raises(RuntimeError):
        ops.stencil2(torch.zeros((4, 4)), 0, "diff", 1, 0, "fill")  # no CPU fallback
    with pytest.raises(TypeError):
        ops.stencil2(x.to(torch.int32), 0, "diff", 1, 0, "fill")


# ----------------------------------------------------------------------------- fused multi-axis
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(6, 8, 64), (5, 7, 33), (3, 4, 6, 40), (4, 130), (2, 3, 260)])
def test_stencil_multi_equals_sequential(dtype, shape):
    """xg_stencil_multi == the per-axis passes of grid.py:800-832, bit for bit: every axis pair /
    triple in every order, every shift, mixed boundaries and ops."""
    from xgcm_b200 import ops

    rng = np.random.default_rng(40)
    a = _field(shape, dtype, seed=41, nan_frac=0.02)
    x = torch.from_numpy(a).to("cuda:0")
    nd = len(shape)
    bcs = [("periodic", 0.0), ("fill", 1.5), ("extend", 0.0), ("fill", float("nan"))]
    combos = list(itertools.permutations(range(nd), 2)) + list(itertools.permutations(range(nd), 3))[:12]
    for axes in combos:
        for trial in range(6):
            specs = []
            ok = True
            same_op = OPS[rng.integers(0, 4)]  # Grid methods apply one operator along all axes;
            for ax in axes:                    # every 6th trial mixes them (chained per-axis path)
                lo, hi = SHIFTS[rng.integers(0, 4)]
                if shape[ax] + lo + hi - 1 <= 0:
                    ok = False
                bc, fill = bcs[rng.integers(0, 4)]
                specs.append((ax, same_op if trial < 5 else OPS[rng.integers(0, 4)], lo, hi, bc, fill))
            if not ok:
                continue
            want = a
            for ax, op, lo, hi, bc, fill in specs:
                want = oracle.stencil2(op, want, ax, lo, hi, bc if (lo or hi) else None, fill)
            got = ops.stencil_multi(x, specs).cpu().numpy()
            assert got.shape == want.shape
            np.testing.assert_array_equal(got, want, err_msg=str(specs))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(5, 9, 452), (9, 6, 904), (3, 2, 5, 676), (2, 3, 1000), (7, 480), (1, 1, 5, 452)])
def test_stencil_multi_tma_tiles_equal_sequential(dtype, shape):
    """The TMA-staged multi-axis kernel (innermost axis first, length-preserving ops, rows of >= 2 tiles):
    every axis subset of the last three dims, every lo / hi choice, every boundary combination (periodic and
    extend are materialised in the tile, fill is applied in the chain), all four operators, ragged tiles in
    x, rows and levels — bit for bit the per-axis passes of grid.py:800-832."""
    from xgcm_b200 import _capi, ops

    a = _field(shape, dtype, seed=43, nan_frac=0.02)
    x = torch.from_numpy(a).to("cuda:0")
    nd = len(shape)
    lead_one = all(s == 1 for s in shape[: max(nd - 3, 0)])
    axis_sets = [(nd - 1, nd - 2)]
    if nd >= 3:
        axis_sets += [(nd - 2, nd - 3), (nd - 1, nd - 3), (nd - 1, nd - 2, nd - 3)]
    bcs = [("periodic", 0.0), ("fill", 1.5), ("extend", 0.0), ("fill", float("nan"))]
    big = shape[-1] >= (448 if dtype == np.float32 else 480)
    for axes in axis_sets:
        if any(shape[ax] < 2 for ax in axes):
            continue
        for los in itertools.product((0, 1), repeat=len(axes)):
            for bcsel in itertools.product(range(3), repeat=len(axes)):
                for op in (("interp", "diff") if sum(bcsel) % 2 else ("interp", "max", "min")):
                    pick = [3 if (op == "min" and b == 1) else b for b in bcsel]  # min: a NaN fill value
                    specs = [(ax, op, lo, 1 - lo, bcs[b][0], bcs[b][1]) for ax, lo, b in zip(axes, los, pick)]
                    want = a
                    for ax, o, lo, hi, bc, fill in specs:
                        want = oracle.stencil2(o, want, ax, lo, hi, bc, fill)
                    got = ops.stencil_multi(x, specs).cpu().numpy()
                    np.testing.assert_array_equal(got, want, err_msg=str(specs))
                    if big and (nd - 3 not in axes or lead_one):
                        assert _capi.last_launch() == "xg_stencil_multi(tile_tma)", str(specs)


def test_stencil_multi_c_grid_interp_to_corner():
    """The notebook idiom grid.interp(da, ['X', 'Y']) on a (Z, Y, X) field, plus the 3-axis corner."""
    from xgcm_b200 import ops

    a = _field((10, 96, 128), np.float32, seed=42)
    x = torch.from_numpy(a).to("cuda:0")
    got = ops.stencil_multi(x, [(2, "interp", 1, 0, "periodic", 0.0), (1, "interp", 1, 0, "fill", 0.0)]).cpu().numpy()
    want = oracle.stencil2("interp", oracle.stencil2("interp", a, 2, 1, 0, "periodic"), 1, 1, 0, "fill", 0.0)
    np.testing.assert_array_equal(got, want)
    got = ops.stencil_multi(x, [(1, "diff", 0, 1, "extend", 0.0), (2, "diff", 1, 0, "periodic", 0.0), (0, "diff", 1, 1, "fill", 2.0)]).cpu().numpy()
    want = oracle.stencil2("diff", oracle.stencil2("diff", oracle.stencil2("diff", a, 1, 0, 1, "extend"), 2, 1, 0, "periodic"), 0, 1, 1, "fill", 2.0)
    np.testing.assert_array_equal(got, want)
    with pytest.raises(ValueError):
        ops.stencil_multi(x, [(2, "interp", 1, 0, None, 0.0), (1, "interp", 1, 0, "fill", 0.0)])
    with pytest.raises(ValueError):
        ops.stencil_multi(x, [(2, "interp", 1, 0, "fill", 0.0), (2, "interp", 1, 0, "fill", 0.0)])
