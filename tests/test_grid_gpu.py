"""Grid-level parity: xgcm_b200.Grid (CUDA kernels via the C-ABI) against the oracle and the
known answers of the reference's test-suite."""

import itertools
import warnings

import numpy as np
import pytest
import torch

import xgcm_b200 as xg
from oracle import stencil as oracle

from _fixtures import POS_SUFFIX, all_positions_1d, all_positions_3d, grid_metric_dataset

pytestmark = pytest.mark.gpu

OPS = ["diff", "interp", "min", "max"]


def _periodic_1d(n=100, dtype=np.float64):
    x_c = np.arange(n) + 0.5
    x_g = np.arange(n) + 0.0
    data = np.sin(2 * np.pi * x_c / n).astype(dtype)
    ds = xg.Dataset(coords={"XC": x_c, "XG": x_g}, data_vars={"data_c": (("XC",), data), "data_g": (("XG",), np.cos(x_g).astype(dtype))})
    return ds


def test_periodic_1d_roll_identity():
    """xgcm/test/test_grid_ufunc.py:345-382: periodic center->left == roll arithmetic."""
    ds = _periodic_1d()
    grid = xg.Grid(ds, coords={"X": {"center": "XC", "left": "XG"}}, padding="periodic")
    a = ds["data_c"].values
    out = grid.diff(ds["data_c"], "X")
    assert out.dims == ("XG",)
    np.testing.assert_array_equal(out.values, a - np.roll(a, 1))
    np.testing.assert_array_equal(out.coords["XG"].values, ds["XG"].values)
    out = grid.interp(ds["data_c"], "X")
    np.testing.assert_array_equal(out.values, 0.5 * (a + np.roll(a, 1)))
    b = ds["data_g"].values
    out = grid.diff(ds["data_g"], "X")
    assert out.dims == ("XC",)
    np.testing.assert_array_equal(out.values, np.roll(b, -1) - b)


def test_config1_1e6_fp64_periodic():
    """BASELINE configs[0] through the Grid API."""
    n = 1_000_000
    rng = np.random.default_rng(0)
    ds = xg.Dataset(coords={"XC": np.arange(n) + 0.5, "XG": np.arange(n) + 0.0},
                    data_vars={"f": (("XC",), rng.random(n))})
    grid = xg.Grid(ds, coords={"X": {"center": "XC", "left": "XG"}}, padding="periodic")
    a = ds["f"].values
    np.testing.assert_array_equal(grid.diff(ds["f"], "X").values, oracle.stencil2("diff", a, 0, 1, 0, "periodic"))
    np.testing.assert_array_equal(grid.interp(ds["f"], "X").values, oracle.stencil2("interp", a, 0, 1, 0, "periodic"))


def test_known_answers_fill_and_extend():
    """test_grid_ufunc.py:1227-1273 (fill 0/1/10 on arange(9)), :1326-1338 (c->outer extend),
    test_grid.py:502-525 (edge = data[0] - fill)."""
    ds, gc = all_positions_1d(9)
    a = np.arange(9.0)
    da = xg.DataArray(a, dims=("x_c",))
    grid = xg.Grid(ds, coords=gc)
    for fill in (0.0, 1.0, 10.0):
        got = grid.diff(da, "X", to="left", padding="fill", fill_value=fill)
        np.testing.assert_array_equal(got.values, np.diff(np.concatenate([[fill], a])))
        assert got.dims == ("x_g",)
    lin = xg.DataArray(np.linspace(1, 10, 9 + 1)[:9] * 0 + np.linspace(1, 9, 9), dims=("x_c",))
    got = grid.interp(lin, "X", to="outer", padding="extend")
    assert got.dims == ("x_o",)
    np.testing.assert_array_equal(got.values, np.concatenate([[1.0], np.arange(1.5, 9, 1.0), [9.0]]))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_all_shifts_ops_paddings_3d(dtype):
    """Every (from, to) of gridops.py:27-215 x op x padding x axis on a small (Z, Y, X) field."""
    ds, gc, rng = all_positions_3d((6, 7, 8))
    grid = xg.Grid(ds, coords=gc, padding="periodic")
    for ax_i, ax in enumerate("ZYX"):
        for (src, dst), (lo, hi) in oracle.PADDING_WIDTH.items():
            dims = [f"{a.lower()}_c" for a in "ZYX"]
            dims[ax_i] = f"{ax.lower()}_{POS_SUFFIX[src]}"
            shape = [ds.sizes[d] for d in dims]
            a = rng.random(shape).astype(dtype)
            da = xg.DataArray(a, dims=dims, name="f")
            for op, (pad, fill) in itertools.product(OPS, [("periodic", None), ("fill", 0.0), ("fill", 2.5), ("extend", None)]):
                got = getattr(grid, op)(da, ax, to=dst, padding=pad, fill_value=fill)
                want = oracle.stencil2(op, a, ax_i, lo, hi, pad if (lo or hi) else None, 0.0 if fill is None else fill)
                exp_dims = list(dims)
                exp_dims[ax_i] = f"{ax.lower()}_{POS_SUFFIX[dst]}"
                assert got.dims == tuple(exp_dims)
                assert got.dtype == dtype
                np.testing.assert_array_equal(got.values, want)
                # coordinates of the shifted dim come from the grid dataset (grid_ufunc.py:1262-1320)
                np.testing.assert_array_equal(got.coords[exp_dims[ax_i]].values, ds[exp_dims[ax_i]].values)


def test_multi_axis_sequential_and_dim_order():
    """grid.py:800: axes are processed in the given order; dim order of the input is kept."""
    ds, gc, rng = all_positions_3d((6, 7, 8))
    grid = xg.Grid(ds, coords=gc, padding={"X": "periodic", "Y": "fill", "Z": "extend"}, fill_value={"Y": 0.0})
    a = rng.random((7, 6, 8)).astype(np.float32)
    da = xg.DataArray(a, dims=("y_c", "z_c", "x_c"))
    got = grid.interp(da, ["X", "Y", "Z"])
    want = oracle.stencil2("interp", a, 2, 1, 0, "periodic")
    want = oracle.stencil2("interp", want, 0, 1, 0, "fill", 0.0)
    want = oracle.stencil2("interp", want, 1, 1, 0, "extend")
    assert got.dims == ("y_g", "z_g", "x_g")
    np.testing.assert_array_equal(got.values, want)
    # per-axis `to` and per-call per-axis padding dicts (grid.py:315-332 precedence)
    got = grid.diff(da, ["X", "Z"], to={"X": "right", "Z": "outer"}, padding={"X": "fill"}, fill_value={"X": 3.0})
    want = oracle.stencil2("diff", a, 2, 0, 1, "fill", 3.0)
    want = oracle.stencil2("diff", want, 1, 1, 1, "extend")
    assert got.dims == ("y_c", "z_o", "x_r")
    np.testing.assert_array_equal(got.values, want)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_multi_axis_fused_path_wide_rows(dtype):
    """Rows >= 32 wide take the fused xg_stencil_multi launch: same values as the per-axis loop
    (grid.py:800-832), for every op, 2 and 3 axes, any order, per-axis `to` / padding / fill."""
    ds, gc, rng = all_positions_3d((6, 20, 40))
    grid = xg.Grid(ds, coords=gc, padding={"X": "periodic", "Y": "fill", "Z": "extend"})
    a = rng.random((6, 20, 40)).astype(dtype)
    da = xg.DataArray(a, dims=("z_c", "y_c", "x_c"))
    bc = {"X": "periodic", "Y": "fill", "Z": "extend"}
    idx = {"Z": 0, "Y": 1, "X": 2}
    for op in OPS:
        for axes in (["X", "Y"], ["Y", "X"], ["Y", "Z"], ["Z", "X"], ["X", "Y", "Z"], ["Z", "Y", "X"]):
            for to in ("left", "right", "outer", "inner"):
                lo, hi = oracle.PADDING_WIDTH[("center", to)]
                want = a
                for ax in axes:
                    want = oracle.stencil2(op, want, idx[ax], lo, hi, bc[ax] if (lo or hi) else None, 0.0)
                got = getattr(grid, op)(da, axes, to=to)
                sfx = POS_SUFFIX[to]
                exp_dims = tuple(f"{d[0]}_{sfx}" if d[0].upper() in axes else d for d in ("z_c", "y_c", "x_c"))
                assert got.dims == exp_dims, (op, axes, to)
                np.testing.assert_array_equal(got.values, want, err_msg=f"{op} {axes} {to}")
                np.testing.assert_array_equal(got.coords[exp_dims[2]].values, ds[exp_dims[2]].values)
    got = grid.diff(da, ["X", "Z"], to={"X": "right", "Z": "outer"}, padding={"X": "fill"}, fill_value={"X": 3.0})
    want = oracle.stencil2("diff", oracle.stencil2("diff", a, 2, 0, 1, "fill", 3.0), 0, 1, 1, "extend")
    np.testing.assert_array_equal(got.values, want)
    with pytest.raises(ValueError, match="No boundary condition"):
        xg.Grid(ds, coords=gc).interp(da, ["X", "Y"])


def test_device_resident_inputs_stay_on_device():
    ds, gc, rng = all_positions_3d((6, 7, 8))
    grid = xg.Grid(ds, coords=gc, padding="periodic")
    a = rng.random((6, 7, 8)).astype(np.float32)
    da = xg.DataArray(torch.from_numpy(a).cuda(), dims=("z_c", "y_c", "x_c"))
    got = grid.diff(da, "Y")
    assert got.is_device and got.dims == ("z_c", "y_g", "x_c")
    np.testing.assert_array_equal(got.values, oracle.stencil2("diff", a, 1, 1, 0, "periodic"))
    got = grid.cumsum(da, "Z", padding="fill")
    assert got.is_device


def test_errors_match_reference():
    ds, gc = all_positions_1d(9)
    grid = xg.Grid(ds, coords=gc)  # no padding anywhere
    da = ds["a_c"]
    with pytest.raises(ValueError, match="No boundary condition was specified"):
        grid.diff(da, "X", to="left")  # padding.py:601-608
    grid.diff(da, "X", to="inner")  # zero halo: fine without a boundary condition
    with pytest.raises(NotImplementedError):
        grid.diff(ds["a_g"], "X", to="inner")  # gridops.py:68-70 / grid.py:1791-1802
    with pytest.raises(KeyError):
        grid.diff(da, "Q")
    with pytest.raises(KeyError):
        grid.diff(xg.DataArray(np.zeros(3), dims=("nope",)), "X")
    with pytest.raises(TypeError):
        grid.diff(np.zeros(9), "X")
    with pytest.raises(ValueError, match="renamed to 'padding'"):
        grid.diff(da, "X", boundary="fill")
    with pytest.raises(ValueError, match="keep_coords"):
        grid.diff(da, "X", keep_coords=True)


def test_input_not_modified_and_coords_preserved():
    """grid.py:791-795 (never mutate), GH#496 (non-core coords from the input survive)."""
    ds, gc, rng = all_positions_3d((4, 5, 6))
    grid = xg.Grid(ds, coords=gc, padding="periodic")
    a = rng.random((4, 5, 6))
    da = xg.DataArray(a.copy(), dims=("z_c", "y_c", "x_c"), coords={"z_c": np.arange(4) * 10.0})
    out = grid.interp(da, "X")
    np.testing.assert_array_equal(da.values, a)
    np.testing.assert_array_equal(out.coords["z_c"].values, np.arange(4) * 10.0)  # user's, not the grid's
    np.testing.assert_array_equal(out.coords["x_g"].values, ds["x_g"].values)


# --------------------------------------------------------------------------- cumsum
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_cumsum_all_shifts(dtype):
    """xgcm/test/test_grid.py:196-285: 4 shifts x fill/extend (+ reverse) vs np.cumsum."""
    ds, gc = all_positions_1d(9, dtype=dtype)
    grid = xg.Grid(ds, coords=gc)
    for (src, dst) in oracle.CUMSUM_TABLE_FWD:
        da = ds[f"a_{POS_SUFFIX[src]}"]
        a = da.values
        for rev, (pad, fill) in itertools.product((False, True), [("fill", 0.0), ("fill", 1.25), ("extend", None), ("periodic", None)]):
            table = oracle.CUMSUM_TABLE_REV if rev else oracle.CUMSUM_TABLE_FWD
            trim, (plo, phi) = table[(src, dst)]
            want = oracle.cumscan(a, 0, rev, trim, plo, phi, pad if (plo or phi) else None, 0.0 if fill is None else fill)
            got = grid.cumsum(da, "X", to=dst, padding=pad, fill_value=fill, reverse=rev)
            assert got.dims == (f"x_{POS_SUFFIX[dst]}",)
            np.testing.assert_array_equal(got.values, want)


def test_cumsum_known_answer_and_errors():
    ds, gc = all_positions_1d(14)
    grid = xg.Grid(ds, coords=gc, padding="fill")
    da = xg.DataArray(np.arange(1.0, 15.0), dims=("x_c",))
    got = grid.cumsum(da, "X", to="outer")
    np.testing.assert_array_equal(got.values, np.concatenate([[0.0], np.cumsum(np.arange(1.0, 15.0))]))
    with pytest.raises(ValueError, match="not a valid position shift"):
        grid.cumsum(ds["a_g"], "X", to="outer")
    with pytest.raises(ValueError, match="reverse"):
        grid.cumsum(da, "X", reverse={"Y": True})
    with pytest.raises(TypeError):
        grid.cumsum(da, "X", bogus=1)


def test_cumsum_multi_axis_3d():
    ds, gc, rng = all_positions_3d((6, 7, 8))
    grid = xg.Grid(ds, coords=gc, padding="fill")
    a = rng.random((6, 7, 8))
    da = xg.DataArray(a, dims=("z_c", "y_c", "x_c"))
    got = grid.cumsum(da, ["Z", "X"], to={"Z": "outer", "X": "left"}, reverse={"Z": True})
    t1, (l1, h1) = oracle.CUMSUM_TABLE_REV[("center", "outer")]
    want = oracle.cumscan(a, 0, True, t1, l1, h1, "fill", 0.0)
    t2, (l2, h2) = oracle.CUMSUM_TABLE_FWD[("center", "left")]
    want = oracle.cumscan(want, 2, False, t2, l2, h2, "fill", 0.0)
    assert got.dims == ("z_o", "y_c", "x_g")
    np.testing.assert_array_equal(got.values, want)


# --------------------------------------------------------------------------- metric-weighted ops
@pytest.mark.parametrize("grid_type", ["B", "C"])
@pytest.mark.parametrize("funcname", ["interp", "diff", "min", "max", "cumsum"])
def test_weighted_metric_bit_exact(funcname, grid_type):
    """xgcm/test/test_metrics_ops.py:35-64: new.equals(func(a*m)/m_new)."""
    ds, gc, metrics = grid_metric_dataset(grid_type)
    for padding_init in ("fill", "periodic", {"X": "periodic", "Y": "fill"}):
        grid = xg.Grid(ds, coords=gc, metrics=metrics, padding=padding_init)
        func = getattr(grid, funcname)
        for variable, axis, mw, padding in itertools.product(
            ["tracer", "u", "v"], ["X", "Y"], ["X", ("Y",), ("X", "Y"), ["X", "Y"]], ["fill", "extend"]
        ):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                metric = grid.get_metric(ds[variable], mw)
                raw = func(ds[variable] * metric, axis, padding=padding)
                metric_new = grid.get_metric(raw, mw)
                expected = raw / metric_new
                new = func(ds[variable], axis, metric_weighted=mw, padding=padding)
            assert new.dims == expected.dims
            assert new.equals(expected), (variable, axis, mw, padding)


@pytest.mark.parametrize("grid_type", ["B", "C"])
def test_weighted_metric_multi_axis(grid_type):
    """test_metrics_ops.py:66-95: multi-axis == the single-axis ops in series."""
    ds, gc, metrics = grid_metric_dataset(grid_type)
    grid = xg.Grid(ds, coords=gc, metrics=metrics)
    for funcname, variable, multi_axis in itertools.product(["interp", "diff", "cumsum"], ["tracer", "u"], [["X"], ["X", "Y"], ("Y", "X")]):
        func = getattr(grid, funcname)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            expected = ds[variable]
            for ax in multi_axis:
                expected = func(expected, ax, metric_weighted=("X", "Y"), padding="fill")
            new = func(ds[variable], multi_axis, metric_weighted=("X", "Y"), padding="fill")
        assert new.equals(expected)


@pytest.mark.parametrize("grid_type", ["B", "C"])
def test_derivative_equals_diff_over_metric(grid_type):
    """test_metrics_ops.py:125-253: derivative == diff / dx (bit-exact)."""
    ds, gc, metrics = grid_metric_dataset(grid_type)
    grid = xg.Grid(ds, coords=gc, metrics=metrics, padding="periodic")
    for var, axis in itertools.product(["tracer", "u", "v", "wt"], ["X", "Y", "Z"]):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            d = grid.diff(ds[var], axis)
            dx = grid.get_metric(d, (axis,))
            expected = d / dx
            got = grid.derivative(ds[var], axis)
        assert got.dims == expected.dims
        assert got.equals(expected), (var, axis)


def test_derivative_known_answer_uniform():
    """test_metrics_ops.py:138-179: 4x4 `foo`, dx = dy = 10, periodic."""
    foo = np.array([[1.0, 2, 4, 3], [4, 7, 1, 2], [3, 3, 0, 9], [8, 5, 1, 1]])
    dx = 10.0
    ds = xg.Dataset(
        data_vars={"foo": (("YC", "XC"), foo)},
        coords={"XC": np.arange(4) + 0.5, "XG": np.arange(4.0), "YC": np.arange(4) + 0.5, "YG": np.arange(4.0),
                "dXC": (("XC",), np.full(4, dx)), "dXG": (("XG",), np.full(4, dx)),
                "dYC": (("YC",), np.full(4, dx)), "dYG": (("YG",), np.full(4, dx))},
    )
    grid = xg.Grid(ds, coords={"X": {"center": "XC", "left": "XG"}, "Y": {"center": "YC", "left": "YG"}},
                   metrics={("X",): ["dXC", "dXG"], ("Y",): ["dYC", "dYG"]}, padding="periodic")
    dfoo_dx = grid.derivative(ds["foo"], "X")
    np.testing.assert_array_equal(dfoo_dx.values, (foo - np.roll(foo, 1, axis=1)) / dx)
    dfoo_dy = grid.derivative(ds["foo"], "Y")
    np.testing.assert_array_equal(dfoo_dy.values, (foo - np.roll(foo, 1, axis=0)) / dx)


@pytest.mark.parametrize("dtype,rtol", [(np.float64, 1e-12), (np.float32, 1e-6)])
def test_integrate_average_cumint(dtype, rtol):
    """test_metrics_ops.py:256-398: (da*metric).sum(dim), weighted mean, cumsum(da*metric)."""
    ds, gc, metrics = grid_metric_dataset("C", dtype=dtype)
    grid = xg.Grid(ds, coords=gc, metrics=metrics, padding="fill")
    for var, axis, dims in [("tracer", "X", ["xt"]), ("u", "X", ["xu"]), ("tracer", "Y", ["yt"]), ("tracer", "Z", ["zt"]),
                            ("wt", "Z", ["zw"]), ("tracer", ["X", "Y"], ["xt", "yt"]), ("tracer", ["X", "Y", "Z"], ["xt", "yt", "zt"])]:
        da = ds[var]
        metric = grid.get_metric(da, axis)
        prod = da.values * metric.broadcast_like(da).transpose(*da.dims).values
        axes = tuple(da.get_axis_num(d) for d in dims)
        got = grid.integrate(da, axis)
        assert got.dims == tuple(d for d in da.dims if d not in dims)
        np.testing.assert_allclose(got.values, prod.sum(axis=axes), rtol=rtol * 10)
        got = grid.average(da, axis)
        w = metric.broadcast_like(da).transpose(*da.dims).values
        np.testing.assert_allclose(got.values, prod.sum(axis=axes) / w.sum(axis=axes), rtol=rtol * 10)
    got = grid.cumint(ds["tracer"], "Z", to="right")
    metric = grid.get_metric(ds["tracer"], "Z")
    want = np.cumsum(ds["tracer"].values * metric.transpose(*ds["tracer"].dims).values, axis=3)
    np.testing.assert_array_equal(got.values, want.astype(dtype))
    assert got.dims == ("xt", "yt", "time", "zw")


def test_integrate_strided_axis_is_bit_exact_with_numpy():
    """np.sum along a non-contiguous axis is sequential: config-3 style integrate('Z') on (Z, Y, X)."""
    rng = np.random.default_rng(3)
    nz, ny, nx = 20, 12, 16
    a = rng.random((nz, ny, nx)).astype(np.float32)
    dz = (10 * 1.05 ** np.arange(nz)).astype(np.float32)
    ds = xg.Dataset(data_vars={"t": (("z", "y", "x"), a)}, coords={"z": np.arange(nz) + 0.5, "dz": (("z",), dz)})
    grid = xg.Grid(ds, coords={"Z": {"center": "z"}}, metrics={("Z",): ["dz"]})
    got = grid.integrate(ds["t"], "Z")
    np.testing.assert_array_equal(got.values, np.nansum(a * dz[:, None, None], axis=0))


def test_average_unmatched_missing():
    """test_metrics_ops.py:98-121: NaNs are skipped together with their weights."""
    x = np.arange(10.0)
    data = np.ones(10)
    ds = xg.Dataset(data_vars={"data": (("x",), data)}, coords={"x": x, "weights": (("x",), data * 30)})
    grid = xg.Grid(ds, coords={"X": {"center": "x"}}, metrics={"X": ["weights"]})
    expected = grid.average(ds["data"], "X")
    masked = data.copy()
    masked[6:8] = np.nan
    got = grid.average(xg.DataArray(masked, dims=("x",)), "X")
    np.testing.assert_allclose(got.values, expected.values)
    assert float(got.values) == 1.0


def test_get_metric_conditions():
    """xgcm/test/test_metrics.py:160-326: exact match, interpolation (warns), product, KeyError."""
    ds, gc, metrics = grid_metric_dataset("C")
    grid = xg.Grid(ds, coords=gc, metrics=metrics, padding="extend")
    m = grid.get_metric(ds["tracer"], ("X", "Y"))
    np.testing.assert_array_equal(m.values, ds["area_t"].values)  # condition 1
    sub = {("X",): metrics[("X",)], ("Y",): metrics[("Y",)]}
    grid2 = xg.Grid(ds, coords=gc, metrics=sub, padding="extend")
    m = grid2.get_metric(ds["tracer"], ("X", "Y"))  # condition 3: dx_t * dy_t
    np.testing.assert_array_equal(m.values, ds["dx_t"].values * ds["dy_t"].values)
    grid3 = xg.Grid(ds, coords=gc, metrics={("X", "Y"): ["area_t"]}, padding="extend")
    with pytest.warns(UserWarning, match="being interpolated"):
        m = grid3.get_metric(ds["u"], ("X", "Y"))  # condition 2: interp area_t to (xu, yt)
    want = oracle.stencil2("interp", ds["area_t"].values, 0, 0, 1, "extend")
    np.testing.assert_array_equal(m.values, want)
    with pytest.raises(KeyError):
        grid3.get_metric(ds["tracer"], ("Z",))


# --------------------------------------------------------------------------- user-defined grid ufuncs
def test_custom_grid_ufunc_pads_on_device_and_calls_user_function():
    """docs/grid_ufuncs.md: as_grid_ufunc decorator with padding_width; user code sees numpy."""
    ds, gc = all_positions_1d(9)
    grid = xg.Grid(ds, coords=gc, padding="periodic")

    @xg.as_grid_ufunc(signature="(X:center)->(X:left)", padding_width={"X": (1, 0)})
    def my_diff(a):
        assert isinstance(a, np.ndarray)
        return a[..., 1:] - a[..., :-1]

    a = ds["a_c"].values
    got = my_diff(grid, ds["a_c"], axis=[("X",)])
    assert got.dims == ("x_g",)
    np.testing.assert_array_equal(got.values, a - np.roll(a, 1))
    got = grid.apply_as_grid_ufunc(lambda a: (a[..., 1:] + a[..., :-1]) / 2, ds["a_c"], axis=[("X",)],
                                   signature="(X:center)->(X:left)", padding_width={"X": (1, 0)}, padding="fill", fill_value=4.0)
    np.testing.assert_array_equal(got.values, oracle.stencil2("interp", a, 0, 1, 0, "fill", 4.0))

    @xg.as_grid_ufunc(signature="(X:center)->(X:left)", padding_width={"X": (1, 0)})
    def wrong_length(a):
        return a
    with pytest.raises(ValueError, match="correctly trim"):
        wrong_length(grid, ds["a_c"], axis=[("X",)])


def test_gridops_raw_ufunc_attribute():
    """The plugin seam: gridops.<name>.ufunc / .signature / .padding_width."""
    from xgcm_b200 import gridops

    gu = gridops.diff_center_to_left
    assert str(gu.signature) == "(X:center)->(X:left)" and gu.padding_width == {"X": (1, 0)}
    a = np.random.default_rng(0).random((3, 10))
    np.testing.assert_array_equal(gu.ufunc(a), a[..., 1:] - a[..., :-1])
    np.testing.assert_array_equal(gridops.interp_center_to_outer.ufunc(a), (a[..., :-1] + a[..., 1:]) / 2.0)
    with pytest.raises(NotImplementedError):
        gridops.diff_left_to_inner.ufunc(a)
