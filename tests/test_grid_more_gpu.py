"""More of the reference's Grid-level tests (xgcm/test/test_grid.py:299-1010), transcribed: kwarg
handling, boundary defaults, coordinate bookkeeping, vector-input validation, 2-D vector wrappers,
interp_like, cumsum / cumint ``reverse`` forms.  The same bodies run on CPU against the mock
backend (tests/test_host_logic.py)."""

import warnings

import numpy as np
import pytest

import xgcm_b200 as xg
from xgcm_b200 import apply_as_grid_ufunc

from _fixtures import grid_metric_dataset

pytestmark = pytest.mark.gpu


def _ds_1d_left(n=9):
    x_c = np.arange(n) + 0.5
    return xg.Dataset(
        data_vars={"data_c": (("XC",), np.sin(2 * np.pi * x_c / n) + 2.0), "data_g": (("XG",), np.cos(np.arange(n) + 0.0))},
        coords={"XC": x_c, "XG": np.arange(n) + 0.0},
    )


COORDS_1D = {"X": {"center": "XC", "left": "XG"}}


def _metric_grid(grid_type="C", **kw):
    ds, coords, metrics = grid_metric_dataset(grid_type)
    return ds, coords, metrics, xg.Grid(ds, coords=coords, autoparse_metadata=False, **kw)


def test_cumsum_reverse_per_axis_dict_and_cumint_reverse():
    """test_grid.py:299-370"""
    ds, coords, metrics, _ = _metric_grid()
    grid = xg.Grid(ds, coords=coords, metrics=metrics, padding="fill", autoparse_metadata=False)
    da = ds["tracer"]
    result = grid.cumsum(da, ["X", "Y"], padding="fill", reverse={"X": True, "Y": False})
    expected = grid.cumsum(grid.cumsum(da, "X", padding="fill", reverse=True), "Y", padding="fill", reverse=False)
    np.testing.assert_array_equal(result.values, expected.values)
    result_all = grid.cumsum(da, ["X", "Y"], padding="fill", reverse=True)
    expected_all = grid.cumsum(grid.cumsum(da, "X", padding="fill", reverse=True), "Y", padding="fill", reverse=True)
    np.testing.assert_array_equal(result_all.values, expected_all.values)
    np.testing.assert_array_equal(grid.cumsum(da, "X", padding="fill").values,
                                  grid.cumsum(da, "X", padding="fill", reverse=False).values)
    # cumint forwards `reverse` (test_grid.py:329-349)
    weight = grid.get_metric(da, ("X",))
    expected = grid.cumsum(da * weight, "X", padding="fill", reverse=True)
    result = grid.cumint(da, "X", padding="fill", reverse=True)
    np.testing.assert_allclose(result.values, expected.values, rtol=1e-12)
    assert not np.allclose(result.values, grid.cumint(da, "X", padding="fill").values)
    for fn in (grid.cumsum, grid.cumint):
        with pytest.raises(ValueError, match="reverse.*not being cumulatively summed"):
            fn(da, "X", padding="fill", reverse={"X": True, "Y": False})


@pytest.mark.parametrize("func", ["diff_2d_vector", "interp_2d_vector"])
@pytest.mark.parametrize("padding", ["fill", "extend"])
def test_2d_vector_dict_input_no_face_connections(func, padding):
    """test_grid.py:399-430 (GH #581): the vector wrappers on a simply connected grid equal the
    scalar operators on each component."""
    ds, coords, _, _ = _metric_grid()
    grid = xg.Grid(ds, coords=coords, padding="periodic", autoparse_metadata=False)
    scalar = getattr(grid, func.replace("_2d_vector", ""))
    expected = {"X": scalar(ds["u"], "X", padding=padding), "Y": scalar(ds["v"], "Y", padding=padding)}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", DeprecationWarning)
        result = getattr(grid, func)({"X": ds["u"], "Y": ds["v"]}, padding=padding)
    for axis, component in result.items():
        assert component.dims == expected[axis].dims
        np.testing.assert_array_equal(component.values, expected[axis].values)


def test_grid_kwargs_dict_and_invalid_values():
    """test_grid.py:433-483,555-568"""
    ds = _ds_1d_left()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", DeprecationWarning)
        grid_direct = xg.Grid(ds, coords=COORDS_1D, padding="fill", fill_value=5, autoparse_metadata=False)
        grid_dict = xg.Grid(ds, coords=COORDS_1D, padding={"X": "fill"}, fill_value={"X": 5}, autoparse_metadata=False)
    assert grid_direct.axes["X"].fill_value == grid_dict.axes["X"].fill_value == 5
    assert grid_direct.axes["X"].padding == grid_dict.axes["X"].padding == "fill"
    for bad in ("bad", {"X": "bad"}, {"X": 0}, 0):
        with pytest.raises(ValueError):
            xg.Grid(ds, coords=COORDS_1D, padding=bad, autoparse_metadata=False)
    with pytest.raises(ValueError, match="periodic.*has been removed"):
        xg.Grid(ds, coords=COORDS_1D, periodic=True, autoparse_metadata=False)
    with pytest.raises(ValueError, match="padding='periodic'"):
        xg.Grid(ds, coords=COORDS_1D, periodic=False, autoparse_metadata=False)
    with pytest.raises(TypeError, match="unexpected keyword"):
        xg.Grid(ds, coords=COORDS_1D, not_a_real_kwarg=True, autoparse_metadata=False)
    for bad in ("bad", {"X": "bad"}):
        with pytest.raises(TypeError), warnings.catch_warnings():
            warnings.simplefilter("ignore", DeprecationWarning)
            xg.Grid(ds, coords=COORDS_1D, fill_value=bad, autoparse_metadata=False)


def test_default_boundary_is_not_periodic_and_declared_fill_does_not_wrap():
    """test_grid.py:485-525 (GH #509 / #604 / #624)"""
    ds = _ds_1d_left()
    grid = xg.Grid(ds, coords=COORDS_1D, autoparse_metadata=False)
    assert grid.axes["X"].padding is None and grid.axes["X"].periodic is False
    with pytest.raises(ValueError, match="No boundary condition was specified"):
        grid.diff(ds["data_c"], "X")
    diff_fill = xg.Grid(ds, coords=COORDS_1D, padding="fill", autoparse_metadata=False).diff(ds["data_c"], "X")
    diff_periodic = xg.Grid(ds, coords=COORDS_1D, padding="periodic", autoparse_metadata=False).diff(ds["data_c"], "X")
    assert not np.allclose(diff_fill.values, diff_periodic.values)
    np.testing.assert_array_equal(diff_fill.values[0], ds["data_c"].values[0])


def test_cumsum_nonperiodic_does_not_wrap():
    """test_grid.py:528-552 (GH #625)"""
    ds = xg.Dataset(coords={"zl": np.arange(1.0, 15.0), "zi": np.arange(0.5, 15.5)})
    coords = {"Z": {"center": "zl", "outer": "zi"}}
    zl = ds["zl"]
    with pytest.raises(ValueError, match="No boundary condition was specified"):
        xg.Grid(ds, coords=coords, autoparse_metadata=False).cumsum(zl, "Z")
    result = xg.Grid(ds, coords=coords, padding="fill", autoparse_metadata=False).cumsum(
        zl, "Z", padding="fill", fill_value=0.0)
    np.testing.assert_array_equal(result.values, np.hstack([0.0, np.cumsum(zl.values)]))


def test_keep_coords_removed():
    """test_grid.py:614-644 (GH #382 / #696)"""
    ds, coords, metrics, _ = _metric_grid("B")
    grid = xg.Grid(ds, coords=coords, metrics=metrics, padding="periodic", autoparse_metadata=False)
    for axis_name in grid.axes:
        with pytest.raises(ValueError, match="has been removed"):
            grid.diff(ds["tracer"], axis_name, keep_coords=False)
        with pytest.raises(ValueError, match="has been removed"):
            grid.cumsum(ds["tracer"], axis_name, keep_coords=False)
    with pytest.raises(ValueError, match="has been removed"):
        apply_as_grid_ufunc(lambda x: x, ds["tracer"], axis=[("X",)], grid=grid,
                            signature="(X:center)->(X:center)", keep_coords=False)


@pytest.mark.parametrize("funcname", ["interp", "diff", "cumsum"])
def test_preserve_input_noncore_coords(funcname):
    """test_grid.py:647-764 (GH #496 / #575): coordinates the user set on the INPUT survive for
    non-core dims; the shifted core dim takes its coordinate from the grid; coordinates living on
    the consumed core dim disappear."""
    N = 8
    ds = xg.Dataset(
        data_vars={"v": (("time", "XC"), np.random.default_rng(0).random((N, N)))},
        coords={"XC": np.arange(N) + 0.5, "XG": np.arange(N) + 0.0, "time": np.arange(N) * 600.0,
                "t_label": (("time",), np.arange(N) + 0.0), "xc_aux": (("XC",), np.arange(N) * 10.0)},
    )
    grid = xg.Grid(ds, coords=COORDS_1D, padding="periodic", autoparse_metadata=False)
    new_time = (np.arange(N) * 600 / 3600.0).astype(np.float32)
    new_t_label = (np.arange(N) + 100).astype(np.float32)
    v = ds["v"].assign_coords(time=new_time, t_label=(("time",), new_t_label),
                              xc_aux=(("XC",), (np.arange(N) + 500).astype(np.float32)))
    out = grid.cumsum(v, "X", to="left") if funcname == "cumsum" else getattr(grid, funcname)(v, "X")
    assert out.coords["time"].values.dtype == np.float32
    np.testing.assert_array_equal(out.coords["time"].values, new_time)
    assert "t_label" in out.coords and out.coords["t_label"].values.dtype == np.float32
    np.testing.assert_array_equal(out.coords["t_label"].values, new_t_label)
    np.testing.assert_array_equal(out.coords["XG"].values, ds["XG"].values)
    assert "XC" not in out.dims and "xc_aux" not in out.coords


def test_boundary_kwarg_same_as_grid_constructor_kwarg():
    """test_grid.py:767-782"""
    ds, coords, _, _ = _metric_grid()
    grid1 = xg.Grid(ds, coords=coords, autoparse_metadata=False)
    grid2 = xg.Grid(ds, coords=coords, padding={"X": "fill", "Y": "fill"}, autoparse_metadata=False)
    actual1 = grid1.interp(ds["tracer"], ("X", "Y"), padding={"X": "fill", "Y": "fill"})
    actual2 = grid2.interp(ds["tracer"], ("X", "Y"))
    assert actual1.dims == actual2.dims
    np.testing.assert_array_equal(actual1.values, actual2.values)


@pytest.mark.parametrize("metric_axes,metric_name", [(["Y", "X"], "area_n"), ("X", "dx_t"), ("Y", "dy_ne"),
                                                     (["Y", "X"], "dy_n"), (["X"], "tracer")])
@pytest.mark.parametrize("padding", [{"X": "fill", "Y": "fill"}, "extend", {"X": "extend", "Y": "fill"}])
@pytest.mark.parametrize("fill_value", [None, 0.1])
def test_interp_like(metric_axes, metric_name, padding, fill_value):
    """test_grid.py:785-830"""
    ds, coords, _, _ = _metric_grid()
    grid = xg.Grid(ds, coords=coords, padding="periodic", autoparse_metadata=False)
    grid.set_metrics(metric_axes, metric_name)
    axes_key = frozenset([metric_axes] if isinstance(metric_axes, str) else metric_axes)
    metric_available = grid._metrics[axes_key][0]
    interp_metric = grid.interp_like(metric_available, ds["u"], padding=padding, fill_value=fill_value)
    expected_metric = grid.interp(ds[metric_name], metric_axes, padding=padding, fill_value=fill_value)
    assert interp_metric.dims == expected_metric.dims
    np.testing.assert_allclose(interp_metric.values, expected_metric.values, rtol=1e-12)


@pytest.mark.parametrize("funcname", ["interp", "diff", "min", "max", "cumsum", "derivative", "cumint"])
@pytest.mark.parametrize("padding", ["fill", "extend"])
@pytest.mark.parametrize("fill_value", [0, 10, None])
def test_boundary_global_input(funcname, padding, fill_value):
    """test_grid.py:851-893: padding / fill_value given to the Grid == given to the method."""
    ds, coords, metrics, _ = _metric_grid()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", DeprecationWarning)
        grid_global = xg.Grid(ds, coords=coords, metrics=metrics, padding=padding, fill_value=fill_value,
                              autoparse_metadata=False)
    global_result = getattr(grid_global, funcname)(ds["tracer"], "X")
    grid_manual = xg.Grid(ds, coords=coords, metrics=metrics, padding=padding, autoparse_metadata=False)
    manual_result = getattr(grid_manual, funcname)(ds["tracer"], "X", padding=padding, fill_value=fill_value)
    assert global_result.dims == manual_result.dims
    np.testing.assert_array_equal(global_result.values, manual_result.values)


def test_vector_input_validation():
    """test_grid.py:896-1010: the same messages from the Grid methods and apply_as_grid_ufunc."""
    ds, coords, _, grid = _metric_grid()
    empty = xg.DataArray(np.zeros(()), dims=())
    calls = (lambda data, **kw: grid.diff(data, "X", **kw),
             lambda data, **kw: grid.apply_as_grid_ufunc(lambda x: x, data, axis="X", **kw))
    for call in calls:
        with pytest.raises(ValueError, match="Vector components provided as dictionaries should contain exactly one key/value pair"):
            call({"X": empty, "Y": empty})
        with pytest.raises(TypeError, match="All data arguments must be either a DataArray or Dictionary"):
            call("not_a_dataarray")
        with pytest.raises(TypeError, match="Dictionary inputs must have a DataArray as value. Got"):
            call({"X": "not_a_dataarray"})
        with pytest.raises(ValueError, match="Vector component with unknown axis provided. Grid has axes"):
            call({"wrong": empty})
    with pytest.raises(ValueError, match="When providing multiple input arguments, `other_component` needs to provide one dictionary per input"):
        grid.apply_as_grid_ufunc(lambda x: x, {"X": empty}, {"Y": empty}, {"Z": empty}, axis="X",
                                 other_component=[{"X": empty}, {"Y": empty}])


def test_axis_validation_at_grid_creation():
    """test_grid.py:838-848 and xgcm/test/test_grid.py:60-110: missing dims, dims reused across axes."""
    ds = xg.Dataset(data_vars={"data": (("x", "y"), np.zeros((4, 5)))}, coords={"x": np.arange(4.0), "y": np.arange(5.0)})
    msg = r"Could not find dimension `other` \(for the `center` position on axis `X`\) in input dataset."
    with pytest.raises(ValueError, match=msg):
        xg.Grid(ds, coords={"X": {"center": "other"}}, autoparse_metadata=False)


# ---------------------------------------------------------------- metrics (xgcm/test/test_metrics_ops.py)
def _run_single_derivative_test(grid, axis, fld, dx):
    """test_metrics_ops.py:125-131: derivative == diff / THE NAMED metric, bit for bit."""
    dvar_dx = grid.derivative(fld, axis)
    expected = grid.diff(fld, axis) / dx
    assert dvar_dx.dims == expected.dims
    np.testing.assert_array_equal(dvar_dx.values, expected.transpose(*dvar_dx.dims).values)


@pytest.mark.parametrize("grid_type", ["C", "B"])
def test_derivative_picks_the_named_metric(grid_type):
    """test_metrics_ops.py:180-252: which metric array ``derivative`` divides by, per variable
    position and axis, on C- and B-grids."""
    ds, coords, metrics, _ = _metric_grid(grid_type)
    grid = xg.Grid(ds, coords=coords, metrics=metrics, padding="periodic", autoparse_metadata=False)
    table = {
        "C": {"tracer": ["dx_e", "dy_n", "dz_w"], "u": ["dx_t", "dy_ne", "dz_w_e"],
              "v": ["dx_ne", "dy_t", "dz_w_n"], "wt": ["dx_e", "dy_n", "dz_t"]},
        "B": {"tracer": ["dx_e", "dy_n", "dz_w"], "u": ["dx_n", "dy_e", "dz_w_ne"],
              "v": ["dx_n", "dy_e", "dz_w_ne"], "wt": ["dx_e", "dy_n", "dz_t"]},
    }[grid_type]
    for var, names in table.items():
        for ax, dx in zip(["X", "Y", "Z"], names):
            _run_single_derivative_test(grid, ax, ds[var], ds[dx])


def _expected_result(da, metric, grid, dims, axes, funcname, padding=None):
    """test_metrics_ops.py:255-265"""
    prod = da * metric
    if funcname == "integrate":
        return prod.sum(dims)
    if funcname == "average":
        return prod.sum(dims) / metric.sum(dims)
    return grid.cumsum(prod, axes, padding=padding)


@pytest.mark.parametrize("funcname", ["integrate", "average", "cumint"])
@pytest.mark.parametrize("padding", ["fill", "extend"])
@pytest.mark.parametrize("padding_init", ["fill", "periodic", {"X": "periodic", "Y": "fill"}, {"X": "fill", "Y": "periodic"}])
@pytest.mark.parametrize("grid_type", ["B", "C"])
def test_metric_reductions_on_every_position(grid_type, funcname, padding, padding_init):
    """test_metrics_ops.py:268-398: integrate / average / cumint with the metric named by the
    reference for tracer, u and v points; list and tuple axis arguments."""
    ds, coords, metrics, _ = _metric_grid(grid_type)
    grid = xg.Grid(ds, coords=coords, metrics=metrics, padding=padding_init, autoparse_metadata=False)
    kwargs = dict(padding=padding) if funcname == "cumint" else {}
    func = getattr(grid, funcname)
    cases = [(ds["tracer"], ["X", "Y", "Z", ["X", "Y"], ["X", "Y", "Z"]],
              ["dx_t", "dy_t", "dz_t", "area_t", "volume_t"],
              ["xt", "yt", "zt", ["xt", "yt"], ["xt", "yt", "zt"]])]
    if grid_type == "B":
        for v in ("u", "v"):
            cases.append((ds[v], ["X", "Y", ["X", "Y"]], ["dx_ne", "dy_ne", "area_ne"], ["xu", "yu", ["xu", "yu"]]))
    else:
        cases.append((ds["u"], ["X", "Y", ["X", "Y"]], ["dx_e", "dy_e", "area_e"], ["xu", "yt", ["xu", "yt"]]))
        cases.append((ds["v"], ["X", "Y", ["X", "Y"]], ["dx_n", "dy_n", "area_n"], ["xt", "yu", ["xt", "yu"]]))
    for da, axes, metric_names, dims in cases:
        for axis, metric_name, dim in zip(axes, metric_names, dims):
            expected = _expected_result(da, ds[metric_name], grid, dim, axis, funcname, **kwargs)
            for ax_arg in ([axis, tuple(axis)] if isinstance(axis, list) else [axis]):
                new = func(da, ax_arg, **kwargs)
                assert set(new.dims) == set(expected.dims)
                np.testing.assert_allclose(new.values, expected.transpose(*new.dims).values, rtol=1e-12)


@pytest.mark.parametrize("funcname", ["integrate", "average", "cumint"])
@pytest.mark.parametrize("axis", ["X", "Y", "Z"])
def test_missingaxis(axis, funcname):
    """test_metrics_ops.py:400-447: an application axis the grid does not have is a KeyError."""
    ds, coords, metrics, _ = _metric_grid("C")
    coords = {k: v for k, v in coords.items() if k != axis}
    metrics = {k: v for k, v in metrics.items() if axis not in k}
    grid = xg.Grid(ds, coords=coords, metrics=metrics, padding="fill", autoparse_metadata=False)
    kwargs = dict(padding="fill") if funcname == "cumint" else {}
    with pytest.raises(KeyError, match="Did not find axis"):
        getattr(grid, funcname)(ds["tracer"], ["X", "Y", "Z"], **kwargs)


@pytest.mark.parametrize("funcname", ["integrate", "average", "cumint"])
def test_metric_axes_missing_from_array(funcname):
    """test_metrics_ops.py:449-479: the array lost the dim of an application axis."""
    ds, coords, metrics, _ = _metric_grid("C")
    grid = xg.Grid(ds, coords=coords, metrics=metrics, padding="fill", autoparse_metadata=False)
    kwargs = dict(padding="fill") if funcname == "cumint" else {}
    reduced = ds["tracer"].mean("xt")
    for axes in ("X", ["X", "Y", "Z"]):
        with pytest.raises(ValueError, match="Did not find single matching dimension"):
            getattr(grid, funcname)(reduced, axes, **kwargs)


# ---------------------------------------------------------------- metric selection (xgcm/test/test_metrics.py)
def _same(a, b):
    assert set(a.dims) == set(b.dims)
    np.testing.assert_allclose(a.values, b.transpose(*a.dims).values, rtol=1e-12)


def test_multiple_metrics_per_axis_and_2d_grid():
    """test_metrics.py:14-86"""
    dx, dy, area, ny, nx = 10.0, 11.0, 120.0, 7, 9
    ds = xg.Dataset(
        data_vars={"foo": (("XC",), np.array([1.0, 2.0, 4.0, 3.0])), "bar": (("XG",), np.array([10.0, 20.0, 30.0, 40.0]))},
        coords={"XC": np.arange(4) + 0.5, "XG": np.arange(4.0), "dXC": (("XC",), np.full(4, dx)), "dXG": (("XG",), np.full(4, dx))},
    )
    grid = xg.Grid(ds, coords={"X": {"center": "XC", "left": "XG"}}, metrics={("X",): ["dXC", "dXG"]},
                   padding="periodic", autoparse_metadata=False)
    assert grid.get_metric(ds["foo"], ("X",)).dims == ("XC",)
    assert grid.get_metric(ds["bar"], ("X",)).dims == ("XG",)
    ds = xg.Dataset(
        data_vars={"foo": (("YC", "XC"), np.ones((ny, nx)))},
        coords={"XC": np.arange(nx) + 0.0, "dX": (("XC",), np.full(nx, dx)), "YC": np.arange(ny) + 0.0,
                "dY": (("YC",), np.full(ny, dy)), "area": (("YC", "XC"), np.full((ny, nx), area))},
    )
    coords = {"X": {"center": "XC"}, "Y": {"center": "YC"}}
    grid = xg.Grid(ds, coords=coords, metrics={("X",): ["dX"], ("Y",): ["dY"], ("X", "Y"): ["area"]}, autoparse_metadata=False)
    _same(grid.get_metric(ds["foo"], ("X",)), ds["dX"])
    _same(grid.get_metric(ds["foo"], ("Y",)), ds["dY"])
    _same(grid.get_metric(ds["foo"], ("X", "Y")), ds["area"])
    _same(grid.get_metric(ds["foo"], ("Y", "X")), ds["area"])
    grid = xg.Grid(ds, coords=coords, metrics={("X",): ["dX"], ("Y",): ["dY"]}, autoparse_metadata=False)
    actual = grid.get_metric(ds["foo"], ("Y", "X")).transpose("YC", "XC")
    np.testing.assert_array_equal(actual.values, np.full((ny, nx), dx * dy))


@pytest.mark.parametrize("key, metric_vars", [(("X",), ["dx_t"]), ("X", "dx_t"), (("X", "Y"), ["area_t"]),
                                              (("X", "Y"), ["area_t", "area_e", "area_n", "area_ne"]),
                                              (("X", "Y", "Z"), ["volume_t"])])
def test_assign_metric(key, metric_vars):
    """test_metrics.py:89-107"""
    ds, coords, _, _ = _metric_grid()
    xg.Grid(ds, coords=coords, metrics={key: metric_vars}, autoparse_metadata=False)


def test_iterate_axis_combinations():
    """test_metrics.py:110-150"""
    from xgcm_b200.metrics import iterate_axis_combinations

    fs = frozenset
    expected = {
        ("X", "Y"): [(fs("XY"),), (fs("X"), fs("Y"))],
        ("X", "Y", "Z"): [(fs("XYZ"),), (fs("X"), fs("Y"), fs("Z")), (fs("YZ"), fs("X")), (fs("XY"), fs("Z")), (fs("XZ"), fs("Y"))],
    }
    for axes, want in expected.items():
        actual = {frozenset(a) for a in iterate_axis_combinations(axes)}
        assert actual == {frozenset(w) for w in want}


@pytest.mark.parametrize("axes, data_var, metric_expected", [("X", "tracer", "dx_t"), (["X", "Y"], "tracer", "area_t"),
                                                             (("X", "Y"), "tracer", "area_t"), (["X", "Y", "Z"], "tracer", "volume_t"),
                                                             (["X"], "u", "dx_e"), (["X", "Y"], "u", "area_e")])
def test_get_metric_orig(axes, data_var, metric_expected):
    """test_metrics.py:153-172"""
    ds, coords, metrics, _ = _metric_grid()
    grid = xg.Grid(ds, coords=coords, metrics=metrics, autoparse_metadata=False)
    _same(grid.get_metric(ds[data_var], axes), ds[metric_expected])


def test_get_metric_with_conditions():
    """test_metrics.py:175-285: the four selection rules of get_metric (grid.py:534-657).  Metrics
    are interpolated with ``extend`` whatever the grid's padding (grid.py:591-593,644-648); the
    reference's test compares with the grid's own padding and only passes because its metrics are
    spatially uniform — the fixture here is not, so the expectation names ``extend``."""
    ds, coords, metrics, _ = _metric_grid()
    # 1: a metric on the right axes and position exists
    grid = xg.Grid(ds, coords=coords, metrics=metrics, autoparse_metadata=False)
    _same(grid.get_metric(ds["v"], ("X", "Y")), ds["area_n"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)
        # 2a / 2b: interpolate the metric with matching AXES, even if others sit on the position
        grid = xg.Grid(ds, coords=coords, padding="extend", autoparse_metadata=False)
        grid.set_metrics(("X", "Y"), "area_e")
        _same(grid.get_metric(ds["v"], ("X", "Y")), grid.interp(ds["area_e"], ("X", "Y")))
        grid = xg.Grid(ds, coords=coords, padding="periodic", autoparse_metadata=False)
        grid.set_metrics(("X", "Y"), "area_e")
        grid.set_metrics(("X"), "dx_n")
        grid.set_metrics(("Y"), "dx_n")
        _same(grid.get_metric(ds["v"], ("X", "Y")), grid.interp(ds["area_e"], ("X", "Y"), padding="extend"))
        # 3a / 3b: multiply metrics that sit on the right position
        grid = xg.Grid(ds, coords=coords, autoparse_metadata=False)
        grid.set_metrics(("X"), "dx_n")
        grid.set_metrics(("Y"), "dy_n")
        _same(grid.get_metric(ds["v"], ("X", "Y")), ds["dx_n"] * ds["dy_n"])
        grid = xg.Grid(ds, coords=coords, autoparse_metadata=False)
        grid.set_metrics(("X", "Y"), "area_t")
        grid.set_metrics(("Z"), "dz_t")
        _same(grid.get_metric(ds["tracer"], ("X", "Y", "Z")), ds["area_t"] * ds["dz_t"])
        # 4a / 4b: interpolate one or both factors first
        grid = xg.Grid(ds, coords=coords, padding="periodic", autoparse_metadata=False)
        grid.set_metrics(("X"), "dx_t")
        grid.set_metrics(("Y"), "dy_n")
        _same(grid.get_metric(ds["v"], ("X", "Y")), grid.interp(ds["dx_t"], "Y", padding="extend") * ds["dy_n"])
        grid = xg.Grid(ds, coords=coords, padding="periodic", autoparse_metadata=False)
        grid.set_metrics(("X"), "dx_t")
        grid.set_metrics(("Y"), "dy_t")
        _same(grid.get_metric(ds["v"], ("X", "Y")),
              grid.interp(ds["dx_t"], "Y", padding="extend") * grid.interp(ds["dy_t"], "Y", padding="extend"))


@pytest.mark.parametrize("data_var, z_metrics, expected_dz", [
    ("tracer", ["dz_w", "dz_w_ne", "dz_w_n", "dz_w_e", "dz_t"], "dz_t"), ("wt", ["dz_t", "dz_w"], "dz_w")])
def test_get_metric_no_spurious_interpolation_warning(data_var, z_metrics, expected_dz):
    """test_metrics.py:288-324 (GH #756)"""
    ds, coords, _, _ = _metric_grid()
    metrics = {("X", "Y"): ["area_t", "area_n", "area_e", "area_ne"], ("Z",): z_metrics}
    grid = xg.Grid(ds, coords=coords, metrics=metrics, autoparse_metadata=False)
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        metric = grid.get_metric(ds[data_var], ("X", "Y", "Z"))
    assert [str(w.message) for w in caught if "being interpolated" in str(w.message)] == []
    _same(metric, ds["area_t"] * ds[expected_dz])


def test_set_metric_and_overwrite_and_errors():
    """test_metrics.py:327-460"""
    ds, coords, metrics, _ = _metric_grid()
    grid = xg.Grid(ds, coords=coords, metrics=metrics, autoparse_metadata=False)
    grid_manual = xg.Grid(ds, coords=coords, autoparse_metadata=False)
    for key, value in metrics.items():
        grid_manual.set_metrics(key, value)
    for k, names in metrics.items():
        for g in (grid, grid_manual):
            assert [m.name for m in g._metrics[frozenset(k)]] == list(names)
            for name, m in zip(names, g._metrics[frozenset(k)]):
                np.testing.assert_array_equal(m.values, ds[name].values)
    # overwrite=True replaces the metric on the same dims and appends new ones
    for metric_axes, exist, add, expected in [
        ("X", ["dx_t", "dx_n", "dx_e", "dx_ne"], ["dx_n_overwrite"], ["dx_t", "dx_n_overwrite", "dx_e", "dx_ne"]),
        (("Y", "X"), ["area_t", "area_n", "area_e", "area_ne"], ["area_n_overwrite"], ["area_t", "area_n_overwrite", "area_e", "area_ne"]),
        ("X", ["dx_t", "dx_n", "dx_e"], ["dx_n_overwrite", "dx_ne"], ["dx_t", "dx_n_overwrite", "dx_e", "dx_ne"]),
    ]:
        ds2 = ds.assign_coords({add[0]: ds[exist[1]] * 10})
        sub = {k: [m for m in v if m in exist] for k, v in metrics.items()}
        g = xg.Grid(ds2, coords=coords, metrics=sub, autoparse_metadata=False)
        for av in add:
            g.set_metrics(metric_axes, av, overwrite=True)
        got = g._metrics[frozenset(list(metric_axes))]
        assert len(got) == len(expected)
        for m, name in zip(got, expected):
            np.testing.assert_array_equal(m.values, ds2[name].values)
    ds3 = ds.assign_coords({"dx_t_overwrite": ds["dx_t"] * 10})
    g = xg.Grid(ds3, coords=coords, metrics=metrics, autoparse_metadata=False)
    for name in ("dx_t_overwrite", "dx_e"):
        with pytest.raises(ValueError, match="setting overwrite=True."):
            g.set_metrics("X", name)
    with pytest.raises(KeyError, match="not found in dataset."):
        grid.set_metrics("X", "foo")
    with pytest.raises(KeyError, match="not compatible with grid axes"):
        grid.set_metrics(("U", "V"), "area_n")


def test_apply_many_equals_single_calls():
    """Extension: Grid.apply_many(da, requests) == the single-operator calls (host arrays go through the
    batched slab pipeline, device arrays loop)."""
    ds, coords, metrics, _ = _metric_grid()
    grid = xg.Grid(ds, coords=coords, metrics=metrics, padding={"X": "periodic", "Y": "fill"}, autoparse_metadata=False)
    da = ds["tracer"]
    reqs = [("diff", "X"), ("interp", "X"), ("diff", "Y"), ("interp", "Y", "right"), ("min", "X"), ("max", "Y")]
    many = grid.apply_many(da, reqs)
    assert len(many) == len(reqs)
    for r, req in zip(many, reqs):
        f, ax = req[0], req[1]
        to = req[2] if len(req) > 2 else None
        want = getattr(grid, f)(da, ax, to=to)
        assert r.dims == want.dims
        np.testing.assert_array_equal(r.values, want.values)
        assert set(r.coords) == set(want.coords)
    with pytest.raises(ValueError, match="apply_many supports"):
        grid.apply_many(da, [("cumsum", "X")])
    with pytest.raises(KeyError):
        grid.apply_many(da, [("diff", "Q")])


def test_average_skipna_false_and_min_count():
    """grid.py:1680-1685 forwards **kwargs to da.weighted(w).mean: skipna=False lets NaN through;
    min_count is refused (not silently ignored)."""
    ds, coords, metrics, _ = _metric_grid()
    grid = xg.Grid(ds, coords=coords, metrics=metrics, padding="fill", autoparse_metadata=False)
    da = ds["tracer"].copy()
    vals = np.array(da.values, dtype=np.float64, copy=True)
    vals[1, 2] = np.nan  # all of (time, z) at x = 1, y = 2
    da = xg.DataArray(vals, dims=da.dims, coords=da.coords)
    for axes in ("X", ["X", "Y"]):
        skip = grid.average(da, axes).values
        keep = grid.average(da, axes, skipna=False).values
        assert not np.isnan(skip).any()
        nan_lines = np.isnan(vals).any(axis=0) if axes == "X" else np.isnan(vals).any(axis=(0, 1))
        np.testing.assert_array_equal(np.isnan(keep), nan_lines)
    with pytest.raises(NotImplementedError, match="min_count"):
        grid.integrate(da, "X", min_count=1)


def test_cumint_equals_cumsum_of_product():
    """grid.py:1656-1660: cumint == cumsum(da * metric); the product rides on the scan kernel's pre operand."""
    ds, coords, metrics, _ = _metric_grid()
    grid = xg.Grid(ds, coords=coords, metrics=metrics, padding="fill", autoparse_metadata=False)
    da = ds["tracer"]
    for axes in ("X", "Y", ["X", "Y"]):
        weight = grid.get_metric(da, (axes,) if isinstance(axes, str) else tuple(axes))
        want = grid.cumsum(da * weight, axes, padding="fill")
        got = grid.cumint(da, axes, padding="fill")
        assert got.dims == want.dims
        np.testing.assert_array_equal(got.values, want.values)


@pytest.mark.parametrize("grid_type", ["C"])
def test_pair_divergence_vorticity_equal_the_explicit_chain(grid_type):
    """Extension: Grid.pair / divergence / vorticity == the chain of Grid.diff calls and array
    arithmetic a user writes with the reference (docs/ufunc_examples.md:105-153), bit for bit."""
    ds, coords, metrics, _ = _metric_grid(grid_type)
    grid = xg.Grid(ds, coords=coords, metrics=metrics, padding={"X": "periodic", "Y": "periodic", "Z": "fill"},
                   autoparse_metadata=False)
    u, v = ds["u"], ds["v"]  # u on (xu, yt), v on (xt, yu): right-shifted C-grid
    # the innermost dim of the fixtures is z; put x last for one variant so that the fused kernel engages
    for order in (("time", "zt", None, None), None):
        if order is None:
            uu, vv = u, v
        else:
            uu = u.transpose("time", "zt", "yt", "xu")
            vv = v.transpose("time", "zt", "yu", "xt")
        dy_u = grid.get_metric(uu, ("Y",))
        dx_v = grid.get_metric(vv, ("X",))
        want = grid.diff(uu * dy_u, "X") + grid.diff(vv * dx_v, "Y")
        want = want / grid.get_metric(want, ("X", "Y"))
        got = grid.divergence(uu, vv)
        assert got.dims == want.dims
        np.testing.assert_array_equal(got.values, want.values)
        dy_v = grid.get_metric(vv, ("Y",))
        dx_u = grid.get_metric(uu, ("X",))
        zw = grid.diff(vv * dy_v, "X") - grid.diff(uu * dx_u, "Y")
        zw = zw / grid.get_metric(zw, ("X", "Y"))
        got = grid.vorticity(uu, vv)
        assert got.dims == zw.dims
        np.testing.assert_array_equal(got.values, zw.values)
        # plain pair without metrics, interp + diff, sub
        want = grid.interp(uu, "X", padding="extend") - grid.diff(vv, "Y", padding="extend")
        got = grid.pair("interp", uu, "X", "diff", vv, "Y", combine="sub", padding="extend")
        np.testing.assert_array_equal(got.values, want.values)
    with pytest.raises(ValueError, match="different positions"):
        grid.pair("diff", u, "X", "diff", u, "Y")
