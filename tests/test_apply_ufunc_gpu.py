"""``apply_as_grid_ufunc`` / ``as_grid_ufunc`` with USER functions: the orchestration row (A4) of the
hot path — validate positions, pad (on the device), hand the core dims last to the user's function,
reattach coordinates, restore the dim order — transcribed from the reference's own tests
(xgcm/test/test_grid_ufunc.py:300-1290; the dask-only cases are out of scope).

On the GPU box the padding runs in ``xg_pad``; the same bodies run on CPU against the mock backend
(tests/test_host_logic.py).
"""

import re
from typing import Annotated, Tuple

import numpy as np
import pytest

import xgcm_b200 as xg
from xgcm_b200 import GridUFunc, apply_as_grid_ufunc, as_grid_ufunc

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------- fixtures (test_grid_ufunc.py:216-297)
def _coords_1d(ax, length):
    return {
        f"{ax}_c": np.arange(1, length + 1) + 0.0,
        f"{ax}_g": np.arange(0.5, length),
        f"{ax}_r": np.arange(1.5, length + 1),
        f"{ax}_i": np.arange(1.5, length),
        f"{ax}_o": np.arange(0.5, length + 1),
    }


def _axis_coords(ax):
    return {"center": f"{ax}_c", "left": f"{ax}_g", "right": f"{ax}_r", "inner": f"{ax}_i", "outer": f"{ax}_o"}


def create_1d_test_grid(ax, length=9):
    ds = xg.Dataset(coords=_coords_1d(ax, length))
    return xg.Grid(ds, coords={ax: _axis_coords(ax)}, padding="periodic", autoparse_metadata=False)


def create_2d_test_grid(ax1, ax2, length1=9, length2=11):
    ds = xg.Dataset(coords={**_coords_1d(ax1, length1), **_coords_1d(ax2, length2)})
    return xg.Grid(ds, coords={ax1: _axis_coords(ax1), ax2: _axis_coords(ax2)}, padding="periodic",
                   autoparse_metadata=False)


def _da(grid, values, dims):
    return xg.DataArray(np.asarray(values, dtype=np.float64), dims=dims,
                        coords={d: grid._ds[d].values for d in dims if d in grid._ds.dims})


def _assert_equal(result, values, dims, grid):
    assert result.dims == tuple(dims)
    np.testing.assert_array_equal(result.values, values)
    for d in dims:
        np.testing.assert_array_equal(result.coords[d].values, grid._ds[d].values)


# ---------------------------------------------------------------- no padding
def test_stores_ufunc_kwarg_info():
    """test_grid_ufunc.py:301-317"""

    @as_grid_ufunc()
    def diff_center_to_left(a: Annotated[np.ndarray, "X:center"]) -> Annotated[np.ndarray, "X:left"]:
        return a - np.roll(a, shift=-1)

    assert isinstance(diff_center_to_left, GridUFunc)
    assert str(diff_center_to_left.signature) == "(X:center)->(X:left)"
    with pytest.raises(TypeError, match="Unsupported keyword argument"):

        @as_grid_ufunc(junk="useless")
        def other(a):
            return a


def test_1d_unchanging_size_three_call_forms():
    """test_grid_ufunc.py:345-382: direct application, Grid method, decorator."""

    def diff_center_to_left(a):
        return a - np.roll(a, shift=-1)

    grid = create_1d_test_grid("depth")
    vals = np.sin(grid._ds["depth_c"].values * 2 * np.pi / 9)
    da = _da(grid, vals, ("depth_c",))
    expected = vals - np.roll(vals, -1)

    result = apply_as_grid_ufunc(diff_center_to_left, da, axis=[("depth",)], grid=grid, signature="(X:center)->(X:left)")
    _assert_equal(result, expected, ("depth_g",), grid)
    result = grid.apply_as_grid_ufunc(diff_center_to_left, da, axis=[("depth",)], signature="(X:center)->(X:left)")
    _assert_equal(result, expected, ("depth_g",), grid)

    @as_grid_ufunc()
    def decorated(a: Annotated[np.ndarray, "X:center"]) -> Annotated[np.ndarray, "X:left"]:
        return a - np.roll(a, shift=-1)

    _assert_equal(decorated(grid, da, axis=[("depth",)]), expected, ("depth_g",), grid)


def test_apply_along_one_axis():
    """test_grid_ufunc.py:481-514: the core dim is moved last for the user's function."""

    def diff_center_to_left(a):
        return a - np.roll(a, shift=-1, axis=-1)

    grid = create_2d_test_grid("lon", "lat")
    lat, lon = grid._ds["lat_c"].values, grid._ds["lon_c"].values
    vals = lat[:, None] ** 2 + lon[None, :] ** 2
    da = _da(grid, vals, ("lat_c", "lon_c"))
    expected = vals - np.roll(vals, -1, axis=1)
    result = apply_as_grid_ufunc(diff_center_to_left, da, axis=[("lon",)], grid=grid, signature="(X:center)->(X:left)")
    _assert_equal(result, expected, ("lat_c", "lon_g"), grid)

    @as_grid_ufunc()
    def decorated(a: Annotated[np.ndarray, "X:center"]) -> Annotated[np.ndarray, "X:left"]:
        return a - np.roll(a, shift=-1, axis=-1)

    _assert_equal(decorated(grid, da, axis=[("lon",)]), expected, ("lat_c", "lon_g"), grid)


def test_preserves_input_dim_order():
    """test_grid_ufunc.py:516-563 (GH #533)."""
    nx, ny, nz = 4, 5, 6
    ds = xg.Dataset(coords={"i": np.arange(nx), "j": np.arange(ny), "jg": np.arange(ny), "k": np.arange(nz)})
    da = xg.DataArray(np.random.default_rng(0).random((nz, ny, nx)), dims=("k", "j", "i"))
    grid = xg.Grid(ds, coords={"Y": {"center": "j", "left": "jg"}}, padding="periodic", autoparse_metadata=False)
    out = grid.apply_as_grid_ufunc(lambda a: a, da, axis=[["Y"]], signature="(Y:center)->(Y:center)",
                                   padding_width={"Y": (0, 0)})
    assert out.dims == ("k", "j", "i")
    np.testing.assert_array_equal(out.values, da.values)
    out_left = grid.apply_as_grid_ufunc(lambda a: a, da, axis=[["Y"]], signature="(Y:center)->(Y:left)",
                                        padding_width={"Y": (0, 0)})
    assert out_left.dims == ("k", "jg", "i")


def test_multiple_inputs():
    """test_grid_ufunc.py:565-606: two inputs on different positions, scalar output."""

    def inner_product_left_right(a, b):
        return np.inner(a, b)

    grid = create_1d_test_grid("depth")
    a = _da(grid, np.sin(grid._ds["depth_g"].values * 2 * np.pi / 9), ("depth_g",))
    b = _da(grid, np.cos(grid._ds["depth_r"].values * 2 * np.pi / 9), ("depth_r",))
    expected = np.inner(a.values, b.values)
    for result in (
        apply_as_grid_ufunc(inner_product_left_right, a, b, axis=[("depth",), ("depth",)], grid=grid,
                            signature="(X:left),(X:right)->()"),
        grid.apply_as_grid_ufunc(inner_product_left_right, a, b, axis=[("depth",), ("depth",)],
                                 signature="(X:left),(X:right)->()"),
    ):
        assert result.dims == ()
        np.testing.assert_array_equal(result.values, expected)

    @as_grid_ufunc()
    def decorated(a: Annotated[np.ndarray, "X:left"], b: Annotated[np.ndarray, "X:right"]):
        return np.inner(a, b)

    np.testing.assert_array_equal(decorated(grid, a, b, axis=[("depth",), ("depth",)]).values, expected)


def test_multiple_outputs():
    """test_grid_ufunc.py:608-658: gradient to the inner positions of two axes."""

    def diff_center_to_inner(a, axis):
        result = a - np.roll(a, shift=1, axis=axis)
        return np.delete(result, 0, axis)

    def grad_to_inner(a):
        return diff_center_to_inner(a, axis=0), diff_center_to_inner(a, axis=1)

    grid = create_2d_test_grid("lon", "lat")
    lon, lat = grid._ds["lon_c"].values, grid._ds["lat_c"].values
    a = _da(grid, lon[:, None] ** 2 + lat[None, :] ** 2, ("lon_c", "lat_c"))
    expected_u = 2 * grid._ds["lon_i"].values[:, None] * np.ones((1, lat.size))
    expected_v = 2 * grid._ds["lat_i"].values[None, :] * np.ones((lon.size, 1))
    sig = "(X:center,Y:center)->(X:inner,Y:center),(X:center,Y:inner)"
    for u, v in (
        apply_as_grid_ufunc(grad_to_inner, a, axis=[("lon", "lat")], grid=grid, signature=sig),
        grid.apply_as_grid_ufunc(grad_to_inner, a, axis=[("lon", "lat")], signature=sig),
    ):
        _assert_equal(u, expected_u, ("lon_i", "lat_c"), grid)
        _assert_equal(v, expected_v, ("lon_c", "lat_i"), grid)

    @as_grid_ufunc()
    def decorated(a: Annotated[np.ndarray, "X:center,Y:center"]) -> Tuple[
        Annotated[np.ndarray, "X:inner,Y:center"], Annotated[np.ndarray, "X:center,Y:inner"]
    ]:
        return diff_center_to_inner(a, axis=0), diff_center_to_inner(a, axis=1)

    u, v = decorated(grid, a, axis=[("lon", "lat")])
    _assert_equal(u, expected_u, ("lon_i", "lat_c"), grid)
    _assert_equal(v, expected_v, ("lon_c", "lat_i"), grid)


# ---------------------------------------------------------------- with padding
def test_1d_padded_but_no_change_in_grid_position():
    """test_grid_ufunc.py:662-720: a width-2 halo on one side, output stays on the centers."""

    def second_order(a):
        return 0.5 * (a[..., 2:] - a[..., :-2])

    grid = create_1d_test_grid("depth")
    vals = np.sin(grid._ds["depth_c"].values * 2 * np.pi / 9)
    da = _da(grid, vals, ("depth_c",))
    expected = 0.5 * (vals - np.roll(vals, 2))
    result = apply_as_grid_ufunc(second_order, da, axis=[("depth",)], grid=grid, signature="(X:center)->(X:center)",
                                 padding_width={"X": (2, 0)})
    _assert_equal(result, expected, ("depth_c",), grid)

    @as_grid_ufunc("(X:center)->(X:center)", padding_width={"X": (2, 0)})
    def decorated(a):
        return 0.5 * (a[..., 2:] - a[..., :-2])

    _assert_equal(decorated(grid, da, axis=[("depth",)]), expected, ("depth_c",), grid)


def test_non_core_coords_survive_and_first_input_wins():
    """test_grid_ufunc.py:752-876 (GH #575): coordinates on surviving dims are kept, also through
    the ``{axis: DataArray}`` form; with several inputs the first one's values win."""

    def diff_center_to_left(a):
        return a[..., 1:] - a[..., :-1]

    grid = create_1d_test_grid("depth")
    rng = np.random.default_rng(0)
    time = np.array([10, 20, 30], dtype="float32")
    da = xg.DataArray(rng.random((3, 9)), dims=("time", "depth_c"),
                      coords={"time": time, "label": (("time",), np.array([1.0, 2.0, 3.0])),
                              "depth_c": grid._ds["depth_c"].values})
    kw = dict(axis=[("depth",)], grid=grid, signature="(X:center)->(X:left)", padding_width={"X": (1, 0)})
    for arg in (da, {"depth": da}):
        result = apply_as_grid_ufunc(diff_center_to_left, arg, **kw)
        assert result.dims == ("time", "depth_g")
        assert result.coords["time"].values.dtype == time.dtype
        np.testing.assert_array_equal(result.coords["time"].values, time)
        np.testing.assert_array_equal(result.coords["label"].values, [1.0, 2.0, 3.0])
        padded = np.concatenate([da.values[:, -1:], da.values], axis=1)
        np.testing.assert_array_equal(result.values, padded[:, 1:] - padded[:, :-1])

    b = xg.DataArray(rng.random((3, 9)), dims=("time", "depth_c"),
                     coords={"time": np.array([99, 98, 97], dtype="float32"), "depth_c": grid._ds["depth_c"].values})
    result = apply_as_grid_ufunc(lambda a, b: (a - b)[..., 1:], da, b, axis=[("depth",), ("depth",)], grid=grid,
                                 signature="(X:center),(X:center)->(X:left)", padding_width={"X": (1, 0)})
    np.testing.assert_array_equal(result.coords["time"].values, time)


def test_2d_padding():
    """test_grid_ufunc.py:878-917: vorticity of (U, V), both axes padded, dummy axis names that are
    the real ones."""

    def diff(a, axis):
        return np.apply_along_axis(lambda r: r[..., 1:] - r[..., :-1], axis, a)

    def vort(u, v):
        return diff(v[..., 1:], axis=-2) - diff(u[..., 1:, :], axis=-1)

    grid = create_2d_test_grid("lon", "lat")
    lon_g, lon_c = grid._ds["lon_g"].values, grid._ds["lon_c"].values
    lat_g, lat_c = grid._ds["lat_g"].values, grid._ds["lat_c"].values
    U = _da(grid, lon_g[:, None] ** 2 + lat_c[None, :] ** 3, ("lon_g", "lat_c"))
    V = _da(grid, lon_c[:, None] ** 3 + lat_g[None, :] ** 2, ("lon_c", "lat_g"))
    expected = (V.values - np.roll(V.values, 1, axis=0)) - (U.values - np.roll(U.values, 1, axis=1))
    result = grid.apply_as_grid_ufunc(
        vort, U, V, axis=2 * [("lon", "lat")],
        signature="(lon:left,lat:center),(lon:center,lat:left)->(lon:left,lat:left)",
        padding_width={"lon": (1, 0), "lat": (1, 0)},
    )
    _assert_equal(result, expected, ("lon_g", "lat_g"), grid)


def test_boundary_constant_and_fill_value_from_decorator():
    """test_grid_ufunc.py:1211-1273 (GH #652): decorator-bound padding / fill_value are used and can
    be overridden per call."""

    def interp(a):
        return 0.5 * (a[..., :-1] + a[..., 1:])

    grid = create_1d_test_grid("lat")
    arr = np.arange(9.0)
    da = _da(grid, arr, ("lat_c",))
    for bound in (0, 10):

        @as_grid_ufunc(signature="(X:center)->(X:left)", padding_width={"X": (1, 0)}, padding="fill", fill_value=bound)
        def interp_center_to_left(a):
            return interp(a)

        result = interp_center_to_left(grid, da, axis=[["lat"]])
        _assert_equal(result, interp(np.concatenate([[bound], arr])), ("lat_g",), grid)
        result = interp_center_to_left(grid, da, axis=[["lat"]], padding="fill", fill_value=1)
        _assert_equal(result, interp(np.concatenate([[1], arr])), ("lat_g",), grid)
        result = interp_center_to_left(grid, da, axis=[["lat"]], fill_value=1)
        _assert_equal(result, interp(np.concatenate([[1], arr])), ("lat_g",), grid)


def test_single_chunk_center_to_outer_known_answer():
    """test_grid_ufunc.py:1312-1338 (GH #518): linspace(1, 10, 10) center -> outer with extend."""
    ds = xg.Dataset(data_vars={"drF": (("Z",), np.linspace(1, 10, num=10))},
                    coords={"Z": np.arange(0.5, 10, 1), "Zp1": np.arange(11) + 0.0})
    grid = xg.Grid(ds, coords={"Z": {"center": "Z", "outer": "Zp1"}}, autoparse_metadata=False)
    result = grid.interp(ds["drF"], "Z", padding="extend", to="outer")
    assert result.dims == ("Zp1",)
    np.testing.assert_array_equal(result.values, np.concatenate(([1.0], np.linspace(1.5, 9.5, num=9), [10.0])))


# ---------------------------------------------------------------- errors
def test_wrong_positions_and_bad_returns():
    """grid_ufunc.py:827-842,954-990: inputs that do not sit on the declared positions, results
    with the wrong number of dims / outputs, sizes that do not match the grid."""
    grid = create_1d_test_grid("depth")
    da = _da(grid, np.arange(9.0), ("depth_c",))
    with pytest.raises(ValueError, match=re.escape("depth:left")):
        apply_as_grid_ufunc(lambda x: x, da, axis=[("depth",)], grid=grid, signature="(X:left)->(X:left)")
    with pytest.raises(ValueError, match="unexpected number of dimensions"):
        apply_as_grid_ufunc(lambda x: x.sum(), da, axis=[("depth",)], grid=grid, signature="(X:center)->(X:center)")
    with pytest.raises(ValueError, match="outputs"):
        apply_as_grid_ufunc(lambda x: (x, x), da, axis=[("depth",)], grid=grid, signature="(X:center)->(X:center)")
    with pytest.raises(ValueError, match="conflicting sizes"):
        apply_as_grid_ufunc(lambda x: x[..., 1:], da, axis=[("depth",)], grid=grid, signature="(X:center)->(X:center)")
    with pytest.raises(ValueError, match="Number of entries in `axis`"):
        apply_as_grid_ufunc(lambda x: x, da, axis=[("depth",), ("depth",)], grid=grid, signature="(X:center)->(X:center)")
