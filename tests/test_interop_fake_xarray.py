"""xarray is not installable in the build image, so the adapter (xgcm_b200/interop.py) is exercised
against a minimal stand-in exposing the xarray attributes it touches (dims, coords, data / values,
name, attrs, data_vars, sizes, the DataArray constructor).  Host logic only (mock backend)."""

import sys
import types

import numpy as np
import pytest
import torch

import xgcm_b200 as xg
from _mock_backend import install


class _Coord:
    def __init__(self, dims, values, attrs=None):
        self.dims, self.values, self.attrs = tuple(dims), np.asarray(values), dict(attrs or {})


class FakeDataArray:
    def __init__(self, data, dims=None, coords=None, name=None, attrs=None):
        self.data = np.asarray(data)
        self.dims = tuple(dims)
        self.name, self.attrs = name, dict(attrs or {})
        self.coords = {}
        for k, v in (coords or {}).items():
            if isinstance(v, tuple):
                self.coords[k] = _Coord(v[0], v[1], v[2] if len(v) > 2 else None)
            else:
                self.coords[k] = _Coord((k,), v)

    @property
    def values(self):
        return self.data

    @property
    def sizes(self):
        return dict(zip(self.dims, self.data.shape))


class FakeDataset:
    def __init__(self, data_vars, coords, attrs=None):
        self.data_vars = {k: FakeDataArray(v[1], dims=v[0]) for k, v in data_vars.items()}
        self.coords = {k: (_Coord((k,), v) if not isinstance(v, tuple) else _Coord(*v)) for k, v in coords.items()}
        self.attrs = dict(attrs or {})
        self.sizes = {}
        for obj in list(self.data_vars.values()) + list(self.coords.values()):
            vals = obj.data if hasattr(obj, "data") else obj.values
            self.sizes.update(dict(zip(obj.dims, vals.shape)))
        self.dims = self.sizes


FakeDataArray.__name__ = "DataArray"
FakeDataset.__name__ = "Dataset"


@pytest.fixture
def fake_xarray(monkeypatch):
    if torch.cuda.is_available():
        pytest.skip("host-logic test (mock backend)")
    install(monkeypatch)
    mod = types.ModuleType("xarray")
    mod.DataArray, mod.Dataset = FakeDataArray, FakeDataset
    monkeypatch.setitem(sys.modules, "xarray", mod)
    return mod


def test_grid_accepts_and_returns_xarray_like_objects(fake_xarray):
    xr = fake_xarray
    n = 8
    a = np.random.default_rng(0).random((3, n))
    ds = xr.Dataset({"t": (("y", "xc"), a)}, {"xc": np.arange(n) + 0.5, "xg": np.arange(n) + 0.0, "y": np.arange(3.0)})
    grid = xg.Grid(ds, coords={"X": {"center": "xc", "left": "xg"}}, padding="periodic")
    da = xr.DataArray(a, dims=("y", "xc"), coords={"y": np.arange(3.0) * 2}, name="t", attrs={"units": "K"})
    out = grid.diff(da, "X")
    assert type(out).__name__ == "DataArray" and isinstance(out, FakeDataArray)  # came back as xarray
    assert out.dims == ("y", "xg") and out.name == "t"
    np.testing.assert_array_equal(out.data, a - np.roll(a, 1, axis=1))
    np.testing.assert_array_equal(out.coords["xg"].values, np.arange(n) + 0.0)   # from the grid dataset
    np.testing.assert_array_equal(out.coords["y"].values, np.arange(3.0) * 2)    # the user's non-core coord
    # native labelled arrays still come back native
    native = grid.interp(xg.DataArray(a, dims=("y", "xc")), "X")
    assert isinstance(native, xg.DataArray)
    out = grid.cumsum(da, "X", to="left")
    assert isinstance(out, FakeDataArray) and out.dims == ("y", "xg")


def test_vector_dict_and_other_component_convert_xarray_values(fake_xarray):
    """`{axis: xr.DataArray}` vector inputs and `other_component` dicts holding xarray objects are
    converted like a bare DataArray (they used to fail the input type check)."""
    xr = fake_xarray
    n = 8
    rng = np.random.default_rng(1)
    u, v = rng.random((n, n)), rng.random((n, n))
    ds = xr.Dataset({}, {"xc": np.arange(n) + 0.5, "xg": np.arange(n) + 0.0, "yc": np.arange(n) + 0.5, "yg": np.arange(n) + 0.0})
    grid = xg.Grid(ds, coords={"X": {"center": "xc", "left": "xg"}, "Y": {"center": "yc", "left": "yg"}}, padding="periodic")
    ux = xr.DataArray(u, dims=("yc", "xg"), name="u")
    vx = xr.DataArray(v, dims=("yg", "xc"), name="v")
    out = grid.interp({"X": ux}, "X", other_component={"Y": vx})
    assert isinstance(out, FakeDataArray) and out.dims == ("yc", "xc")
    np.testing.assert_array_equal(out.data, 0.5 * (u + np.roll(u, -1, axis=1)))
    # extension methods accept xarray objects too
    div = grid.pair("diff", ux, "X", "diff", vx, "Y")
    assert isinstance(div, FakeDataArray) and div.dims == ("yc", "xc")
    np.testing.assert_array_equal(div.data, (np.roll(u, -1, axis=1) - u) + (np.roll(v, -1, axis=0) - v))
