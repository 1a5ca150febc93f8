"""Conversion to / from real ``xarray`` objects (used only when xarray is importable).

xarray is absent from the build image, so this path could not be exercised
there; it is a thin, lossless mapping (data, dims, coords, name, attrs).
"""

from __future__ import annotations

from .labeled import DataArray, Dataset


def _require_xarray():
    try:
        import xarray as xr
    except ImportError as err:  # pragma: no cover
        raise ImportError("xarray is not installed; use xgcm_b200.DataArray / Dataset") from err
    return xr


def dataarray_from_xarray(obj) -> DataArray:
    coords = {}
    for name, c in obj.coords.items():
        coords[name] = (tuple(c.dims), c.values, dict(c.attrs))
    return DataArray(obj.data, dims=tuple(obj.dims), coords=coords, name=obj.name, attrs=dict(obj.attrs))


def dataset_from_xarray(obj) -> Dataset:
    coords = {n: (tuple(c.dims), c.values, dict(c.attrs)) for n, c in obj.coords.items()}
    data_vars = {n: (tuple(v.dims), v.data, dict(v.attrs)) for n, v in obj.data_vars.items()}
    ds = Dataset(data_vars=data_vars, coords=coords, attrs=dict(obj.attrs))
    for d, s in obj.sizes.items():  # dims without coordinate variables
        ds._sizes.setdefault(d, int(s))
    return ds


def dataarray_to_xarray(da: DataArray):
    xr = _require_xarray()
    coords = {n: (tuple(c.dims), c.values, dict(c.attrs)) for n, c in da.coords.items()}
    return xr.DataArray(da.values if da.is_device else da.data, dims=da.dims, coords=coords,
                        name=da.name, attrs=dict(da.attrs))
