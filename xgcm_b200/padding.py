"""``pad``: halo padding of labelled arrays (reference ``xgcm/padding.py:765-871``).

The built-in operators never call this (their halo is fused into the stencil
kernel); it exists for user-defined grid ufuncs and for API parity.  The copy
runs on the device in ``xg_pad``.  Face-connection and north-fold padding
(padding.py:230-572, :21-181) are out of scope.
"""

from __future__ import annotations

from typing import Dict, Mapping, Optional, Tuple, Union

from .labeled import DataArray

# reference padding.py:15-19
_XGCM_BOUNDARY_KWARG_TO_XARRAY_PAD_KWARG = {
    "periodic": "wrap",
    "fill": "constant",
    "extend": "edge",
}


def _strip_all_coords(data):
    """Padding cannot invent coordinate values: drop them all (padding.py:220-227)."""
    if isinstance(data, dict):
        return {k: v.drop_vars(list(v.coords)) for k, v in data.items()}
    return data.drop_vars(list(data.coords))


def _pad_basic(da: DataArray, grid, padding_width, padding, fill_value):
    from . import ops
    from .device import as_device_tensor, result_like

    out = da.copy(deep=False)
    for ax, widths in padding_width.items():
        if all(w == 0 for w in widths):
            continue  # padding.py:592-593
        axis = grid.axes[ax]
        _, dim = axis._get_position_name(out)
        ax_padding = padding[ax]
        if ax_padding is None:
            raise ValueError(
                f"No boundary condition was specified for axis {ax!r}, but the "
                f"requested operation needs to pad it. Set a boundary condition, "
                f"e.g. ``padding='fill'`` (or 'extend'/'periodic'), on the Grid "
                f"(``Grid(..., padding=...)``) or pass ``padding=`` to the "
                f"grid method."
            )
        if isinstance(ax_padding, Mapping):
            raise NotImplementedError("fold padding is outside the scope of xgcm_b200")
        x, was_host = as_device_tensor(out.data, grid._device_for(out))
        fv = fill_value[ax] if fill_value[ax] is not None else 0.0
        y = ops.pad(x, out.get_axis_num(dim), int(widths[0]), int(widths[1]), ax_padding, fv)
        out = DataArray(result_like(y, was_host), dims=out.dims, name=out.name, attrs=out.attrs)
    return out


def pad(
    data: Union[DataArray, Dict[str, DataArray]],
    grid,
    padding_width: Optional[Dict[str, Tuple[int, int]]],
    padding: Optional[Union[str, Mapping[str, str]]] = None,
    fill_value: Optional[Union[float, Mapping[str, float]]] = None,
    other_component: Optional[Dict[str, DataArray]] = None,
    **kwargs,
):
    """Pad ``data`` along the given grid axes according to the boundary conditions."""
    if "boundary" in kwargs:
        raise ValueError(
            "Argument 'boundary' has been renamed to 'padding'. Please use 'padding' instead."
        )
    if "boundary_width" in kwargs:
        raise ValueError(
            "Argument 'boundary_width' has been renamed to 'padding_width'. "
            "Please use 'padding_width' instead."
        )
    padding = grid._complete_user_kwargs_using_axis_defaults(padding, "padding")
    fill_value = grid._complete_user_kwargs_using_axis_defaults(fill_value, "fill_value")

    if padding_width is None or all(tuple(w) == (0, 0) for w in padding_width.values()):
        return data  # padding.py:831-836

    data = _strip_all_coords(data)
    if grid._face_connections is not None:
        raise NotImplementedError("face-connection padding is outside the scope of xgcm_b200")
    if isinstance(data, dict):
        [data] = list(data.values())
    return _pad_basic(data, grid, padding_width, padding, fill_value)
