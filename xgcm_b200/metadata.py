"""Minimal COMODO attribute parsing (reference ``xgcm/comodo.py:16-100``).

Only what ``Grid(ds)`` without ``coords=`` needs: dims whose coordinate carries
``attrs['axis']`` form an axis, ``attrs['c_grid_axis_shift']`` (+-0.5) puts the
dim on the left / right position — or inner / outer when the shifted dim is one
shorter / longer than the centre dim.  SGRID parsing and the rest of the
metadata machinery (comodo.py / sgrid.py / metadata_parsers.py) are out of scope.
"""

from __future__ import annotations

from collections import OrderedDict


def parse_comodo(ds):
    axes = OrderedDict()
    for name in list(ds.dims):
        if name not in ds.variables:
            continue
        var = ds.variables[name]
        attrs = getattr(var, "attrs", {})
        if "axis" not in attrs or var.dims != (name,):
            continue
        axes.setdefault(attrs["axis"], []).append((name, attrs.get("c_grid_axis_shift")))
    coords = OrderedDict()
    for ax, members in axes.items():
        centers = [n for n, shift in members if shift is None]
        if len(centers) != 1:
            continue
        center = centers[0]
        n_center = ds.sizes[center]
        positions = {"center": center}
        for n, shift in members:
            if shift is None:
                continue
            size = ds.sizes[n]
            if size == n_center:
                positions["left" if float(shift) < 0 else "right"] = n
            elif size == n_center + 1:
                positions["outer"] = n
            elif size == n_center - 1:
                positions["inner"] = n
        coords[ax] = positions
    return coords or None
