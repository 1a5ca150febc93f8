"""CPU oracle for face-connection padding (cubed sphere / LLC tiles).

TEST INFRASTRUCTURE ONLY.  Nothing under ``xgcm_b200/`` may import this.

A numpy restatement of ``xgcm/padding.py:260-572`` (``_pad_face_connections``) that follows the
reference step by step — pre-pad every face, then for every face / axis / side with a connection
slice the neighbour, swap the dimension names, flip, change sign, concatenate, finally trim — on
"named arrays" (an ndarray plus a tuple of dim names), i.e. with the operations xarray performs
(``isel``, ``rename``, ``concat``, ``transpose``).  The product instead derives one signed-stride
index map per edge (``xgcm_b200/padding.py``); the two share no code, which is the point.

Pinning: the reference cannot run here (its padding module needs xarray), so this file is pinned
by ``tests/test_faces.py`` against the expectations of the reference's own tests
(``xgcm/test/test_padding.py:172-1205`` construct the expected arrays by hand with ``pad`` /
``isel`` / ``concat``; ``xgcm/test/test_faceconnections.py:164-230,410-478`` assert individual
seams) re-expressed in numpy.

The reference iterates the axes in ``set`` order (hash-seed dependent); halo corners depend on
it.  Here, like in the product, the order is that of ``grid_axes``.
"""

from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np

_PAD_MODE = {"periodic": "wrap", "fill": "constant", "extend": "edge"}


class Named:
    """ndarray + dim names: the little of xarray the reference's algorithm uses."""

    def __init__(self, data, dims):
        self.data = np.asarray(data)
        self.dims = tuple(dims)
        assert self.data.ndim == len(self.dims)

    def isel(self, **idx):
        key = tuple(idx.get(d, slice(None)) for d in self.dims)
        dims = tuple(d for d in self.dims if not isinstance(idx.get(d, slice(None)), (int, np.integer)))
        return Named(self.data[key], dims)

    def rename(self, mapping):
        return Named(self.data, tuple(mapping.get(d, d) for d in self.dims))

    def transpose(self, *dims):
        return Named(np.transpose(self.data, [self.dims.index(d) for d in dims]), dims)

    def flip(self, dim):
        return Named(np.flip(self.data, self.dims.index(dim)), self.dims)

    def neg(self):
        return Named(-self.data, self.dims)

    def pad(self, dim, widths, mode, **kw):
        pw = [(0, 0)] * self.data.ndim
        pw[self.dims.index(dim)] = tuple(widths)
        return Named(np.pad(self.data, pw, mode, **kw), self.dims)


def concat(arrs: Sequence[Named], dim: str) -> Named:
    first = arrs[0]
    if dim in first.dims:
        aligned = [a.transpose(*first.dims) for a in arrs]
        return Named(np.concatenate([a.data for a in aligned], axis=first.dims.index(dim)), first.dims)
    return Named(np.stack([a.transpose(*first.dims).data for a in arrs], axis=0), (dim,) + first.dims)


def _axis_dim(axes_coords: Dict[str, Sequence[str]], axname: str, dims) -> str:
    for d in axes_coords[axname]:
        if d in dims:
            return d
    raise KeyError(axname)


def pad_basic(da: Named, axes_coords, padding_width, padding, fill_value) -> Named:
    """padding.py:575-616"""
    out = da
    for ax, widths in padding_width.items():
        if all(w == 0 for w in widths):
            continue
        dim = _axis_dim(axes_coords, ax, out.dims)
        mode = _PAD_MODE[padding[ax]]
        kw = dict(constant_values=fill_value[ax]) if mode == "constant" else {}
        out = out.pad(dim, widths, mode, **kw)
    return out


def _rename_grid_positions(axes_coords, source: Named, target: Named) -> Named:
    """padding.py:183-198: name the partner component's dims like the target's."""
    mapping = {}
    for di in target.dims:
        if di not in source.dims:
            for positions in axes_coords.values():
                if di in positions:
                    mapping[[p for p in positions if p in source.dims][0]] = di
    return source.rename(mapping)


def _swap_dimension_names(da: Named, from_name, to_name) -> Named:
    """padding.py:201-210"""
    if to_name in da.dims:
        da = da.rename({to_name: to_name + "dummy"})
        if from_name in da.dims:
            da = da.rename({from_name: to_name})
        return da.rename({to_name + "dummy": from_name})
    return da.rename({from_name: to_name})


def pad_face_connections(
    data: np.ndarray,
    dims: Sequence[str],
    axes_coords: Dict[str, Sequence[str]],
    facedim: str,
    face_links: Dict[int, Dict[str, Tuple]],
    padding_width: Dict[str, Tuple[int, int]],
    padding: Dict[str, Optional[str]],
    fill_value: Dict[str, float],
    vector_axis: Optional[str] = None,
    partner: Optional[np.ndarray] = None,
    partner_dims: Optional[Sequence[str]] = None,
) -> np.ndarray:
    """padding.py:260-572.  ``axes_coords``: axis name -> dim names of its positions (in the
    order of ``grid.axes``); ``face_links``: ``connections[facedim]``."""
    da = Named(data, dims)
    isvector = vector_axis is not None
    connection_axes = []
    for c in face_links.values():
        connection_axes.extend(c.keys())
    wanted = set(connection_axes) | set(padding_width)
    pad_axes = [ax for ax in axes_coords if ax in wanted]
    padding_width = {ax: tuple(padding_width.get(ax, (0, 0))) for ax in pad_axes}
    width = max(max(w) for w in padding_width.values())  # padding.py:320-326
    maxw = {ax: (width, width) for ax in pad_axes}

    prepad_padding = dict(padding)
    for ax in pad_axes:  # padding.py:349-380
        if prepad_padding.get(ax) is None:
            prepad_padding[ax] = "fill"
    pre = pad_basic(da, axes_coords, maxw, prepad_padding, fill_value)
    pre_partner = None
    if isvector:
        pre_partner = pad_basic(Named(partner, partner_dims), axes_coords, maxw, prepad_padding, fill_value)

    n_face = da.data.shape[da.dims.index(facedim)]
    faces = []
    for i in range(n_face):
        target = pre.isel(**{facedim: i})
        single = face_links.get(i, {})
        for axname in pad_axes:
            left, right = single.get(axname, (None, None))
            target_dim = _axis_dim(axes_coords, axname, target.dims)
            for connection, is_right in [(left, False), (right, True)]:
                if width == 0 or not connection:
                    continue
                source_face, source_axis, reverse = connection
                swap_axis = axname != source_axis
                source = pre.isel(**{facedim: source_face})
                if isvector and swap_axis:
                    source = _rename_grid_positions(
                        axes_coords, pre_partner.isel(**{facedim: source_face}), target
                    )
                source_dim = _axis_dim(axes_coords, source_axis, source.dims)
                if is_right:  # padding.py:443-459
                    s_idx = slice(-2 * width, -width) if reverse else slice(width, 2 * width)
                    t_idx = slice(0, -width)
                else:
                    s_idx = slice(width, 2 * width) if reverse else slice(-2 * width, -width)
                    t_idx = slice(width, None)
                s_slice = source.isel(**{source_dim: s_idx})
                t_slice = target.isel(**{target_dim: t_idx})
                if swap_axis:
                    s_slice = _swap_dimension_names(s_slice, source_dim, target_dim)
                ortho_dim, tangential_dim = target_dim, source_dim
                if reverse:  # padding.py:478-487
                    s_slice = s_slice.flip(ortho_dim)
                    if isvector and vector_axis == axname:
                        s_slice = s_slice.neg()
                if swap_axis and not reverse:  # padding.py:489-498
                    s_slice = s_slice.flip(tangential_dim)
                    if isvector and vector_axis != axname:
                        s_slice = s_slice.neg()
                s_slice = s_slice.transpose(*t_slice.dims)
                parts = [t_slice, s_slice] if is_right else [s_slice, t_slice]
                target = concat(parts, target_dim)
        faces.append(target)
    padded = concat(faces, facedim).transpose(*pre.dims)

    out = padded  # padding.py:557-572
    for axname in padding_width:
        dim = _axis_dim(axes_coords, axname, out.dims)
        start = width - padding_width[axname][0]
        stop = width - padding_width[axname][1]
        out = out.isel(**{dim: slice(start, -stop if stop else None)})
    return out.data
