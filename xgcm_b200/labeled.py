"""Minimal labelled arrays: the subset of xarray's DataArray / Dataset the Grid API needs.

The reference consumes and produces ``xarray`` objects.  xarray is not available
in the build image, so the Grid API here is written against this small duck-typed
layer (same attribute / method names and semantics as xarray for the subset
used: dims, coords, isel, transpose, rename, broadcasting arithmetic by
dimension name ...).  When xarray *is* importable, ``xgcm_b200.interop``
converts in both directions so ``Grid`` accepts and returns real xarray objects.

``DataArray.data`` is either a ``numpy.ndarray`` (host) or a CUDA
``torch.Tensor`` (device resident, stays resident through Grid operations).

Label algebra (``a * b`` with broadcasting by dim name) on device-resident data
runs in the library's ``xg_binary`` kernel; on host arrays it is plain numpy,
exactly what xarray does in the reference — it is not part of the xgcm hot path,
which never calls these operators (see ``grid.py``).
"""

from __future__ import annotations

from collections import OrderedDict
from typing import Mapping

import numpy as np

try:  # torch is only needed for device-resident data
    import torch
except Exception:  # pragma: no cover
    torch = None  # type: ignore


def is_device_array(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


def to_numpy(x) -> np.ndarray:
    if is_device_array(x):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def _as_data(x):
    if is_device_array(x):
        return x
    if isinstance(x, DataArray):
        return x.data
    return np.asarray(x)


class Coordinates(Mapping):
    def __init__(self, mapping: "OrderedDict[str, DataArray]"):
        self._m = mapping

    def __getitem__(self, k):
        return self._m[k]

    def __iter__(self):
        return iter(self._m)

    def __len__(self):
        return len(self._m)

    def __repr__(self):
        return "Coordinates(" + ", ".join(f"{k}{tuple(v.dims)}" for k, v in self._m.items()) + ")"


class DataArray:
    """N-d array with named dimensions and coordinates (xarray.DataArray look-alike)."""

    __array_priority__ = 60

    def __init__(self, data, coords=None, dims=None, name=None, attrs=None):
        data = _as_data(data)
        if dims is None:
            if data.ndim == 0:
                dims = ()
            elif coords is not None and not isinstance(coords, Mapping):
                dims = tuple(c[0] for c in coords)
            else:
                dims = tuple(f"dim_{i}" for i in range(data.ndim))
        if isinstance(dims, str):
            dims = (dims,)
        dims = tuple(dims)
        if len(dims) != data.ndim:
            raise ValueError(
                f"different number of dimensions on data and dims: {data.ndim} vs {len(dims)}"
            )
        if len(set(dims)) != len(dims):
            raise ValueError(f"duplicate dimension names: {dims}")
        self._data = data
        self._dims = dims
        self.name = name
        self.attrs = dict(attrs) if attrs else {}
        self._coords: "OrderedDict[str, DataArray]" = OrderedDict()
        if coords is not None:
            items = coords.items() if isinstance(coords, Mapping) else coords
            for cname, cval in items:
                self._set_coord(cname, cval)

    # -- construction helpers -------------------------------------------------
    def _set_coord(self, cname, cval):
        if isinstance(cval, DataArray):
            c = DataArray(cval.data, dims=cval.dims, name=cname, attrs=cval.attrs)
        elif isinstance(cval, tuple):
            cdims, cdata = cval[0], cval[1]
            cattrs = cval[2] if len(cval) > 2 else None
            c = DataArray(cdata, dims=cdims, name=cname, attrs=cattrs)
        else:
            cdata = np.asarray(cval)
            if cdata.ndim == 0:
                c = DataArray(cdata, dims=(), name=cname)
            elif cdata.ndim == 1:
                c = DataArray(cdata, dims=(cname,), name=cname)
            else:
                raise ValueError(f"coordinate {cname!r} needs explicit dims")
        sizes = self.sizes
        for d, s in zip(c.dims, c.shape):
            if d not in sizes:
                raise ValueError(f"coordinate {cname!r} has dim {d!r} not present on the array {self.dims}")
            if sizes[d] != s:
                raise ValueError(
                    f"conflicting sizes for dimension {d!r}: length {s} on coordinate {cname!r} "
                    f"and length {sizes[d]} on the data"
                )
        self._coords[cname] = c

    def _replace(self, data=None, dims=None, coords=None, name="__keep__", attrs=None):
        out = DataArray.__new__(DataArray)
        out._data = self._data if data is None else data
        out._dims = self._dims if dims is None else tuple(dims)
        out._coords = OrderedDict(self._coords if coords is None else coords)
        out.name = self.name if name == "__keep__" else name
        out.attrs = dict(self.attrs if attrs is None else attrs)
        return out

    # -- basic properties ------------------------------------------------------
    @property
    def data(self):
        return self._data

    @data.setter
    def data(self, value):
        value = _as_data(value)
        if tuple(value.shape) != tuple(self.shape):
            raise ValueError("replacement data must match the existing shape")
        self._data = value

    @property
    def values(self) -> np.ndarray:
        return to_numpy(self._data)

    @property
    def dims(self) -> Tuple[str, ...]:
        return self._dims

    @property
    def shape(self) -> Tuple[int, ...]:
        return tuple(int(s) for s in self._data.shape)

    @property
    def ndim(self) -> int:
        return len(self._dims)

    @property
    def size(self) -> int:
        return int(np.prod(self.shape, dtype=np.int64))

    @property
    def sizes(self) -> Dict[str, int]:
        return dict(zip(self._dims, self.shape))

    @property
    def dtype(self) -> np.dtype:
        if is_device_array(self._data):
            return np.dtype(str(self._data.dtype).replace("torch.", ""))
        return self._data.dtype

    @property
    def coords(self) -> Coordinates:
        return Coordinates(self._coords)

    @property
    def chunks(self):
        return None  # never dask-backed

    @property
    def variable(self):
        return self

    @property
    def is_device(self) -> bool:
        return is_device_array(self._data)

    @property
    def device(self):
        return self._data.device if self.is_device else "cpu"

    def __len__(self):
        return self.shape[0]

    def __array__(self, dtype=None, copy=None):
        v = self.values
        return v.astype(dtype) if dtype is not None else v

    def __repr__(self):
        where = f"cuda:{self._data.device.index}" if self.is_device else "host"
        head = f"<xgcm_b200.DataArray {self.name or ''} ({', '.join(f'{d}: {s}' for d, s in self.sizes.items())}) {self.dtype} [{where}]>"
        if self.size <= 20:
            head += "\n" + repr(self.values)
        if self._coords:
            head += "\nCoordinates: " + ", ".join(f"{k}{tuple(v.dims)}" for k, v in self._coords.items())
        return head

    # -- residency ----------------------------------------------------------------
    def to_device(self, device="cuda"):
        """Return a copy of this array resident on a CUDA device (torch tensor)."""
        if torch is None:
            raise RuntimeError("torch is required for device-resident arrays")
        if self.is_device:
            return self._replace(data=self._data.to(device))
        return self._replace(data=torch.from_numpy(np.ascontiguousarray(self._data)).to(device))

    def to_host(self):
        return self._replace(data=self.values) if self.is_device else self

    # -- xarray-like methods ----------------------------------------------------------
    def get_axis_num(self, dim):
        if isinstance(dim, (list, tuple)):
            return tuple(self.get_axis_num(d) for d in dim)
        try:
            return self._dims.index(dim)
        except ValueError:
            raise ValueError(f"{dim!r} not found in array dimensions {self._dims!r}")

    def copy(self, deep=True, data=None):
        if data is not None:
            return self._replace(data=_as_data(data))
        if deep:
            d = self._data.clone() if self.is_device else self._data.copy()
            return self._replace(data=d)
        return self._replace()

    def astype(self, dtype):
        if self.is_device:
            tdt = getattr(torch, np.dtype(dtype).name)
            return self._replace(data=self._data.to(tdt))
        return self._replace(data=self._data.astype(dtype))

    def rename(self, new_name_or_name_dict=None, **names):
        if new_name_or_name_dict is None or isinstance(new_name_or_name_dict, Mapping):
            mapping = dict(new_name_or_name_dict or {})
            mapping.update(names)
            dims = tuple(mapping.get(d, d) for d in self._dims)
            coords = OrderedDict()
            for k, c in self._coords.items():
                nk = mapping.get(k, k)
                coords[nk] = c._replace(dims=tuple(mapping.get(d, d) for d in c.dims), name=nk)
            return self._replace(dims=dims, coords=coords)
        return self._replace(name=new_name_or_name_dict)

    def transpose(self, *dims):
        if not dims:
            dims = self._dims[::-1]
        if Ellipsis in dims:
            rest = [d for d in self._dims if d not in dims]
            i = dims.index(Ellipsis)
            dims = tuple(dims[:i]) + tuple(rest) + tuple(dims[i + 1:])
        if set(dims) != set(self._dims) or len(dims) != len(self._dims):
            raise ValueError(f"{dims} must be a permutation of {self._dims}")
        perm = [self._dims.index(d) for d in dims]
        if perm == list(range(self.ndim)):
            return self._replace()
        data = self._data.permute(*perm) if self.is_device else np.transpose(self._data, perm)
        return self._replace(data=data, dims=dims)

    def isel(self, indexers=None, **kw):
        indexers = dict(indexers or {})
        indexers.update(kw)
        for d in indexers:
            if d not in self._dims:
                raise ValueError(f"Dimensions {set(indexers) - set(self._dims)} do not exist.")
        key = []
        new_dims = []
        for d in self._dims:
            idx = indexers.get(d, slice(None))
            if is_device_array(self._data) and isinstance(idx, slice) and idx.step is not None and idx.step < 0:
                raise NotImplementedError("negative-step slices of device arrays are not supported")
            if isinstance(idx, (list, np.ndarray)):
                idx = np.asarray(idx)
            key.append(idx)
            if not isinstance(idx, (int, np.integer)):
                new_dims.append(d)
        data = self._data[tuple(key)]
        coords = OrderedDict()
        for k, c in self._coords.items():
            sub = {d: indexers[d] for d in c.dims if d in indexers}
            coords[k] = c.isel(sub) if sub else c
        return self._replace(data=data, dims=tuple(new_dims), coords=coords)

    def __getitem__(self, key):
        if isinstance(key, Mapping):
            return self.isel(key)
        if not isinstance(key, tuple):
            key = (key,)
        key = key + (slice(None),) * (self.ndim - len(key))
        return self.isel(dict(zip(self._dims, key)))

    def assign_coords(self, coords=None, **kw):
        coords = dict(coords or {})
        coords.update(kw)
        out = self._replace()
        for k, v in coords.items():
            out._set_coord(k, v)
        return out

    def drop_vars(self, names, errors="raise"):
        if isinstance(names, str):
            names = [names]
        names = list(names)
        coords = OrderedDict((k, v) for k, v in self._coords.items() if k not in names)
        return self._replace(coords=coords)

    def reset_coords(self, names=None, drop=False):
        if not drop:
            raise NotImplementedError("reset_coords(drop=False)")
        keep = OrderedDict((k, v) for k, v in self._coords.items() if k in self._dims)
        return self._replace(coords=keep)

    def expand_dims(self, dim, axis=0):
        data = self._data.unsqueeze(axis) if self.is_device else np.expand_dims(self._data, axis)
        dims = list(self._dims)
        dims.insert(axis, dim)
        return self._replace(data=data, dims=dims)

    def squeeze(self, dim=None):
        dims = [d for d, s in self.sizes.items() if s == 1] if dim is None else ([dim] if isinstance(dim, str) else list(dim))
        out = self
        for d in dims:
            out = out.isel({d: 0})
        return out

    def broadcast_like(self, other: "DataArray"):
        (a, _), dims = _broadcast_pair(self, other)
        shape = tuple({**other.sizes, **self.sizes}[d] for d in dims)
        data = a.expand(shape).contiguous() if is_device_array(a) else np.broadcast_to(a, shape).copy()
        return DataArray(data, dims=dims, name=self.name)

    def equals(self, other) -> bool:
        if not isinstance(other, DataArray):
            return False
        if self.dims != other.dims or self.shape != other.shape:
            return False
        a, b = self.values, other.values
        return bool(np.array_equal(a, b, equal_nan=a.dtype.kind == "f"))

    def identical(self, other) -> bool:
        return self.equals(other) and self.name == other.name

    def isnull(self):
        return self._replace(data=np.isnan(self.values))

    def notnull(self):
        return self._replace(data=~np.isnan(self.values))

    # reductions -- label conveniences only; the Grid hot path uses the fused kernels
    def _reduce(self, fn, dim, **kw):
        if dim is None:
            dims = list(self._dims)
        elif isinstance(dim, str):
            dims = [dim]
        else:
            dims = list(dim)
        axes = tuple(self.get_axis_num(d) for d in dims)
        data = fn(self.values, axis=axes, **kw)
        keep = tuple(d for d in self._dims if d not in dims)
        coords = OrderedDict((k, c) for k, c in self._coords.items() if all(d in keep for d in c.dims))
        return DataArray(data, dims=keep, coords=coords, name=self.name)

    def sum(self, dim=None, skipna=None):
        fn = np.nansum if (skipna or (skipna is None and self.dtype.kind == "f")) else np.sum
        return self._reduce(fn, dim)

    def mean(self, dim=None, skipna=None):
        fn = np.nanmean if (skipna or (skipna is None and self.dtype.kind == "f")) else np.mean
        return self._reduce(fn, dim)

    def max(self, dim=None):
        return self._reduce(np.nanmax, dim)

    def min(self, dim=None):
        return self._reduce(np.nanmin, dim)

    # -- arithmetic (broadcast by dim name) ---------------------------------------------
    def _binary(self, other, opname, reflexive=False):
        if isinstance(other, DataArray):
            a, b = (other, self) if reflexive else (self, other)
            (da, db), dims = _broadcast_pair(a, b)
            sizes = {**b.sizes, **a.sizes}
            data = _apply_binary(opname, da, db, tuple(sizes[d] for d in dims))
            coords = OrderedDict()
            for src in (a, b):
                for k, c in src._coords.items():
                    if k not in coords and all(d in dims for d in c.dims):
                        coords[k] = c
            name = self.name if self.name == other.name else None
            return DataArray(data, dims=dims, coords=coords, name=name)
        # scalar / raw array
        if is_device_array(self._data):
            o = other
            if not is_device_array(o):
                o = torch.as_tensor(np.asarray(other), device=self._data.device).to(self._data.dtype) if np.ndim(other) else other
            a, b = (o, self._data) if reflexive else (self._data, o)
            data = _torch_scalar_op(opname, a, b)
        else:
            a, b = (other, self._data) if reflexive else (self._data, other)
            data = _NP_OPS[opname](a, b)
        return self._replace(data=data)

    def __mul__(self, o):
        return self._binary(o, "mul")

    def __rmul__(self, o):
        return self._binary(o, "mul", True)

    def __truediv__(self, o):
        return self._binary(o, "div")

    def __rtruediv__(self, o):
        return self._binary(o, "div", True)

    def __add__(self, o):
        return self._binary(o, "add")

    def __radd__(self, o):
        return self._binary(o, "add", True)

    def __sub__(self, o):
        return self._binary(o, "sub")

    def __rsub__(self, o):
        return self._binary(o, "sub", True)

    def __neg__(self):
        return self._replace(data=-self._data)

    def __pow__(self, p):
        return self._replace(data=self._data ** p)


_NP_OPS = {"mul": np.multiply, "div": np.true_divide, "add": np.add, "sub": np.subtract}


def _torch_scalar_op(opname, a, b):
    return {"mul": lambda: a * b, "div": lambda: a / b, "add": lambda: a + b, "sub": lambda: a - b}[opname]()


def _broadcast_pair(a: DataArray, b: DataArray):
    """Align two arrays by dim name; returns views with a common dim order (size-1 inserted)."""
    dims = list(a.dims) + [d for d in b.dims if d not in a.dims]
    for d in set(a.dims) & set(b.dims):
        if a.sizes[d] != b.sizes[d]:
            raise ValueError(
                f"cannot broadcast: conflicting sizes for dimension {d!r}: {a.sizes[d]} vs {b.sizes[d]}"
            )
    out = []
    for x in (a, b):
        data = x.data
        present = [d for d in dims if d in x.dims]
        perm = [x.dims.index(d) for d in present]
        if perm != list(range(len(perm))):
            data = data.permute(*perm) if is_device_array(data) else np.transpose(data, perm)
        shape = [x.sizes[d] if d in x.dims else 1 for d in dims]
        data = data.reshape(shape)
        out.append(data)
    return tuple(out), tuple(dims)


def _apply_binary(opname, a, b, shape):
    dev_a, dev_b = is_device_array(a), is_device_array(b)
    if dev_a or dev_b:
        from . import ops  # device path: the library's xg_binary kernel

        if not dev_a:
            a = torch.from_numpy(np.ascontiguousarray(a)).to(b.device)
        if not dev_b:
            b = torch.from_numpy(np.ascontiguousarray(b)).to(a.device)
        return ops.binary(opname, a, b, shape)
    return _NP_OPS[opname](a, b)


class Dataset:
    """Dict of variables sharing dimensions (xarray.Dataset look-alike)."""

    def __init__(self, data_vars=None, coords=None, attrs=None):
        self._vars: "OrderedDict[str, DataArray]" = OrderedDict()
        self._coord_names = []
        self.attrs = dict(attrs) if attrs else {}
        self._sizes: Dict[str, int] = {}
        for cname, cval in (coords or {}).items():
            self._add(cname, cval, is_coord=True)
        for vname, vval in (data_vars or {}).items():
            self._add(vname, vval, is_coord=False)

    def _add(self, name, val, is_coord):
        if isinstance(val, DataArray):
            da = DataArray(val.data, dims=val.dims, name=name, attrs=val.attrs)
            extra = val._coords
        elif isinstance(val, tuple):
            da = DataArray(val[1], dims=val[0], name=name, attrs=val[2] if len(val) > 2 else None)
            extra = {}
        else:
            arr = np.asarray(val)
            if arr.ndim == 1:
                da = DataArray(arr, dims=(name,), name=name)
            elif arr.ndim == 0:
                da = DataArray(arr, dims=(), name=name)
            else:
                raise ValueError(f"variable {name!r} needs explicit dims")
            extra = {}
        for d, s in da.sizes.items():
            if d in self._sizes and self._sizes[d] != s:
                raise ValueError(f"conflicting sizes for dimension {d!r}: {self._sizes[d]} vs {s} ({name})")
            self._sizes[d] = s
        self._vars[name] = da
        if is_coord and name not in self._coord_names:
            self._coord_names.append(name)
        for k, c in extra.items():
            if k not in self._vars:
                self._add(k, c, is_coord=True)

    @property
    def dims(self):
        return dict(self._sizes)

    @property
    def sizes(self):
        return dict(self._sizes)

    @property
    def variables(self):
        return self._vars

    @property
    def coords(self) -> Coordinates:
        return Coordinates(OrderedDict((k, self[k]) for k in self._coord_names))

    @property
    def data_vars(self):
        return OrderedDict((k, self[k]) for k in self._vars if k not in self._coord_names)

    def __contains__(self, k):
        return k in self._vars or k in self._sizes

    def __iter__(self):
        return iter(self.data_vars)

    def keys(self):
        return self.data_vars.keys()

    def __getitem__(self, name) -> DataArray:
        if name not in self._vars:
            if name in self._sizes:  # dimension without coordinate: index values
                return DataArray(np.arange(self._sizes[name]), dims=(name,), name=name)
            raise KeyError(name)
        v = self._vars[name]
        coords = OrderedDict()
        for cn in self._coord_names:
            c = self._vars[cn]
            if cn != name and all(d in v.dims for d in c.dims):
                coords[cn] = c
        if name in self._coord_names and v.dims == (name,):
            coords[name] = v
        return v._replace(coords=coords, name=name)

    def __setitem__(self, name, val):
        self._add(name, val, is_coord=False)

    def assign_coords(self, coords=None, **kw):
        coords = dict(coords or {})
        coords.update(kw)
        out = self.copy()
        for k, v in coords.items():
            out._add(k, v, is_coord=True)
        return out

    def copy(self, deep=False):
        out = Dataset()
        out._vars = OrderedDict((k, v.copy(deep=deep)) for k, v in self._vars.items())
        out._coord_names = list(self._coord_names)
        out._sizes = dict(self._sizes)
        out.attrs = dict(self.attrs)
        return out

    def __repr__(self):
        return (
            "<xgcm_b200.Dataset dims=" + str(self._sizes) + " coords=" + str(self._coord_names)
            + " data_vars=" + str(list(self.data_vars)) + ">"
        )
