"""Array-level operators: torch CUDA tensors in/out, numerics in libxgcm_b200.so.

torch is plumbing only (device allocation, streams); every value is produced by
a hand-written sm_100a kernel reached through the C-ABI (``include/xgcm_b200.h``).
There is no CPU path: a non-CUDA tensor raises.
"""

from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _capi

_TORCH_DTYPE_CODE = {torch.float32: _capi.XG_F32, torch.float64: _capi.XG_F64}


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{what} must be a torch.Tensor, got {type(t)}")
    if not t.is_cuda:
        raise RuntimeError(
            f"{what} lives on {t.device}; xgcm_b200 kernels run on CUDA devices only "
            "(there is no CPU fallback)"
        )


def _dtype_code(t: torch.Tensor) -> int:
    try:
        return _TORCH_DTYPE_CODE[t.dtype]
    except KeyError:
        raise TypeError(f"xgcm_b200 kernels support float32/float64 fields, got {t.dtype}")


def _stream_ptr(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _norm_axis(axis: int, ndim: int) -> int:
    if not -ndim <= axis < ndim:
        raise ValueError(f"axis {axis} out of range for {ndim}-d field")
    return axis % ndim


def _operand(m: Optional[torch.Tensor], shape: Sequence[int], like: torch.Tensor, what: str):
    """Return (keepalive tensor, data_ptr, int64[ndim] strides) for a broadcast operand."""
    if m is None:
        return None, None, None
    _require_cuda(m, what)
    if m.dtype != like.dtype:
        m = m.to(like.dtype)
    if m.device != like.device:
        raise RuntimeError(f"{what} is on {m.device}, field on {like.device}")
    try:
        mb = m.expand(tuple(shape))
    except RuntimeError as err:
        raise ValueError(f"{what} of shape {tuple(m.shape)} does not broadcast to {tuple(shape)}") from err
    strides = [0 if s == 1 else st for s, st in zip(shape, mb.stride())]
    if any(st < 0 for st in strides):
        mb = m.contiguous().expand(tuple(shape))
        strides = [0 if s == 1 else st for s, st in zip(shape, mb.stride())]
    return mb, mb.data_ptr(), _capi.i64_array(strides)


def stencil2(
    x: torch.Tensor,
    axis: int,
    op: str,
    lo: int,
    hi: int,
    padding: Optional[str],
    fill_value: float = 0.0,
    pre: Optional[torch.Tensor] = None,
    post: Optional[torch.Tensor] = None,
    halo_lo: Optional[torch.Tensor] = None,
    halo_hi: Optional[torch.Tensor] = None,
    out: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """Fused pad + {diff, interp, min, max} (+ metric) along ``axis``.

    Mirrors padding.py:575-616 + gridops.py:23-24,76-77,123-126,172-175 +
    grid.py:806-808,830-832 of the reference in one HBM pass.
    """
    lib = _capi.load()
    _require_cuda(x, "field")
    if op not in _capi.OPS:
        raise ValueError(f"unknown op {op!r}")
    if padding not in _capi.BCS:
        raise ValueError(
            f"padding must be one of ['periodic', 'fill', 'extend'] or None, but got {padding}"
        )
    x = x.contiguous()
    axis = _norm_axis(axis, x.dim())
    shape = list(x.shape)
    out_shape = list(shape)
    out_shape[axis] = shape[axis] + lo + hi - 1
    if out_shape[axis] < 0:
        raise ValueError("operated axis too short")
    if out is None:
        out = torch.empty(out_shape, dtype=x.dtype, device=x.device)
    else:
        if list(out.shape) != out_shape or out.dtype != x.dtype or not out.is_contiguous():
            raise ValueError("out has wrong shape/dtype/layout")
    keep_pre, pre_ptr, pre_st = _operand(pre, shape, x, "pre metric")
    keep_post, post_ptr, post_st = _operand(post, out_shape, x, "post metric")
    plane = [s for d, s in enumerate(shape) if d != axis]
    hl = hh = None
    if halo_lo is not None:
        _require_cuda(halo_lo, "halo_lo")
        hl = halo_lo.to(x.dtype).contiguous()
        if hl.numel() != int(np.prod(plane, dtype=np.int64)):
            raise ValueError("halo_lo has wrong size")
    if halo_hi is not None:
        _require_cuda(halo_hi, "halo_hi")
        hh = halo_hi.to(x.dtype).contiguous()
        if hh.numel() != int(np.prod(plane, dtype=np.int64)):
            raise ValueError("halo_hi has wrong size")
    with torch.cuda.device(x.device):
        rc = lib.xg_stencil2(
            _capi.OPS[op], _dtype_code(x), x.data_ptr(), out.data_ptr(), x.dim(),
            _capi.i64_array(shape), axis, lo, hi, _capi.BCS[padding], float(fill_value),
            pre_ptr, pre_st, post_ptr, post_st,
            hl.data_ptr() if hl is not None else None,
            hh.data_ptr() if hh is not None else None,
            _stream_ptr(x),
        )
    _capi.check(rc)
    return out
