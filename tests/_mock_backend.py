"""CPU stand-in for the CUDA kernels, for HOST-LOGIC tests only.

On a machine without a GPU the Grid façade cannot compute anything (by design there is no CPU
path in the product).  To still exercise the Python host logic here — kwarg precedence, metric
selection, dims / coords bookkeeping, error behaviour — this fixture swaps ``xgcm_b200.ops`` for
oracle-backed functions operating on CPU torch tensors.  It lives under tests/, is never
imported by the package, and is not used on the GPU box, where the same test bodies run against
the real kernels (tests/test_grid_gpu.py, tests/test_transform_gpu.py).
"""

import numpy as np
import torch

from oracle import stencil as oracle


def _np(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy()
    return np.asarray(t)


def _t(a):
    return torch.from_numpy(np.array(a, copy=True, order="C"))


def install(monkeypatch):
    from xgcm_b200 import device, ops

    cpu = torch.device("cpu")
    monkeypatch.setattr(device, "default_device", lambda: cpu)

    def as_device_tensor(data, dev=None):
        if isinstance(data, torch.Tensor):
            t = data if data.dtype in (torch.float32, torch.float64) else data.to(torch.float64)
            return t.contiguous(), False
        arr = np.asarray(data)
        if arr.dtype not in (np.float32, np.float64):
            arr = arr.astype(np.float64)
        return _t(arr), True

    monkeypatch.setattr(device, "as_device_tensor", as_device_tensor)
    monkeypatch.setattr(device, "as_device_constant", lambda data, dev=None: as_device_tensor(data, dev)[0])
    monkeypatch.setattr(device, "result_like", lambda t, was_host: t.numpy() if was_host else t)

    def stencil2(x, axis, op, lo, hi, padding, fill_value=0.0, pre=None, post=None, halo_lo=None,
                 halo_hi=None, out=None):
        if (lo or hi) and padding is None:
            raise ValueError("no boundary condition was specified")
        if halo_lo is not None or halo_hi is not None:
            # explicit halo planes replace the boundary condition on their side (xgcm_b200.h:109-112)
            a = _np(x) if pre is None else _np(x) * _np(pre)
            padded = oracle.pad_axis(a, axis, lo if halo_lo is None else 0, hi if halo_hi is None else 0,
                                     padding, fill_value)
            parts = ([_np(halo_lo).reshape([1 if d == axis % a.ndim else n for d, n in enumerate(a.shape)])]
                     if halo_lo is not None and lo else []) + [padded] + (
                [_np(halo_hi).reshape([1 if d == axis % a.ndim else n for d, n in enumerate(a.shape)])]
                if halo_hi is not None and hi else [])
            padded = np.concatenate(parts, axis=axis)
            r = np.moveaxis(oracle.KERNELS[op](np.moveaxis(padded, axis, -1)), -1, axis)
            if post is not None:
                r = r / _np(post)
            return _t(r.astype(_np(x).dtype))
        r = oracle.stencil2(op, _np(x), axis, lo, hi, padding if (lo or hi) else None, fill_value,
                            _np(pre), _np(post))
        return _t(r.astype(_np(x).dtype))

    def stencil2_host(x, axis, op, lo, hi, padding, fill_value=0.0, pre=None, post=None, out=None,
                      device=None):
        return stencil2(x, axis, op, lo, hi, padding, fill_value, pre, post).numpy()

    def pad(x, axis, lo, hi, padding, fill_value=0.0):
        return _t(oracle.pad_axis(_np(x), axis, lo, hi, padding, fill_value))

    def binary(opname, a, b, shape=None):
        def divnz(x, y):
            with np.errstate(invalid="ignore", divide="ignore"):
                return np.where(y != 0, x / y, np.nan).astype(np.result_type(x, y))

        fn = {"mul": np.multiply, "div": np.true_divide, "add": np.add, "sub": np.subtract, "divnz": divnz}[opname]
        with np.errstate(invalid="ignore", divide="ignore"):
            return _t(np.asarray(fn(_np(a), _np(b))))

    def cumscan(x, axis, reverse=False, trim="none", pad_lo=0, pad_hi=0, padding=None,
                fill_value=0.0, pre=None, post=None, skipna=True):
        r = oracle.cumscan(_np(x), axis, reverse, trim, pad_lo, pad_hi,
                           padding if (pad_lo or pad_hi) else None, fill_value, _np(pre), _np(post), skipna)
        return _t(r)

    def wreduce(x, axis, weight=None, mode="sum", skipna=True):
        return _t(np.asarray(oracle.wreduce(_np(x), _np(weight), axis, mode, skipna)))

    def vinterp_linear(phi, theta, target, axis, mask_edges=False, bypass_checks=False,
                       logarithmic=False):
        p = _np(phi)
        th = np.broadcast_to(_np(theta), p.shape)
        tg = _np(target)
        if tg.ndim <= 1:
            return _t(oracle.vinterp_linear(p, th, tg.reshape(-1), axis, mask_edges, bypass_checks, logarithmic))
        # one level vector per column: loop the oracle over the columns
        pm = np.moveaxis(p, axis, -1)
        tm = np.moveaxis(th, axis, -1)
        tgb = np.broadcast_to(tg, pm.shape[:-1] + (tg.shape[-1],))
        cols = pm.reshape(-1, pm.shape[-1])
        out = [oracle.vinterp_linear(cols[c], tm.reshape(-1, tm.shape[-1])[c], tgb.reshape(-1, tg.shape[-1])[c],
                                     0, mask_edges, bypass_checks, logarithmic) for c in range(cols.shape[0])]
        return _t(np.stack(out).reshape(pm.shape[:-1] + (tg.shape[-1],)))

    def stencil_multi(x, specs):
        r = _np(x)
        for axis, op, lo, hi, padding, fill in specs:
            r = oracle.stencil2(op, r, axis, lo, hi, padding if (lo or hi) else None, fill)
        return _t(r.astype(_np(x).dtype))

    monkeypatch.setattr(ops, "stencil_multi", stencil_multi)

    def vinterp_conservative(phi, theta, target_bins, axis):
        p = _np(phi)
        tshape = list(p.shape)
        tshape[axis] += 1
        return _t(oracle.vinterp_conservative(p, np.broadcast_to(_np(theta), tshape), _np(target_bins), axis))

    monkeypatch.setattr(ops, "vinterp_conservative", vinterp_conservative)

    def strided_copy(dst, dst_offset, dst_strides, src, src_offset, src_strides, shape, negate=False):
        """The definition of xg_strided_copy, index tuple by index tuple (halo slabs are small)."""
        d, s = dst.numpy().reshape(-1), src.numpy().reshape(-1)
        idx = np.indices([int(n) for n in shape]).reshape(len(shape), -1)
        di = int(dst_offset) + (np.asarray(dst_strides, dtype=np.int64)[:, None] * idx).sum(0)
        si = int(src_offset) + (np.asarray(src_strides, dtype=np.int64)[:, None] * idx).sum(0)
        d[di] = -s[si] if negate else s[si]

    monkeypatch.setattr(ops, "strided_copy", strided_copy)
    monkeypatch.setattr(ops, "strided_copy_batch", lambda copies: [strided_copy(*c) for c in copies] and None)

    def cumscan_host(x, axis, reverse=False, trim="none", pad_lo=0, pad_hi=0, padding=None, fill_value=0.0,
                     pre=None, post=None, skipna=True, device=None):
        return cumscan(x, axis, reverse, trim, pad_lo, pad_hi, padding, fill_value, pre, post, skipna).numpy()

    def wreduce_host(x, axis, weight=None, mode="sum", skipna=True, device=None):
        return wreduce(x, axis, weight, mode, skipna).numpy()

    def vinterp_linear_host(phi, theta, target, axis, mask_edges=False, bypass_checks=False, logarithmic=False,
                            device=None):
        return vinterp_linear(phi, theta, target, axis, mask_edges, bypass_checks, logarithmic).numpy()

    def stencil_pair(a, b, spec_a, spec_b, subtract=0, pre_a=None, pre_b=None, post=None):
        op_a, lo_a, hi_a, pad_a, fill_a = spec_a
        axis_b, op_b, lo_b, hi_b, pad_b, fill_b = spec_b
        aa = _np(a)
        r = oracle.stencil_pair(op_a, aa, aa.ndim - 1, lo_a, hi_a, pad_a, fill_a, _np(pre_a), op_b, _np(b), axis_b, lo_b,
                                hi_b, pad_b, fill_b, _np(pre_b), subtract, _np(post))
        return _t(r)

    monkeypatch.setattr(ops, "stencil_pair", stencil_pair)

    def stencil2_host_multi(x, specs, outs=None, device=None):
        return [stencil2(x, ax, op, lo, hi, pad_ if (lo or hi) else None, 0.0 if fv is None else fv).numpy()
                for ax, op, lo, hi, pad_, fv in specs]

    for name, fn in dict(stencil2=stencil2, stencil2_host=stencil2_host, pad=pad, binary=binary,
                         cumscan=cumscan, wreduce=wreduce, vinterp_linear=vinterp_linear,
                         cumscan_host=cumscan_host, wreduce_host=wreduce_host,
                         vinterp_linear_host=vinterp_linear_host,
                         stencil2_host_multi=stencil2_host_multi).items():
        monkeypatch.setattr(ops, name, fn)
