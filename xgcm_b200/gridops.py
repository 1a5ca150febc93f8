"""The built-in grid ufuncs: ``<op>_<from>_to_<to>`` for op in diff, interp, min, max.

Same names, signatures and ``padding_width`` table as the reference's
``xgcm/gridops.py:27-215`` so ``_select_grid_ufunc`` (name prefix + signature
equivalence, grid.py:1779-1824) keeps working.  Each object carries
``kernel_op`` so the engine runs halo + operator + metrics as one fused
``xg_stencil2`` launch; ``.ufunc(a)`` applies the bare pairwise operator along
the last axis of an already padded array (numpy in -> numpy out through the GPU,
CUDA tensor in -> CUDA tensor out).
"""

from __future__ import annotations

from .grid_ufunc import GridUFunc

# (from, to) -> halo widths; identical for every operator (gridops.py:27-65)
_SHIFTS = {
    ("center", "left"): (1, 0),
    ("left", "center"): (0, 1),
    ("center", "right"): (0, 1),
    ("right", "center"): (1, 0),
    ("center", "outer"): (1, 1),
    ("outer", "center"): (0, 0),
    ("center", "inner"): (0, 0),
    ("inner", "center"): (1, 1),
}


def _make_raw(op):
    def raw(a):
        import numpy as np
        import torch

        from . import ops
        from .device import as_device_tensor, result_like

        if not isinstance(a, (np.ndarray, torch.Tensor)):
            raise TypeError(f"grid ufunc kernels take arrays, got {type(a)}")
        x, was_host = as_device_tensor(a)
        return result_like(ops.stencil2(x, -1, op, 0, 0, None), was_host)

    raw.kernel_op = op
    raw.__name__ = f"{op}_forward"
    return raw


def _define(op):
    raw = _make_raw(op)
    for (src, dst), width in _SHIFTS.items():
        name = f"{op}_{src}_to_{dst}"
        globals()[name] = GridUFunc(
            raw, signature=f"(X:{src})->(X:{dst})", padding_width={"X": width}, kernel_op=op
        )
    globals()[f"{op}_forward" if op in ("diff", "interp") else f"pairwise_forward_{op}"] = raw


for _op in ("diff", "interp", "min", "max"):
    _define(_op)


def _left_to_inner(a):
    raise NotImplementedError


# reference gridops.py:68-70: declared but unimplemented
diff_left_to_inner = GridUFunc(_left_to_inner, signature="(X:left)->(X:inner)")
