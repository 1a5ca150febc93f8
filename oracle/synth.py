"""Synthetic bench / test fields on the host WITHOUT the product library (test infrastructure).

numpy restatement of the counter-based U(0,1) generator the device uses (`xg_uniform`,
xgcm_b200/csrc/xg_elementwise.cu: splitmix64 finaliser of seed and linear index), so that the reference
arm of bench.py builds bit-identical inputs without importing `xgcm_b200` (inputs keyed by
(seed, linear index) so any sub-block can be produced anywhere).  Only tests/, smoke() and bench.py may
import this package.
"""

from __future__ import annotations

import numpy as np

_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def _mix64(z: np.ndarray) -> np.ndarray:
    z = z + _GOLD
    z = (z ^ (z >> np.uint64(30))) * _M1
    z = (z ^ (z >> np.uint64(27))) * _M2
    return z ^ (z >> np.uint64(31))


def fill_uniform(out: np.ndarray, seed: int, offset: int = 0, block: int = 1 << 22) -> np.ndarray:
    """Fill the C-contiguous array `out` (float32 / float64) in place; element g gets
    uniform(seed, offset + g).  Same bits as xg_fill_uniform / xg_fill_uniform_host."""
    flat = out.reshape(-1)
    if flat.base is None and flat is not out and not np.shares_memory(flat, out):
        raise ValueError("fill_uniform needs a C-contiguous array")
    with np.errstate(over="ignore"):
        s = _mix64(np.array([seed], dtype=np.uint64))[0]
        for g0 in range(0, flat.size, block):
            g1 = min(flat.size, g0 + block)
            idx = np.arange(offset + g0, offset + g1, dtype=np.uint64)
            h = _mix64(s ^ (idx * _GOLD))
            if flat.dtype == np.float32:
                flat[g0:g1] = (h >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)
            elif flat.dtype == np.float64:
                flat[g0:g1] = (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
            else:
                raise TypeError("fill_uniform: float32 or float64 only")
    return out
