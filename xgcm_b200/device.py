"""Device residency helpers (torch = allocator / stream owner only).

Host arrays handed to a Grid method are copied to the GPU, processed by the
CUDA kernels and copied back; CUDA tensors stay resident.  There is no CPU
compute path: without a CUDA device these helpers raise.
"""

from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

_FLOAT = (np.dtype("float32"), np.dtype("float64"))


def default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "xgcm_b200 needs a CUDA device: the stencil engine has no CPU fallback"
        )
    return torch.device("cuda", torch.cuda.current_device())


def as_device_tensor(data, device=None) -> Tuple[torch.Tensor, bool]:
    """Return ``(contiguous CUDA tensor, was_host)``.

    Integer / bool fields are promoted to float64 (the kernels are fp32/fp64).
    """
    if isinstance(data, torch.Tensor):
        if not data.is_cuda:
            raise RuntimeError("CPU torch tensors are not supported; pass numpy arrays or CUDA tensors")
        t = data
        if t.dtype not in (torch.float32, torch.float64):
            t = t.to(torch.float64)
        return t.contiguous(), False
    arr = np.asarray(data)
    if arr.dtype not in _FLOAT:
        arr = arr.astype(np.float64)
    if not arr.flags.c_contiguous:
        arr = np.ascontiguousarray(arr)
    if not arr.flags.writeable:
        arr = arr.copy()
    dev = device if device is not None else default_device()
    return torch.from_numpy(arr).to(dev, non_blocking=True), True


def result_like(t: torch.Tensor, was_host: bool):
    """Give the result the residency of the input: numpy for host inputs."""
    if was_host:
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)  # torch's caching host allocator
        host.copy_(t)
        return host.numpy()
    return t
