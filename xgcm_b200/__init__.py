"""xgcm_b200 — B200-native (sm_100a) stencil engine behind xgcm's Grid / Axis API.

Public surface mirrors ``xgcm/__init__.py:6-7`` of the reference
(``Grid``, ``as_grid_ufunc``, ``apply_as_grid_ufunc``) plus the labelled-array
stand-ins used when xarray is not installed.
"""

from .axis import Axis  # noqa: F401
from .grid import Grid  # noqa: F401
from .grid_ufunc import GridUFunc, apply_as_grid_ufunc, as_grid_ufunc  # noqa: F401
from .labeled import DataArray, Dataset  # noqa: F401
from .padding import pad  # noqa: F401
from . import ingest  # noqa: F401  (chunked ingest: disk -> page-locked ring -> device)

__version__ = "0.1.0"
__all__ = [
    "Grid",
    "Axis",
    "GridUFunc",
    "as_grid_ufunc",
    "apply_as_grid_ufunc",
    "pad",
    "DataArray",
    "Dataset",
]
