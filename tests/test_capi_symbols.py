"""The C-ABI library loads here (no GPU) and exports every symbol the header declares."""

import ctypes
import os
import re

from xgcm_b200 import _build, _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "xgcm_b200.h")).read()
    return sorted(set(re.findall(r"XG_API\s+[\w\s\*]+?\b(xg_\w+)\s*\(", text)))


def test_header_declares_the_whole_boundary():
    names = _declared()
    for required in ("xg_stencil2", "xg_cumscan", "xg_wreduce", "xg_vinterp_linear", "xg_pad",
                     "xg_binary", "xg_stencil2_host", "xg_last_error", "xg_version"):
        assert required in names


def test_library_builds_and_exports_every_declared_symbol():
    path = _build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(str(path))
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in include/xgcm_b200.h but not exported"
    assert set(_capi.SIGNATURES) == set(_declared())


def test_loader_attaches_prototypes_and_reports_errors_without_gpu():
    lib = _capi.load()
    assert lib.xg_version() == 100
    # argument validation happens before any CUDA call: usable on a CPU-only box
    rc = lib.xg_stencil2(0, 0, None, None, 1, None, 0, 1, 0, 2, 0.0, None, None, None, None, None, None, None)
    assert rc == -1
    assert "null" in _capi.last_error()
    try:
        _capi.check(rc)
    except ValueError:
        pass
    else:  # pragma: no cover
        raise AssertionError("XG_EINVAL must map to ValueError")
