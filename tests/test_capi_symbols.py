"""The C-ABI library loads here (no GPU) and exports every symbol the header declares."""

import ctypes
import os
import re

import pytest

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


def test_header_is_plain_c_and_links(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (no C++-isms, no torch types) and a
    C program must link against the library and call it (argument validation needs no GPU)."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:  # pragma: no cover
        import pytest

        pytest.skip("gcc not available")
    lib = _build.build()
    src = tmp_path / "capi_c.c"
    src.write_text(
        '#include "xgcm_b200.h"\n'
        "#include <stdio.h>\n"
        "#include <string.h>\n"
        "int main(void) {\n"
        "  int64_t shape[1] = {4};\n"
        "  int rc = xg_pad(XG_F32, 0, 0, 1, shape, 0, 1, 0, XG_BC_FILL, 0.0, 0);\n"
        '  printf("%d %d %d\\n", xg_version(), rc, (int)(strstr(xg_last_error(), "null") != 0));\n'
        "  return rc == XG_EINVAL ? 0 : 1;\n"
        "}\n"
    )
    exe = tmp_path / "capi_c"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src),
                    "-L", os.path.dirname(str(lib)), "-l:" + os.path.basename(str(lib)),
                    "-Wl,-rpath," + os.path.dirname(str(lib)), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert out == ["100", "-1", "1"]


def test_strided_copy_argument_validation_without_gpu():
    """xg_strided_copy(_batch) validate their arguments before any CUDA call: null pointers, bad
    rank, negative extents, and (batch) a copy that does not collapse to 5 dims -> XG_ENOTIMPL, the
    signal on which ops.strided_copy_batch falls back to one launch per copy."""
    import ctypes as C

    lib = _capi.load()
    i64 = _capi.i64_array
    one = (C.c_void_p * 1)(1)  # never dereferenced: validation fails (or nothing launches) first
    assert lib.xg_strided_copy(0, None, i64([1]), None, i64([1]), 1, i64([4]), 0, None) == -1
    assert lib.xg_strided_copy(0, 1, i64([1]), 1, i64([1]), 9, i64([4]), 0, None) == -1
    assert lib.xg_strided_copy(0, 1, i64([1]), 1, i64([1]), 1, i64([-4]), 0, None) == -1
    assert lib.xg_strided_copy(7, 1, i64([1]), 1, i64([1]), 1, i64([4]), 0, None) == -1
    assert lib.xg_strided_copy_batch(0, 0, None, None, 1, None, None, None, None, None) == 0
    assert lib.xg_strided_copy_batch(0, -1, None, None, 1, None, None, None, None, None) == -1
    assert lib.xg_strided_copy_batch(0, 1, None, None, 1, None, None, None, None, None) == -1
    # six dims with pairwise incompatible strides cannot be collapsed below six
    shape = i64([2, 2, 2, 2, 2, 2])
    dst_strides = i64([32, 16, 8, 4, 2, 1])
    src_strides = i64([1, 2, 4, 8, 16, 32])
    rc = lib.xg_strided_copy_batch(0, 1, one, one, 6, shape, dst_strides, src_strides, (C.c_int * 1)(0), None)
    assert rc == -2 and "5" in _capi.last_error()


def test_comm_error_paths_without_a_gpu():
    """XG_ENCCL is reachable: a NCCL library that cannot be loaded, and argument validation of the sharded
    entry points, without touching a device."""
    import ctypes as C

    from xgcm_b200 import _capi

    lib = _capi.load()
    rc = lib.xg_nccl_load(b"/nonexistent/libnccl.so.2")
    if rc != 0:  # (0 only if an earlier test in this process already loaded a real NCCL)
        assert rc == -4
        assert "NCCL" in _capi.last_error()
        with pytest.raises(RuntimeError):
            _capi.check(rc)
    shape = (C.c_int64 * 2)(4, 8)
    buf = (C.c_float * 32)()
    out = (C.c_float * 32)()
    rc = lib.xg_stencil2_sharded(None, 0, 0, buf, out, 2, shape, 0, 1, 0, 2, 0.0, None, None, None, None, None, 0, None)
    assert rc == -1 and "null communicator" in _capi.last_error()
    assert lib.xg_halo_exchange(None, None, None, None, None, 0, 0, None) == -1
    assert lib.xg_comm_init(None, 2, 0, None) == -1
