"""Build recipe for the C-ABI shared library (nvcc, sm_100a only, in-tree).

``python -m xgcm_b200._build`` or ``__graft_entry__.build()``.  The library
links only against the (static) CUDA runtime: no torch, no python.
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
INCLUDE = PKG_DIR.parent / "include"
BUILD_DIR = PKG_DIR / "csrc" / "build"
LIB_PATH = PKG_DIR / "libxgcm_b200.so"

NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-O3",
    "-std=c++17",
    # numpy rounds after every ufunc: no FMA contraction, IEEE div/sqrt
    "--fmad=false",
    "--prec-div=true",
    "--prec-sqrt=true",
    "--ftz=false",
    "-Xcompiler",
    "-fPIC",
    "-Xcompiler",
    "-fvisibility=hidden",
    "-DXG_BUILDING",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libxgcm_b200.so")
    return nvcc


def sources():
    return sorted(CSRC.glob("*.cu"))


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(src: Path, verbose: bool) -> Path:
    obj = BUILD_DIR / (src.stem + ".o")
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", str(INCLUDE), "-I", str(CSRC), "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stdout}\n{res.stderr}")
    if verbose:
        sys.stderr.write(res.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every ``csrc/*.cu`` for sm_100a and link ``libxgcm_b200.so``."""
    srcs = sources()
    deps = srcs + sorted(CSRC.glob("*.cuh")) + sorted(INCLUDE.glob("*.h"))
    stamp = BUILD_DIR / "digest.txt"
    digest = _digest(deps)
    if not force and LIB_PATH.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB_PATH
    BUILD_DIR.mkdir(parents=True, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile_one(s, verbose), srcs))
    cmd = [
        _nvcc(),
        "-shared",
        "-gencode",
        "arch=compute_100a,code=sm_100a",
        "-o",
        str(LIB_PATH),
        *[str(o) for o in objs],
        "-cudart",
        "static",
    ]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    stamp.write_text(digest)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
