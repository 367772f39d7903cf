"""Build libsdtf.so (the C-ABI engine) in-tree with nvcc for sm_100a.  Cross-compiles without a GPU."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsdtf.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    out = [os.path.join(ROOT, "include", "sdtf.h")]
    for f in sorted(os.listdir(CSRC)):
        out.append(os.path.join(CSRC, f))
    return out


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources())


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-shared", "-o", LIB, os.path.join(CSRC, "sdtf.cu")]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True, cwd=ROOT)
    return LIB


def build_test_gemm() -> str:
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    out = os.path.join(ROOT, "build", "test_gemm")
    src = os.path.join(ROOT, "tests", "cuda", "test_gemm.cu")
    if os.path.exists(out) and os.path.getmtime(out) > max(os.path.getmtime(s) for s in _sources() + [src]):
        return out
    cmd = [_nvcc(), *[f for f in NVCC_FLAGS if f not in ("-Xcompiler", "-fPIC")], "-I", CSRC, src, "-o", out]
    subprocess.run(cmd, check=True, cwd=ROOT)
    return out


if __name__ == "__main__":
    print(build_lib(force=True, verbose=True))
