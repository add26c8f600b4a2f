"""Builds libfdm_b200.so in-tree with nvcc for sm_100a (no torch extension machinery needed:
the library has a plain C ABI and is loaded with ctypes)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "fdm_b200.cu")
OUT = os.path.join(HERE, "libfdm_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def build(force: bool = False, out: str = OUT, extra_flags=()) -> str:
    hdr = os.path.join(ROOT, "include", "fdm_b200.h")
    csrc = os.path.dirname(SRC)
    deps = [hdr] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh", ".h"))]
    if not force and os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(d) for d in deps):
        return out
    cmd = [nvcc_path(), *NVCC_FLAGS, *extra_flags, "-I", os.path.join(ROOT, "include"), "-o", out, SRC]
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    print(build(force=True))
