"""Builds libfdm_b200.so in-tree with nvcc for sm_100a (no torch extension machinery needed:
the library has a plain C ABI and is loaded with ctypes).

The kernels are compiled as one translation unit per (family, size class) -- fdm_warp_inst.cu,
fdm_lean_inst.cu and fdm_cta_inst.cu with ``-DFDM_INST_C`` / ``-DFDM_INST_W`` -- in parallel, objects
cached under ``build/obj`` by source time stamps, then linked with fdm_api.cu (the C ABI)."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libfdm_b200.so")
OBJ_DIR = os.path.join(ROOT, "build", "obj")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]

LEAN_CLASSES = (10, 13, 16)
LEAN_ALL_CLASSES = (4, 7, 8, 10, 13, 16)
CTA_CLASSES = ((4, 12), (4, 16), (5, 14), (5, 16), (8, 12), (8, 16), (16, 12), (16, 16))   # = FDM_CTA_CLASSES, fdm_internal.h      # with -DFDM_LEAN_ALL_CLASSES (layout parity tests at the other sizes)

# (object name, source, defines)
def units(lean_classes=LEAN_CLASSES):
    u = [("api", "fdm_api.cu", [])]
    u += [(f"warp_{c}", "fdm_warp_inst.cu", [f"-DFDM_INST_C={c}"]) for c in (4, 8, 13, 16)]
    u += [(f"lean_{c}", "fdm_lean_inst.cu", [f"-DFDM_INST_C={c}"]) for c in lean_classes]
    u += [(f"cta_{w}" if c == 16 else f"cta_{w}x{c}", "fdm_cta_inst.cu", [f"-DFDM_INST_W={w}", f"-DFDM_INST_C={c}"])
          for w, c in CTA_CLASSES]
    u += [(f"rowvar_{c}", "fdm_rowvar_inst.cu", [f"-DFDM_INST_C={c}"]) for c in (7, 10, 13, 19, 26, 32)]
    # the large instances first: the longest compile bounds the build
    order = {"cta_16": 0, "cta_16x12": 0, "cta_8": 1, "cta_8x12": 1, "cta_5": 2, "cta_5x14": 2, "cta_4": 3, "cta_4x12": 3}
    return sorted(u, key=lambda x: order.get(x[0], 10))


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _deps():
    hdr = os.path.join(ROOT, "include", "fdm_b200.h")
    return [hdr] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]


def build(force: bool = False, out: str = OUT, extra_flags=(), jobs: int | None = None, only=None) -> str:
    """Compile what is out of date and link ``out``.  ``extra_flags`` (e.g. ``-DFDM_LEAN_ALL_CLASSES``) select an
    alternative build with its own object cache.  ``only``: restrict the compile step to the named units (dev use)."""
    extra_flags = list(extra_flags)
    lean_classes = LEAN_ALL_CLASSES if "-DFDM_LEAN_ALL_CLASSES" in extra_flags else LEAN_CLASSES
    tag = hashlib.sha1(" ".join(extra_flags).encode()).hexdigest()[:8] if extra_flags else "default"
    obj_dir = os.path.join(OBJ_DIR, tag)
    os.makedirs(obj_dir, exist_ok=True)
    newest = max(os.path.getmtime(d) for d in _deps())
    nvcc = nvcc_path()
    todo, objs = [], []
    for name, src, defs in units(lean_classes):
        obj = os.path.join(obj_dir, name + ".o")
        objs.append(obj)
        if only is not None and name not in only and os.path.exists(obj):
            continue
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < newest:
            todo.append([nvcc, *NVCC_FLAGS, *extra_flags, *defs, "-I", os.path.join(ROOT, "include"), "-I", CSRC,
                         "-c", os.path.join(CSRC, src), "-o", obj])
    if not todo and os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(o) for o in objs):
        return out
    jobs = jobs or max(1, min(len(todo), os.cpu_count() or 1))
    if todo:
        with ThreadPoolExecutor(jobs) as pool:
            for res in pool.map(lambda c: subprocess.run(c, capture_output=True, text=True), todo):
                if res.returncode != 0:
                    raise RuntimeError("nvcc failed:\n" + " ".join(res.args) + "\n" + res.stdout + res.stderr)
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "shared",
                    "-o", out, *objs], check=True)
    return out


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv))
