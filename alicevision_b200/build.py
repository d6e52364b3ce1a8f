"""In-tree build of the native engine: nvcc -> alicevision_b200/libb200match.so (sm_100a only).

The .so is git-ignored but travels to the GPU box with the repo snapshot, so the box never compiles.
"""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libb200match.so")
SOURCES = ["engine.cu", "voctree.cu", "io.cpp"]
def _deps():
    """Every file the library is built from: all of csrc/ and the public headers (a forgotten header here once shipped a stale .so)."""
    import glob
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.cpp"))
                  + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(_HERE, "..", "include", "*.h")) + [os.path.abspath(__file__)])

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",   # explicit form: `-arch=sm_100a` also emits compute_100 PTX, which rejects tcgen05
    "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    if os.environ.get("B200M_NO_BUILD"):       # GPU box: use the library that travelled with the snapshot, whatever the source timestamps say
        return False
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in _deps())


def build_native(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA engine if sources are newer than the library. Returns the library path."""
    if not force and not stale():
        return LIB
    env = dict(os.environ)
    env.pop("CXX", None); env.pop("CC", None)       # the image exports a gcc wrapper without OpenMP specs; nvcc finds /usr/bin/g++
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build_native(force=True, verbose=True))
