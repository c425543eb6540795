"""Build liblaud_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m laudnet_b200.build [--force]
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "liblaud_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


# sources the mask-conditioned convolution kernels are built from (the kernels profiles/conv_traffic.json describes)
CONV_SOURCES = ("conv_tma.cu", "conv_umma.cu", "conv_ref.cu", "c_api.cu", "umma_ptx.cuh", "laud_common.cuh")


def source_hash() -> str:
    """sha256 over the sources of the convolution kernels (16 hex digits): ties a committed ncu capture
    (profiles/conv_traffic.json) to the build of those kernels it was taken from."""
    import hashlib
    h = hashlib.sha256()
    paths = sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + sorted(glob.glob(os.path.join(ROOT, "include", "*.h")))
    for path in paths:
        if os.path.basename(path) not in CONV_SOURCES:
            continue
        h.update(os.path.basename(path).encode())
        h.update(open(path, "rb").read())
    return h.hexdigest()[:16]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_prof() -> str:
    """Diagnostic build with in-kernel lap timers (-DLAUD_KPROF) -> lib/liblaud_b200_prof.so;
    load it with LAUD_LIB=<path> (scripts/kprof.py)."""
    out = os.path.join(LIBDIR, "liblaud_b200_prof.so")
    cmd = [_nvcc(), *[f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")], "-DLAUD_KPROF", "-I",
           os.path.join(ROOT, "include"), *sources(), "-o", out]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + (proc.stdout + proc.stderr)[-4000:])
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", os.path.join(ROOT, "include"), *sources(), "-o", LIB + ".tmp"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    with open(os.path.join(LIBDIR, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log[-4000:])
    os.replace(LIB + ".tmp", LIB)
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    if "--prof" in sys.argv:
        print(build_prof())
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
