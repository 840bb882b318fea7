"""Builds libdips_b200.so in-tree with nvcc for sm_100a (no torch / setuptools involved: the product is a plain C-ABI
shared library).  `python -m dips_b200._build [--force]`."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "libdips_b200.so")
SOURCES = ["clip_kernel.cu", "aux_kernels.cu", "ring_clip.cu", "api.cu", "comm.cu", "host_copy.cu"]
HEADERS = [os.path.join(CSRC, "dipsb_internal.h"), os.path.join(CSRC, "dipsb_ctx.h"), os.path.join(CSRC, "intensity.cuh"), os.path.join(ROOT, "include", "dips_b200.h")]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC,-pthread", "-shared", "-cudart", "static",
]
LINK_FLAGS = ["-ldl"]    # NCCL is loaded with dlopen when a communicator is made: no link-time dependency


def nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libdips_b200.so")


def is_stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return SO
    cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO] + [os.path.join(CSRC, s) for s in SOURCES] + LINK_FLAGS
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
