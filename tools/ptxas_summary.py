#!/usr/bin/env python
"""Registers / spills / shared memory of every kernel of libdips_b200.so as ptxas reports them (no GPU needed).
    python tools/ptxas_summary.py [source.cu ...]        default: all sources of the library"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dips_b200 import _build  # noqa: E402


def main():
    srcs = sys.argv[1:] or [os.path.join(_build.CSRC, s) for s in _build.SOURCES if s.endswith(".cu")]
    for src in srcs:
        cmd = [_build.nvcc(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xptxas", "-v",
               "-c", "-o", "/dev/null", src]
        err = subprocess.run(cmd, capture_output=True, text=True).stderr
        name = None
        for line in err.splitlines():
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
                name = re.sub(r"dipsb::\(anonymous namespace\)::|\(anonymous namespace\)::", "", name).split("(")[0]
                continue
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and name:
                stack = m.groups()
                continue
            m = re.search(r"Used (\d+) registers", line)
            if m and name:
                print(f"{os.path.basename(src):18s} {name:70s} regs {m.group(1):>3s}  stack {stack[0]:>4s}  spill st/ld {stack[1]:>3s}/{stack[2]:<3s}")
                name = None


if __name__ == "__main__":
    main()
