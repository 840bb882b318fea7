#!/usr/bin/env python
"""Hot-loop statistics of a clip kernel instantiation from the SASS of libdips_b200.so: the longest backward branch whose
body emits per-frame scalars (REDUX) and never flushes is the stage-unrolled frame loop (one trip = U frames); prints its
length, opcode histogram and the local-memory instructions inside it.
    python tools/sass_loop.py 'clip_kernel_wsILi3ELin1ELi0ELi4E' [path/to/lib.so]"""
import collections
import re
import subprocess
import sys


def main():
    pat = sys.argv[1]
    so = sys.argv[2] if len(sys.argv) > 2 else "dips_b200/libdips_b200.so"
    names = subprocess.run(["cuobjdump", "-elf", so], capture_output=True, text=True).stdout
    fn = sorted(set(re.findall(r"(_ZN5dipsb\w*" + re.escape(pat) + r"\w*)", names)), key=len)[0]
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn, so], capture_output=True, text=True).stdout
    ins = []
    for line in sass.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    best = None
    for k, (addr, text) in enumerate(ins):
        m = re.search(r"BRA(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", text)
        if m:
            tgt = int(m.group(1), 16)
            if tgt >= addr:
                continue
            inner = [t for a, t in ins if tgt <= a <= addr]
            # the stage-unrolled trip loop: emits scalars (REDUX) but never flushes (no REDG / ATOMG inside)
            if not any("REDUX" in t for t in inner) or any(re.search(r"\b(REDG|ATOMG|RED)\b", t.split(".")[0]) for t in inner):
                continue
            if best is None or addr - tgt > best[1] - best[0]:
                best = (tgt, addr)
    lo, hi = best
    body = [t for a, t in ins if lo <= a <= hi]
    frames = sum(1 for t in body if "REDUX" in t)
    print(f"frames per trip: {frames}; instructions per frame: {len(body) / max(frames, 1):.1f}")
    ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for t in body)
    print(f"{fn[-60:]}: total {len(ins)} instructions; frame loop 0x{lo:x}..0x{hi:x} = {len(body)} instructions")
    print("  " + ", ".join(f"{k} {v}" for k, v in ops.most_common(24)))
    loc = [t for t in body if re.search(r"\b(LDL|STL)\b", t)]
    print(f"  local-memory instructions in the loop: {len(loc)}; in the whole kernel: {sum(1 for _, t in ins if re.search(r'(LDL|STL)', t))}")


if __name__ == "__main__":
    main()
