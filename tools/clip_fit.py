#!/usr/bin/env python
"""Clip kernel time against the number of frames of the launch (fixed cost per launch vs cost per frame).
python tools/clip_fit.py [workload geometry: 4k|8k|1080p] [extra tuning: stages tile_px]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dips_b200

GEO = {"4k": (3840, 2160, 1), "8k": (7680, 4320, 0), "1080p": (1920, 1080, 0)}


def main():
    geo = sys.argv[1] if len(sys.argv) > 1 else "4k"
    stages = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    tile_px = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    mode = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    w, h, fmt = GEO[geo]
    fb = w * h * dips_b200.bytes_per_pixel(fmt)
    nmax = int(min(900, 30e9 // fb))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    clip = torch.empty(nmax * fb, dtype=torch.uint8, device="cuda")
    dips_b200.synth_fill_device(0, clip.data_ptr(), 0, nmax, w, h, fmt, stream=stream.cuda_stream)
    torch.cuda.synchronize()
    with dips_b200.Context(w, h, fmt, mode, 32) as ctx:
        ctx.set_stream(stream.cuda_stream)
        if stages or tile_px:
            ctx.set_tuning(stages, tile_px, 0, 0)
        ctx.enable_timing(True)
        rows = []
        for n in [x for x in (4, 16, 64, 128, 256, 450, 900) if x <= nmax]:
            for _ in range(3):
                ctx.reset(); ctx.run_clip_device(clip.data_ptr(), n, fb, 0)
            ctx.clip_kernel_time()
            for _ in range(10):
                ctx.reset(); ctx.run_clip_device(clip.data_ptr(), n, fb, 0)
            ms, k = ctx.clip_kernel_time()
            rows.append((n, ms / k))
            print(f"{geo} n={n:4d}  kernel {1e3 * ms / k:8.1f} us   {n * fb / (ms / k) / 1e6:7.0f} GB/s   plan {ctx.last_plan()['tiles']} tiles x{ctx.last_plan()['segments']} seg, {ctx.last_plan()['stages']} stages")
        (n0, t0), (n1, t1) = rows[-3], rows[-1]
        b = (t1 - t0) / (n1 - n0)
        print(f"per frame {1e3 * b:.2f} us ({fb / b / 1e6:.0f} GB/s asymptotic), fixed {1e3 * (t1 - b * n1):.1f} us per launch")


if __name__ == "__main__":
    main()
