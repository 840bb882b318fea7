#!/usr/bin/env python
"""Device times of the once-per-clip kernels around the clip kernel (prime, reset, finalize is inside run) -- CUDA events on
the context's stream, 20 repetitions after a warm-up.  python tools/aux_rate.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dips_b200


def timed(stream, fn, reps=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    for name, w, h, fmt in (("1080p RGB8", 1920, 1080, 0), ("1080p RGBx8", 1920, 1080, 1), ("4K RGBx8", 3840, 2160, 1), ("8K RGB8", 7680, 4320, 0)):
        fb = w * h * dips_b200.bytes_per_pixel(fmt)
        frames = torch.empty(4 * fb, dtype=torch.uint8, device="cuda")
        dips_b200.synth_fill_device(0, frames.data_ptr(), 0, 4, w, h, fmt, stream=stream.cuda_stream)
        with dips_b200.Context(w, h, fmt, 0, 32) as ctx:
            ctx.set_stream(stream.cuda_stream)
            t_prime = timed(stream, lambda: ctx.prime_device(frames.data_ptr()))
            t_prime_un = timed(stream, lambda: ctx.prime_device(frames.data_ptr() + fb + 1)) if fb % 16 == 0 else float("nan")
            t_reset = timed(stream, ctx.reset)
            ctx.reset()
            t_run4 = timed(stream, lambda: (ctx.reset(), ctx.run_clip_device(frames.data_ptr(), 4, fb, 0)))
            print(f"{name:12s} prime {t_prime:7.1f} us ({(fb + w * h * 2) / t_prime / 1e3:6.0f} GB/s)   prime(unaligned) {t_prime_un:7.1f} us   "
                  f"reset {t_reset:6.1f} us   reset+4-frame pass {t_run4:7.1f} us")


if __name__ == "__main__":
    main()
