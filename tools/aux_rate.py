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


def ring_rate():
    """batch call of the ring flavours and of windowed contexts over a device-resident 1080p clip: ring_clip_kernel (one launch
    per run of frames, ring in registers) next to the per-frame kernels back to back (DIPSB_RING_BATCH=0; their u16 / u32
    planes stay in L2 between frames, unlike in an ncu launch list)"""
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    for fmt_name, fmt, n in (("RGBx8", 1, 600), ("RGB8", 0, 600)):
        w, h = 1920, 1080
        fb = w * h * dips_b200.bytes_per_pixel(fmt)
        clip = torch.empty(n * fb, dtype=torch.uint8, device="cuda")
        dips_b200.synth_fill_device(0, clip.data_ptr(), 0, n, w, h, fmt, stream=stream.cuda_stream)
        for name, flavor, window in (("dips ring-of-4", dips_b200.FLAVOR_DIPS_RING4, 1), ("dips_alt ring-of-2", dips_b200.FLAVOR_ALT_RING2, 1),
                                     ("frame0, 3x3 median", dips_b200.FLAVOR_FRAME0, 3), ("frame0, 7x7 median", dips_b200.FLAVOR_FRAME0, 7)):
            if window > 1 and fmt != 1:
                continue
            for batch in ((True, False) if window == 1 else (False,)):
                os.environ["DIPSB_RING_BATCH"] = "1" if batch else "0"
                with dips_b200.Context(w, h, fmt, 0, 32, flavor=flavor, spatial_window=window) as ctx:
                    ctx.set_stream(stream.cuda_stream)
                    m = n if batch else 200
                    t = timed(stream, lambda: (ctx.reset(), ctx.run_clip_device(clip.data_ptr(), m, fb, 0)), reps=3)
                    how = "ring_clip_kernel" if ctx.last_plan()["ring_clip"] else "per-frame kernels"
                    print(f"1080p {fmt_name} batch, {name:20s} {how:18s}: {t / m:6.2f} us per frame ({m * 1e6 / t:8.0f} frames/s, "
                          f"{fb * m / t / 1e3:6.0f} GB/s of frame bytes)")
        os.environ.pop("DIPSB_RING_BATCH", None)
        del clip


if __name__ == "__main__":
    ring_rate()
    main()
