"""Per-frame call rates (1080p RGBx8 in, RGBA8 out) without the rest of bench.py: pageable / page-locked buffers,
synchronous / pipelined call.  Prints one line; environment knobs (DIPSB_FRAME_BANDS, DIPSB_COPY_THREADS, ...) apply."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dips_b200  # noqa: E402

W, H, N = 1920, 1080, 200


def main():
    rng = np.random.default_rng(1)
    frames = rng.integers(0, 256, (8, W * H * 4), dtype=np.uint8)
    pin_in = dips_b200.PinnedBuffer(8 * W * H * 4)
    pin_out = dips_b200.PinnedBuffer(2 * W * H * 4)
    pin_in.array[:] = frames.reshape(-1)
    res = {}
    for kind in ("pageable", "pinned"):
        src = frames if kind == "pageable" else pin_in.array.reshape(8, -1)
        dst = np.zeros((2, W * H * 4), np.uint8) if kind == "pageable" else pin_out.array.reshape(2, -1)
        for name in ("sync", "pipelined"):
            with dips_b200.Context(W, H, dips_b200.FMT_RGBX8, 0, 16) as ctx:
                fn = ctx.push_frame if name == "sync" else ctx.push_frame_pipelined
                for k in range(8):
                    fn(src[k % 8], out=dst[k & 1])
                t0 = time.perf_counter()
                for k in range(N):
                    fn(src[k % 8], out=dst[k & 1])
                if name == "pipelined":
                    ctx.flush_frame(out=dst[N & 1])
                res[f"{kind}_{name}"] = round(N / (time.perf_counter() - t0))
    print(" ".join(f"{k}={v}" for k, v in res.items()))


if __name__ == "__main__":
    main()
