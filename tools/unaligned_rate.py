"""Batch rate of an UNALIGNED device clip (pitch = frame bytes + 8) against the aligned one: 1080p RGB8, 600 frames."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import dips_b200  # noqa: E402

W, H, N = 1920, 1080, 600
fb = W * H * 3


def rate(pitch):
    buf = torch.empty(N * pitch + 64, dtype=torch.uint8, device="cuda")
    tight = torch.empty(N * fb, dtype=torch.uint8, device="cuda")
    dips_b200.synth_fill_device(0, tight.data_ptr(), 0, N, W, H, dips_b200.FMT_RGB8)
    buf[: N * pitch].view(N, pitch)[:, :fb] = tight.view(N, fb)
    del tight
    with dips_b200.Context(W, H, dips_b200.FMT_RGB8, 0, 32) as ctx:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s = torch.cuda.Stream()
        ctx.set_stream(s.cuda_stream)
        with torch.cuda.stream(s):
            for _ in range(3):
                ctx.reset(); ctx.run_clip_device(buf.data_ptr(), N, pitch, 0)
            e0.record(s)
            for _ in range(10):
                ctx.reset(); ctx.run_clip_device(buf.data_ptr(), N, pitch, 0)
            e1.record(s)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        sad = int(ctx.get_scalars(0, N)[0].sum())
    return N / ms * 1e3, ms, sad


if __name__ == "__main__":
    a = rate(fb)
    u = rate(fb + 8)
    assert a[2] == u[2]
    print(f"aligned {a[0]:.0f} frames/s ({a[1]:.3f} ms)  unaligned pitch {u[0]:.0f} frames/s ({u[1]:.3f} ms)  ratio {u[0] / a[0]:.2f}")
