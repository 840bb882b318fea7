"""A short per-frame session for the ncu launch list of the streaming kernels (frame_kernel, ring_kernel,
spatial_median_kernel, passthrough): 1080p RGBx8, 6 frames per configuration."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dips_b200  # noqa: E402

W, H, N = 1920, 1080, 6


def main():
    import torch
    dev = torch.empty(N * W * H * 4, dtype=torch.uint8, device="cuda")
    dips_b200.synth_fill_device(0, dev.data_ptr(), 0, N, W, H, dips_b200.FMT_RGBX8)
    torch.cuda.synchronize()
    clip = dev.cpu().numpy().reshape(N, -1)
    out = np.empty(W * H * 4, np.uint8)
    for name, kw in (("frame0 grey", dict()),
                     ("frame0 sigmoid colour", dict(colorize=True, filt=dips_b200.FILTER_SIGMOID)),
                     ("dips ring-of-4", dict(flavor=dips_b200.FLAVOR_DIPS_RING4, colorize=True, filt=dips_b200.FILTER_SIGMOID)),
                     ("dips_alt ring-of-2", dict(flavor=dips_b200.FLAVOR_ALT_RING2, colorize=True, filt=dips_b200.FILTER_SIGMOID)),
                     ("frame0 window 3", dict(spatial_window=3)),
                     ("frame0 window 7", dict(spatial_window=7))):
        with dips_b200.Context(W, H, dips_b200.FMT_RGBX8, 0, 16, **kw) as ctx:
            for t in range(N):
                ctx.push_frame(clip[t], out=out)
        print(name, "ok", int(out[:64].sum()))


if __name__ == "__main__":
    main()
