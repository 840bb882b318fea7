"""NumPy twin of the C oracle -- an independent second restatement used by tests/ to cross-check
oracle/dips_oracle.c on small frames.  TEST INFRASTRUCTURE ONLY (see oracle/dips_oracle.h).
PARITY UNPINNED.  Citations are to the reference's WGSL, relative to the reference root.
"""
from __future__ import annotations

import numpy as np

MASK64 = (1 << 64) - 1
GOLD = 0x9E3779B97F4A7C15
BG_SALT = 0xB5AD4ECEDA1CE2A9


def _bpp(fmt):
    return 3 if fmt in (0, 2) else 4


def _rgb(frame, fmt):
    px = np.asarray(frame, dtype=np.uint8).reshape(-1, _bpp(fmt)).astype(np.int32)
    if fmt in (2, 3):
        return px[:, 2], px[:, 1], px[:, 0]
    return px[:, 0], px[:, 1], px[:, 2]


def intensity2(frame, fmt, chroma=0):
    """dips_shader.wgsl:64-82 as the integer max+min (or 2*channel)."""
    r, g, b = _rgb(frame, fmt)
    if chroma == 1:
        return (2 * r).astype(np.uint16)
    if chroma == 2:
        return (2 * g).astype(np.uint16)
    if chroma == 3:
        return (2 * b).astype(np.uint16)
    return (np.maximum(np.maximum(r, g), b) + np.minimum(np.minimum(r, g), b)).astype(np.uint16)


def median4(frames4, fmt, chroma=0):
    """pre_compute_shader.wgsl:103-131: element [2] of the ascending sort of 4 intensities."""
    st = np.stack([intensity2(f, fmt, chroma) for f in frames4])
    return np.sort(st, axis=0)[2].astype(np.uint16)


def run_clip(frames, fmt, mode, tau, chroma=0, state=None):
    frames = np.asarray(frames, dtype=np.uint8)
    n = frames.shape[0]
    i2 = np.stack([intensity2(frames[t], fmt, chroma) for t in range(n)]).astype(np.int64)
    if state is None:
        state = i2[0].copy()
    state = np.asarray(state, dtype=np.int64)
    if mode == 0:
        d = np.abs(i2 - state[None, :])
        out_state = state
    else:
        prev = np.concatenate([state[None, :], i2[:-1]], axis=0)
        d = np.abs(i2 - prev)
        out_state = i2[-1]
    m = d > tau
    return dict(acc_sum=d.sum(0).astype(np.uint32), acc_cnt=m.sum(0).astype(np.uint32),
                sad=d.sum(1).astype(np.uint64), cnt=m.sum(1).astype(np.uint64), state=out_state.astype(np.uint16))


def mix64(z):
    z &= MASK64
    z ^= z >> 30
    z = (z * 0xBF58476D1CE4E5B9) & MASK64
    z ^= z >> 27
    z = (z * 0x94D049BB133111EB) & MASK64
    z ^= z >> 31
    return z


def _hash_byte(seed, index):
    v = mix64((seed + ((index >> 3) + 1) * GOLD) & MASK64)
    return (v >> (8 * (index & 7))) & 0xFF


def synth_byte(seed, profile, t, i, width, height, fmt):
    """Pure-Python definition of the synthetic clip (small cases only)."""
    bpp = _bpp(fmt)
    fb = width * height * bpp
    g = t * fb + i
    if profile == 0:
        return _hash_byte(seed, g)
    bg = _hash_byte(seed ^ BG_SALT, i)
    noise = (_hash_byte(seed, g) % 17) - 8
    p = i // bpp
    x, y = p % width, p // width
    bx, by = (t * 7) % width, (t * 3) % height
    bw, bh = width * 5 // 16, height * 5 // 16
    dx, dy = (x + width - bx) % width, (y + height - by) % height
    v = bg + noise + (64 if (dx < bw and dy < bh) else 0)
    return min(255, max(0, v))


def visual_diff(s_i2, filt, sig_scalar=5.0):
    """dips_shader.wgsl:213-229 in f32."""
    f = np.float32
    d = f(s_i2) / f(510.0)
    d = f(d * f(0.5))
    with np.errstate(all="ignore"):
        if filt == 0:
            d = f(f(1.0) / (f(1.0) + np.exp(f(-sig_scalar) * d, dtype=np.float32)) - f(0.5))
        elif filt == 1:
            d = f(-np.log(f(f(1.0) / f(d + f(0.5))) - f(1.0), dtype=np.float32) / f(sig_scalar))
    return f(d * f(5.0))
