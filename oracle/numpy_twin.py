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


def ring_clip(i2, flavor, tau, snapshot_before=(), want_planes=False):
    """Integer restatement of the reference's temporal rings over a clip of I2 planes (frames x pixels): per-pixel sums and
    counts, per-frame sad / cnt -- what dipsb_run_clip_* returns on a ring-flavour context.
    flavor 1, `dips` (dips/src/gpu/mod.rs:170-216, bind_groups.rs:407-427, dips_shader.wgsl:187-214, pre_compute_shader.wgsl:
    103-131): 3 pass-through frames, start = grey(upper median of the first 4), every later frame overwrites the oldest ring
    slot with its grey-quantised intensity, D = |start - sorted4(ring)[2]|; a frame index in `snapshot_before` restarts the
    warm-up.  flavor 2 / 3, `dips_alt` (dips_alt/src/dips_compute/shaders/pre_compute_shader.wgsl:212-262): ring of 2,
    min (as shipped) / max (in-bounds median) of the two, D = |snapshot - median|; at a frame index in `snapshot_before` the
    snapshot becomes grey(median) and the frame contributes nothing.  grey(v) = 2 * ((v + 1) >> 1): the rgba8unorm store of an
    intensity, in I2 units.  With want_planes the signed S = start - median of every frame is returned too (None where no
    difference is produced)."""
    i2 = np.asarray(i2, dtype=np.int64)
    n, npx = i2.shape
    grey = lambda v: 2 * ((v + 1) >> 1)
    s, c = np.zeros(npx, np.int64), np.zeros(npx, np.int64)
    sad, cnt = np.zeros(n, np.int64), np.zeros(n, np.int64)
    planes = [None] * n

    def account(t, signed):
        d = np.abs(signed)
        s[:] += d
        c[:] += d > tau
        sad[t] = d.sum()
        cnt[t] = (d > tau).sum()
        planes[t] = signed

    if flavor == 1:
        ring, start, idx, seen = np.zeros((4, npx), np.int64), None, 0, 0
        for t in range(n):
            if t in snapshot_before:
                seen, idx = 0, 0
            seen += 1
            if seen < 4:
                ring[seen - 1] = i2[t]
                continue
            if seen == 4:
                ring[3] = i2[t]
                start = grey(np.sort(ring, axis=0)[2])
                ring[0] = grey(ring[0])
            else:
                ring[idx] = grey(i2[t])
                idx = (idx + 1) % 4
            account(t, start - np.sort(ring, axis=0)[2])
    else:
        ring, snap, idx = np.zeros((2, npx), np.int64), np.zeros(npx, np.int64), 0
        for t in range(n):
            ring[idx] = i2[t]
            idx ^= 1
            med = ring.max(axis=0) if flavor == 3 else ring.min(axis=0)
            if t in snapshot_before:
                snap = grey(med)
                continue
            account(t, snap - med)
    return (s, c, sad, cnt, planes) if want_planes else (s, c, sad, cnt)


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
