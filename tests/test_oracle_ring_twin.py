"""The integer restatement of the reference's temporal rings (oracle/numpy_twin.py::ring_clip -- the checker of the batch
execution of the ring flavours, ring_clip_kernel) against the oracle's f32 restatement of the shaders' state machines
(oracle/dips_oracle.c: dipso_cs_frame = dips ComputeState, dipso_alt_frame = dips_alt DiPsCompute).  CPU only.

The f32 machines return the visual frame, not the difference; with the filter off and colourising off the grey level is
round(255 * clamp(0.5 - 2.5 * S / 510)) = 127.5 - 1.25 * S (dips_shader.wgsl:217-237: x0.5, x5, 0.5 - diff), so S = start -
median is recovered to within one grey level = 0.8 I2 units while |S| < 102 (a little more where the two restatements break
a rounding tie of an rgba8unorm store differently: one grey level of the start plane or of a ring slot = 2 I2 units)."""
import numpy as np
import pytest


def _estimate_s(rgba):
    grey = rgba.reshape(-1, 4)[:, 0].astype(np.float64)
    return (127.5 - grey) / 1.25


@pytest.mark.parametrize("chroma", [0, 2])
def test_dips_ring_of_4_twin_follows_the_shader_state_machine(oracle, twin, chroma):
    w, h, n = 96, 40, 14
    clip = oracle.synth_clip(n, w, h, 1, profile=oracle.SYNTH_SCENE)
    i2 = np.stack([oracle.i2_plane(clip[t], 1, chroma) for t in range(n)]).astype(np.int64)
    _, _, sad, cnt, planes = twin.ring_clip(i2, 1, 10, want_planes=True)
    cs = oracle.ComputeStateOracle(w, h, colorize=False, filt=oracle.FILTER_NONE, chroma=chroma)
    produced = 0
    for t in range(n):
        out, passthrough = cs.frame(clip[t])
        assert passthrough == (planes[t] is None), t                       # three warm-up frames, then a difference per frame
        if passthrough:
            assert sad[t] == 0 and cnt[t] == 0
            continue
        s_est = _estimate_s(out)
        inside = np.abs(planes[t]) < 98                                    # grey level not clamped
        err = np.abs(s_est - planes[t])[inside]
        assert inside.mean() > 0.8 and err.max() <= 4.5 and (err <= 0.9).mean() > 0.97, (t, err.max(), (err <= 0.9).mean())
        produced += 1
    assert produced == n - 3 and int(sad.sum()) > 0


@pytest.mark.parametrize("flavor,intended", [(2, False), (3, True)])
def test_dips_alt_ring_of_2_twin_follows_the_shader_state_machine(oracle, twin, flavor, intended):
    w, h, n = 96, 40, 12
    clip = oracle.synth_clip(n, w, h, 1, profile=oracle.SYNTH_SCENE)
    i2 = np.stack([oracle.i2_plane(clip[t], 1, 0) for t in range(n)]).astype(np.int64)
    snaps = (2, 7)
    _, _, sad, cnt, planes = twin.ring_clip(i2, flavor, 10, snapshot_before=snaps, want_planes=True)
    alt = oracle.DiPsComputeOracle(w, h, colorize=False, filt=oracle.FILTER_NONE, intended_median=intended)
    for t in range(n):
        out = alt.send_frame(clip[t], snapshot=t in snaps)
        if t in snaps:                                                      # the snapshot frame returns the grey snapshot itself
            assert planes[t] is None and sad[t] == 0
            continue
        s_est = _estimate_s(out)
        inside = np.abs(planes[t]) < 98
        err = np.abs(s_est - planes[t])[inside]
        assert err.max() <= 4.5 and (err <= 0.9).mean() > 0.97, (t, err.max(), (err <= 0.9).mean())
    assert int(sad[3:].sum()) > 0
