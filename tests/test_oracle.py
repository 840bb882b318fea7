"""CPU tests of the oracle: hand-derived KATs (tests/golden/kats.json), committed synthetic fixtures, the numpy twin, and
the size-independent properties of SURVEY.md 8(c).  The reference ships no vectors (parity unpinned)."""
import hashlib
import json
import math
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")
KATS = json.load(open(os.path.join(GOLD, "kats.json")))
SYNTH = json.load(open(os.path.join(GOLD, "synth_golden.json")))

FILTERS = {"none": 255, "sigmoid5": 0, "inv_sigmoid5": 1}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("kat", KATS["pixels"])
def test_kat_pixels(oracle, kat):
    s = np.array(kat["start"], np.uint8)
    c = np.array(kat["current"], np.uint8)
    i_s = int(oracle.i2_plane(s, oracle.FMT_RGB8)[0])
    i_c = int(oracle.i2_plane(c, oracle.FMT_RGB8)[0])
    assert (i_s, i_c) == (kat["i2_start"], kat["i2_current"])
    assert abs(i_s - i_c) == kat["d"]
    assert abs(abs(i_s - i_c) / 510.0 - kat["d_norm"]) < 1e-6
    # the same pixel through the clip path: frame 0 = start, frame 1 = current
    clip = np.stack([s, c])
    r = oracle.run_clip(clip, oracle.FMT_RGB8, oracle.MODE_OVERALL, 0)
    assert int(r.sad[1]) == kat["d"] and int(r.acc_sum[0]) == kat["d"] and int(r.cnt[1]) == (1 if kat["d"] > 0 else 0)
    for name, filt in FILTERS.items():
        want = kat["pre_color"][name]
        got = oracle.visual_diff(i_s - i_c, filt, 5.0)
        if isinstance(want, str):
            assert math.isinf(got) and (got > 0) == (want == "inf")
        else:
            assert abs(got - want) < 2e-6
        grey = oracle.visual_pixel(i_s - i_c, False, filt, 5.0)
        assert grey[3] == 255 and grey[0] == grey[1] == grey[2]
        assert abs(grey[0] - kat["grey"][name]) <= kat["grey_tol"]
    if kat["color_none"] is not None:
        assert list(oracle.visual_pixel(i_s - i_c, True, 255, 5.0)[:3]) == kat["color_none"]


@pytest.mark.parametrize("kat", KATS["chroma"])
def test_kat_chroma(oracle, kat):
    px = np.array(kat["pixel"], np.uint8)
    for name, chroma in (("none", 0), ("red", 1), ("green", 2), ("blue", 3)):
        assert int(oracle.i2_plane(px, oracle.FMT_RGB8, chroma)[0]) == kat["i2"][name]
        assert int(oracle.i2_plane(px[::-1].copy(), oracle.FMT_BGR8, chroma)[0]) == kat["i2"][name]
        rgbx = np.array(list(kat["pixel"]) + [77], np.uint8)
        assert int(oracle.i2_plane(rgbx, oracle.FMT_RGBX8, chroma)[0]) == kat["i2"][name]


@pytest.mark.parametrize("kat", KATS["median4"])
def test_kat_median4(oracle, kat):
    # grey pixels v,v,v have I2 = 2v; use I2/2 as the grey level (all KAT values are even or we double)
    frames = np.array([[v // 2, v // 2, v // 2] for v in kat["i2_of_4_frames"]], np.uint8)
    got = int(oracle.median4_plane(frames, oracle.FMT_RGB8)[0])
    assert got == 2 * (kat["upper_median"] // 2)


@pytest.mark.parametrize("case", SYNTH["cases"], ids=lambda c: c["name"])
def test_synth_golden(oracle, case):
    clip = oracle.synth_clip(case["n_frames"], case["width"], case["height"], case["fmt"], profile=case["profile"])
    assert sha(clip) == case["clip_sha256"]
    r = oracle.run_clip(clip, case["fmt"], case["mode"], case["tau"], case["chroma"])
    assert sha(r.acc_sum) == case["acc_sum_sha256"] and sha(r.acc_cnt) == case["acc_cnt_sha256"]
    assert [int(v) for v in r.sad] == case["sad"] and [int(v) for v in r.cnt] == case["cnt"]
    assert sha(r.state) == case["state_sha256"]


@pytest.mark.parametrize("fmt", [0, 1, 2, 3])
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("chroma", [0, 1, 2, 3])
def test_oracle_vs_numpy_twin(oracle, twin, fmt, mode, chroma):
    clip = oracle.synth_clip(6, 29, 13, fmt, profile=oracle.SYNTH_SCENE, seed=7)
    a = oracle.run_clip(clip, fmt, mode, 20, chroma, nthreads=3)
    b = twin.run_clip(clip, fmt, mode, 20, chroma)
    for k in ("acc_sum", "acc_cnt", "sad", "cnt", "state"):
        assert np.array_equal(getattr(a, k), b[k]), k
    m4 = oracle.median4_plane(clip[:4], fmt, chroma)
    assert np.array_equal(m4, twin.median4(clip[:4], fmt, chroma))


def test_synth_matches_python_definition(oracle, twin):
    for fmt, profile in ((0, 0), (1, 1), (2, 1)):
        w, h, n = 7, 5, 3
        clip = oracle.synth_clip(n, w, h, fmt, profile=profile, seed=99)
        fb = w * h * oracle.bpp(fmt)
        for t in range(n):
            for i in range(0, fb, 3):
                assert clip[t, i] == twin.synth_byte(99, profile, t, i, w, h, fmt)
    # first_frame offset gives the same bytes as slicing
    full = oracle.synth_clip(5, 16, 8, 0)
    part = oracle.synth_clip(2, 16, 8, 0, first_frame=3)
    assert np.array_equal(full[3:], part)


def test_threads_do_not_change_results(oracle):
    clip = oracle.synth_clip(10, 64, 48, 0)
    base = oracle.run_clip(clip, 0, 1, 8, nthreads=1)
    for nt in (2, 5, 8):
        r = oracle.run_clip(clip, 0, 1, 8, nthreads=nt)
        assert np.array_equal(r.acc_sum, base.acc_sum) and np.array_equal(r.sad, base.sad) and np.array_equal(r.cnt, base.cnt)


def test_properties(oracle):
    fmt = oracle.FMT_RGB8
    clip = oracle.synth_clip(12, 40, 30, fmt, profile=oracle.SYNTH_SCENE)
    ov = oracle.run_clip(clip, fmt, oracle.MODE_OVERALL, 16)
    pf = oracle.run_clip(clip, fmt, oracle.MODE_PERFRAME, 16)
    # checksum of checksums
    for r in (ov, pf):
        assert int(r.acc_sum.astype(np.uint64).sum()) == int(r.sad.sum())
        assert int(r.acc_cnt.astype(np.uint64).sum()) == int(r.cnt.sum())
    assert int(ov.sad[0]) == 0 and int(pf.sad[0]) == 0
    # channel permutation invariance without chroma filter
    perm = clip.reshape(12, -1, 3)[:, :, [2, 0, 1]].reshape(12, -1).copy()
    ovp = oracle.run_clip(perm, fmt, oracle.MODE_OVERALL, 16)
    assert np.array_equal(ovp.acc_sum, ov.acc_sum) and np.array_equal(ovp.cnt, ov.cnt)
    # BGR with the same bytes == RGB when chroma is none
    ovb = oracle.run_clip(clip, oracle.FMT_BGR8, oracle.MODE_OVERALL, 16)
    assert np.array_equal(ovb.acc_sum, ov.acc_sum)
    # identical frames -> zero
    same = np.repeat(clip[:1], 5, axis=0)
    z = oracle.run_clip(same, fmt, oracle.MODE_OVERALL, 0)
    assert not z.acc_sum.any() and not z.sad.any() and not z.cnt.any()
    # telescoping: overall D_t <= sum_{k<=t} perframe D_k (per frame totals)
    assert np.all(ov.sad <= np.cumsum(pf.sad))
    # counts are monotone non-increasing in tau
    prev = None
    for tau in (0, 1, 8, 64, 255, 510, 600):
        c = oracle.run_clip(clip, fmt, oracle.MODE_OVERALL, tau).cnt
        if prev is not None:
            assert np.all(c <= prev)
        prev = c
    assert not prev.any()
    # chunked processing (state chaining) == one pass, both modes
    for mode, whole in ((0, ov), (1, pf)):
        a = oracle.run_clip(clip[:5], fmt, mode, 16)
        b = oracle.run_clip(clip[5:], fmt, mode, 16, state=a.state, acc_sum=a.acc_sum, acc_cnt=a.acc_cnt)
        assert np.array_equal(b.acc_sum, whole.acc_sum) and np.array_equal(b.acc_cnt, whole.acc_cnt)
        assert np.array_equal(np.concatenate([a.sad, b.sad]), whole.sad)
    # float outputs
    im = oracle.intensity_map(ov.acc_sum, 12)
    assert np.allclose(im, ov.acc_sum / (510.0 * 12), rtol=1e-6)
    fm = oracle.frame_means(ov.sad, 40 * 30)
    assert np.allclose(fm, ov.sad / (510.0 * 1200), rtol=1e-6)


def test_empty_and_single(oracle):
    fmt = oracle.FMT_RGBX8
    one = oracle.synth_clip(1, 8, 8, fmt)
    r = oracle.run_clip(one, fmt, oracle.MODE_PERFRAME, 0)
    assert int(r.sad[0]) == 0 and not r.acc_sum.any()
    assert np.array_equal(r.state, oracle.i2_plane(one[0], fmt))


def test_visual_frame_matches_twin(oracle, twin):
    rng = np.random.default_rng(5)
    ref = rng.integers(0, 511, 500).astype(np.uint16)
    cur = rng.integers(0, 511, 500).astype(np.uint16)
    for filt in (255, 0, 1):
        out = oracle.visual_frame(ref, cur, False, filt, 5.0).reshape(-1, 4)
        for p in range(0, 500, 37):
            d = float(twin.visual_diff(int(ref[p]) - int(cur[p]), filt, 5.0))
            want = 0.5 - d
            want = 0 if not (want > 0) else (255 if want >= 1 else math.floor(want * 255 + 0.5))
            assert abs(int(out[p, 0]) - want) <= 1 and out[p, 3] == 255


def test_compute_state_flavour(oracle):
    """dips ComputeState: first 3 frames pass through, then start - median-of-4 (dips/src/gpu/mod.rs:170-216, :394-396)."""
    w, h = 16, 8
    clip = oracle.synth_clip(8, w, h, oracle.FMT_RGBX8, profile=oracle.SYNTH_SCENE)
    cs = oracle.ComputeStateOracle(w, h)
    for t in range(3):
        out, passthrough = cs.frame(clip[t])
        assert passthrough and np.array_equal(out, clip[t])
    out, passthrough = cs.frame(clip[3])
    assert not passthrough
    # static clip: every later frame equal to the first four -> diff 0 -> mid grey
    cs2 = oracle.ComputeStateOracle(w, h)
    outs = [cs2.frame(clip[0])[0] for _ in range(6)]
    grey = outs[-1].reshape(-1, 4)
    assert np.all(np.abs(grey[:, :3].astype(int) - 128) <= 1) and np.all(grey[:, 3] == 255)


def test_dips_alt_flavour(oracle):
    """dips_alt DiPsCompute: zero snapshot until the first snapshot, grey snapshot frame, min-of-two quirk vs intended."""
    w, h = 12, 6
    clip = oracle.synth_clip(5, w, h, oracle.FMT_RGBX8, profile=oracle.SYNTH_SCENE)
    for intended in (False, True):
        a = oracle.DiPsComputeOracle(w, h, colorize=False, filt=oracle.FILTER_NONE, intended_median=intended)
        i2 = [oracle.i2_plane(clip[t], oracle.FMT_RGBX8).astype(int) for t in range(5)]
        out0 = a.send_frame(clip[0]).reshape(-1, 4)
        # ring = [frame0, zeros]; as shipped the median is min(I0, 0) = 0 -> diff 0 -> mid grey; intended: max = I0
        if not intended:
            assert np.all(np.abs(out0[:, 0].astype(int) - 128) <= 1)
        else:
            want = np.clip(np.floor((0.5 + 2.5 * i2[0] / 510.0) * 255 + 0.5), 0, 255)
            assert np.all(np.abs(out0[:, 0].astype(int) - want) <= 1)
        a.send_frame(clip[1])
        snap = a.send_frame(clip[2], snapshot=True).reshape(-1, 4)
        med = np.maximum(i2[1], i2[2]) if intended else np.minimum(i2[1], i2[2])
        assert np.all(np.abs(snap[:, 0].astype(int) - (med + 1) // 2) <= 1) and np.all(snap[:, 3] == 255)
        assert np.all(snap[:, 0] == snap[:, 1]) and np.all(snap[:, 1] == snap[:, 2])


@pytest.mark.parametrize("window", [1, 3, 5, 7])
def test_spatial_median_is_the_true_zero_padded_median(oracle, window):
    """N4: against numpy (zero padding, full symmetric window, middle element) on ragged sizes."""
    rng = np.random.default_rng(window)
    for w, h in ((1, 1), (2, 5), (9, 4), (31, 17)):
        plane = rng.integers(0, 511, w * h).astype(np.uint16)
        got = oracle.spatial_median(plane, w, h, window).reshape(h, w)
        r = window // 2
        pad = np.pad(plane.reshape(h, w), r, constant_values=0)
        win = np.lib.stride_tricks.sliding_window_view(pad, (window, window)).reshape(h, w, -1)
        want = np.sort(win, axis=-1)[..., (window * window) // 2]
        assert np.array_equal(got, want)
    # a constant plane keeps its value in the interior and drops to 0 in corners where zeros are the majority
    flat = np.full(11 * 11, 300, np.uint16)
    out = oracle.spatial_median(flat, 11, 11, window).reshape(11, 11)
    assert out[5, 5] == 300
    if window > 1:
        assert out[0, 0] == 0
