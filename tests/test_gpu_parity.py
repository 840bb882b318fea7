"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every test drives the CUDA path through the C ABI
(dips_b200.Context -> libdips_b200.so) and checks it bit-exactly against the CPU oracle / the committed fixtures."""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SYNTH = json.load(open(os.path.join(GOLD, "synth_golden.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    torch.cuda.init()
    return torch


def to_device(torch, clip: np.ndarray):
    return torch.from_numpy(np.ascontiguousarray(clip)).cuda()


def run_gpu(torch, clip, w, h, fmt, mode, tau, chroma=0, chunks=None, tuning=None, stride_pad=0, offset=0):
    """Run `clip` ([n, frame_bytes] uint8) through the library; returns (acc_sum, acc_cnt, sad, cnt, state, plan)."""
    import dips_b200
    n, fb = clip.shape
    stride = fb + stride_pad
    if stride_pad or offset:
        host = np.zeros(offset + n * stride + 64, np.uint8)
        for t in range(n):
            host[offset + t * stride: offset + t * stride + fb] = clip[t]
        dev = torch.from_numpy(host).cuda()
        base = dev.data_ptr() + offset
    else:
        dev = to_device(torch, clip)
        base = dev.data_ptr()
    with dips_b200.Context(w, h, fmt, mode, tau, chroma) as ctx:
        if tuning:
            tuning = dict(tuning)
            kernel = tuning.pop("kernel", None)
            if kernel is not None:
                ctx.set_kernel(kernel)
            if tuning:
                ctx.set_tuning(**tuning)
        bounds = [0, n] if not chunks else chunks
        for a, b in zip(bounds, bounds[1:]):
            ctx.run_clip_device(base + a * stride, b - a, stride, a)
        ctx.synchronize()
        acc_sum, acc_cnt = ctx.get_accumulators()
        sad, cnt = ctx.get_scalars(0, n)
        state = ctx.get_state_plane()
        plan = ctx.last_plan()
        assert ctx.frames_processed == n
    return acc_sum, acc_cnt, sad, cnt, state, plan


def check(oracle, got, clip, fmt, mode, tau, chroma=0):
    want = oracle.run_clip(clip, fmt, mode, tau, chroma)
    acc_sum, acc_cnt, sad, cnt, state, _ = got
    assert np.array_equal(sad, want.sad), "per-frame sad"
    assert np.array_equal(cnt, want.cnt), "per-frame count"
    assert np.array_equal(acc_sum, want.acc_sum), "acc_sum"
    assert np.array_equal(acc_cnt, want.acc_cnt), "acc_cnt"
    assert np.array_equal(state, want.state), "state plane"


def test_synth_generator_matches_oracle(torch_cuda, oracle):
    import dips_b200
    torch = torch_cuda
    for fmt, profile, w, h, n, first in ((0, 1, 96, 64, 5, 0), (1, 0, 33, 17, 4, 3), (2, 1, 61, 37, 3, 1000), (3, 1, 80, 60, 2, 7)):
        fb = w * h * dips_b200.bytes_per_pixel(fmt)
        dev = torch.empty(n * fb, dtype=torch.uint8, device="cuda")
        dips_b200.synth_fill_device(0, dev.data_ptr(), first, n, w, h, fmt, 0x44695073, profile,
                                    torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        want = oracle.synth_clip(n, w, h, fmt, profile=profile, first_frame=first)
        assert np.array_equal(dev.cpu().numpy().reshape(n, fb), want)


@pytest.mark.parametrize("case", SYNTH["cases"], ids=lambda c: c["name"])
def test_committed_golden_fixtures(torch_cuda, oracle, case):
    clip = oracle.synth_clip(case["n_frames"], case["width"], case["height"], case["fmt"], profile=case["profile"])
    assert sha(clip) == case["clip_sha256"]
    acc_sum, acc_cnt, sad, cnt, state, plan = run_gpu(torch_cuda, clip, case["width"], case["height"], case["fmt"],
                                                      case["mode"], case["tau"], case["chroma"])
    assert plan["tma_path"] == (clip.shape[1] % 16 == 0 or case["n_frames"] >= 2)   # unaligned clips are re-packed
    assert [int(v) for v in sad] == case["sad"] and [int(v) for v in cnt] == case["cnt"]
    assert sha(acc_sum) == case["acc_sum_sha256"] and sha(acc_cnt) == case["acc_cnt_sha256"]
    assert sha(state) == case["state_sha256"]


@pytest.mark.parametrize("fmt", [0, 1, 2, 3])
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("chroma", [0, 1, 2, 3])
def test_all_variants_small(torch_cuda, oracle, fmt, mode, chroma):
    w, h, n = 200, 75, 9     # 15000 px: several tiles with a ragged last one
    clip = oracle.synth_clip(n, w, h, fmt, profile=oracle.SYNTH_SCENE, seed=11 + fmt)
    check(oracle, run_gpu(torch_cuda, clip, w, h, fmt, mode, 24, chroma), clip, fmt, mode, 24, chroma)


@pytest.mark.parametrize("tau", [0, 1, 32, 128, 509, 510, 511, 70000])
def test_thresholds(torch_cuda, oracle, tau):
    w, h, n = 128, 64, 6
    clip = oracle.synth_clip(n, w, h, 0, profile=oracle.SYNTH_UNIFORM)
    check(oracle, run_gpu(torch_cuda, clip, w, h, 0, 0, tau), clip, 0, 0, tau)


@pytest.mark.parametrize("w,h", [(1, 1), (5, 3), (16, 1), (17, 3), (511, 2), (513, 7), (640, 480), (1000, 9)])
@pytest.mark.parametrize("fmt", [0, 1])
def test_ragged_sizes(torch_cuda, oracle, w, h, fmt):
    n = 5
    clip = oracle.synth_clip(n, w, h, fmt, profile=oracle.SYNTH_UNIFORM, seed=w * 131 + h)
    for mode in (0, 1):
        got = run_gpu(torch_cuda, clip, w, h, fmt, mode, 100)
        check(oracle, got, clip, fmt, mode, 100)


def test_extremes_saturate_correctly(torch_cuda, oracle):
    """all-black reference vs all-white frames: D = 510 everywhere; 300 frames cross the 128-frame flush twice."""
    w, h, n = 64, 32, 300
    clip = np.full((n, w * h * 3), 255, np.uint8)
    clip[0] = 0
    got = run_gpu(torch_cuda, clip, w, h, 0, 0, 509)
    check(oracle, got, clip, 0, 0, 509)
    assert int(got[0][0]) == 510 * (n - 1) and int(got[1][0]) == n - 1
    alt = np.zeros((n, w * h * 3), np.uint8)
    alt[1::2] = 255
    got = run_gpu(torch_cuda, alt, w, h, 0, 1, 0)
    check(oracle, got, alt, 0, 1, 0)


def test_single_frame_and_empty_call(torch_cuda, oracle):
    import dips_b200
    w, h = 40, 20
    clip = oracle.synth_clip(1, w, h, 1)
    check(oracle, run_gpu(torch_cuda, clip, w, h, 1, 1, 0), clip, 1, 1, 0)
    with dips_b200.Context(w, h, 1) as ctx:
        ctx.run_clip_device(0, 0)           # empty clip: no-op, no error
        assert ctx.frames_processed == 0
        s, c = ctx.get_accumulators()
        assert not s.any() and not c.any()
        with pytest.raises(dips_b200.DipsError):
            ctx.get_scalars(0, 1)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("fmt", [0, 1])
def test_chunked_calls_chain(torch_cuda, oracle, mode, fmt):
    w, h, n = 320, 100, 40
    clip = oracle.synth_clip(n, w, h, fmt, profile=oracle.SYNTH_SCENE)
    got = run_gpu(torch_cuda, clip, w, h, fmt, mode, 16, chunks=[0, 1, 2, 17, 33, 40])
    check(oracle, got, clip, fmt, mode, 16)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("segments", [2, 3, 7])
def test_forced_frame_segments(torch_cuda, oracle, mode, segments):
    w, h, n = 256, 64, 45
    clip = oracle.synth_clip(n, w, h, 0, profile=oracle.SYNTH_SCENE)
    got = run_gpu(torch_cuda, clip, w, h, 0, mode, 8, tuning=dict(segments=segments))
    assert got[5]["segments"] == segments
    check(oracle, got, clip, 0, mode, 8)


@pytest.mark.parametrize("tile_px,stages,regs", [(512, 2, 0), (1040, 3, 72), (2048, 8, 64), (7008, 4, 80), (3504, 3, 96),
                                                  (14336, 2, 72), (16384, 2, 64), (16, 3, 0), (1040, 3, 128), (7008, 4, 128),
                                                  (16384, 3, 128)])
def test_forced_tile_geometry(torch_cuda, oracle, tile_px, stages, regs):
    w, h, n = 300, 77, 12
    for fmt in (0, 1):
        clip = oracle.synth_clip(n, w, h, fmt, profile=oracle.SYNTH_SCENE)
        got = run_gpu(torch_cuda, clip, w, h, fmt, 1, 8, tuning=dict(stages=stages, tile_px=tile_px, regs=regs))
        assert got[5]["tile_px"] == tile_px and got[5]["stages"] == stages and (not regs or got[5]["regs"] == regs)
        check(oracle, got, clip, fmt, 1, 8)


@pytest.mark.parametrize("regs", [64, 72, 80, 96, 128])
@pytest.mark.parametrize("fmt,mode", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_register_variants(torch_cuda, oracle, regs, fmt, mode):
    w, h, n = 640, 360, 140          # > 128 frames: crosses the packed-accumulator flush
    clip = oracle.synth_clip(n, w, h, fmt, profile=oracle.SYNTH_SCENE)
    got = run_gpu(torch_cuda, clip, w, h, fmt, mode, 20, tuning=dict(regs=regs))
    assert got[5]["regs"] == regs
    check(oracle, got, clip, fmt, mode, 20)


def test_padded_stride_and_unaligned_clips(torch_cuda, oracle):
    """16-byte aligned base and pitch go straight to the clip kernel; anything else (odd pitch, odd base, frame size not a
    multiple of 16 bytes) is re-packed on the device into an aligned, zero-padded scratch and streamed from there; a single
    unaligned frame takes the per-frame kernel."""
    w, h, n = 112, 30, 7       # frame bytes are a multiple of 16 for both formats
    for fmt in (0, 1):
        clip = oracle.synth_clip(n, w, h, fmt, profile=oracle.SYNTH_SCENE)
        pad = 48 + (-clip.shape[1]) % 16
        for mode in (0, 1):
            got = run_gpu(torch_cuda, clip, w, h, fmt, mode, 8, stride_pad=pad)    # aligned pitch: direct
            assert got[5]["tma_path"]
            check(oracle, got, clip, fmt, mode, 8)
            got = run_gpu(torch_cuda, clip, w, h, fmt, mode, 8, stride_pad=5)      # unaligned pitch: re-packed
            assert got[5]["tma_path"]
            check(oracle, got, clip, fmt, mode, 8)
            got = run_gpu(torch_cuda, clip, w, h, fmt, mode, 8, offset=3)          # unaligned base: re-packed
            assert got[5]["tma_path"]
            check(oracle, got, clip, fmt, mode, 8)
        got = run_gpu(torch_cuda, clip[:1], w, h, fmt, 0, 8, offset=3)             # one unaligned frame: per-frame kernel
        assert not got[5]["tma_path"]
        check(oracle, got, clip[:1], fmt, 0, 8)
    # frame size not a multiple of 16 bytes: the last bulk copy of a frame is rounded up into the zero padding
    for (w, h, fmt) in ((37, 5, 0), (37, 5, 1), (1001, 33, 0), (333, 3, 1), (2047, 129, 0)):
        clip = oracle.synth_clip(n, w, h, fmt, profile=oracle.SYNTH_UNIFORM)
        assert clip.shape[1] % 16
        for mode in (0, 1):
            for kw in (dict(), dict(stride_pad=(-clip.shape[1]) % 16), dict(offset=1), dict(chunks=[0, 1, 2, 5, 7])):
                got = run_gpu(torch_cuda, clip, w, h, fmt, mode, 8, **kw)
                check(oracle, got, clip, fmt, mode, 8)
            for tuning in (dict(kernel=0), dict(kernel=0, regs=128), dict(segments=2)):
                got = run_gpu(torch_cuda, clip, w, h, fmt, mode, 8, tuning=tuning)
                assert got[5]["tma_path"]
                check(oracle, got, clip, fmt, mode, 8)


def test_config1_full_640x480x300(torch_cuda, oracle):
    """BASELINE.json configs[0]: 640x480 RGB8, 300 frames, overall difference vs first frame."""
    w, h, n = 640, 480, 300
    clip = oracle.synth_clip(n, w, h, 0, profile=oracle.SYNTH_SCENE)
    for mode in (0, 1):
        got = run_gpu(torch_cuda, clip, w, h, 0, mode, 32)
        check(oracle, got, clip, 0, mode, 32)


def test_reset_prime_and_accumulator_roundtrip(torch_cuda, oracle):
    import dips_b200
    torch = torch_cuda
    w, h, n = 160, 90, 10
    clip = oracle.synth_clip(n, w, h, 0, profile=oracle.SYNTH_SCENE)
    dev = to_device(torch, clip)
    with dips_b200.Context(w, h, 0, 0, 16) as ctx:
        # explicit reference = frame 3 (not the first frame of the call)
        ctx.prime_device(dev[3].data_ptr())
        ctx.run_clip_device(dev.data_ptr(), n)
        want = oracle.run_clip(clip, 0, 0, 16, state=oracle.i2_plane(clip[3], 0))
        s, c = ctx.get_accumulators()
        assert np.array_equal(s, want.acc_sum) and np.array_equal(c, want.acc_cnt)
        # float outputs (tolerance 1e-5 relative, north star)
        im = ctx.get_intensity_map(n)
        assert np.allclose(im, oracle.intensity_map(want.acc_sum, n), rtol=1e-5, atol=0)
        fm = ctx.get_frame_means(0, n)
        assert np.allclose(fm, oracle.frame_means(want.sad, w * h), rtol=1e-5, atol=0)
        # median-of-4 start plane (reference pre_compute_main)
        ctx.reset()
        ctx.prime_median4_device(dev.data_ptr(), clip.shape[1])
        assert np.array_equal(ctx.get_state_plane(), oracle.median4_plane(clip[:4], 0))
        # set/get accumulators round trip (checkpoint/resume) then continue accumulating
        ctx.reset()
        ctx.set_accumulators(want.acc_sum, want.acc_cnt)
        s, c = ctx.get_accumulators()
        assert np.array_equal(s, want.acc_sum) and np.array_equal(c, want.acc_cnt)
        ctx.prime_device(dev[3].data_ptr())
        ctx.run_clip_device(dev.data_ptr(), n)
        s, c = ctx.get_accumulators()
        assert np.array_equal(s, 2 * want.acc_sum) and np.array_equal(c, 2 * want.acc_cnt)
        # reset really clears
        ctx.reset()
        s, c = ctx.get_accumulators()
        assert not s.any() and not c.any() and ctx.frames_processed == 0


def test_run_clip_host_pageable_and_pinned(torch_cuda, oracle):
    import dips_b200
    torch = torch_cuda
    w, h, n = 320, 180, 25
    for fmt, mode in ((0, 0), (1, 1), (0, 1)):
        clip = oracle.synth_clip(n, w, h, fmt, profile=oracle.SYNTH_SCENE)
        want = oracle.run_clip(clip, fmt, mode, 20)
        for pinned in (False, True):
            with dips_b200.Context(w, h, fmt, mode, 20) as ctx:
                if pinned:
                    t = torch.from_numpy(clip).pin_memory()
                    ctx.run_clip_host(t.data_ptr(), n, clip.shape[1])
                else:
                    ctx.run_clip_host(clip)
                ctx.synchronize()
                s, c = ctx.get_accumulators()
                sad, cnt = ctx.get_scalars(0, n)
            assert np.array_equal(s, want.acc_sum) and np.array_equal(c, want.acc_cnt)
            assert np.array_equal(sad, want.sad) and np.array_equal(cnt, want.cnt)


@pytest.mark.parametrize("mode", [0, 1])
def test_logical_shards_merge_exactly(torch_cuda, oracle, mode):
    """R logical shards run one after another on one GPU through the same sharding code; the all-reduce is replaced by
    an in-process integer sum.  Result must equal the single-shard run for every R."""
    import dips_b200
    from dips_b200 import sharding
    torch = torch_cuda
    w, h, n, fmt, tau = 256, 120, 37, 0, 12
    clip = oracle.synth_clip(n, w, h, fmt, profile=oracle.SYNTH_SCENE)
    want = oracle.run_clip(clip, fmt, mode, tau)
    dev = to_device(torch, clip)
    for R in (1, 2, 4, 8):
        total = None
        sad = np.zeros(n, np.uint64)
        cnt = np.zeros(n, np.uint64)
        for r in range(R):
            t0, t1 = sharding.shard_range(r, R, n)
            with dips_b200.Context(w, h, fmt, mode, tau) as ctx:
                eng = sharding.GpuShardEngine(ctx, dev[t0:t1], torch)
                if r > 0 or mode == 0:
                    eng.prime(dev[0] if mode == 0 else dev[t0 - 1])      # what broadcast / halo exchange deliver
                sharding.run_sharded(eng, mode, t0)
                ctx.synchronize()
                acc = eng.acc_tensor().clone()
                total = acc if total is None else total + acc
                s, c = ctx.get_scalars(t0, t1 - t0)
                sad[t0:t1], cnt[t0:t1] = s, c
        with dips_b200.Context(w, h, fmt, mode, tau) as ctx:      # un-permute the merged accumulators
            ptr, ne = ctx.accumulators_device()
            view = torch.as_tensor(sharding._DeviceBuffer(ptr, 2 * ne, "<i4"), device="cuda")
            view.copy_(total)
            torch.cuda.synchronize()
            s, c = ctx.get_accumulators()
        assert np.array_equal(s, want.acc_sum) and np.array_equal(c, want.acc_cnt), f"R={R}"
        assert np.array_equal(sad, want.sad) and np.array_equal(cnt, want.cnt), f"R={R}"


def test_full_size_1080p_properties(torch_cuda, oracle):
    """BASELINE configs[1]/[2] geometry (1920x1080 RGB8) on a device-generated clip: size-independent properties on
    all frames, bit-exact oracle comparison on a bounded prefix."""
    import dips_b200
    torch = torch_cuda
    w, h, n, fmt, tau = 1920, 1080, 200, 0, 32
    fb = w * h * 3
    dev = torch.empty(n * fb, dtype=torch.uint8, device="cuda")
    dips_b200.synth_fill_device(0, dev.data_ptr(), 0, n, w, h, fmt, 0x44695073, dips_b200.SYNTH_SCENE,
                                torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    prefix = 12
    host = dev[: prefix * fb].cpu().numpy().reshape(prefix, fb)
    res = {}
    for mode in (0, 1):
        with dips_b200.Context(w, h, fmt, mode, tau) as ctx:
            ctx.run_clip_device(dev.data_ptr(), n)
            ctx.synchronize()
            s, c = ctx.get_accumulators()
            sad, cnt = ctx.get_scalars(0, n)
            res[mode] = (s, c, sad, cnt)
            # checksum of checksums
            assert int(s.astype(np.uint64).sum()) == int(sad.sum())
            assert int(c.astype(np.uint64).sum()) == int(cnt.sum())
            assert int(sad[0]) == 0 and int(cnt[0]) == 0
            # prefix against the oracle (scalars are per frame, so they compare directly)
            want = oracle.run_clip(host, fmt, mode, tau)
            assert np.array_equal(sad[:prefix], want.sad) and np.array_equal(cnt[:prefix], want.cnt)
            # idempotence: a second identical pass doubles the accumulators and repeats the scalars
            ctx.mark_state_valid(False)
            ctx.run_clip_device(dev.data_ptr(), n)
            s2, c2 = ctx.get_accumulators()
            sad2, _ = ctx.get_scalars(0, n)
            assert np.array_equal(s2, 2 * s) and np.array_equal(c2, 2 * c) and np.array_equal(sad2, sad)
    # telescoping bound between the two modes
    assert np.all(res[0][2] <= np.cumsum(res[1][2]))


@pytest.mark.parametrize("w,h,fmt,n", [(3840, 2160, 1, 20), (7680, 4320, 0, 6), (2560, 1440, 2, 12), (1280, 720, 3, 140)])
@pytest.mark.parametrize("mode", [0, 1])
def test_baseline_geometries_bit_exact(torch_cuda, oracle, w, h, fmt, n, mode):
    """The tile plans of the BASELINE resolutions (4K RGBx: 896 threads x 4 waves; 8K RGB8: 1024 threads / 64 registers x
    14 waves; 1440p: 2 waves; 720p: frame segments + a flush at 128 frames) on device-generated clips, every output
    compared bit for bit with the oracle."""
    import dips_b200
    torch = torch_cuda
    fb = w * h * dips_b200.bytes_per_pixel(fmt)
    dev = torch.empty(n * fb, dtype=torch.uint8, device="cuda")
    dips_b200.synth_fill_device(0, dev.data_ptr(), 0, n, w, h, fmt, 0x44695073, dips_b200.SYNTH_SCENE,
                                torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    host = dev.cpu().numpy().reshape(n, fb)
    want = oracle.run_clip(host, fmt, mode, 32)
    with dips_b200.Context(w, h, fmt, mode, 32) as ctx:
        ctx.run_clip_device(dev.data_ptr(), n)
        ctx.synchronize()
        s, c = ctx.get_accumulators()
        sad, cnt = ctx.get_scalars(0, n)
        state = ctx.get_state_plane()
        assert ctx.last_plan()["tma_path"]
    assert np.array_equal(sad, want.sad) and np.array_equal(cnt, want.cnt)
    assert np.array_equal(s, want.acc_sum) and np.array_equal(c, want.acc_cnt)
    assert np.array_equal(state, want.state)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("n,segments", [(127, 1), (128, 1), (129, 1), (257, 1), (300, 2), (385, 3)])
def test_flush_boundaries(torch_cuda, oracle, mode, n, segments):
    """packed 16-bit accumulators are flushed every 128 accumulated frames: counts around the boundary, with the halo frame
    of per-frame segments shifting the phase, and saturated differences (D = 510) so that an off-by-one overflows."""
    w, h = 96, 64
    clip = np.zeros((n, w * h * 3), np.uint8)
    clip[1::2] = 255                       # overall: D alternates 510/0; per-frame: D = 510 every frame
    clip[0] = 0
    got = run_gpu(torch_cuda, clip, w, h, 0, mode, 509, tuning=dict(segments=segments))
    check(oracle, got, clip, 0, mode, 509)
    full = np.full((n, w * h * 3), 255, np.uint8)
    full[0] = 0                            # overall: D = 510 on every frame after the first
    got = run_gpu(torch_cuda, full, w, h, 0, 0, 0, tuning=dict(segments=segments))
    check(oracle, got, full, 0, 0, 0)


# ---- the warp-specialised clip kernel (clip_kernel_ws): same results as clip_kernel on every path ----------------------
@pytest.mark.parametrize("fmt", [0, 1, 2, 3])
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("chroma", [0, 1, 3])
def test_ws_kernel_all_variants(torch_cuda, oracle, fmt, mode, chroma):
    w, h, n = 208, 75, 11      # frame bytes are a multiple of 16 for 3 and 4 B/px (TMA path)
    clip = oracle.synth_clip(n, w, h, fmt, profile=oracle.SYNTH_SCENE, seed=23 + fmt)
    if fmt in (1, 3) and mode == 1 and chroma:
        # the one combination clip_kernel_ws is not built for (it would spill at 64 registers): asking for it is an error,
        # and the automatic choice is clip_kernel
        import dips_b200
        with dips_b200.Context(w, h, fmt, mode, 24, chroma) as ctx:
            with pytest.raises(dips_b200.DipsError):
                ctx.set_kernel(1)
        got = run_gpu(torch_cuda, clip, w, h, fmt, mode, 24, chroma)
        assert got[5]["kernel"] == 0 and got[5]["tma_path"]
    else:
        got = run_gpu(torch_cuda, clip, w, h, fmt, mode, 24, chroma, tuning=dict(kernel=1))
        assert got[5]["kernel"] == 1 and got[5]["tma_path"]
    check(oracle, got, clip, fmt, mode, 24, chroma)


@pytest.mark.parametrize("stages", [3, 4])
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("n,segments,w,h", [(1, 1, 64, 32), (2, 1, 64, 32), (5, 1, 512, 7), (13, 3, 96, 64), (129, 1, 96, 64),
                                            (300, 2, 96, 64), (385, 3, 48, 32), (140, 1, 640, 360)])
def test_ws_kernel_frame_counts_segments_and_flush(torch_cuda, oracle, stages, mode, n, segments, w, h):
    """lead-in (halo frame of a segment), unrolled trips and the run-time tail of clip_kernel_ws for both stage counts;
    saturated differences across the 128-frame flush"""
    clip = np.zeros((n, w * h * 3), np.uint8)
    clip[1::2] = 255
    clip[0] = 0
    got = run_gpu(torch_cuda, clip, w, h, 0, mode, 509, tuning=dict(kernel=1, stages=stages, segments=segments))
    assert got[5]["kernel"] == 1 and got[5]["stages"] == stages
    check(oracle, got, clip, 0, mode, 509)
    scene = oracle.synth_clip(n, w, h, 1, profile=oracle.SYNTH_SCENE)
    got = run_gpu(torch_cuda, scene, w, h, 1, mode, 20, tuning=dict(kernel=1, stages=stages, segments=segments))
    check(oracle, got, scene, 1, mode, 20)


@pytest.mark.parametrize("w,h,fmt,n", [(1920, 1080, 0, 24), (3840, 2160, 1, 10), (7680, 4320, 0, 5)])
@pytest.mark.parametrize("mode", [0, 1])
def test_ws_kernel_baseline_geometries(torch_cuda, oracle, w, h, fmt, n, mode):
    import dips_b200
    torch = torch_cuda
    fb = w * h * dips_b200.bytes_per_pixel(fmt)
    dev = torch.empty(n * fb, dtype=torch.uint8, device="cuda")
    dips_b200.synth_fill_device(0, dev.data_ptr(), 0, n, w, h, fmt, 0x44695073, dips_b200.SYNTH_SCENE,
                                torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    host = dev.cpu().numpy().reshape(n, fb)
    want = oracle.run_clip(host, fmt, mode, 32)
    with dips_b200.Context(w, h, fmt, mode, 32) as ctx:
        ctx.set_kernel(1)
        ctx.run_clip_device(dev.data_ptr(), n)
        ctx.synchronize()
        s, c = ctx.get_accumulators()
        sad, cnt = ctx.get_scalars(0, n)
        state = ctx.get_state_plane()
        assert ctx.last_plan()["kernel"] == 1
    assert np.array_equal(sad, want.sad) and np.array_equal(cnt, want.cnt)
    assert np.array_equal(s, want.acc_sum) and np.array_equal(c, want.acc_cnt) and np.array_equal(state, want.state)


@pytest.mark.parametrize("total_frames,expect_words", [(40, 1.0), (3000, 1.5), (70000, 2.0)])
def test_packed_accumulator_exchange(torch_cuda, oracle, total_frames, expect_words):
    """dipsb_pack_accumulators_device: summing the packed buffers of two shards as int32 and unpacking equals the
    combined accumulators, for the 4-byte, 6-byte and unpacked exchange formats."""
    import dips_b200
    from dips_b200 import sharding
    torch = torch_cuda
    w, h, n, fmt, tau = 256, 120, 40, 0, 12
    clip = oracle.synth_clip(n, w, h, fmt, profile=oracle.SYNTH_SCENE)
    want = oracle.run_clip(clip, fmt, 0, tau)
    dev = to_device(torch, clip)
    ctxs = [dips_b200.Context(w, h, fmt, 0, tau) for _ in range(2)]
    try:
        packed = []
        for r, ctx in enumerate(ctxs):
            t0, t1 = sharding.shard_range(r, 2, n)
            ctx.prime_device(dev[0].data_ptr())
            ctx.run_clip_device(dev[t0].data_ptr(), t1 - t0, clip.shape[1], t0)
            ptr, nw = ctx.pack_accumulators_device(total_frames)
            _, ne = ctx.accumulators_device()
            assert nw == int(expect_words * ne)
            ctx.synchronize()
            packed.append(torch.as_tensor(sharding._DeviceBuffer(ptr, nw, "<i4"), device="cuda"))
        total = packed[0] + packed[1]                      # what the all-reduce computes
        for ctx, buf in zip(ctxs, packed):
            buf.copy_(total)
            torch.cuda.synchronize()
            ctx.unpack_accumulators_device()
            s, c = ctx.get_accumulators()
            assert np.array_equal(s, want.acc_sum) and np.array_equal(c, want.acc_cnt)
    finally:
        for ctx in ctxs:
            ctx.close()


def test_threshold_change_between_chunks_and_thread_migration(torch_cuda, oracle):
    """tau may change between calls (accumulators are kept), and a context may be driven from another thread than the one
    that created it (the reference moves its ComputeState to a GStreamer streaming thread, frame_extractor.rs:76, :232)."""
    import threading
    import dips_b200
    torch = torch_cuda
    w, h, n, fmt = 192, 100, 30, 0
    clip = oracle.synth_clip(n, w, h, fmt, profile=oracle.SYNTH_SCENE)
    a = oracle.run_clip(clip[:12], fmt, 1, 5)
    b = oracle.run_clip(clip[12:], fmt, 1, 40, state=a.state, acc_sum=a.acc_sum, acc_cnt=a.acc_cnt)
    dev = to_device(torch, clip)
    ctx = dips_b200.Context(w, h, fmt, 1, 5)
    errors = []

    def second_half():
        try:
            ctx.set_threshold(40)
            ctx.run_clip_device(dev[12].data_ptr(), n - 12, clip.shape[1], 12)
            ctx.synchronize()
        except Exception as e:      # noqa: BLE001
            errors.append(e)

    try:
        ctx.run_clip_device(dev.data_ptr(), 12, clip.shape[1], 0)
        t = threading.Thread(target=second_half)
        t.start()
        t.join()
        assert not errors, errors
        s, c = ctx.get_accumulators()
        sad, cnt = ctx.get_scalars(0, n)
    finally:
        ctx.close()
    assert np.array_equal(s, b.acc_sum) and np.array_equal(c, b.acc_cnt)
    assert np.array_equal(sad, np.concatenate([a.sad, b.sad])) and np.array_equal(cnt, np.concatenate([a.cnt, b.cnt]))


def test_two_contexts_on_two_streams(torch_cuda, oracle):
    """independent contexts on their own streams do not interfere"""
    import dips_b200
    torch = torch_cuda
    w, h, n = 320, 180, 50
    clips = [oracle.synth_clip(n, w, h, f, profile=oracle.SYNTH_SCENE, seed=5 + f) for f in (0, 1)]
    devs = [to_device(torch, c) for c in clips]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    ctxs = [dips_b200.Context(w, h, f, m, 16) for f, m in ((0, 0), (1, 1))]
    try:
        for ctx, st in zip(ctxs, streams):
            ctx.set_stream(st.cuda_stream)
        for rep in range(3):
            for ctx, dev in zip(ctxs, devs):
                ctx.reset()
                ctx.run_clip_device(dev.data_ptr(), n)
        for ctx, clip, (f, m) in zip(ctxs, clips, ((0, 0), (1, 1))):
            ctx.synchronize()
            want = oracle.run_clip(clip, f, m, 16)
            s, c = ctx.get_accumulators()
            sad, cnt = ctx.get_scalars(0, n)
            assert np.array_equal(s, want.acc_sum) and np.array_equal(c, want.acc_cnt)
            assert np.array_equal(sad, want.sad) and np.array_equal(cnt, want.cnt)
    finally:
        for ctx in ctxs:
            ctx.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_run_clip_host_multi_chunk(torch_cuda, oracle, mode):
    """dipsb_run_clip_host splits long clips into <=256 MB chunks that are uploaded on a copy stream while the previous
    chunk is processed: 100 frames of 1080p RGB8 = 622 MB = 3 chunks; results must equal the one-shot oracle run."""
    import dips_b200
    torch = torch_cuda
    w, h, n, fmt = 1920, 1080, 100, 0
    fb = w * h * 3
    dev = torch.empty(n * fb, dtype=torch.uint8, device="cuda")
    dips_b200.synth_fill_device(0, dev.data_ptr(), 0, n, w, h, fmt, 0x44695073, dips_b200.SYNTH_SCENE,
                                torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    host = dev.cpu().numpy().reshape(n, fb)
    del dev
    want = oracle.run_clip(host, fmt, mode, 32)
    pinned = torch.from_numpy(host).pin_memory()
    for src in ("pageable", "pinned"):
        with dips_b200.Context(w, h, fmt, mode, 32) as ctx:
            if src == "pinned":
                ctx.run_clip_host(pinned.data_ptr(), n, fb)
            else:
                ctx.run_clip_host(host)
            ctx.synchronize()
            s, c = ctx.get_accumulators()
            sad, cnt = ctx.get_scalars(0, n)
            assert ctx.frames_processed == n
        assert np.array_equal(sad, want.sad) and np.array_equal(cnt, want.cnt), src
        assert np.array_equal(s, want.acc_sum) and np.array_equal(c, want.acc_cnt), src


@pytest.mark.parametrize("seed", range(6))
def test_randomised_geometry_and_call_patterns(torch_cuda, oracle, seed):
    """Seeded fuzz: random frame sizes (aligned and ragged), formats, modes, chroma filters, thresholds, pitches, call
    splits, kernels and tunings -- every case bit-exact against the oracle."""
    rng = np.random.default_rng(0xD1B5 + seed)
    for case in range(8):
        fmt = int(rng.integers(0, 4))
        mode = int(rng.integers(0, 2))
        chroma = int(rng.integers(0, 4))
        tau = int(rng.choice([0, 1, 3, 17, 64, 200, 509, 600]))
        w = int(rng.integers(1, 41)) * 16 if rng.random() < 0.6 else int(rng.integers(1, 700))
        h = int(rng.integers(1, 48))
        n = int(rng.integers(1, 36))
        profile = oracle.SYNTH_SCENE if rng.random() < 0.5 else oracle.SYNTH_UNIFORM
        clip = oracle.synth_clip(n, w, h, fmt, seed=int(rng.integers(1, 1 << 30)), profile=profile)
        fb = clip.shape[1]
        pad = int(rng.choice([0, 0, (-fb) % 16, (-fb) % 16 + 32, 7]))
        cuts = sorted(set(int(x) for x in rng.integers(1, n + 1, size=int(rng.integers(0, 4)))) - {n})
        chunks = [0] + cuts + [n]
        tuning = {}
        if rng.random() < 0.5:
            tuning["kernel"] = int(rng.integers(0, 2))
            if tuning["kernel"] == 1 and fmt in (1, 3) and mode == 1 and chroma:
                tuning["kernel"] = 0              # the one combination clip_kernel_ws is not built for
        if rng.random() < 0.4:
            tuning["segments"] = int(rng.integers(1, 5))
        if rng.random() < 0.3 and tuning.get("kernel") != 1:
            tuning["regs"] = int(rng.choice([64, 72, 80, 96, 128]))
        if rng.random() < 0.3:
            tuning["stages"] = int(rng.choice([3, 4]))
        got = run_gpu(torch_cuda, clip, w, h, fmt, mode, tau, chroma, chunks=chunks, tuning=tuning or None, stride_pad=pad)
        try:
            check(oracle, got, clip, fmt, mode, tau, chroma)
        except AssertionError as e:
            raise AssertionError(f"seed {seed} case {case}: {w}x{h}x{n} fmt {fmt} mode {mode} chroma {chroma} tau {tau} "
                                 f"pad {pad} chunks {chunks} tuning {tuning} plan {got[5]}: {e}") from e


def test_accumulator_range_guard(torch_cuda, oracle):
    """A context refuses to accumulate more frames than its u32 per-pixel sums can hold (DIPSB_MAX_ACCUMULATED_FRAMES)
    instead of wrapping; results so far stay intact and dipsb_reset clears the condition."""
    import dips_b200
    limit = 8421504
    w, h, fmt = 4, 1, 1                                        # 16-byte frames
    clip = oracle.synth_clip(10, w, h, fmt, profile=oracle.SYNTH_UNIFORM)
    want = oracle.run_clip(clip, fmt, 0, 5)
    big = torch_cuda.zeros((limit + 1) * 16, dtype=torch_cuda.uint8, device="cuda")
    big[:160] = torch_cuda.from_numpy(clip.reshape(-1)).cuda()
    with dips_b200.Context(w, h, fmt, 0, 5) as ctx:
        with pytest.raises(dips_b200.DipsError) as e:
            ctx.run_clip_device(big.data_ptr(), limit + 1, 16, 0)
        assert e.value.code == -4 and "overflow" in str(e.value)
        ctx.run_clip_device(big.data_ptr(), 10, 16, 0)
        with pytest.raises(dips_b200.DipsError):
            ctx.run_clip_device(big.data_ptr(), limit - 9, 16, 10)
        acc_sum, acc_cnt = ctx.get_accumulators()
        assert np.array_equal(acc_sum, want.acc_sum) and np.array_equal(acc_cnt, want.acc_cnt)
        assert ctx.frames_processed == 10
        ctx.reset()
        ctx.run_clip_device(big.data_ptr(), 10, 16, 0)
        assert np.array_equal(ctx.get_accumulators()[0], want.acc_sum)


def test_run_clip_host_back_to_back_without_sync(torch_cuda, oracle):
    """Several dipsb_run_clip_host calls with no synchronisation in between (a long video fed in pieces; reset + run loops):
    the staging slots are guarded across calls, pageable and page-locked clips alike."""
    import dips_b200
    torch = torch_cuda
    w, h, fmt, tau = 512, 288, 0, 9
    n = 48
    clip = oracle.synth_clip(n, w, h, fmt, profile=oracle.SYNTH_SCENE)
    pinned = torch.from_numpy(clip.copy()).pin_memory()
    for mode in (0, 1):
        want = oracle.run_clip(clip, fmt, mode, tau)
        for src in (clip, pinned.numpy()):
            with dips_b200.Context(w, h, fmt, mode, tau) as ctx:
                for rep in range(3):                       # reset + run, three times, never synchronising
                    ctx.reset()
                    for a, b in ((0, 7), (7, 8), (8, 31), (31, n)):
                        ctx.run_clip_host(src[a:b], first_frame=a)
                s, c = ctx.get_accumulators()
                sad, cnt = ctx.get_scalars(0, n)
            assert np.array_equal(s, want.acc_sum) and np.array_equal(c, want.acc_cnt)
            assert np.array_equal(sad, want.sad) and np.array_equal(cnt, want.cnt)


def test_shard_engine_on_a_default_stream_context(torch_cuda, oracle):
    """GpuShardEngine puts a Context that still runs on its private stream onto torch's current stream, so that torch's
    collectives and copies are ordered with the library's kernels."""
    import dips_b200
    from dips_b200 import sharding
    torch = torch_cuda
    w, h, fmt, tau, n = 256, 128, 1, 5, 10
    clip = oracle.synth_clip(n, w, h, fmt)
    dev = to_device(torch, clip)
    want = oracle.run_clip(clip, fmt, 0, tau)
    with dips_b200.Context(w, h, fmt, 0, tau) as ctx:          # private stream, never set_stream by the caller
        eng = sharding.GpuShardEngine(ctx, dev, torch)
        eng.prime(dev[0])
        sharding.run_sharded(eng, 0, 0)
        acc = eng.acc_tensor().clone()                          # torch op on torch's stream right behind the kernels
        torch.cuda.current_stream().synchronize()
        ptr, ne = ctx.accumulators_device()
        s, c = ctx.get_accumulators()
    assert np.array_equal(s, want.acc_sum) and np.array_equal(c, want.acc_cnt)
    assert int(acc[:ne].to(torch.int64).sum()) == int(want.acc_sum.astype(np.int64).sum())
