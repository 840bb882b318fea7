"""GPU tests of the streaming boundary (dipsb_push_frame): the reference's per-frame callback contract
(dips/src/lib.rs:233-246): same-frame RGBA8 output, passthrough while there is no reference, alpha 255."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rgba_of(frame, fmt, npx):
    bpp = 3 if fmt in (0, 2) else 4
    px = frame.reshape(npx, bpp)
    order = [2, 1, 0] if fmt in (2, 3) else [0, 1, 2]
    out = np.full((npx, 4), 255, np.uint8)
    out[:, :3] = px[:, order]
    if bpp == 4:
        out[:, 3] = px[:, 3]
    return out.reshape(-1)


@pytest.mark.parametrize("fmt", [1, 0, 3])
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("filt,colorize", [(255, False), (0, False), (1, False), (255, True), (0, True)])
def test_push_frame_matches_oracle(oracle, fmt, mode, filt, colorize):
    import dips_b200
    w, h, n, tau = 96, 54, 7, 10
    npx = w * h
    clip = oracle.synth_clip(n, w, h, fmt, profile=oracle.SYNTH_SCENE)
    want = oracle.run_clip(clip, fmt, mode, tau)
    i2 = [oracle.i2_plane(clip[t], fmt) for t in range(n)]
    with dips_b200.Context(w, h, fmt, mode, tau, colorize=colorize, filt=filt, sigmoid_scalar=5.0) as ctx:
        for t in range(n):
            rc, rgba, (idx, sad, cnt) = ctx.push_frame(clip[t])
            assert idx == t and sad == int(want.sad[t]) and cnt == int(want.cnt[t])
            if t == 0:
                assert rc == dips_b200.NOT_READY
                assert np.array_equal(rgba, rgba_of(clip[0], fmt, npx))          # passthrough
            else:
                assert rc == 0
                ref = i2[0] if mode == 0 else i2[t - 1]
                vis = oracle.visual_frame(ref, i2[t], colorize, filt, 5.0)
                diff = np.abs(vis.astype(int) - rgba.astype(int))
                assert diff.max() <= 1, f"visual frame differs by {diff.max()} LSB"      # X7: +-1 LSB
                assert np.all(rgba.reshape(-1, 4)[:, 3] == 255)
        s, c = ctx.get_accumulators()
        assert np.array_equal(s, want.acc_sum) and np.array_equal(c, want.acc_cnt)
        sad, cnt = ctx.get_scalars(0, n)
        assert np.array_equal(sad, want.sad) and np.array_equal(cnt, want.cnt)


def test_snapshot_and_row_stride(oracle):
    import dips_b200
    w, h, fmt = 50, 20, 1
    clip = oracle.synth_clip(6, w, h, fmt, profile=oracle.SYNTH_SCENE)
    with dips_b200.Context(w, h, fmt, 0, 0) as ctx:
        ctx.push_frame(clip[0], want_rgba=False)
        _, _, (_, sad1, _) = ctx.push_frame(clip[1], want_rgba=False)
        assert sad1 == int(np.abs(oracle.i2_plane(clip[1], fmt).astype(int) - oracle.i2_plane(clip[0], fmt).astype(int)).sum())
        ctx.snapshot()                                   # dips_alt refresh marker: next frame becomes the reference
        rc, _, (_, sad2, _) = ctx.push_frame(clip[2], want_rgba=False)
        assert rc == dips_b200.NOT_READY and sad2 == 0
        _, _, (_, sad3, _) = ctx.push_frame(clip[3], want_rgba=False)
        assert sad3 == int(np.abs(oracle.i2_plane(clip[3], fmt).astype(int) - oracle.i2_plane(clip[2], fmt).astype(int)).sum())
        # padded rows
        stride = w * 4 + 12
        padded = np.zeros((h, stride), np.uint8)
        padded[:, : w * 4] = clip[4].reshape(h, w * 4)
        _, _, (_, sad4, _) = ctx.push_frame(padded, stride=stride, want_rgba=False)
        assert sad4 == int(np.abs(oracle.i2_plane(clip[4], fmt).astype(int) - oracle.i2_plane(clip[2], fmt).astype(int)).sum())
        # wrong geometry is an error, not a crash
        with pytest.raises(dips_b200.DipsError):
            ctx._ck(ctx._lib.dipsb_push_frame(ctx._h, clip[0].ctypes.data, w + 1, h, (w + 1) * 4, fmt, None, None))


def test_push_then_batch_interop(oracle):
    """frames pushed one by one and a following device batch continue the same accumulators and indices"""
    import torch
    import dips_b200
    w, h, fmt, mode, tau = 64, 40, 0, 1, 5
    clip = oracle.synth_clip(20, w, h, fmt, profile=oracle.SYNTH_SCENE)
    want = oracle.run_clip(clip, fmt, mode, tau)
    dev = torch.from_numpy(clip).cuda()
    with dips_b200.Context(w, h, fmt, mode, tau) as ctx:
        for t in range(6):
            ctx.push_frame(clip[t], want_rgba=False)
        ctx.run_clip_device(dev[6].data_ptr(), 14, clip.shape[1], 6)
        ctx.synchronize()
        s, c = ctx.get_accumulators()
        sad, cnt = ctx.get_scalars(0, 20)
    assert np.array_equal(s, want.acc_sum) and np.array_equal(c, want.acc_cnt)
    assert np.array_equal(sad, want.sad) and np.array_equal(cnt, want.cnt)


def _within(a, b, tol):
    return np.abs(a.astype(int) - b.astype(int)) <= tol


@pytest.mark.parametrize("filt,colorize", [(255, False), (0, False), (0, True)])
@pytest.mark.parametrize("chroma", [0, 2])
def test_dips_ring4_flavour_matches_reference_state_machine(oracle, filt, colorize, chroma):
    """N1: `dips` crate semantics -- 3 passthrough frames, start = grey(upper median of 4), output = start - median of the
    4-frame ring with in-place grey quantisation.  The GPU keeps everything in exact integers (ties of the rgba8unorm
    store round half up); the oracle restates the shader in f32, where a .5 tie may land on either side (the reference
    leaves it to the driver).  One grey level of the start/ring plane is 2.5 LSB of output (x0.5 x5), hence: every pixel
    within 3 LSB, and at least 97 % within 1 LSB."""
    import dips_b200
    w, h, n = 80, 45, 11
    clip = oracle.synth_clip(n, w, h, oracle.FMT_RGBX8, profile=oracle.SYNTH_SCENE)
    cs = oracle.ComputeStateOracle(w, h, colorize, filt, 5.0, chroma)
    with dips_b200.Context(w, h, dips_b200.FMT_RGBX8, 0, 0, chroma, colorize=colorize, filt=filt,
                           flavor=dips_b200.FLAVOR_DIPS_RING4) as ctx:
        for t in range(n):
            want, passthrough = cs.frame(clip[t])
            rc, got, _ = ctx.push_frame(clip[t])
            assert (rc == dips_b200.NOT_READY) == passthrough == (t < 3)
            if passthrough:
                assert np.array_equal(got, clip[t])
            else:
                assert _within(got, want, 3).all(), f"frame {t}: max diff {np.abs(got.astype(int) - want.astype(int)).max()}"
                assert _within(got, want, 1).mean() >= 0.97
        with pytest.raises(dips_b200.DipsError):
            ctx.run_clip_device(0, 1)                  # a null clip is rejected (the batch call itself accepts ring flavours)


@pytest.mark.parametrize("intended", [False, True])
@pytest.mark.parametrize("filt,colorize", [(0, True), (255, False), (1, False)])
def test_dips_alt_ring2_flavour_matches_reference_state_machine(oracle, intended, filt, colorize):
    """N1: `dips_alt` semantics -- 2-frame ring, min-of-two (as shipped) or in-bounds median, snapshot on the 3rd frame and
    at a refresh marker (dips_alt/src/lib.rs:222-225, :668-670).  Same tolerance reasoning as the ring-of-4 test."""
    import dips_b200
    w, h, n = 64, 36, 10
    clip = oracle.synth_clip(n, w, h, oracle.FMT_RGBX8, profile=oracle.SYNTH_SCENE)
    ref = oracle.DiPsComputeOracle(w, h, colorize, filt, 5.0, 0, intended)
    flavor = dips_b200.FLAVOR_ALT_RING2_MEDIAN if intended else dips_b200.FLAVOR_ALT_RING2
    with dips_b200.Context(w, h, dips_b200.FMT_RGBX8, 0, 0, colorize=colorize, filt=filt, flavor=flavor) as ctx:
        for t in range(n):
            snap = t in (2, 6)
            want = ref.send_frame(clip[t], snap)
            if snap:
                ctx.snapshot()
            rc, got, _ = ctx.push_frame(clip[t])
            assert rc == 0
            assert _within(got, want, 3).all(), f"frame {t}: max diff {np.abs(got.astype(int) - want.astype(int)).max()}"
            assert _within(got, want, 1).mean() >= 0.97


@pytest.mark.parametrize("flavor", [0, 1, 2])
def test_pipelined_push_equals_synchronous(oracle, flavor):
    """N2: the one-frame-latency pipelined call returns exactly what the synchronous call returns, one call later."""
    import dips_b200
    w, h, n, fmt = 96, 54, 9, 1
    clip = oracle.synth_clip(n, w, h, fmt, profile=oracle.SYNTH_SCENE)
    sync = []
    with dips_b200.Context(w, h, fmt, 0, 7, flavor=flavor) as ctx:
        for t in range(n):
            sync.append(ctx.push_frame(clip[t]))
        acc_sync = ctx.get_accumulators()
    with dips_b200.Context(w, h, fmt, 0, 7, flavor=flavor) as ctx:
        got = []
        for t in range(n):
            rc, rgba, st = ctx.push_frame_pipelined(clip[t])
            if t == 0:
                assert rc == dips_b200.NOT_READY and rgba is None
            else:
                got.append((rc, rgba, st))
        with pytest.raises(dips_b200.DipsError):
            ctx.push_frame(clip[0])                      # mixing without a flush is refused
        got.append(ctx.flush_frame())
        assert ctx.flush_frame()[0] == dips_b200.NOT_READY
        acc_pipe = ctx.get_accumulators()
    assert len(got) == n
    for t in range(n):
        rc_s, rgba_s, st_s = sync[t]
        rc_p, rgba_p, st_p = got[t]
        assert rc_p == (2 if rc_s == dips_b200.NOT_READY else 0)
        assert st_p == st_s and np.array_equal(rgba_p, rgba_s), t
    assert np.array_equal(acc_sync[0], acc_pipe[0]) and np.array_equal(acc_sync[1], acc_pipe[1])


def _expected_from_planes(planes, mode, tau, ref_plane=None):
    """difference path on already filtered I2 planes (numpy restatement of SURVEY X1-X5)"""
    planes = np.stack(planes).astype(np.int64)
    if mode == 0:
        ref = planes[0] if ref_plane is None else ref_plane.astype(np.int64)
        d = np.abs(planes - ref[None])
        s_signed = ref[None] - planes
    else:
        prev = np.concatenate([planes[:1], planes[:-1]])
        d = np.abs(planes - prev)
        s_signed = prev - planes
    m = d > tau
    return d.sum(0).astype(np.uint32), m.sum(0).astype(np.uint32), d.sum(1).astype(np.uint64), m.sum(1).astype(np.uint64), s_signed


@pytest.mark.parametrize("window", [3, 5, 7])
@pytest.mark.parametrize("fmt,mode", [(1, 0), (0, 1)])
def test_spatial_window_streaming_and_batch(oracle, window, fmt, mode):
    """N4: window > 1 filters every frame's intensity with the correct zero-padded median before differencing --
    the streaming call (stats, visual frame) and the batch call (accumulators, scalars) against the oracle's filter."""
    import torch
    import dips_b200
    w, h, n, tau = 70, 41, 6, 9
    clip = oracle.synth_clip(n, w, h, fmt, profile=oracle.SYNTH_SCENE)
    planes = [oracle.filtered_i2(clip[t], w, h, fmt, window) for t in range(n)]
    acc_sum, acc_cnt, sad, cnt, s_signed = _expected_from_planes(planes, mode, tau)
    with dips_b200.Context(w, h, fmt, mode, tau, spatial_window=window) as ctx:
        for t in range(n):
            rc, rgba, (idx, s, c) = ctx.push_frame(clip[t])
            assert (s, c) == (int(sad[t]), int(cnt[t])), (t, s, c)
            if t > 0:
                ref = planes[0] if mode == 0 else planes[t - 1]
                vis = oracle.visual_frame(ref, planes[t], False, oracle.FILTER_NONE, 5.0)
                assert np.abs(vis.astype(int) - rgba.astype(int)).max() <= 1
        gs, gc = ctx.get_accumulators()
        assert np.array_equal(gs, acc_sum) and np.array_equal(gc, acc_cnt)
        assert np.array_equal(ctx.get_state_plane(), planes[0] if mode == 0 else planes[-1])
    dev = torch.from_numpy(clip).cuda()
    with dips_b200.Context(w, h, fmt, mode, tau, spatial_window=window) as ctx:
        ctx.run_clip_device(dev.data_ptr(), n)
        ctx.synchronize()
        assert not ctx.last_plan()["tma_path"]            # windowed contexts run frame by frame
        gs, gc = ctx.get_accumulators()
        gsad, gcnt = ctx.get_scalars(0, n)
        assert np.array_equal(gs, acc_sum) and np.array_equal(gc, acc_cnt)
        assert np.array_equal(gsad, sad) and np.array_equal(gcnt, cnt)
        # reference "start" plane with a window: each of the 4 frames filtered, then the upper median
        ctx.reset()
        ctx.prime_median4_device(dev.data_ptr(), clip.shape[1])
        want = np.sort(np.stack(planes[:4]), axis=0)[2]
        assert np.array_equal(ctx.get_state_plane(), want)


@pytest.mark.parametrize("flavor", [0, 1, 2])
@pytest.mark.parametrize("pin_in,pin_out", [(True, True), (True, False), (False, True)])
def test_page_locked_buffers_take_the_direct_path_with_identical_results(oracle, flavor, pin_in, pin_out):
    """dipsb_host_alloc buffers are read and written by the copy engine directly (no staging memcpy); results are the
    same bytes as with ordinary memory, and the input buffer is free for reuse as soon as each call returns (the
    reference borrows the slice for the callback only, frame_extractor.rs:224-226)."""
    import dips_b200
    w, h, n, fmt = 128, 72, 10, 1
    clip = oracle.synth_clip(n, w, h, fmt, profile=oracle.SYNTH_SCENE)
    with dips_b200.Context(w, h, fmt, 0, 7, flavor=flavor) as ctx:
        want = [ctx.push_frame(clip[t]) for t in range(n)]
        acc_want = ctx.get_accumulators()
    fb = w * h * 4
    with dips_b200.PinnedBuffer(fb) as pin, dips_b200.PinnedBuffer(2 * fb) as pout:
        src = pin.array if pin_in else np.empty(fb, np.uint8)
        outs = [pout.array[:fb], pout.array[fb:]] if pin_out else [np.empty(fb, np.uint8), np.empty(fb, np.uint8)]
        # synchronous
        with dips_b200.Context(w, h, fmt, 0, 7, flavor=flavor) as ctx:
            for t in range(n):
                src[:] = clip[t]
                rc, rgba, st = ctx.push_frame(src, out=outs[0])
                src[:] = 0xA5                               # scribble: the library must be done with it
                assert rgba is outs[0]
                assert rc == want[t][0] and st == want[t][2] and np.array_equal(rgba, want[t][1]), t
            acc = ctx.get_accumulators()
            assert np.array_equal(acc[0], acc_want[0]) and np.array_equal(acc[1], acc_want[1])
        # pipelined
        with dips_b200.Context(w, h, fmt, 0, 7, flavor=flavor) as ctx:
            for t in range(n + 1):
                if t < n:
                    src[:] = clip[t]
                    rc, rgba, st = ctx.push_frame_pipelined(src, out=outs[t & 1])
                    src[:] = 0x5A
                else:
                    rc, rgba, st = ctx.flush_frame(out=outs[t & 1])
                if t == 0:
                    assert rc == dips_b200.NOT_READY
                    continue
                assert rc == (2 if want[t - 1][0] == dips_b200.NOT_READY else 0)
                assert st == want[t - 1][2] and np.array_equal(rgba, want[t - 1][1]), t
            acc = ctx.get_accumulators()
            assert np.array_equal(acc[0], acc_want[0]) and np.array_equal(acc[1], acc_want[1])


def test_pinned_buffer_lifecycle_and_errors():
    import dips_b200
    from dips_b200 import _lib
    import ctypes as C
    b = dips_b200.PinnedBuffer(1 << 20)
    b.array[:] = 7
    assert int(b.array.sum()) == 7 << 20
    b.close()
    b.close()                                               # idempotent
    lib = _lib.load()
    p = C.c_void_p()
    assert lib.dipsb_host_alloc(0, 0, C.byref(p)) == -1     # empty request
    assert lib.dipsb_host_alloc(0, 16, None) == -1
    assert lib.dipsb_host_alloc(9999, 16, C.byref(p)) == -2 and not p.value
    assert lib.dipsb_host_free(None) == 0


@pytest.mark.parametrize("flavor", [0, 1, 2])
@pytest.mark.parametrize("w,h,fmt,pad", [(1920, 1080, 1, 0), (1280, 720, 2, 96), (1024, 601, 0, 0)])
def test_large_frames_take_the_row_band_path_with_identical_results(oracle, flavor, w, h, fmt, pad):
    """Frames of 2 MB and more go through the synchronous call in row bands (upload / kernels / read-back overlapped).
    Same bytes as the whole-frame sequence of the pipelined call, with pageable and page-locked buffers, padded rows
    included; for the frame-0 flavour the integers are also checked against the oracle."""
    import dips_b200
    n = 7
    bpp = 3 if fmt in (0, 2) else 4
    clip = oracle.synth_clip(n, w, h, fmt, profile=oracle.SYNTH_SCENE)
    row, stride = w * bpp, w * bpp + pad
    padded = np.zeros((n, h, stride), np.uint8)
    padded[:, :, :row] = clip.reshape(n, h, row)
    padded = padded.reshape(n, -1)
    kw = dict(colorize=True, filt=dips_b200.FILTER_SIGMOID, flavor=flavor)
    with dips_b200.Context(w, h, fmt, 0, 9, **kw) as ctx:                     # whole-frame sequence
        want = []
        for t in range(n):
            rc, rgba, st = ctx.push_frame_pipelined(padded[t], stride=stride)
            if t:
                want.append((rc, rgba, st))
        want.append(ctx.flush_frame())
        acc_want = ctx.get_accumulators()
    if flavor == 0:
        ref = oracle.run_clip(clip, fmt, 0, 9)
        assert np.array_equal(acc_want[0], ref.acc_sum) and np.array_equal(acc_want[1], ref.acc_cnt)
        assert [s[2][1] for s in want] == [int(x) for x in ref.sad]
    with dips_b200.PinnedBuffer(h * stride) as pin, dips_b200.PinnedBuffer(w * h * 4) as pout:
        for pinned in (False, True):
            if pinned and pad:
                continue                                              # the direct path needs tightly packed rows anyway
            with dips_b200.Context(w, h, fmt, 0, 9, **kw) as ctx:
                for t in range(n):
                    if pinned:
                        pin.array[:] = padded[t]
                        rc, rgba, st = ctx.push_frame(pin.array, stride=stride, out=pout.array)
                    else:
                        rc, rgba, st = ctx.push_frame(padded[t], stride=stride)
                    assert (2 if rc == dips_b200.NOT_READY else 0) == want[t][0]
                    assert st == want[t][2], (t, st, want[t][2])
                    assert np.array_equal(rgba, want[t][1]), t
                acc = ctx.get_accumulators()
                assert np.array_equal(acc[0], acc_want[0]) and np.array_equal(acc[1], acc_want[1])


def test_stage_and_dispatch_equal_push_frame(oracle):
    """dipsb_stage_frame + dipsb_dispatch_staged (the reference's add_texture + dispatch split, dips/src/gpu/mod.rs:170, :306)
    give exactly what dipsb_push_frame gives, frame by frame, for the north-star and the ring flavours."""
    import dips_b200
    w, h, n = 320, 184, 9
    clip = oracle.synth_clip(n, w, h, 1, profile=oracle.SYNTH_SCENE)
    for flavor in (dips_b200.FLAVOR_FRAME0, dips_b200.FLAVOR_DIPS_RING4, dips_b200.FLAVOR_ALT_RING2):
        with dips_b200.Context(w, h, 1, 0, 12, colorize=True, filt=0, flavor=flavor) as a, \
                dips_b200.Context(w, h, 1, 0, 12, colorize=True, filt=0, flavor=flavor) as b:
            for t in range(n):
                if flavor == dips_b200.FLAVOR_ALT_RING2 and t == 2:
                    a.snapshot(); b.snapshot()
                ra, oa, sa = a.push_frame(clip[t])
                b.stage_frame(clip[t])
                rb, ob, sb = b.dispatch_staged()
                assert ra == rb and sa == sb and np.array_equal(oa, ob), (flavor, t)
            sa_, ca_ = a.get_accumulators()
            sb_, cb_ = b.get_accumulators()
            assert np.array_equal(sa_, sb_) and np.array_equal(ca_, cb_)
        with dips_b200.Context(w, h, 1, 0, 12) as c:
            with pytest.raises(dips_b200.DipsError):
                c.dispatch_staged()                      # nothing staged
            c.stage_frame(clip[0])
            c.stage_frame(clip[1])                       # re-staging replaces the frame
            rc, _, st = c.dispatch_staged()
            assert rc == dips_b200.NOT_READY and st[0] == 0


@pytest.mark.parametrize("flavor", [1, 2, 3])
def test_batch_call_runs_the_ring_flavours(oracle, flavor):
    """dipsb_run_clip_device / _host on a DIPS_RING4 / ALT_RING2[_MEDIAN] context: the reference-exact state machine over a
    resident clip gives the accumulators and per-frame scalars of the same frames pushed one by one."""
    import torch

    import dips_b200
    w, h, n, tau = 256, 144, 14, 10
    clip = oracle.synth_clip(n, w, h, 1, profile=oracle.SYNTH_SCENE)
    with dips_b200.Context(w, h, 1, 0, tau, flavor=flavor) as a:
        for t in range(n):
            a.push_frame(clip[t], want_rgba=False)
        want_s, want_c = a.get_accumulators()
        want_sad, want_cnt = a.get_scalars(0, n)
    assert int(want_sad.sum()) > 0
    dev = torch.from_numpy(clip).cuda()
    for host in (False, True):
        with dips_b200.Context(w, h, 1, 0, tau, flavor=flavor) as b:
            if host:
                b.run_clip_host(clip[:5], first_frame=0)
                b.run_clip_host(clip[5:], first_frame=5)
            else:
                b.run_clip_device(dev.data_ptr(), 9, first_frame=0)
                b.run_clip_device(dev.data_ptr() + 9 * clip.shape[1], n - 9, first_frame=9)       # the ring carries over
            s, c = b.get_accumulators()
            sad, cnt = b.get_scalars(0, n)
            assert b.frames_processed == n and not b.last_plan()["tma_path"]
        assert np.array_equal(s, want_s) and np.array_equal(c, want_c)
        assert np.array_equal(sad, want_sad) and np.array_equal(cnt, want_cnt)


def _ring_twin(i2, flavor, tau, snapshot_before=()):
    """oracle/numpy_twin.py::ring_clip -- the integer restatement of the two ring machines (checked against the oracle's f32
    restatement of the shaders in tests/test_oracle_ring_twin.py)"""
    from oracle import numpy_twin
    return numpy_twin.ring_clip(i2, flavor, tau, snapshot_before)


@pytest.mark.parametrize("flavor", [1, 2, 3])
@pytest.mark.parametrize("fmt,chroma", [(0, 0), (1, 0), (1, 2), (2, 3), (3, 1), (0, 1)])
def test_ring_clip_kernel_matches_integer_twin(oracle, monkeypatch, flavor, fmt, chroma):
    """The one-launch ring kernel (ring_clip_kernel: ring slots in registers, frame segments that rebuild the ring from the
    frames before them, packed accumulators flushed every 128 frames) against an independent numpy restatement and against
    the per-frame kernel, over a clip long enough for three flushes, with snapshots between batch calls and with forced
    short segments (ragged last one)."""
    import torch

    import dips_b200
    w, h, n, tau = 200, 72, 301, 24
    clip = oracle.synth_clip(n, w, h, fmt, profile=oracle.SYNTH_SCENE)
    i2 = np.stack([oracle.i2_plane(clip[t], fmt, chroma) for t in range(n)]).astype(np.int64)
    cuts = [0, 150, n]                                                    # two batch calls; a snapshot before the second
    snaps = (2, 150) if flavor != 1 else (150,)
    want = _ring_twin(i2, flavor, tau, snapshot_before=snaps)
    assert int(want[2].sum()) > 0 and int(want[3].sum()) > 0
    dev = torch.from_numpy(clip).cuda()
    fb = clip.shape[1]

    def run(ctx):
        if flavor != 1:                                                   # dips_alt: frames 0, 1 and the snapshot frame 2 one by one
            for t in range(3):
                if t == 2:
                    ctx.snapshot()
                ctx.push_frame(clip[t], want_rgba=False)
            ctx.run_clip_device(dev.data_ptr() + 3 * fb, cuts[1] - 3, first_frame=3)
        else:
            ctx.run_clip_device(dev.data_ptr(), cuts[1], first_frame=0)
        ctx.snapshot()
        ctx.run_clip_device(dev.data_ptr() + cuts[1] * fb, n - cuts[1], first_frame=cuts[1])
        return ctx.get_accumulators() + ctx.get_scalars(0, n), ctx.last_plan()

    for seg in (None, "36"):
        if seg:
            monkeypatch.setenv("DIPSB_RING_SEG_FRAMES", seg)
        with dips_b200.Context(w, h, fmt, 0, tau, chroma=chroma, flavor=flavor) as ctx:
            got, plan = run(ctx)
        assert plan["ring_clip"] and not plan["tma_path"]
        for g_, w_, name in zip(got, want, ("sum", "count", "sad", "cnt")):
            assert np.array_equal(np.asarray(g_, dtype=np.int64).ravel(), w_), (name, seg)
    monkeypatch.delenv("DIPSB_RING_SEG_FRAMES")
    monkeypatch.setenv("DIPSB_RING_BATCH", "0")                           # the same calls through the per-frame kernel
    with dips_b200.Context(w, h, fmt, 0, tau, chroma=chroma, flavor=flavor) as ctx:
        got, plan = run(ctx)
    assert not plan["ring_clip"]
    for g_, w_, name in zip(got, want, ("sum", "count", "sad", "cnt")):
        assert np.array_equal(np.asarray(g_, dtype=np.int64).ravel(), w_), (name, "per-frame")


def test_ring_clip_kernel_mixes_with_per_frame_calls_and_odd_layouts(oracle):
    """Per-frame calls after a batch call see the ring the batch left (stored back by the last segment); a geometry whose
    pixel count is not a multiple of 8 and an unaligned clip base stay on the per-frame kernel and agree."""
    import torch

    import dips_b200
    tau = 16
    for (w, h, off) in ((128, 96, 0), (99, 51, 0), (128, 96, 4)):
        n = 41
        clip = oracle.synth_clip(n, w, h, 1, profile=oracle.SYNTH_SCENE)
        i2 = np.stack([oracle.i2_plane(clip[t], 1, 0) for t in range(n)]).astype(np.int64)
        want = _ring_twin(i2, 1, tau)
        raw = torch.zeros(clip.size + 64, dtype=torch.uint8, device="cuda")
        raw[off: off + clip.size] = torch.from_numpy(clip).cuda().view(-1)
        base = raw.data_ptr() + off
        with dips_b200.Context(w, h, 1, 0, tau, flavor=1) as ctx:
            ctx.push_frame(clip[0], want_rgba=False)
            ctx.push_frame(clip[1], want_rgba=False)
            ctx.run_clip_device(base + 2 * clip.shape[1], 30, first_frame=2)
            used = ctx.last_plan()["ring_clip"]
            for t in range(32, n):
                ctx.push_frame(clip[t], want_rgba=False)
            s, c = ctx.get_accumulators()
            sad, cnt = ctx.get_scalars(0, n)
        assert used == ((w * h) % 8 == 0 and off % 16 == 0), (w, h, off)
        assert np.array_equal(s.ravel(), want[0]) and np.array_equal(c.ravel(), want[1])
        assert np.array_equal(np.asarray(sad, np.int64), want[2]) and np.array_equal(np.asarray(cnt, np.int64), want[3])


def test_registered_caller_buffer_takes_the_direct_path(oracle):
    """dipsb_host_register: a buffer the caller owns, page-locked in place, gives the same frames as ordinary memory (and is
    recognised as page-locked by the frame calls); unregistering twice / registering nothing are errors, not crashes."""
    import ctypes as C

    import dips_b200
    from dips_b200 import _lib
    w, h, n = 320, 200, 6
    clip = oracle.synth_clip(n, w, h, 1, profile=oracle.SYNTH_SCENE)
    pool = np.empty((2, w * h * 4), np.uint8)                      # a two-buffer "decoder pool"
    out_pool = np.empty((2, w * h * 4), np.uint8)
    with dips_b200.Context(w, h, 1, 0, 12, colorize=True, filt=0) as a, dips_b200.Context(w, h, 1, 0, 12, colorize=True, filt=0) as b, \
            dips_b200.RegisteredBuffer(pool) as rin, dips_b200.RegisteredBuffer(out_pool):
        for t in range(n):
            ra, oa, sa = a.push_frame(clip[t])
            rin.array[t & 1] = clip[t]
            rb, ob, sb = b.push_frame(pool[t & 1], out=out_pool[t & 1])
            assert ra == rb and sa == sb and np.array_equal(oa, ob), t
        assert np.array_equal(a.get_accumulators()[0], b.get_accumulators()[0])
    lib = _lib.load()
    assert lib.dipsb_host_unregister(C.c_void_p(pool.ctypes.data)) != 0          # already unregistered by the with block
    assert lib.dipsb_host_register(0, None, 16) != 0 and lib.dipsb_host_register(0, C.c_void_p(pool.ctypes.data), 0) != 0
    assert b"host_register" in lib.dipsb_last_error(None)
