"""BASELINE.json's clips at their FULL length against the oracle, bit for bit (pytest -m gpu).

C2 / C3: all 1800 frames of the 1920x1080 RGB8 clip, overall + threshold and per-frame + scalars -- accumulators, counts,
every per-frame scalar and the final state plane.  C4 / C5: what one of 8 GPUs runs, a 450-frame 3840x2160 RGBx8 shard and a
150-frame 7680x4320 RGB8 shard (both modes), through the same clip kernel plan as the full clips.  The clips are generated on
the device (the generator's bytes are pinned against the oracle's in test_gpu_parity.py) and copied to the host for the
oracle, which runs on every host core: a few seconds per case."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 0x44695073


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    torch.cuda.init()
    return torch


def cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def run_case(torch, oracle, w, h, fmt, mode, tau, n, first_frame=0, prime_with=None):
    """n frames starting at logical frame `first_frame` of the synthetic clip; prime_with: logical index of the frame the
    state plane is primed from before the pass (what a shard receives), or None for a clip that starts at its own frame 0."""
    import dips_b200
    fb = w * h * dips_b200.bytes_per_pixel(fmt)
    dev = torch.empty(n * fb, dtype=torch.uint8, device="cuda")
    dips_b200.synth_fill_device(0, dev.data_ptr(), first_frame, n, w, h, fmt, SEED, dips_b200.SYNTH_SCENE,
                                torch.cuda.current_stream().cuda_stream)
    ref_dev = None
    if prime_with is not None:
        ref_dev = torch.empty(fb, dtype=torch.uint8, device="cuda")
        dips_b200.synth_fill_device(0, ref_dev.data_ptr(), prime_with, 1, w, h, fmt, SEED, dips_b200.SYNTH_SCENE,
                                    torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    with dips_b200.Context(w, h, fmt, mode, tau) as ctx:
        if ref_dev is not None:
            ctx.prime_device(ref_dev.data_ptr())
        ctx.run_clip_device(dev.data_ptr(), n, fb, first_frame)
        ctx.synchronize()
        acc_sum, acc_cnt = ctx.get_accumulators()
        sad, cnt = ctx.get_scalars(first_frame, n)
        state = ctx.get_state_plane()
        plan = ctx.last_plan()
        imap = ctx.get_intensity_map(n)
        means = ctx.get_frame_means(first_frame, n)
    assert plan["tma_path"] and plan["kernel"] == 1
    host = dev.cpu().numpy().reshape(n, fb)
    del dev
    state0 = None if ref_dev is None else oracle.i2_plane(ref_dev.cpu().numpy(), fmt)
    want = oracle.run_clip(host, fmt, mode, tau, state=state0, nthreads=cores())
    assert np.array_equal(sad, want.sad), "per-frame sad"
    assert np.array_equal(cnt, want.cnt), "per-frame count"
    assert np.array_equal(acc_sum, want.acc_sum), "acc_sum"
    assert np.array_equal(acc_cnt, want.acc_cnt), "acc_cnt"
    assert np.array_equal(state, want.state), "state plane"
    # X6 float outputs, 1e-5 relative (one division of the exact integers)
    np.testing.assert_allclose(imap, oracle.intensity_map(want.acc_sum, n), rtol=1e-5, atol=0)
    np.testing.assert_allclose(means, oracle.frame_means(want.sad, w * h), rtol=1e-5, atol=0)
    assert int(acc_sum.astype(np.uint64).sum()) == int(sad.sum()) and int(acc_cnt.astype(np.uint64).sum()) == int(cnt.sum())


def test_config2_all_1800_frames_overall_threshold(torch_cuda, oracle):
    run_case(torch_cuda, oracle, 1920, 1080, 0, 0, 32, 1800)


def test_config3_all_1800_frames_perframe_scalars(torch_cuda, oracle):
    run_case(torch_cuda, oracle, 1920, 1080, 0, 1, 32, 1800)


def test_config4_450_frame_shard_of_the_4k_clip(torch_cuda, oracle):
    # rank 5 of 8: frames 2250 .. 2699, differenced against the clip's frame 0
    run_case(torch_cuda, oracle, 3840, 2160, 1, 0, 32, 450, first_frame=2250, prime_with=0)


@pytest.mark.parametrize("mode", [0, 1])
def test_config5_150_frame_shard_of_the_8k_clip(torch_cuda, oracle, mode):
    # rank 3 of 8: frames 450 .. 599; overall: against frame 0; per-frame: primed with the halo frame 449
    run_case(torch_cuda, oracle, 7680, 4320, 0, mode, 32, 150, first_frame=450, prime_with=0 if mode == 0 else 449)
