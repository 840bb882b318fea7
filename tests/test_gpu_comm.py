"""GPU tests of the multi-GPU entry points of the C ABI (dipsb_comm_*, dipsb_create_group, dipsb_run_clip_sharded_*).

On a one-GPU box the ranks of a group share the device (loopback group: same kernels over the same "peer" windows, the
reference plane travels by copy-engine pushes instead of NCCL); with two or more GPUs visible the same checks run on a
real group (ncclCommInitAll, peer access over NVLink).  Everything is compared bit for bit with the CPU oracle run over the
WHOLE clip on one thread of control: the sharded result must not depend on the number of ranks."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    torch.cuda.init()
    return torch


def shard_bounds(n, ranks):
    import dips_b200
    b = [dips_b200.shard_range(n, ranks, r) for r in range(ranks)]
    assert b[0][0] == 0 and all(b[i][0] + b[i][1] == b[i + 1][0] for i in range(ranks - 1)) and b[-1][0] + b[-1][1] == n
    return b


def run_group(torch, oracle, devices, w, h, fmt, mode, tau, n, profile=None, reduce_path=None, passes=1):
    import dips_b200
    clip = oracle.synth_clip(n, w, h, fmt, profile=oracle.SYNTH_SCENE if profile is None else profile)
    want = oracle.run_clip(clip, fmt, mode, tau)
    bounds = shard_bounds(n, len(devices))
    shards = [torch.from_numpy(np.ascontiguousarray(clip[t0:t0 + k])).to(f"cuda:{d}") for (t0, k), d in zip(bounds, devices)]
    torch.cuda.synchronize()
    with dips_b200.Group(devices, w, h, fmt, mode, tau) as grp:
        if reduce_path is not None:
            for r in grp.ranks:
                r.comm_set_reduce(reduce_path)
        for _ in range(passes):                      # back-to-back passes exercise the epoch / parity logic of the windows
            grp.reset()
            grp.run_clip_device([s.data_ptr() for s in shards], [k for _, k in bounds])
        info = grp.ranks[0].comm_info()
        acc_sum, acc_cnt = grp.get_accumulators()
        sad, cnt = grp.get_scalars(0, n)
        # after the gather every rank holds the complete planes
        for r in grp.ranks[1:]:
            s2, c2 = r.get_accumulators()
            assert np.array_equal(s2, acc_sum) and np.array_equal(c2, acc_cnt)
    assert np.array_equal(sad, want.sad), "per-frame sad"
    assert np.array_equal(cnt, want.cnt), "per-frame count"
    assert np.array_equal(acc_sum, want.acc_sum), "acc_sum"
    assert np.array_equal(acc_cnt, want.acc_cnt), "acc_cnt"
    return info


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("fmt", [0, 1])
@pytest.mark.parametrize("ranks", [2, 3, 4])
def test_loopback_group_matches_oracle(torch_cuda, oracle, ranks, fmt, mode):
    info = run_group(torch_cuda, oracle, [0] * ranks, 320, 176, fmt, mode, 20, 41)
    assert info["nranks"] == ranks and info["peer_memory"] and info["reduce_path"] == "p2p" and info["single_process"]


@pytest.mark.parametrize("mode", [0, 1])
def test_loopback_group_repeated_passes_and_eight_ranks(torch_cuda, oracle, mode):
    run_group(torch_cuda, oracle, [0] * 8, 256, 96, 0, mode, 5, 67, passes=3)


@pytest.mark.parametrize("mode", [0, 1])
def test_loopback_group_wide_exchange_format(torch_cuda, oracle, mode):
    """more than 2056 frames per rank: sum and count no longer share a u32 and travel as two words"""
    import dips_b200
    n = 2 * 2100
    assert dips_b200.xchg_plan_query(n, 2)["bytes_per_element"] == 8
    run_group(torch_cuda, oracle, [0, 0], 48, 32, 1, mode, 3, n, profile=0)


@pytest.mark.parametrize("mode", [0, 1])
def test_loopback_group_unaligned_frames(torch_cuda, oracle, mode):
    """37x5 RGB8 frames (555 bytes): the clip is re-packed and the extra trailing frame goes through the one-frame path"""
    run_group(torch_cuda, oracle, [0, 0, 0], 37, 5, 0, mode, 9, 23)


def test_single_rank_communicator_is_the_plain_path(torch_cuda, oracle):
    import dips_b200
    torch = torch_cuda
    w, h, fmt, mode, tau, n = 200, 64, 0, 1, 7, 19
    clip = oracle.synth_clip(n, w, h, fmt, profile=oracle.SYNTH_SCENE)
    want = oracle.run_clip(clip, fmt, mode, tau)
    dev = torch.from_numpy(clip).cuda()
    with dips_b200.Group([0], w, h, fmt, mode, tau) as grp:
        grp.run_clip_device([dev.data_ptr()], [n])
        s, c = grp.get_accumulators()
        sad, cnt = grp.get_scalars(0, n)
    assert np.array_equal(s, want.acc_sum) and np.array_equal(c, want.acc_cnt)
    assert np.array_equal(sad, want.sad) and np.array_equal(cnt, want.cnt)


def test_sharded_totals_must_be_gathered_before_reading(torch_cuda, oracle):
    import dips_b200
    torch = torch_cuda
    w, h, fmt, n = 128, 64, 1, 12
    clip = oracle.synth_clip(n, w, h, fmt)
    bounds = shard_bounds(n, 2)
    shards = [torch.from_numpy(np.ascontiguousarray(clip[t0:t0 + k])).cuda() for t0, k in bounds]
    with dips_b200.Group([0, 0], w, h, fmt, 0, 10) as grp:
        grp.run_clip_device([s.data_ptr() for s in shards], [k for _, k in bounds])
        assert grp.ranks[0].comm_info()["acc_sharded"]
        with pytest.raises(dips_b200.DipsError) as e:
            grp.ranks[0].get_accumulators()
        assert "gather" in str(e.value)
        with pytest.raises(dips_b200.DipsError):          # one pass per reset
            grp.run_clip_device([s.data_ptr() for s in shards], [k for _, k in bounds])
        with pytest.raises(dips_b200.DipsError):          # geometry is frozen once the planes are mapped by the peers
            grp.ranks[0].set_tuning(3, 0, 0, 0)
        grp.gather_accumulators()
        grp.synchronize()
        assert not grp.ranks[1].comm_info()["acc_sharded"]


def _gpus(torch):
    return torch.cuda.device_count()


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("reduce_path", [1, 2])
def test_real_group_matches_oracle(torch_cuda, oracle, mode, reduce_path):
    """two or more GPUs: ncclCommInitAll + peer access; both accumulator paths (peer-memory kernels, NCCL all-reduce)"""
    torch = torch_cuda
    if _gpus(torch) < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    devices = list(range(min(_gpus(torch), 8)))
    info = run_group(torch, oracle, devices, 640, 360, 1, mode, 24, 16 * len(devices) + 5, reduce_path=reduce_path, passes=2)
    assert info["nccl"] and info["nranks"] == len(devices)
    assert info["reduce_path"] == ("p2p" if reduce_path == 1 else "nccl")


def test_real_group_baseline_geometry_shards(torch_cuda, oracle):
    """C4 geometry (3840x2160 RGBx8), a few frames per GPU, overall + per-frame"""
    torch = torch_cuda
    if _gpus(torch) < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    devices = list(range(min(_gpus(torch), 8)))
    for mode in (0, 1):
        run_group(torch, oracle, devices, 3840, 2160, 1, mode, 32, 5 * len(devices) + 1)
