"""CPU tests (gloo, world_size 2) of the frame-range sharding host logic in dips_b200/sharding.py.  The per-shard engine
is a numpy/oracle stand-in -- what is under test is the orchestration: ranges, frame-0 broadcast, halo send/recv and the
accumulator all-reduce must reproduce the single-rank result bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dips_b200 import sharding


def test_shard_ranges_cover_and_are_disjoint():
    for n in (1, 7, 300, 1800, 3601):
        for world in (1, 2, 3, 4, 8):
            edges = [sharding.shard_range(r, world, n) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for a, b in zip(edges, edges[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(2, 2, 10)


class OracleEngine:
    """ShardEngine over numpy arrays; the math is the oracle's (this test is about the exchange, not the kernel)."""

    def __init__(self, O, frames, fmt, mode, tau):
        self.O, self.frames, self.fmt, self.mode, self.tau = O, frames, fmt, mode, tau
        npx = frames.shape[1] // O.bpp(fmt)
        self.acc = np.zeros(2 * npx, np.int32)
        self._state_buf = np.zeros(npx, np.int16)
        self.primed_by_broadcast = False
        self.reduced = False
        self.state = None
        self.sad = self.cnt = None

    def first_frame(self):
        return torch.from_numpy(self.frames[0])

    def last_frame(self):
        return torch.from_numpy(self.frames[-1])

    def frame_buffer(self):
        return torch.empty(self.frames.shape[1], dtype=torch.uint8)

    def prime(self, frame):
        self.state = self.O.i2_plane(frame.numpy(), self.fmt)
        self._state_buf[:] = self.state.view(np.int16)

    def state_tensor(self):
        return torch.from_numpy(self._state_buf.view(np.uint8))

    def mark_primed(self):
        self.state = self._state_buf.view(np.uint16).copy()
        self.primed_by_broadcast = True

    def run(self, first_frame_index):
        r = self.O.run_clip(self.frames, self.fmt, self.mode, self.tau, state=self.state)
        npx = r.acc_sum.size
        self.acc[:npx] = r.acc_sum.view(np.int32)
        self.acc[npx:] = r.acc_cnt.view(np.int32)
        self.sad, self.cnt, self.first = r.sad, r.cnt, first_frame_index

    def acc_tensor(self):
        return torch.from_numpy(self.acc)

    def after_reduce(self):
        self.reduced = True


class ReplicatedEngine(OracleEngine):
    """The shard carries its own reference / halo frame (replicated at load time): no exchange before the pass."""

    replicated = True

    def __init__(self, O, clip, t0, t1, fmt, mode, tau):
        super().__init__(O, clip[t0:t1], fmt, mode, tau)
        self._ref = None if t0 == 0 else torch.from_numpy(clip[0] if mode == sharding.MODE_OVERALL else clip[t0 - 1])

    def local_reference(self, mode):
        return self._ref

    def state_tensor(self):
        raise AssertionError("a replicated shard must not take part in a broadcast")

    def frame_buffer(self):
        raise AssertionError("a replicated shard must not take part in a halo exchange")


class NotReplicatedEngine(OracleEngine):
    """Has the optional method but is not in replicated mode (dips_b200.sharding.GpuShardEngine without a reference, as
    bench.py's end-to-end leg builds it): every rank, rank 0 included, must take the collective path."""
    replicated = False

    def local_reference(self, mode):
        return None


def _worker(rank, world, port, mode, tmpdir, replicated=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        fmt, tau, n, w, h = O.FMT_RGB8, 12, 11, 24, 10
        clip = O.synth_clip(n, w, h, fmt, profile=O.SYNTH_SCENE)
        t0, t1 = sharding.shard_range(rank, world, n)
        if replicated == "off":
            eng = NotReplicatedEngine(O, clip[t0:t1], fmt, mode, tau)
            replicated = False
        else:
            eng = ReplicatedEngine(O, clip, t0, t1, fmt, mode, tau) if replicated else OracleEngine(O, clip[t0:t1], fmt, mode, tau)
        sharding.run_sharded(eng, mode, t0, rank, world, dist)
        whole = O.run_clip(clip, fmt, mode, tau)
        npx = w * h
        assert np.array_equal(eng.acc[:npx].view(np.uint32), whole.acc_sum), "acc_sum differs after all-reduce"
        assert np.array_equal(eng.acc[npx:].view(np.uint32), whole.acc_cnt), "acc_cnt differs after all-reduce"
        assert np.array_equal(eng.sad, whole.sad[t0:t1]) and np.array_equal(eng.cnt, whole.cnt[t0:t1])
        assert eng.reduced and eng.primed_by_broadcast == (mode == sharding.MODE_OVERALL and rank > 0 and not replicated)
        open(os.path.join(tmpdir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("mode", [sharding.MODE_OVERALL, sharding.MODE_PERFRAME])
def test_multi_rank_sharding_matches_single_rank(mode, world, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), mode, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


@pytest.mark.parametrize("mode", [sharding.MODE_OVERALL, sharding.MODE_PERFRAME])
def test_replicated_reference_needs_no_exchange_before_the_pass(mode, tmp_path):
    world = 3
    mp.spawn(_worker, args=(world, _free_port(), mode, str(tmp_path), True), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


@pytest.mark.parametrize("mode", [sharding.MODE_OVERALL, sharding.MODE_PERFRAME])
def test_engine_with_the_optional_method_but_not_replicated_still_meets_in_the_collective(mode, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), mode, str(tmp_path), "off"), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_gpu_shard_engine_defaults_to_the_collective_exchange():
    """bench.py's end-to-end engine is a GpuShardEngine built without `replicated`: it must not skip the broadcast."""
    eng = sharding.GpuShardEngine(ctx=None, frames=None, torch=None, total_frames=10)
    assert eng.replicated is False and eng.local_reference(sharding.MODE_OVERALL) is None
    eng = sharding.GpuShardEngine(ctx=None, frames=None, torch=None, total_frames=10, replicated=True, reference={0: "x"})
    assert eng.replicated is True and eng.local_reference(0) == "x" and eng.local_reference(1) is None


def test_single_rank_needs_no_dist():
    from oracle import oracle as O
    clip = O.synth_clip(5, 12, 6, O.FMT_RGBX8)
    eng = OracleEngine(O, clip, O.FMT_RGBX8, 1, 3)
    sharding.run_sharded(eng, 1, 0)
    whole = O.run_clip(clip, O.FMT_RGBX8, 1, 3)
    assert np.array_equal(eng.acc[: 12 * 6].view(np.uint32), whole.acc_sum)
