"""CPU tests of the host-only parts of the multi-GPU entry points (no device, no NCCL): the frame ranges of the ranks, the
exchange format of the peer-memory accumulator reduce, and that the sharded calls fail cleanly instead of falling back."""
import ctypes

import pytest

import dips_b200
from dips_b200 import _lib, sharding


def test_shard_range_matches_the_floor_rule_and_covers_the_clip():
    for total in (1, 7, 300, 1200, 1800, 3600, 3601, 2**40 + 5):
        for world in (1, 2, 3, 4, 5, 8, 16):
            edges = [dips_b200.shard_range(total, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][0] + edges[-1][1] == total
            for (a, n), (b, _) in zip(edges, edges[1:]):
                assert a + n == b
            for r, (a, n) in enumerate(edges):
                assert a == (r * total) // world                          # rank r owns [floor(rN/R), floor((r+1)N/R))
                if total < 2**32:
                    assert (a, a + n) == sharding.shard_range(r, world, total)
            sizes = [n for _, n in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        dips_b200.shard_range(10, 2, 2)


def test_exchange_format_of_the_peer_memory_reduce():
    # BASELINE configurations: C4 (3600 frames) and C5 (1200 frames) over 2/4/8 ranks pack sum | count << bits into one u32
    for total, ranks in ((3600, 2), (3600, 4), (3600, 8), (1200, 8), (1800, 2)):
        p = dips_b200.xchg_plan_query(total, ranks, 148 * 896 * 16)
        bound = -(-total // ranks) + 1
        assert p["frames_per_rank_bound"] == bound
        assert p["bytes_per_element"] == 4 and (510 * bound).bit_length() == p["sum_bits"]
        assert p["sum_bits"] + bound.bit_length() <= 32
        assert p["owned_elements"] % 4 == 0 and p["owned_elements"] * ranks >= 148 * 896 * 16
    # a rank that differences more than 2056 frames needs two words per element
    assert dips_b200.xchg_plan_query(2 * 2055, 2)["bytes_per_element"] == 4
    assert dips_b200.xchg_plan_query(2 * 2056, 2)["bytes_per_element"] == 8
    assert dips_b200.xchg_plan_query(3600, 1)["bytes_per_element"] == 8
    with pytest.raises(dips_b200.DipsError):
        dips_b200.xchg_plan_query(0, 2)
    with pytest.raises(dips_b200.DipsError):
        dips_b200.xchg_plan_query(10, 17)


def test_sharded_entry_points_reject_null_handles_without_touching_a_device():
    L = _lib.load()
    assert L.dipsb_run_clip_sharded_device(None, None, 1, 1, 0, 1) == -1
    assert L.dipsb_run_clip_sharded_host(None, None, 1, 1, 0, 1) == -1
    assert L.dipsb_gather_accumulators(None) == -1
    assert L.dipsb_comm_init_rank(None, 2, 0, None) == -1
    assert L.dipsb_comm_check(None) == -1
    assert L.dipsb_group_size(None) == 0 and L.dipsb_group_ctx(None, 0) is None
    out = (ctypes.c_uint32 * 8)()
    assert L.dipsb_comm_info(None, ctypes.byref(out)) == -1


def test_create_group_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(dips_b200.DipsError) as e:
        dips_b200.Group([0, 0], 64, 48)
    assert "no CUDA device" in str(e.value)
