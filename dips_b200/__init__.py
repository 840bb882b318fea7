"""dips_b200 -- B200-native (sm_100a) implementation of the DiPs per-pixel frame-difference hot path.

The product is the C-ABI shared library `libdips_b200.so` (include/dips_b200.h).  This package is the thin Python host
layer used by the tests and bench.py: a ctypes binding (`api.Context`) and the in-tree build (`_build.build`).
"""
from .api import (  # noqa: F401
    CHROMA_BLUE, CHROMA_GREEN, CHROMA_NONE, CHROMA_RED, FILTER_INV_SIGMOID, FILTER_NONE, FILTER_SIGMOID, FMT_BGR8,
    FLAVOR_ALT_RING2, FLAVOR_ALT_RING2_MEDIAN, FLAVOR_DIPS_RING4, FLAVOR_FRAME0, FMT_BGRX8, FMT_RGB8, FMT_RGBX8, MODE_OVERALL, MODE_PERFRAME, NOT_READY, SYNTH_SCENE, SYNTH_UNIFORM, Context,
    DipsError, Group, PinnedBuffer, RegisteredBuffer, REDUCE_AUTO, REDUCE_NCCL, REDUCE_P2P, SnapshotSchedule, bytes_per_pixel, comm_unique_id, launch_count,
    run_dips_on_frames, shard_range, synth_fill_device, xchg_plan_query,
)
