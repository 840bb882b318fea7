"""ctypes loader for libdips_b200.so.  There is no fallback: if the library is missing or cannot be loaded the import of
anything that needs it raises, loudly."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libdips_b200.so")


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("device", C.c_int32), ("width", C.c_uint32), ("height", C.c_uint32),
        ("format", C.c_int32), ("mode", C.c_int32), ("chroma", C.c_int32), ("threshold", C.c_uint32),
        ("colorize", C.c_int32), ("filter", C.c_int32), ("sigmoid_scalar", C.c_float), ("spatial_window", C.c_int32),
        ("flavor", C.c_int32), ("reserved", C.c_uint32 * 3),
    ]


class FrameStats(C.Structure):
    _fields_ = [("frame_index", C.c_uint64), ("sad", C.c_uint64), ("count", C.c_uint64)]


# every symbol include/dips_b200.h declares: name -> (restype, argtypes)
_vp, _u64, _u32, _i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32
SYMBOLS = {
    "dipsb_abi_version": (_i32, []),
    "dipsb_default_config": (None, [C.POINTER(Config)]),
    "dipsb_create": (_i32, [C.POINTER(Config), C.POINTER(_vp)]),
    "dipsb_destroy": (None, [_vp]),
    "dipsb_last_error": (C.c_char_p, [_vp]),
    "dipsb_reset": (_i32, [_vp]),
    "dipsb_set_threshold": (_i32, [_vp, _u32]),
    "dipsb_set_stream": (_i32, [_vp, _vp]),
    "dipsb_adopt_stream": (_i32, [_vp, _vp]),
    "dipsb_use_private_stream": (_i32, [_vp]),
    "dipsb_synchronize": (_i32, [_vp]),
    "dipsb_prime_device": (_i32, [_vp, _vp]),
    "dipsb_prime_median4_device": (_i32, [_vp, _vp, _u64]),
    "dipsb_prime_host": (_i32, [_vp, _vp]),
    "dipsb_state_plane_device": (_i32, [_vp, C.POINTER(_vp)]),
    "dipsb_mark_state_valid": (_i32, [_vp, _i32]),
    "dipsb_get_state_plane": (_i32, [_vp, _vp]),
    "dipsb_run_clip_device": (_i32, [_vp, _vp, _u64, _u64, _u64]),
    "dipsb_run_clip_host": (_i32, [_vp, _vp, _u64, _u64, _u64]),
    "dipsb_push_frame": (_i32, [_vp, _vp, _u32, _u32, _u32, _i32, _vp, C.POINTER(FrameStats)]),
    "dipsb_push_frame_pipelined": (_i32, [_vp, _vp, _u32, _u32, _u32, _i32, _vp, C.POINTER(FrameStats)]),
    "dipsb_flush_frame": (_i32, [_vp, _vp, C.POINTER(FrameStats)]),
    "dipsb_stage_frame": (_i32, [_vp, _vp, _u32, _u32, _u32, _i32]),
    "dipsb_dispatch_staged": (_i32, [_vp, _vp, C.POINTER(FrameStats)]),
    "dipsb_snapshot": (_i32, [_vp]),
    "dipsb_frames_processed": (_u64, [_vp]),
    "dipsb_get_accumulators": (_i32, [_vp, _vp, _vp]),
    "dipsb_set_accumulators": (_i32, [_vp, _vp, _vp]),
    "dipsb_accumulators_device": (_i32, [_vp, C.POINTER(_vp), C.POINTER(_u64)]),
    "dipsb_pack_accumulators_device": (_i32, [_vp, _u64, C.POINTER(_vp), C.POINTER(_u64)]),
    "dipsb_unpack_accumulators_device": (_i32, [_vp]),
    "dipsb_get_scalars": (_i32, [_vp, _u64, _u64, _vp, _vp]),
    "dipsb_get_intensity_map": (_i32, [_vp, _u64, _vp]),
    "dipsb_get_frame_means": (_i32, [_vp, _u64, _u64, _vp]),
    "dipsb_host_alloc": (_i32, [_i32, _u64, C.POINTER(_vp)]),
    "dipsb_host_free": (_i32, [_vp]),
    "dipsb_host_register": (_i32, [_i32, _vp, _u64]),
    "dipsb_host_unregister": (_i32, [_vp]),
    "dipsb_host_copy2d": (_i32, [_vp, _u64, _vp, _u64, _u64, _u64]),
    "dipsb_host_copy_threads": (_u32, []),
    "dipsb_synth_fill_device": (_i32, [_i32, _vp, _u64, _u64, _u32, _u32, _i32, _u64, _i32, _vp]),
    "dipsb_launch_count": (_u64, []),
    "dipsb_last_plan": (_i32, [_vp, C.POINTER(_u32 * 8)]),
    "dipsb_enable_timing": (_i32, [_vp, _i32]),
    "dipsb_clip_kernel_time": (_i32, [_vp, C.POINTER(C.c_double), C.POINTER(_u64)]),
    "dipsb_stream_probe": (_i32, [_vp, _vp, _u64, _u64, _u32, C.POINTER(C.c_float)]),
    "dipsb_set_kernel": (_i32, [_vp, _i32]),
    "dipsb_plan_query": (_i32, [_u32, _u32, _i32, _u32, C.POINTER(_u32 * 8)]),
    "dipsb_set_tuning": (_i32, [_vp, _u32, _u32, _u32, _u32]),
    # several GPUs
    "dipsb_shard_range": (None, [_u64, _u32, _u32, C.POINTER(_u64), C.POINTER(_u64)]),
    "dipsb_xchg_plan_query": (_i32, [_u64, _u32, _u64, C.POINTER(_u64 * 4)]),
    "dipsb_comm_unique_id": (_i32, [_vp]),
    "dipsb_comm_init_rank": (_i32, [_vp, _u32, _u32, _vp]),
    "dipsb_comm_destroy": (_i32, [_vp]),
    "dipsb_comm_info": (_i32, [_vp, C.POINTER(_u32 * 8)]),
    "dipsb_comm_set_reduce": (_i32, [_vp, _i32]),
    "dipsb_comm_check": (_i32, [_vp]),
    "dipsb_run_clip_sharded_device": (_i32, [_vp, _vp, _u64, _u64, _u64, _u64]),
    "dipsb_run_clip_sharded_host": (_i32, [_vp, _vp, _u64, _u64, _u64, _u64]),
    "dipsb_comm_phase_times": (_i32, [_vp, C.POINTER(C.c_double * 3), C.POINTER(_u64)]),
    "dipsb_comm_probe": (_i32, [_vp, _i32, _u64, _u32, C.POINTER(C.c_float)]),
    "dipsb_gather_accumulators": (_i32, [_vp]),
    "dipsb_create_group": (_i32, [C.POINTER(Config), _u32, C.POINTER(_i32), C.POINTER(_vp)]),
    "dipsb_destroy_group": (None, [_vp]),
    "dipsb_group_size": (_u32, [_vp]),
    "dipsb_group_ctx": (_vp, [_vp, _u32]),
    "dipsb_group_last_error": (C.c_char_p, [_vp]),
    "dipsb_group_reset": (_i32, [_vp]),
    "dipsb_group_run_clip_device": (_i32, [_vp, C.POINTER(_vp), C.POINTER(_u64), _u64]),
    "dipsb_group_gather_accumulators": (_i32, [_vp]),
    "dipsb_group_synchronize": (_i32, [_vp]),
    "dipsb_group_get_accumulators": (_i32, [_vp, _vp, _vp]),
    "dipsb_group_get_scalars": (_i32, [_vp, _u64, _u64, _vp, _vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the library and bind every declared symbol (raises if the .so or a symbol is missing)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            raise ImportError(
                f"{SO} is missing: build it with `python -m dips_b200._build` (or __graft_entry__.build()); "
                "dips_b200 has no CPU or PyTorch fallback")
        lib = C.CDLL(SO)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)      # AttributeError if the export is missing
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
