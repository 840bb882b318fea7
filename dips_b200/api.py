"""Python host layer over the C ABI (include/dips_b200.h).  One `Context` == one `dipsb_ctx`.

Every method is a 1:1 call into libdips_b200.so; arrays cross the boundary as raw pointers (numpy for host memory,
integer device addresses -- e.g. `torch.Tensor.data_ptr()` -- for device memory).  Nothing here computes pixels.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

FMT_RGB8, FMT_RGBX8, FMT_BGR8, FMT_BGRX8 = 0, 1, 2, 3
MODE_OVERALL, MODE_PERFRAME = 0, 1
CHROMA_NONE, CHROMA_RED, CHROMA_GREEN, CHROMA_BLUE = 0, 1, 2, 3
FILTER_SIGMOID, FILTER_INV_SIGMOID, FILTER_NONE = 0, 1, 255
SYNTH_UNIFORM, SYNTH_SCENE = 0, 1
FLAVOR_FRAME0, FLAVOR_DIPS_RING4, FLAVOR_ALT_RING2, FLAVOR_ALT_RING2_MEDIAN = 0, 1, 2, 3
OK, NOT_READY = 0, 1


class DipsError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"dips_b200 error {code}: {message}")
        self.code = code


def bytes_per_pixel(fmt: int) -> int:
    return 3 if fmt in (FMT_RGB8, FMT_BGR8) else 4


def launch_count() -> int:
    return int(_lib.load().dipsb_launch_count())


def synth_fill_device(device: int, d_dst: int, first_frame: int, n_frames: int, width: int, height: int, fmt: int,
                      seed: int = 0x44695073, profile: int = SYNTH_SCENE, stream: int = 0) -> None:
    rc = _lib.load().dipsb_synth_fill_device(device, d_dst, first_frame, n_frames, width, height, fmt, seed, profile,
                                             stream)
    if rc != 0:
        raise DipsError(rc, _lib.load().dipsb_last_error(None).decode())


REDUCE_AUTO, REDUCE_P2P, REDUCE_NCCL = 0, 1, 2
UNIQUE_ID_BYTES = 128


def shard_range(total_frames: int, nranks: int, rank: int):
    """(first, count): the contiguous frames of a total_frames clip that `rank` of `nranks` owns (dipsb_shard_range)."""
    if not (0 <= rank < nranks):
        raise ValueError("rank outside the communicator")
    first, count = C.c_uint64(), C.c_uint64()
    _lib.load().dipsb_shard_range(total_frames, nranks, rank, C.byref(first), C.byref(count))
    return int(first.value), int(count.value)


def xchg_plan_query(total_frames: int, nranks: int, n_elems: int = 0) -> dict:
    """host-only: the exchange format of the peer-memory accumulator reduce for such a clip"""
    out = (C.c_uint64 * 4)()
    rc = _lib.load().dipsb_xchg_plan_query(total_frames, nranks, n_elems, C.byref(out))
    if rc != 0:
        raise DipsError(rc, "xchg_plan_query: invalid arguments")
    return dict(bytes_per_element=int(out[0]), sum_bits=int(out[1]), frames_per_rank_bound=int(out[2]), owned_elements=int(out[3]))


def comm_unique_id() -> bytes:
    """128 bytes rank 0 hands to every rank of a new communicator (ncclGetUniqueId inside the library)"""
    buf = C.create_string_buffer(UNIQUE_ID_BYTES)
    rc = _lib.load().dipsb_comm_unique_id(buf)
    if rc != 0:
        raise DipsError(rc, _lib.load().dipsb_last_error(None).decode())
    return bytes(buf.raw)


FRAME_COUNT = 2      # dips_alt/src/lib.rs:36


class SnapshotSchedule:
    """When the caller loops of dips_alt ask for a snapshot: on the frame where index == FRAME_COUNT (the third frame, and
    the third after every refresh marker); index saturates one past it and a marker -- a 1-based count of frames processed
    so far -- resets it (dips_alt/src/lib.rs:222-232 live mode, :560-561 and :662-670 file mode)."""

    def __init__(self, refresh_markers=()):
        self.refresh_markers = set(int(m) for m in refresh_markers)
        self.index = 0
        self.overall_frame = 0

    def snapshot_now(self) -> bool:
        """Ask before sending the frame."""
        return self.index == FRAME_COUNT

    def frame_sent(self) -> None:
        """Call after sending it."""
        if self.index <= FRAME_COUNT:
            self.index += 1
        self.overall_frame += 1
        if self.overall_frame in self.refresh_markers:
            self.index = 0


def run_dips_on_frames(ctx: "Context", frames, refresh_markers=(), sink=None):
    """The compute part of run_dips_on_file (dips_alt/src/lib.rs:553-690) without the OpenCV capture / writer around it:
    every frame goes through push_frame on a FLAVOR_ALT_RING2[_MEDIAN] context with the reference's snapshot schedule.
    Returns the list of difference frames, or hands each (t, rgba) to `sink`."""
    schedule = SnapshotSchedule(refresh_markers)
    outs = []
    for t, frame in enumerate(frames):
        if schedule.snapshot_now():
            ctx.snapshot()
        _, rgba, _ = ctx.push_frame(frame)
        schedule.frame_sent()
        if sink is None:
            outs.append(rgba)
        else:
            sink(t, rgba)
    return outs


class PinnedBuffer:
    """A page-locked host buffer from dipsb_host_alloc, exposed as a numpy uint8 array (`.array`).  The frame calls
    recognise it and skip their staging copy.  Free it with close() (or a `with` block) once no array view is in use."""

    def __init__(self, nbytes: int, device: int = 0):
        self._lib = _lib.load()
        p = C.c_void_p()
        rc = self._lib.dipsb_host_alloc(device, nbytes, C.byref(p))
        if rc != 0:
            raise DipsError(rc, self._lib.dipsb_last_error(None).decode())
        self._p, self.nbytes = p, nbytes
        self.array = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(p.value))

    def close(self) -> None:
        if self._p is not None:
            self.array = None
            self._lib.dipsb_host_free(self._p)
            self._p = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RegisteredBuffer:
    """Page-locks a numpy array the caller owns, in place (dipsb_host_register), for as long as the object lives: frames
    inside it take the direct copy-engine path of the frame calls.  close() (or a `with` block) unlocks it."""

    def __init__(self, array: np.ndarray, device: int = 0):
        if not array.flags["C_CONTIGUOUS"]:
            raise ValueError("RegisteredBuffer needs a C-contiguous array")
        self._lib = _lib.load()
        self.array = array
        self._p = C.c_void_p(array.ctypes.data)
        rc = self._lib.dipsb_host_register(device, self._p, array.nbytes)
        if rc != 0:
            self._p = None
            raise DipsError(rc, self._lib.dipsb_last_error(None).decode())

    def close(self) -> None:
        if self._p is not None:
            self._lib.dipsb_host_unregister(self._p)
            self._p = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _make_config(lib, width, height, fmt, mode, threshold, chroma, device, colorize, filt, sigmoid_scalar, spatial_window,
                 flavor):
    cfg = _lib.Config()
    lib.dipsb_default_config(C.byref(cfg))
    cfg.device, cfg.width, cfg.height, cfg.format, cfg.mode = device, width, height, fmt, mode
    cfg.chroma, cfg.threshold, cfg.colorize, cfg.filter = chroma, threshold, int(colorize), filt
    cfg.sigmoid_scalar, cfg.spatial_window, cfg.flavor = sigmoid_scalar, spatial_window, flavor
    return cfg


def _host_ptr(a: np.ndarray) -> int:
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("array must be C-contiguous")
    return a.ctypes.data


class Context:
    """Owner of one dipsb_ctx.  Mirrors the reference's ComputeState / DiPsCompute lifetime (create once per clip
    geometry and property set; dips/src/gpu/mod.rs:59, dips_alt/src/dips_compute/mod.rs:270)."""

    def __init__(self, width: int, height: int, fmt: int = FMT_RGBX8, mode: int = MODE_OVERALL, threshold: int = 0,
                 chroma: int = CHROMA_NONE, device: int = 0, colorize: bool = False, filt: int = FILTER_NONE,
                 sigmoid_scalar: float = 5.0, spatial_window: int = 1, flavor: int = FLAVOR_FRAME0, _borrowed=None):
        self._lib = _lib.load()
        self._owned = _borrowed is None
        if _borrowed is None:
            cfg = _make_config(self._lib, width, height, fmt, mode, threshold, chroma, device, colorize, filt,
                               sigmoid_scalar, spatial_window, flavor)
            h = C.c_void_p()
            rc = self._lib.dipsb_create(C.byref(cfg), C.byref(h))
            if rc != 0:
                raise DipsError(rc, self._lib.dipsb_last_error(None).decode())
        else:
            h = C.c_void_p(_borrowed)      # a rank of a Group: the group owns the handle
        self._h = h
        self.width, self.height, self.fmt, self.mode = width, height, fmt, mode
        self.npx = width * height
        self.bpp = bytes_per_pixel(fmt)
        self.frame_bytes = self.npx * self.bpp

    # -- plumbing ---------------------------------------------------------------------------------------------
    def _ck(self, rc: int) -> int:
        if rc < 0:
            raise DipsError(rc, self._lib.dipsb_last_error(self._h).decode())
        return rc

    def close(self) -> None:
        if getattr(self, "_h", None):
            if self._owned:
                self._lib.dipsb_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def reset(self) -> None:
        self._ck(self._lib.dipsb_reset(self._h))

    def set_threshold(self, tau: int) -> None:
        self._ck(self._lib.dipsb_set_threshold(self._h, tau))

    def set_stream(self, stream: int) -> None:
        self._ck(self._lib.dipsb_set_stream(self._h, stream))

    def adopt_stream(self, stream: int) -> None:
        """switch streams without ordering (the caller orders the streams with its own events)"""
        self._ck(self._lib.dipsb_adopt_stream(self._h, stream))

    def use_private_stream(self) -> None:
        self._ck(self._lib.dipsb_use_private_stream(self._h))

    def synchronize(self) -> None:
        self._ck(self._lib.dipsb_synchronize(self._h))

    def set_tuning(self, stages: int = 0, tile_px: int = 0, segments: int = 0, regs: int = 0) -> None:
        self._ck(self._lib.dipsb_set_tuning(self._h, stages, tile_px, segments, regs))

    def set_kernel(self, kernel: int) -> None:
        """0 = clip_kernel, 1 = clip_kernel_ws (producer warp, stage-unrolled)"""
        self._ck(self._lib.dipsb_set_kernel(self._h, kernel))

    def enable_timing(self, on: bool = True) -> None:
        self._ck(self._lib.dipsb_enable_timing(self._h, int(on)))

    def clip_kernel_time(self):
        """(total milliseconds, launches) of the clip kernel since the last call; synchronises."""
        ms, n = C.c_double(), C.c_uint64()
        self._ck(self._lib.dipsb_clip_kernel_time(self._h, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def stream_probe(self, d_frames: int, n_frames: int, stride: int | None = None, reps: int = 5) -> float:
        """mean milliseconds of the compute-free TMA streaming probe over the clip (bandwidth ceiling of the pattern)"""
        ms = C.c_float()
        self._ck(self._lib.dipsb_stream_probe(self._h, d_frames, n_frames, stride or self.frame_bytes, reps, C.byref(ms)))
        return float(ms.value)

    def last_plan(self) -> dict:
        out = (C.c_uint32 * 8)()
        self._ck(self._lib.dipsb_last_plan(self._h, C.byref(out)))
        return dict(tiles=out[0], segments=out[1], threads=out[2], stages=out[3] & 0xFFFF, kernel=out[3] >> 16, blocks_per_sm=out[4],
                    tile_px=out[5], smem_bytes=out[6] & 0xFFFFFF, regs=out[6] >> 24, tma_path=out[7] == 1, ring_clip=out[7] == 2)

    # -- state plane ------------------------------------------------------------------------------------------
    def prime_device(self, d_frame: int) -> None:
        self._ck(self._lib.dipsb_prime_device(self._h, d_frame))

    def prime_median4_device(self, d_frames: int, stride: int) -> None:
        self._ck(self._lib.dipsb_prime_median4_device(self._h, d_frames, stride))

    def prime_host(self, frame: np.ndarray) -> None:
        frame = np.ascontiguousarray(frame, dtype=np.uint8)
        if frame.size != self.frame_bytes:
            raise ValueError("frame size mismatch")
        self._ck(self._lib.dipsb_prime_host(self._h, _host_ptr(frame)))

    def state_plane_device(self) -> int:
        p = C.c_void_p()
        self._ck(self._lib.dipsb_state_plane_device(self._h, C.byref(p)))
        return int(p.value)

    def mark_state_valid(self, valid: bool = True) -> None:
        self._ck(self._lib.dipsb_mark_state_valid(self._h, int(valid)))

    def get_state_plane(self) -> np.ndarray:
        out = np.empty(self.npx, np.uint16)
        self._ck(self._lib.dipsb_get_state_plane(self._h, _host_ptr(out)))
        return out

    # -- batch ------------------------------------------------------------------------------------------------
    def run_clip_device(self, d_frames: int, n_frames: int, stride: int | None = None, first_frame: int = 0) -> None:
        self._ck(self._lib.dipsb_run_clip_device(self._h, d_frames, n_frames, stride or self.frame_bytes, first_frame))

    def run_clip_host(self, frames, n_frames: int | None = None, stride: int | None = None, first_frame: int = 0) -> None:
        """frames: numpy uint8 array [n, frame_bytes] or a raw host address (then n_frames is required)."""
        if isinstance(frames, np.ndarray):
            if frames.dtype != np.uint8 or not frames.flags["C_CONTIGUOUS"]:
                raise ValueError("frames must be a C-contiguous uint8 array")
            n = frames.shape[0] if n_frames is None else n_frames
            st = stride or (frames.strides[0] if frames.ndim > 1 else self.frame_bytes)
            ptr = _host_ptr(frames)
        else:
            ptr, n, st = int(frames), int(n_frames), stride or self.frame_bytes
        self._ck(self._lib.dipsb_run_clip_host(self._h, ptr, n, st, first_frame))

    # -- several GPUs: this context as one rank of a communicator ----------------------------------------------
    def comm_init_rank(self, nranks: int, rank: int, unique_id: bytes) -> None:
        """collective over all ranks (ncclCommInitRank + mapping of the peers' windows)"""
        if len(unique_id) != UNIQUE_ID_BYTES:
            raise ValueError("unique_id must be the 128 bytes of comm_unique_id()")
        self._ck(self._lib.dipsb_comm_init_rank(self._h, nranks, rank, C.create_string_buffer(unique_id, UNIQUE_ID_BYTES)))

    def comm_destroy(self) -> None:
        self._ck(self._lib.dipsb_comm_destroy(self._h))

    def comm_info(self) -> dict:
        out = (C.c_uint32 * 8)()
        self._ck(self._lib.dipsb_comm_info(self._h, C.byref(out)))
        return dict(nranks=out[0], rank=out[1], peer_memory=bool(out[2]), nccl_version=out[3],
                    reduce_path={REDUCE_P2P: "p2p", REDUCE_NCCL: "nccl"}.get(out[4], "none"), single_process=bool(out[5]),
                    acc_sharded=bool(out[6]), nccl=bool(out[7]))

    def comm_set_reduce(self, path: int) -> None:
        self._ck(self._lib.dipsb_comm_set_reduce(self._h, path))

    def comm_check(self) -> None:
        self._ck(self._lib.dipsb_comm_check(self._h))

    def run_clip_sharded_device(self, d_frames: int, n_frames: int, first_frame: int, total_frames: int,
                                stride: int | None = None) -> None:
        self._ck(self._lib.dipsb_run_clip_sharded_device(self._h, d_frames, n_frames, stride or self.frame_bytes,
                                                         first_frame, total_frames))

    def run_clip_sharded_host(self, frames, n_frames: int, first_frame: int, total_frames: int,
                              stride: int | None = None) -> None:
        ptr = _host_ptr(frames) if isinstance(frames, np.ndarray) else int(frames)
        self._ck(self._lib.dipsb_run_clip_sharded_host(self._h, ptr, n_frames, stride or self.frame_bytes, first_frame,
                                                       total_frames))

    def comm_phase_times(self):
        """((exchange_ms, pass_ms, reduce_ms) summed, passes) of the sharded passes since the last call (enable_timing)"""
        ms, n = (C.c_double * 3)(), C.c_uint64()
        self._ck(self._lib.dipsb_comm_phase_times(self._h, C.byref(ms), C.byref(n)))
        return (float(ms[0]), float(ms[1]), float(ms[2])), int(n.value)

    def comm_probe(self, what: int, total_frames: int, reps: int = 10) -> float:
        """collective measurement aid: mean ms of back-to-back exchanges (0 reduce-scatter, 1 broadcast, 2 all-gather,
        3 NCCL all-reduce path); leaves the accumulators undefined"""
        ms = C.c_float()
        self._ck(self._lib.dipsb_comm_probe(self._h, what, total_frames, reps, C.byref(ms)))
        return float(ms.value)

    def gather_accumulators(self) -> None:
        """collective: complete the accumulator planes on every rank after a sharded pass"""
        self._ck(self._lib.dipsb_gather_accumulators(self._h))

    # -- streaming --------------------------------------------------------------------------------------------
    def _out_buffer(self, out):
        if out is None:
            return np.empty(self.npx * 4, np.uint8)
        if out.dtype != np.uint8 or out.size < self.npx * 4 or not out.flags["C_CONTIGUOUS"] or not out.flags["WRITEABLE"]:
            raise ValueError("out must be a writable C-contiguous uint8 array of width*height*4 bytes")
        return out

    def push_frame(self, frame: np.ndarray, fmt: int | None = None, want_rgba: bool = True, stride: int | None = None,
                   out: np.ndarray | None = None):
        """Returns (status, rgba or None, (frame_index, sad, count)).  status NOT_READY == passthrough frame.
        `out`: caller-owned RGBA8 buffer to fill (a PinnedBuffer's array avoids the staging copies)."""
        fmt = self.fmt if fmt is None else fmt
        frame = np.ascontiguousarray(frame, dtype=np.uint8)
        stride = stride or self.width * bytes_per_pixel(fmt)
        if frame.size < stride * self.height:
            raise ValueError("frame smaller than height*stride")
        out = self._out_buffer(out) if want_rgba else None
        st = _lib.FrameStats()
        rc = self._ck(self._lib.dipsb_push_frame(self._h, _host_ptr(frame), self.width, self.height, stride, fmt,
                                                 _host_ptr(out) if want_rgba else None, C.byref(st)))
        return rc, out, (int(st.frame_index), int(st.sad), int(st.count))

    def stage_frame(self, frame: np.ndarray, fmt: int | None = None, stride: int | None = None) -> None:
        """first half of push_frame (the reference's add_texture): copy the frame into the library's slot, start its upload"""
        fmt = self.fmt if fmt is None else fmt
        frame = np.ascontiguousarray(frame, dtype=np.uint8)
        stride = stride or self.width * bytes_per_pixel(fmt)
        if frame.size < stride * self.height:
            raise ValueError("frame smaller than height*stride")
        self._ck(self._lib.dipsb_stage_frame(self._h, _host_ptr(frame), self.width, self.height, stride, fmt))

    def dispatch_staged(self, want_rgba: bool = True, out: np.ndarray | None = None):
        """second half (the reference's dispatch): same return value as push_frame"""
        out = self._out_buffer(out) if want_rgba else None
        st = _lib.FrameStats()
        rc = self._ck(self._lib.dipsb_dispatch_staged(self._h, _host_ptr(out) if want_rgba else None, C.byref(st)))
        return rc, out, (int(st.frame_index), int(st.sad), int(st.count))

    def push_frame_pipelined(self, frame: np.ndarray, fmt: int | None = None, stride: int | None = None,
                             out: np.ndarray | None = None):
        """Submit `frame`, get back the previous frame's result: (status, rgba or None, stats or None).
        status: NOT_READY (first call), 0 (difference frame of t-1), 2 (t-1 was a passed-through warm-up frame)."""
        fmt = self.fmt if fmt is None else fmt
        frame = np.ascontiguousarray(frame, dtype=np.uint8)
        stride = stride or self.width * bytes_per_pixel(fmt)
        out = self._out_buffer(out)
        st = _lib.FrameStats()
        rc = self._ck(self._lib.dipsb_push_frame_pipelined(self._h, _host_ptr(frame), self.width, self.height, stride,
                                                           fmt, _host_ptr(out), C.byref(st)))
        if rc == NOT_READY:
            return rc, None, None
        return rc, out, (int(st.frame_index), int(st.sad), int(st.count))

    def flush_frame(self, out: np.ndarray | None = None):
        out = self._out_buffer(out)
        st = _lib.FrameStats()
        rc = self._ck(self._lib.dipsb_flush_frame(self._h, _host_ptr(out), C.byref(st)))
        if rc == NOT_READY:
            return rc, None, None
        return rc, out, (int(st.frame_index), int(st.sad), int(st.count))

    def snapshot(self) -> None:
        self._ck(self._lib.dipsb_snapshot(self._h))

    # -- results ----------------------------------------------------------------------------------------------
    @property
    def frames_processed(self) -> int:
        return int(self._lib.dipsb_frames_processed(self._h))

    def get_accumulators(self):
        s = np.empty(self.npx, np.uint32)
        c = np.empty(self.npx, np.uint32)
        self._ck(self._lib.dipsb_get_accumulators(self._h, _host_ptr(s), _host_ptr(c)))
        return s, c

    def get_accumulators_into(self, host_sum: int, host_cnt: int) -> None:
        """same, into caller-owned host memory given as raw addresses (e.g. pinned torch tensors' data_ptr())"""
        self._ck(self._lib.dipsb_get_accumulators(self._h, host_sum, host_cnt))

    def set_accumulators(self, acc_sum: np.ndarray, acc_cnt: np.ndarray) -> None:
        s = np.ascontiguousarray(acc_sum, dtype=np.uint32)
        c = np.ascontiguousarray(acc_cnt, dtype=np.uint32)
        if s.size != self.npx or c.size != self.npx:
            raise ValueError("accumulator size mismatch")
        self._ck(self._lib.dipsb_set_accumulators(self._h, _host_ptr(s), _host_ptr(c)))

    def accumulators_device(self):
        """(device address, n_elems): one u32[2*n_elems] buffer, sum plane then count plane, internal tile order."""
        p, n = C.c_void_p(), C.c_uint64()
        self._ck(self._lib.dipsb_accumulators_device(self._h, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def pack_accumulators_device(self, total_frames: int):
        """(device address, n_words) of the packed exchange buffer: all-reduce it as int32, then unpack_accumulators_device()"""
        p, n = C.c_void_p(), C.c_uint64()
        self._ck(self._lib.dipsb_pack_accumulators_device(self._h, total_frames, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def unpack_accumulators_device(self) -> None:
        self._ck(self._lib.dipsb_unpack_accumulators_device(self._h))

    def get_scalars(self, first: int, n: int):
        sad = np.empty(n, np.uint64)
        cnt = np.empty(n, np.uint64)
        self._ck(self._lib.dipsb_get_scalars(self._h, first, n, _host_ptr(sad), _host_ptr(cnt)))
        return sad, cnt

    def get_intensity_map(self, n_eff: int) -> np.ndarray:
        out = np.empty(self.npx, np.float32)
        self._ck(self._lib.dipsb_get_intensity_map(self._h, n_eff, _host_ptr(out)))
        return out

    def get_frame_means(self, first: int, n: int) -> np.ndarray:
        out = np.empty(n, np.float32)
        self._ck(self._lib.dipsb_get_frame_means(self._h, first, n, _host_ptr(out)))
        return out


class Group:
    """All GPUs of one process as the ranks of one clip (dipsb_create_group): one context per device, ncclCommInitAll and
    peer access inside the library.  devices=[0, 0, ...] gives a loopback group on one GPU (tests)."""

    def __init__(self, devices, width: int, height: int, fmt: int = FMT_RGBX8, mode: int = MODE_OVERALL, threshold: int = 0,
                 chroma: int = CHROMA_NONE):
        self._lib = _lib.load()
        devices = list(devices)
        cfg = _make_config(self._lib, width, height, fmt, mode, threshold, chroma, 0, False, FILTER_NONE, 5.0, 1, FLAVOR_FRAME0)
        arr = (C.c_int32 * len(devices))(*devices)
        h = C.c_void_p()
        rc = self._lib.dipsb_create_group(C.byref(cfg), len(devices), arr, C.byref(h))
        if rc != 0:
            raise DipsError(rc, self._lib.dipsb_group_last_error(None).decode())
        self._h = h
        self.devices, self.npx = devices, width * height
        self.frame_bytes = self.npx * bytes_per_pixel(fmt)
        self.ranks = [Context(width, height, fmt, mode, threshold, chroma, device=d,
                              _borrowed=self._lib.dipsb_group_ctx(h, i)) for i, d in enumerate(devices)]

    def _ck(self, rc: int) -> int:
        if rc < 0:
            raise DipsError(rc, self._lib.dipsb_group_last_error(self._h).decode())
        return rc

    def close(self) -> None:
        if getattr(self, "_h", None):
            for r in self.ranks:
                r.close()
            self._lib.dipsb_destroy_group(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self) -> None:
        self._ck(self._lib.dipsb_group_reset(self._h))

    def run_clip_device(self, d_frames, n_frames, stride: int | None = None) -> None:
        """d_frames[i] / n_frames[i]: device address and frame count of shard i (resident on device i), in clip order"""
        n = len(self.ranks)
        ptrs = (C.c_void_p * n)(*[int(p) for p in d_frames])
        cnts = (C.c_uint64 * n)(*[int(k) for k in n_frames])
        self._ck(self._lib.dipsb_group_run_clip_device(self._h, ptrs, cnts, stride or self.frame_bytes))

    def gather_accumulators(self) -> None:
        self._ck(self._lib.dipsb_group_gather_accumulators(self._h))

    def synchronize(self) -> None:
        self._ck(self._lib.dipsb_group_synchronize(self._h))

    def get_accumulators(self):
        s = np.empty(self.npx, np.uint32)
        c = np.empty(self.npx, np.uint32)
        self._ck(self._lib.dipsb_group_get_accumulators(self._h, _host_ptr(s), _host_ptr(c)))
        return s, c

    def get_scalars(self, first: int, n: int):
        sad = np.empty(n, np.uint64)
        cnt = np.empty(n, np.uint64)
        self._ck(self._lib.dipsb_group_get_scalars(self._h, first, n, _host_ptr(sad), _host_ptr(cnt)))
        return sad, cnt
