"""ctypes binding of the CPU oracle (oracle/libdips_oracle.so).

TEST INFRASTRUCTURE ONLY -- may be imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py, never by the product package dips_b200.
PARITY UNPINNED: see oracle/dips_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdips_oracle.so")

FMT_RGB8, FMT_RGBX8, FMT_BGR8, FMT_BGRX8 = 0, 1, 2, 3
CHROMA_NONE, CHROMA_RED, CHROMA_GREEN, CHROMA_BLUE = 0, 1, 2, 3
MODE_OVERALL, MODE_PERFRAME = 0, 1
FILTER_SIGMOID, FILTER_INV_SIGMOID, FILTER_NONE = 0, 1, 255
SYNTH_UNIFORM, SYNTH_SCENE = 0, 1


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("dips_oracle.c", "dips_oracle.h", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        u8p, u16p, u32p, u64p, f32p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint16, C.c_uint32, C.c_uint64, C.c_float))
        L.dipso_bytes_per_pixel.argtypes = [C.c_int]
        L.dipso_intensity2.argtypes = [u8p, C.c_int, C.c_int]
        L.dipso_intensity2.restype = C.c_uint16
        L.dipso_i2_plane.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
        L.dipso_i2_plane.restype = None
        L.dipso_median4_plane.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_int, C.c_int, C.c_void_p]
        L.dipso_median4_plane.restype = None
        L.dipso_run_clip.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                     C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.dipso_run_clip.restype = None
        L.dipso_spatial_median_plane.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
        L.dipso_spatial_median_plane.restype = None
        L.dipso_intensity_map.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p]
        L.dipso_intensity_map.restype = None
        L.dipso_frame_means.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
        L.dipso_frame_means.restype = None
        L.dipso_visual_pixel.argtypes = [C.c_int32, C.c_int, C.c_int, C.c_float, u8p]
        L.dipso_visual_pixel.restype = None
        L.dipso_visual_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_float, C.c_void_p]
        L.dipso_visual_frame.restype = None
        L.dipso_visual_diff.argtypes = [C.c_int32, C.c_int, C.c_float]
        L.dipso_visual_diff.restype = C.c_float
        L.dipso_mix64.argtypes = [C.c_uint64]
        L.dipso_mix64.restype = C.c_uint64
        L.dipso_synth_fill.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int,
                                       C.c_uint64, C.c_int, C.c_int]
        L.dipso_synth_fill.restype = None
        L.dipso_cs_new.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_float, C.c_int]
        L.dipso_cs_new.restype = C.c_void_p
        L.dipso_cs_free.argtypes = [C.c_void_p]
        L.dipso_cs_free.restype = None
        L.dipso_cs_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.dipso_cs_frame.restype = C.c_int
        L.dipso_alt_new.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int]
        L.dipso_alt_new.restype = C.c_void_p
        L.dipso_alt_free.argtypes = [C.c_void_p]
        L.dipso_alt_free.restype = None
        L.dipso_alt_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.dipso_alt_frame.restype = None
        L.dipso_num_threads.restype = C.c_int
        _lib = L
    return _lib


def bpp(fmt: int) -> int:
    return 3 if fmt in (FMT_RGB8, FMT_BGR8) else 4


def _ptr(a: np.ndarray) -> int:
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def num_threads() -> int:
    return int(lib().dipso_num_threads())


def synth_clip(n_frames, width, height, fmt, seed=0x44695073, profile=SYNTH_SCENE, first_frame=0, nthreads=0):
    """uint8 array [n_frames, height*width*bpp] of the deterministic synthetic clip."""
    out = np.empty((n_frames, width * height * bpp(fmt)), dtype=np.uint8)
    lib().dipso_synth_fill(_ptr(out), first_frame, n_frames, width, height, fmt, seed, profile, nthreads)
    return out


def i2_plane(frame: np.ndarray, fmt: int, chroma: int = CHROMA_NONE) -> np.ndarray:
    frame = np.ascontiguousarray(frame, dtype=np.uint8).reshape(-1)
    npx = frame.size // bpp(fmt)
    out = np.empty(npx, dtype=np.uint16)
    lib().dipso_i2_plane(_ptr(frame), npx, fmt, chroma, _ptr(out))
    return out


def median4_plane(frames4: np.ndarray, fmt: int, chroma: int = CHROMA_NONE) -> np.ndarray:
    frames4 = np.ascontiguousarray(frames4, dtype=np.uint8)
    assert frames4.shape[0] == 4
    npx = frames4[0].size // bpp(fmt)
    ptrs = (C.c_void_p * 4)(*[frames4[k].ctypes.data for k in range(4)])
    out = np.empty(npx, dtype=np.uint16)
    lib().dipso_median4_plane(ptrs, npx, fmt, chroma, _ptr(out))
    return out


def spatial_median(i2: np.ndarray, width: int, height: int, window: int) -> np.ndarray:
    """correct zero-padded median of the window x window neighbourhood of an I2 plane (N4)"""
    i2 = np.ascontiguousarray(i2, dtype=np.uint16).reshape(-1)
    assert i2.size == width * height
    out = np.empty_like(i2)
    lib().dipso_spatial_median_plane(_ptr(i2), width, height, window, _ptr(out))
    return out


def filtered_i2(frame: np.ndarray, width: int, height: int, fmt: int, window: int, chroma: int = CHROMA_NONE) -> np.ndarray:
    return spatial_median(i2_plane(frame, fmt, chroma), width, height, window)


class ClipResult:
    __slots__ = ("acc_sum", "acc_cnt", "sad", "cnt", "state")

    def __init__(self, acc_sum, acc_cnt, sad, cnt, state):
        self.acc_sum, self.acc_cnt, self.sad, self.cnt, self.state = acc_sum, acc_cnt, sad, cnt, state


def run_clip(frames: np.ndarray, fmt: int, mode: int, tau: int, chroma: int = CHROMA_NONE, state=None,
             acc_sum=None, acc_cnt=None, nthreads: int = 0) -> ClipResult:
    """frames: uint8 [n, frame_bytes].  state None -> I2 of frame 0 (ref in overall mode; D_0 = 0 in per-frame)."""
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    n, fb = frames.shape
    npx = fb // bpp(fmt)
    if state is None:
        state = i2_plane(frames[0], fmt, chroma)
    else:
        state = np.array(state, dtype=np.uint16, copy=True)
    acc_sum = np.zeros(npx, np.uint32) if acc_sum is None else np.array(acc_sum, dtype=np.uint32, copy=True)
    acc_cnt = np.zeros(npx, np.uint32) if acc_cnt is None else np.array(acc_cnt, dtype=np.uint32, copy=True)
    sad = np.zeros(n, np.uint64)
    cnt = np.zeros(n, np.uint64)
    lib().dipso_run_clip(_ptr(frames), n, fb, npx, fmt, chroma, mode, tau, _ptr(state), _ptr(acc_sum),
                         _ptr(acc_cnt), _ptr(sad), _ptr(cnt), nthreads)
    return ClipResult(acc_sum, acc_cnt, sad, cnt, state)


def intensity_map(acc_sum: np.ndarray, n_eff: int) -> np.ndarray:
    out = np.empty(acc_sum.size, np.float32)
    lib().dipso_intensity_map(_ptr(np.ascontiguousarray(acc_sum, dtype=np.uint32)), acc_sum.size, n_eff, _ptr(out))
    return out


def frame_means(sad: np.ndarray, npx: int) -> np.ndarray:
    out = np.empty(sad.size, np.float32)
    lib().dipso_frame_means(_ptr(np.ascontiguousarray(sad, dtype=np.uint64)), sad.size, npx, _ptr(out))
    return out


def visual_pixel(s_i2: int, colorize: bool, filt: int, sig_scalar: float = 5.0):
    out = (C.c_uint8 * 4)()
    lib().dipso_visual_pixel(int(s_i2), int(colorize), filt, sig_scalar, out)
    return tuple(out)


def visual_diff(s_i2: int, filt: int, sig_scalar: float = 5.0) -> float:
    return float(lib().dipso_visual_diff(int(s_i2), filt, sig_scalar))


def visual_frame(ref: np.ndarray, cur: np.ndarray, colorize: bool, filt: int, sig_scalar: float = 5.0) -> np.ndarray:
    ref = np.ascontiguousarray(ref, dtype=np.uint16)
    cur = np.ascontiguousarray(cur, dtype=np.uint16)
    out = np.empty(ref.size * 4, np.uint8)
    lib().dipso_visual_frame(_ptr(ref), _ptr(cur), ref.size, int(colorize), filt, sig_scalar, _ptr(out))
    return out


class ComputeStateOracle:
    """Reference-flavour `dips` ComputeState (dips/src/gpu/mod.rs) on RGBA8 frames."""

    def __init__(self, width, height, colorize=False, filt=FILTER_NONE, sig_scalar=5.0, chroma=CHROMA_NONE):
        self.w, self.h = width, height
        self._h = lib().dipso_cs_new(width, height, int(colorize), filt, sig_scalar, chroma)

    def frame(self, rgba: np.ndarray):
        rgba = np.ascontiguousarray(rgba, dtype=np.uint8).reshape(-1)
        out = np.empty_like(rgba)
        passthrough = lib().dipso_cs_frame(self._h, _ptr(rgba), _ptr(out))
        return out, bool(passthrough)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().dipso_cs_free(self._h)
            self._h = None


class DiPsComputeOracle:
    """Reference-flavour `dips_alt` DiPsCompute (dips_alt/src/dips_compute/mod.rs) on RGBA8 frames."""

    def __init__(self, width, height, colorize=True, filt=FILTER_SIGMOID, sig_scalar=5.0, chroma=CHROMA_NONE,
                 intended_median=False):
        self.w, self.h = width, height
        self._h = lib().dipso_alt_new(width, height, int(colorize), filt, sig_scalar, chroma, int(intended_median))

    def send_frame(self, rgba: np.ndarray, snapshot: bool = False) -> np.ndarray:
        rgba = np.ascontiguousarray(rgba, dtype=np.uint8).reshape(-1)
        out = np.empty_like(rgba)
        lib().dipso_alt_frame(self._h, _ptr(rgba), int(snapshot), _ptr(out))
        return out

    def __del__(self):
        if getattr(self, "_h", None):
            lib().dipso_alt_free(self._h)
            self._h = None
