"""CPU tests of the drop-in boundary: the C-ABI library builds, loads, exports every symbol include/dips_b200.h declares,
and fails loudly (no fallback) without a GPU.  No compute is called here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dips_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dipsb_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from dips_b200 import _build, _lib
    so = _build.build()
    assert os.path.exists(so)
    lib = ctypes.CDLL(so)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dips_b200.h but not exported"
    assert set(names) == set(_lib.SYMBOLS), set(names) ^ set(_lib.SYMBOLS)
    assert _lib.load().dipsb_abi_version() == 2


def test_config_struct_layout():
    from dips_b200 import _lib
    cfg = _lib.Config()
    _lib.load().dipsb_default_config(ctypes.byref(cfg))
    assert cfg.struct_size == ctypes.sizeof(_lib.Config) == 64
    assert cfg.filter == 255 and cfg.spatial_window == 1 and abs(cfg.sigmoid_scalar - 5.0) < 1e-9


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import dips_b200
    with pytest.raises(dips_b200.DipsError) as e:
        dips_b200.Context(64, 48)
    assert "no CUDA device" in str(e.value) and "no CPU fallback" in str(e.value)


def test_invalid_arguments_are_rejected_before_touching_the_device():
    from dips_b200 import _lib
    L = _lib.load()
    h = ctypes.c_void_p()
    cfg = _lib.Config()
    L.dipsb_default_config(ctypes.byref(cfg))
    cfg.width, cfg.height = 0, 10
    assert L.dipsb_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b"empty frame" in L.dipsb_last_error(None)
    cfg.width, cfg.height, cfg.format = 8, 8, 9
    assert L.dipsb_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    cfg.format, cfg.spatial_window = 0, 4
    assert L.dipsb_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b"spatial_window" in L.dipsb_last_error(None)
    cfg.spatial_window, cfg.struct_size = 1, 12
    assert L.dipsb_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert L.dipsb_reset(None) == -1 and L.dipsb_frames_processed(None) == 0


def test_product_does_not_import_oracle():
    """The product path must not route through the oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "dips_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.replace("the oracle's generator", "").replace("as the oracle does", "").replace("oracle/dips_oracle.c", ""), f


def test_planner_fills_the_sms_evenly():
    """dipsb_plan_query is host-only: for the BASELINE geometries the tiles must fill 148 SMs in (nearly) whole waves."""
    from dips_b200 import _lib
    L = _lib.load()
    for (w, h, fmt) in [(1920, 1080, 0), (3840, 2160, 1), (7680, 4320, 0), (640, 480, 0), (1, 1, 0), (37, 5, 0), (8192, 8192, 1)]:
        out = (ctypes.c_uint32 * 8)()
        assert L.dipsb_plan_query(w, h, fmt, 148, ctypes.byref(out)) == 0
        tiles, threads, stages, occ, tile_px = out[0], out[2], out[3] & 0xFFFF, out[4], out[5]
        threads += 32 * (out[3] >> 16)             # clip_kernel_ws adds a producer warp to the block
        smem, regs = out[6] & 0xFFFFFF, out[6] >> 24
        npx = w * h
        assert threads % 32 == 0 and 32 <= threads <= 1024 and tile_px % 16 == 0 and tile_px <= 16 * threads
        assert tiles == -(-npx // tile_px) and 2 <= stages <= 8 and occ >= 1 and regs in (64, 72, 80, 96)
        assert occ * threads * regs <= 65536 and occ * (smem + 1024) <= 227 * 1024
        resident = occ * 148
        if tiles >= resident:                      # large frames: whole waves, < 3 % of the last wave idle
            waves = -(-tiles // resident)
            assert tiles / (waves * resident) > 0.97, (w, h, tiles, resident)
    out = (ctypes.c_uint32 * 8)()
    assert L.dipsb_plan_query(0, 10, 0, 148, ctypes.byref(out)) == -1


def test_rust_sys_crate_declares_every_symbol():
    """rust/dips_b200_sys (shipped as source: no Rust toolchain here) must stay 1:1 with include/dips_b200.h."""
    text = open(os.path.join(ROOT, "rust", "dips_b200_sys", "src", "lib.rs")).read()
    rust = set(re.findall(r"pub fn (dipsb_[a-z0-9_]+)\s*\(", text))
    assert rust == set(declared_symbols()), rust ^ set(declared_symbols())


def test_header_is_valid_c_and_links(tmp_path):
    """include/dips_b200.h compiles as C11 and a C program can call the library (what any FFI relies on)."""
    import subprocess
    from dips_b200 import _build
    so_dir = os.path.dirname(_build.build())
    exe = str(tmp_path / "c_abi_smoke")
    subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-o", exe, "-L", so_dir, "-ldips_b200",
                           "-Wl,-rpath," + so_dir])
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "plan:" in res.stdout


def test_host_copy_pool_is_exact_and_thread_safe():
    """The staging copy (host only): contiguous, pitched and small copies are byte-exact, also from concurrent callers."""
    import threading
    import numpy as np
    from dips_b200 import _lib
    lib = _lib.load()
    assert 1 <= lib.dipsb_host_copy_threads() <= 64
    rng = np.random.default_rng(5)

    def one(rows, row_bytes, spitch, dpitch, seed):
        r = np.random.default_rng(seed)
        src = r.integers(0, 256, rows * spitch, dtype=np.uint8)
        dst = np.full(rows * dpitch, 0xEE, np.uint8)
        assert lib.dipsb_host_copy2d(dst.ctypes.data, dpitch, src.ctypes.data, spitch, row_bytes, rows) == 0
        s2, d2 = src.reshape(rows, spitch), dst.reshape(rows, dpitch)
        assert np.array_equal(d2[:, :row_bytes], s2[:, :row_bytes])
        assert (d2[:, row_bytes:] == 0xEE).all()           # padding untouched

    one(1, 100, 100, 100, 1)                               # small: caller only
    one(1, 8_294_400, 8_294_400, 8_294_400, 2)             # one 1080p RGBA frame, contiguous
    one(1080, 7680, 7680 + 64, 7680, 3)                    # pitched source
    one(7, 1_000_003, 1_000_003, 1_000_003 + 13, 4)        # odd sizes, pitched destination
    for _ in range(4):
        rows = int(rng.integers(1, 40)); rb = int(rng.integers(1, 300_000))
        one(rows, rb, rb + int(rng.integers(0, 100)), rb + int(rng.integers(0, 100)), int(rng.integers(1 << 30)))
    errs = []

    def worker(k):
        try:
            for i in range(5):
                one(3, 1_500_000 + k, 1_500_000 + k, 1_500_000 + k + 8, 100 * k + i)
        except BaseException as e:                          # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    assert lib.dipsb_host_copy2d(None, 8, None, 8, 8, 1) == -1
    src = np.zeros(64, np.uint8)
    assert lib.dipsb_host_copy2d(src.ctypes.data, 4, src.ctypes.data, 8, 8, 2) == -1       # pitch < row
    assert lib.dipsb_host_copy2d(None, 0, None, 0, 0, 0) == 0                              # empty copy


def test_tools_and_entry_points_compile():
    """Every helper script that travels to the GPU box at least parses (they only run there)."""
    import glob
    import py_compile
    files = glob.glob(os.path.join(ROOT, "tools", "*.py")) + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    assert len(files) >= 6
    for f in files:
        py_compile.compile(f, doraise=True)


def test_planner_invariants_on_random_geometries():
    """Host-only planner, 400 random frame sizes and SM counts: the plan always covers the frame, fits a block, fits the
    SM's registers and shared memory, and the tile-order index (host restatement below) is a bijection onto the
    accumulator elements for a sample of them."""
    import numpy as np
    from dips_b200 import _lib
    L = _lib.load()
    rng = np.random.default_rng(20261018)
    for i in range(400):
        w = int(rng.integers(1, 8193)) if i % 3 else int(rng.integers(1, 513)) * 16
        h = int(rng.integers(1, 4321)) if i % 5 else int(rng.integers(1, 65))
        fmt = int(rng.integers(0, 4))
        sms = int(rng.choice([148, 148, 148, 132, 74, 8]))
        out = (ctypes.c_uint32 * 8)()
        assert L.dipsb_plan_query(w, h, fmt, sms, ctypes.byref(out)) == 0, (w, h, fmt, sms)
        tiles, threads, stages, occ, tile_px = out[0], out[2], out[3] & 0xFFFF, out[4], out[5]
        block = threads + 32 * (out[3] >> 16)
        smem, regs = out[6] & 0xFFFFFF, out[6] >> 24
        npx = w * h
        assert tile_px % 16 == 0 and tile_px >= 16 and tiles * tile_px >= npx > (tiles - 1) * tile_px
        assert threads % 32 == 0 and 32 <= block <= 1024 and tile_px <= 16 * threads
        assert 2 <= stages <= 8 and occ >= 1 and regs in (64, 72)
        assert occ * block * regs <= 65536 and occ * (smem + 1024) <= 227 * 1024
        assert out[7] >= 1 and out[7] * 32 <= threads          # active warps
        if i % 20 == 0 and npx <= 1 << 18:                      # the accumulator order is a bijection
            bpp = 3 if fmt in (0, 2) else 4
            p = np.arange(npx, dtype=np.int64)
            tile, q = p // tile_px, p % tile_px
            if bpp == 3:
                grp = q // 16
                thread, k = grp % threads, (grp // threads) * 16 + q % 16
            else:
                quad = q // 4
                thread, k = quad % threads, 4 * (quad // threads) + q % 4
            idx = tile * (threads * 16) + k * threads + thread
            assert k.max() < 16 and idx.max() < tiles * threads * 16 and np.unique(idx).size == npx
