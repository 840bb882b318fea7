"""The caller-side snapshot schedule of dips_alt (N1): index / refresh-marker logic of dips_alt/src/lib.rs:222-232,
:560-561, :662-670, restated by hand here as known answers and checked against the Python and C++ host mirrors."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# (n_frames, refresh markers) -> 0-based frames sent with snapshot = Some(()), derived by hand from the reference loop:
# snapshot when index == 2; index += 1 while index <= 2; overall_frame += 1; a marker equal to overall_frame resets index
KATS = [
    (6, [], [2]),
    (2, [], []),
    (12, [5], [2, 7]),
    (8, [1], [3]),               # reset after the first frame: indices 0,1,2 fall on frames 1,2,3
    (9, [3], [2, 5]),            # marker right on the snapshot frame
    (12, [4, 5], [2, 7]),        # a second marker inside the warm-up restarts it
    (10, [100], [2]),            # marker beyond the clip
    (12, [6, 9], [2, 8, 11]),
]


def reference_loop(n, markers):
    """Transliteration of the loop body's bookkeeping (dips_alt/src/lib.rs:636-670)."""
    index, overall, snaps = 0, 0, []
    for t in range(n):
        if index == 2:
            snaps.append(t)
        if index <= 2:
            index += 1
        overall += 1
        if overall in markers:
            index = 0
    return snaps


@pytest.mark.parametrize("n,markers,want", KATS)
def test_snapshot_schedule_known_answers(n, markers, want):
    import dips_b200
    assert reference_loop(n, markers) == want
    s = dips_b200.SnapshotSchedule(markers)
    got = []
    for t in range(n):
        if s.snapshot_now():
            got.append(t)
        s.frame_sent()
    assert got == want and s.overall_frame == n


def test_cpp_snapshot_schedule_matches(tmp_path):
    src = tmp_path / "sched.cpp"
    cases = "".join("{%d, {%s}}," % (n, ",".join(map(str, m))) for n, m, _ in KATS)
    src.write_text('#include "dips_host.hpp"\n#include <cstdio>\nint main() {\n'
                   "  struct C { size_t n; std::vector<size_t> m; };\n  std::vector<C> cases = {" + cases + "};\n"
                   "  for (auto& c : cases) { dips_alt::SnapshotSchedule s(c.m);\n"
                   '    for (size_t t = 0; t < c.n; ++t) { if (s.snapshot_now()) printf("%zu ", t); s.frame_sent(); }\n'
                   '    printf("\\n"); }\n  return 0;\n}\n')
    exe = tmp_path / "sched"
    so_dir = os.path.join(ROOT, "dips_b200")
    from dips_b200 import _build
    _build.build()
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(so_dir, "host"),
                    str(src), "-o", str(exe), "-L", so_dir, "-ldips_b200", "-Wl,-rpath," + so_dir], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines()
    got = [[int(x) for x in line.split()] for line in out]
    assert got == [w for _, _, w in KATS]


@pytest.mark.gpu
@pytest.mark.parametrize("intended", [False, True])
def test_run_dips_on_frames_matches_reference_loop(oracle, intended):
    """The whole file-mode compute loop: schedule + ring-of-2 state machine, against the oracle driven by the
    transliterated reference loop (same tolerance as tests/test_gpu_stream.py's ring tests)."""
    import dips_b200
    w, h, n, markers = 80, 44, 14, [6, 9]
    clip = oracle.synth_clip(n, w, h, oracle.FMT_RGBX8, profile=oracle.SYNTH_SCENE)
    snaps = set(reference_loop(n, markers))
    ref = oracle.DiPsComputeOracle(w, h, True, 0, 5.0, 0, intended)
    want = [ref.send_frame(clip[t], t in snaps) for t in range(n)]
    flavor = dips_b200.FLAVOR_ALT_RING2_MEDIAN if intended else dips_b200.FLAVOR_ALT_RING2
    with dips_b200.Context(w, h, dips_b200.FMT_RGBX8, 0, 0, colorize=True, filt=0, flavor=flavor) as ctx:
        got = dips_b200.run_dips_on_frames(ctx, clip, markers)
    assert len(got) == n
    for t in range(n):
        d = np.abs(got[t].astype(int) - want[t].astype(int))
        assert d.max() <= 3 and (d <= 1).mean() >= 0.97, (t, d.max())
