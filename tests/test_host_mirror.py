"""C++ host mirror of the reference API (dips_b200/host/dips_host.hpp): compiles on CPU, and on the GPU box runs the
reference's own call sequence (frame_callback per frame / send_frame with the snapshot rule) against the oracle's
reference-flavour state machines."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_driver(tmp_path):
    from dips_b200 import _build
    so_dir = os.path.dirname(_build.build())
    exe = str(tmp_path / "host_mirror_main")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "dips_b200", "host"), os.path.join(ROOT, "tests", "host_mirror_main.cpp"),
                           "-o", exe, "-L", so_dir, "-ldips_b200", "-Wl,-rpath," + so_dir])
    return exe


def test_host_mirror_compiles_and_links(tmp_path):
    exe = build_driver(tmp_path)
    assert subprocess.run([exe, "--compile-check"], capture_output=True, text=True).stdout.strip() == "ok"


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["dips", "alt"])
def test_host_mirror_matches_reference_state_machines(tmp_path, oracle, mode):
    exe = build_driver(tmp_path)
    w, h, n = 64, 40, 9
    clip = oracle.synth_clip(n, w, h, oracle.FMT_RGBX8, profile=oracle.SYNTH_SCENE)
    fin, fout = str(tmp_path / "in.rgba"), str(tmp_path / "out.rgba")
    clip.tofile(fin)
    res = subprocess.run([exe, mode, str(w), str(h), str(n), fin, fout], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    got = np.fromfile(fout, np.uint8).reshape(n, -1)
    if mode == "dips":
        ref = oracle.ComputeStateOracle(w, h, False, oracle.FILTER_NONE, 5.0, 0)
        want = [ref.frame(clip[t])[0] for t in range(n)]
    else:
        ref = oracle.DiPsComputeOracle(w, h, True, oracle.FILTER_SIGMOID, 5.0, 0, False)
        want, index = [], 0
        for t in range(n):
            want.append(ref.send_frame(clip[t], index == 2))
            if index <= 2:
                index += 1
    for t in range(n):
        d = np.abs(got[t].astype(int) - want[t].astype(int))
        assert d.max() <= 3 and (d <= 1).mean() >= 0.97, (t, d.max())
