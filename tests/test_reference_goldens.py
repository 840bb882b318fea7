"""Consumes outputs of the REAL reference (tests/golden/reference_out_<case>.bin, produced once on a machine with cargo and a
wgpu adapter by the recipe in tools/reference_goldens/README.md).  With them the oracle -- and through it every parity claim
of this project -- is pinned to the reference itself; without them these tests skip and parity stays "unpinned".

    CPU: the oracle's dipso_cs_* / dipso_alt_* state machines vs the reference, +-1 LSB
    GPU: dipsb_push_frame in the DIPS_RING4 / ALT_RING2 flavours vs the reference, <= 3 LSB, >= 97 % within 1 LSB
         (the library keeps the ring in exact integers where the shader rounds through rgba8unorm: DESIGN.md section 9)
"""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = json.load(open(os.path.join(GOLD, "reference_cases.json")))


def load_reference(case):
    path = os.path.join(GOLD, f"reference_out_{case['name']}.bin")
    if not os.path.exists(path):
        pytest.skip(f"{os.path.basename(path)} not committed yet: parity unpinned (tools/reference_goldens/README.md)")
    fb = case["width"] * case["height"] * 4
    raw = np.fromfile(path, np.uint8)
    assert raw.size == case["frames"] * (fb + 1), "unexpected size of the reference dump"
    raw = raw.reshape(case["frames"], fb + 1)
    return raw[:, 0].astype(bool), raw[:, 1:]


def inputs(oracle, case):
    clip = oracle.synth_clip(case["frames"], case["width"], case["height"], oracle.FMT_RGBX8, profile=oracle.SYNTH_SCENE)
    path = os.path.join(GOLD, f"reference_in_{case['name']}.bin")
    if os.path.exists(path):
        assert np.array_equal(np.fromfile(path, np.uint8), clip.reshape(-1)), "make_inputs.py and the generator disagree"
    return clip


def test_case_table_matches_the_rust_dump_modules():
    """the Rust dump modules carry the same case table as make_inputs.py (they cannot read JSON without a dependency)"""
    root = os.path.dirname(GOLD.rstrip("/"))
    root = os.path.dirname(root)
    text = {c: open(os.path.join(root, "tools", "reference_goldens", f"golden_dump_{c}.rs")).read() for c in ("dips", "dips_alt")}
    for case in CASES:
        line = f'("{case["name"]}", {case["width"]}, {case["height"]}, {case["frames"]}, {"true" if case["colorize"] else "false"}, {case["filter"]}, {case["sigmoid_scalar"]:.1f}, {case["chroma"]})'
        assert line in text[case["crate"]], line


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_the_reference(oracle, case):
    computed, want = load_reference(case)
    clip = inputs(oracle, case)
    w, h = case["width"], case["height"]
    if case["crate"] == "dips":
        ref = oracle.ComputeStateOracle(w, h, bool(case["colorize"]), case["filter"], case["sigmoid_scalar"], case["chroma"])
        for t in range(case["frames"]):
            out, passthrough = ref.frame(clip[t])
            assert passthrough == (not computed[t]), f"frame {t}: warm-up / passthrough pattern differs from the reference"
            assert np.abs(out.astype(int) - want[t].astype(int)).max() <= 1, f"frame {t}"
    else:
        ref = oracle.DiPsComputeOracle(w, h, bool(case["colorize"]), case["filter"], case["sigmoid_scalar"], case["chroma"])
        for t in range(case["frames"]):
            out = ref.send_frame(clip[t], t in case["snapshot_frames"])
            assert np.abs(out.astype(int) - want[t].astype(int)).max() <= 1, f"frame {t}"


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_library_matches_the_reference(oracle, case):
    import dips_b200
    computed, want = load_reference(case)
    clip = inputs(oracle, case)
    w, h = case["width"], case["height"]
    flavor = dips_b200.FLAVOR_DIPS_RING4 if case["crate"] == "dips" else dips_b200.FLAVOR_ALT_RING2
    with dips_b200.Context(w, h, dips_b200.FMT_RGBX8, 0, 0, chroma=case["chroma"], colorize=bool(case["colorize"]), filt=case["filter"],
                           sigmoid_scalar=case["sigmoid_scalar"], flavor=flavor) as ctx:
        for t in range(case["frames"]):
            if t in case["snapshot_frames"]:
                ctx.snapshot()
            rc, out, _ = ctx.push_frame(clip[t])
            if case["crate"] == "dips":
                assert (rc == dips_b200.NOT_READY) == (not computed[t]), f"frame {t}: passthrough pattern"
            d = np.abs(out.astype(int) - want[t].astype(int))
            assert d.max() <= 3 and (d <= 1).mean() >= 0.97, f"frame {t}: max {d.max()}, within 1 LSB {(d <= 1).mean():.3f}"
