#!/usr/bin/env python
"""Writes the deterministic RGBA8 input clips of the reference-golden recipe (tools/reference_goldens/README.md) and the
case table both Rust dump modules and tests/test_reference_goldens.py read.  Seed 0x44695073, oracle generator."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
# crate, name, width, height, frames, colorize, filter (dips: 0 sigmoid / 1 inverse / 255 unfiltered; dips_alt: 0 / 1), sigmoid scalar, chroma
CASES = [
    ("dips", "dips_grey_unfiltered", 64, 48, 10, 0, 255, 5.0, 0),
    ("dips", "dips_colour_sigmoid", 64, 48, 10, 1, 0, 5.0, 0),
    ("dips", "dips_grey_inverse_red", 48, 32, 9, 0, 1, 3.0, 1),
    ("dips_alt", "alt_colour_sigmoid", 64, 48, 10, 1, 0, 5.0, 0),
    ("dips_alt", "alt_grey_inverse", 48, 32, 9, 0, 1, 7.0, 0),
]


def main():
    table = []
    for crate, name, w, h, n, colorize, filt, sig, chroma in CASES:
        clip = O.synth_clip(n, w, h, O.FMT_RGBX8, profile=O.SYNTH_SCENE)
        clip.tofile(os.path.join(GOLD, f"reference_in_{name}.bin"))
        table.append(dict(crate=crate, name=name, width=w, height=h, frames=n, colorize=colorize, filter=filt,
                          sigmoid_scalar=sig, chroma=chroma,
                          snapshot_frames=[2] if crate == "dips_alt" else []))   # dips_alt/src/lib.rs:222-225: index == FRAME_COUNT
    with open(os.path.join(GOLD, "reference_cases.json"), "w") as f:
        json.dump(table, f, indent=1)
    print("wrote", len(table), "input clips to", GOLD)


if __name__ == "__main__":
    main()
