//! Golden dump for the `dips` crate (see tools/reference_goldens/README.md in the dips-b200 repository).
//! Lives inside the crate (`dips/src/golden_dump.rs`, `#[cfg(test)] mod golden_dump;` in lib.rs) because `gpu::ComputeState`
//! is private.  Feeds every input clip through the crate's own per-frame path -- exactly what `frame_callback`
//! (dips/src/lib.rs:233-246) does -- and writes what `dispatch()` returned.
use std::{env, fs, io::Write, path::PathBuf};

use crate::gpu::ComputeState;
use crate::{ChromaFilter, DiPsFilter};

// (name, width, height, frames, colorize, filter, sigmoid scalar, chroma) -- keep in sync with make_inputs.py
const CASES: &[(&str, u32, u32, usize, bool, u32, f32, u32)] = &[
    ("dips_grey_unfiltered", 64, 48, 10, false, 255, 5.0, 0),
    ("dips_colour_sigmoid", 64, 48, 10, true, 0, 5.0, 0),
    ("dips_grey_inverse_red", 48, 32, 9, false, 1, 3.0, 1),
];

#[test]
fn golden_dump() {
    let dir = PathBuf::from(env::var("DIPS_GOLDEN_DIR").expect("set DIPS_GOLDEN_DIR to dips-b200/tests/golden"));
    for &(name, w, h, n, colorize, filter, sig, chroma) in CASES {
        let input = fs::read(dir.join(format!("reference_in_{name}.bin"))).expect("run make_inputs.py first");
        let fb = (w * h * 4) as usize;
        assert_eq!(input.len(), fb * n);
        let filter = match filter { 0 => DiPsFilter::Sigmoid, 1 => DiPsFilter::InverseSigmoid, _ => DiPsFilter::Unfiltered };
        let chroma = match chroma { 1 => ChromaFilter::Red, 2 => ChromaFilter::Green, 3 => ChromaFilter::Blue, _ => ChromaFilter::None };
        let mut cs = ComputeState::new(colorize, 1, sig, filter, chroma).expect("no wgpu adapter");
        let mut out = fs::File::create(dir.join(format!("reference_out_{name}.bin"))).unwrap();
        for t in 0..n {
            let frame = &input[t * fb..(t + 1) * fb];
            cs.add_texture(w, h, frame);
            match cs.dispatch() {
                Some(px) => { out.write_all(&[1u8]).unwrap(); out.write_all(&px).unwrap(); }
                None => { out.write_all(&[0u8]).unwrap(); out.write_all(frame).unwrap(); }
            }
        }
        println!("wrote reference_out_{name}.bin");
    }
}
