//! Golden dump for the `dips_alt` crate (see tools/reference_goldens/README.md in the dips-b200 repository).
//! Lives inside the crate (`dips_alt/src/golden_dump.rs`, `#[cfg(test)] mod golden_dump;` in lib.rs): `dips_compute` and
//! `gpu_controller` are private.  Drives `DiPsCompute::send_frame` the way `run_dips_on_file` does
//! (dips_alt/src/lib.rs:599-642): no window, no surface texture, snapshot on the frame where index == FRAME_COUNT.
use std::{env, fs, io::Write, path::PathBuf};

use crate::dips_compute::{ChromaFilter, DiPsCompute, DiPsProperties, Filter};
use crate::gpu_controller::GpuController;

// (name, width, height, frames, colorize, filter, sigmoid scalar, chroma) -- keep in sync with make_inputs.py
const CASES: &[(&str, u32, u32, usize, bool, u32, f32, u32)] = &[
    ("alt_colour_sigmoid", 64, 48, 10, true, 0, 5.0, 0),
    ("alt_grey_inverse", 48, 32, 9, false, 1, 7.0, 0),
];
const FRAME_COUNT: usize = 2; // dips_alt/src/lib.rs:36

#[test]
fn golden_dump() {
    let dir = PathBuf::from(env::var("DIPS_GOLDEN_DIR").expect("set DIPS_GOLDEN_DIR to dips-b200/tests/golden"));
    let gpu = GpuController::new().expect("no wgpu adapter");
    for &(name, w, h, n, colorize, filter, sig, chroma) in CASES {
        let input = fs::read(dir.join(format!("reference_in_{name}.bin"))).expect("run make_inputs.py first");
        let fb = (w * h * 4) as usize;
        assert_eq!(input.len(), fb * n);
        let mut props = DiPsProperties::default();
        props.set_colorize(colorize);
        props.set_filter(if filter == 1 { Filter::InverseSigmoid } else { Filter::Sigmoid });
        props.set_sigmoid_horizontal_scalar(sig);
        props.set_chroma_filter(match chroma { 1 => ChromaFilter::Red, 2 => ChromaFilter::Green, 3 => ChromaFilter::Blue, _ => ChromaFilter::All });
        let mut dc = DiPsCompute::new(FRAME_COUNT, w, h, None, gpu.device.clone(), gpu.queue.clone(), props).unwrap();
        let mut out = fs::File::create(dir.join(format!("reference_out_{name}.bin"))).unwrap();
        for t in 0..n {
            let frame = &input[t * fb..(t + 1) * fb];
            let px = dc.send_frame(frame, if t == FRAME_COUNT { Some(()) } else { None }, None);
            out.write_all(&[1u8]).unwrap();
            out.write_all(&px).unwrap();
        }
        println!("wrote reference_out_{name}.bin");
    }
}
