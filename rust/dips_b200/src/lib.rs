//! Safe wrappers that keep the reference's per-frame API so the callers compile unchanged:
//!   * `ComputeState::{new, add_texture, dispatch}`  == dips/src/gpu/mod.rs:59, :170, :306
//!   * `frame_callback`                              == dips/src/lib.rs:233-246
//!   * `DiPsCompute::{new, send_frame}`              == dips_alt/src/dips_compute/mod.rs:270, :498
//! The wgpu device/queue/bind-group machinery is gone: the state lives in a `dipsb_ctx` (CUDA, sm_100a).
//! Shipped as source only (no Rust toolchain in the build image); see INTEGRATION.md for how a maintainer wires it in.
use std::ffi::CStr;
use std::ptr;

use dips_b200_sys as sys;

/// dips/src/lib.rs:26-41 (numbering of `Into<f64>`)
#[derive(Copy, Clone, Debug)]
pub enum DiPsFilter {
    Unfiltered,
    Sigmoid,
    InverseSigmoid,
}

impl DiPsFilter {
    fn as_ffi(self) -> i32 {
        match self {
            DiPsFilter::Unfiltered => sys::DIPSB_FILTER_NONE,
            DiPsFilter::Sigmoid => sys::DIPSB_FILTER_SIGMOID,
            DiPsFilter::InverseSigmoid => sys::DIPSB_FILTER_INV_SIGMOID,
        }
    }
}

/// dips/src/lib.rs:44-61
#[derive(Copy, Clone, Debug)]
pub enum ChromaFilter {
    None,
    Red,
    Green,
    Blue,
}

impl ChromaFilter {
    fn as_ffi(self) -> i32 {
        match self {
            ChromaFilter::None => 0,
            ChromaFilter::Red => 1,
            ChromaFilter::Green => 2,
            ChromaFilter::Blue => 3,
        }
    }
}

fn last_error(ctx: *const sys::dipsb_ctx) -> String {
    unsafe { CStr::from_ptr(sys::dipsb_last_error(ctx)).to_string_lossy().into_owned() }
}

/// Same constructor arguments as the reference's `ComputeState::new` (dips/src/gpu/mod.rs:59-65).  The CUDA context is
/// created lazily on the first `add_texture`, because the reference only learns the frame size there (:170).
pub struct ComputeState {
    colorize: bool,
    spatial_window_size: i32,
    sensitivity: f32,
    filter_type: DiPsFilter,
    chroma_filter: ChromaFilter,
    ctx: *mut sys::dipsb_ctx,
    width: u32,
    height: u32,
    have_frame: bool,
    /// true (default): the crate's own temporal semantics (median-of-4 ring); false: reference = first frame
    pub reference_exact: bool,
}

// One caller at a time, but the owner thread may change (GStreamer streaming thread vs. the smol executor thread,
// dips/src/frame_extractor.rs:76, :232-234): Send, not Sync.  Every entry point of the library selects its device.
unsafe impl Send for ComputeState {}

impl ComputeState {
    pub fn new(
        colorize: bool,
        spatial_window_size: i32,
        sensitivity: f32,
        filter_type: DiPsFilter,
        chroma_filter: ChromaFilter,
    ) -> anyhow::Result<Self> {
        if ![1, 3, 5, 7].contains(&spatial_window_size) {
            // windows > 1 run a correct zero-padded median (the reference shader's own loop is defective)
            anyhow::bail!("spatial_window_size {} is not one of 1, 3, 5, 7", spatial_window_size);
        }
        Ok(Self {
            colorize,
            spatial_window_size,
            sensitivity,
            filter_type,
            chroma_filter,
            ctx: ptr::null_mut(),
            width: 0,
            height: 0,
            have_frame: false,
            reference_exact: true,
        })
    }

    fn ensure_ctx(&mut self, width: u32, height: u32) -> anyhow::Result<()> {
        if !self.ctx.is_null() && (width, height) == (self.width, self.height) {
            return Ok(());
        }
        if !self.ctx.is_null() {
            unsafe { sys::dipsb_destroy(self.ctx) };
            self.ctx = ptr::null_mut();
        }
        let mut cfg: sys::dipsb_config = unsafe { std::mem::zeroed() };
        unsafe { sys::dipsb_default_config(&mut cfg) };
        cfg.width = width;
        cfg.height = height;
        cfg.format = sys::DIPSB_FMT_RGBX8; // the decoder hands over tightly packed RGBA (frame_extractor.rs:141-148)
        cfg.mode = sys::DIPSB_MODE_OVERALL;
        cfg.chroma = self.chroma_filter.as_ffi();
        cfg.colorize = self.colorize as i32;
        cfg.filter = self.filter_type.as_ffi();
        cfg.sigmoid_scalar = self.sensitivity;
        cfg.spatial_window = self.spatial_window_size;
        // DIPSB_FLAVOR_DIPS_RING4 reproduces the crate's median-of-4 ring and 3 passthrough frames exactly;
        // DIPSB_FLAVOR_FRAME0 is the north-star semantics (reference = first frame, 1 passthrough frame)
        cfg.flavor = if self.reference_exact { sys::DIPSB_FLAVOR_DIPS_RING4 } else { sys::DIPSB_FLAVOR_FRAME0 };
        let mut ctx = ptr::null_mut();
        let rc = unsafe { sys::dipsb_create(&cfg, &mut ctx) };
        if rc != sys::DIPSB_OK {
            anyhow::bail!("dipsb_create failed ({}): {}", rc, last_error(ptr::null()));
        }
        self.ctx = ctx;
        self.width = width;
        self.height = height;
        Ok(())
    }

    /// dips/src/gpu/mod.rs:170 -- the slice is borrowed for the call only.  No copy of our own (the reference's `to_vec`,
    /// :171): `dipsb_stage_frame` copies it straight into the library's page-locked input slot and starts the upload.
    pub fn add_texture(&mut self, width: u32, height: u32, frame_data: &[u8]) {
        self.have_frame = false;
        if self.ensure_ctx(width, height).is_err() || frame_data.len() < (width as usize) * (height as usize) * 4 {
            return;
        }
        let rc = unsafe { sys::dipsb_stage_frame(self.ctx, frame_data.as_ptr(), width, height, width * 4, sys::DIPSB_FMT_RGBX8) };
        self.have_frame = rc == sys::DIPSB_OK;
    }

    /// dips/src/gpu/mod.rs:306 -- `None` while there is no reference yet (the caller passes the input through).
    pub fn dispatch(&mut self) -> Option<Vec<u8>> {
        if !self.have_frame || self.ctx.is_null() {
            return None;
        }
        self.have_frame = false;
        let mut out = vec![0u8; (self.width * self.height * 4) as usize];
        let rc = unsafe { sys::dipsb_dispatch_staged(self.ctx, out.as_mut_ptr(), ptr::null_mut()) };
        match rc {
            sys::DIPSB_OK => Some(out),
            _ => None, // DIPSB_NOT_READY (reference frame) or an error: passthrough, like the reference's warm-up
        }
    }
}

impl Drop for ComputeState {
    fn drop(&mut self) {
        if !self.ctx.is_null() {
            unsafe { sys::dipsb_destroy(self.ctx) };
        }
    }
}

/// dips/src/lib.rs:233-246, unchanged.
pub fn frame_callback(width: u32, height: u32, frame_data: &[u8], compute: &mut ComputeState) -> Vec<u8> {
    compute.add_texture(width, height, frame_data);
    if let Some(new_frame) = compute.dispatch() {
        new_frame
    } else {
        frame_data.to_vec()
    }
}

/// dips_alt/src/dips_compute/mod.rs:151-234
#[derive(Debug, Default, Copy, Clone, PartialEq)]
pub enum Filter {
    #[default]
    Sigmoid = 0,
    InverseSigmoid = 1,
}

#[derive(Debug, Copy, Clone)]
pub struct DiPsProperties {
    pub colorize: bool,
    pub window_size: u8,
    pub sigmoid_horizontal_scalar: f32,
    pub filter_type: Filter,
    pub chroma_filter: ChromaFilter,
}

/// `DiPsCompute` with the reference's own signatures (dips_alt/src/dips_compute/mod.rs:270-278, :498-503): the window, the
/// wgpu device / queue handles and the swap-chain texture are accepted -- as generic parameters, so this crate needs neither
/// wgpu nor winit -- and ignored: they have no meaning on the CUDA path.  `run_dips_on_file` (dips_alt/src/lib.rs:599-642)
/// therefore compiles against this type unchanged; the live-render branch of `send_frame` (`surface_texture: Some(..)`,
/// :565-595) is not reproduced: the difference frame is returned, presenting it stays with the caller.
pub struct DiPsCompute {
    ctx: *mut sys::dipsb_ctx,
    width: u32,
    height: u32,
}

impl DiPsCompute {
    pub fn new<W, D, Q>(
        _num_textures: usize,
        textures_width: u32,
        textures_height: u32,
        _dips_window: Option<&W>,
        _device: D,
        _queue: Q,
        props: DiPsProperties,
    ) -> anyhow::Result<Self> {
        let mut cfg: sys::dipsb_config = unsafe { std::mem::zeroed() };
        unsafe { sys::dipsb_default_config(&mut cfg) };
        cfg.width = textures_width;
        cfg.height = textures_height;
        cfg.format = sys::DIPSB_FMT_RGBX8;
        cfg.mode = sys::DIPSB_MODE_OVERALL;
        cfg.chroma = props.chroma_filter.as_ffi();
        cfg.colorize = props.colorize as i32;
        cfg.filter = props.filter_type as i32;
        cfg.sigmoid_scalar = props.sigmoid_horizontal_scalar;
        cfg.spatial_window = props.window_size as i32;
        cfg.flavor = sys::DIPSB_FLAVOR_ALT_RING2; // as shipped: 2-frame ring, snapshot on request
        let mut ctx = ptr::null_mut();
        let rc = unsafe { sys::dipsb_create(&cfg, &mut ctx) };
        if rc != sys::DIPSB_OK {
            anyhow::bail!("dipsb_create failed ({}): {}", rc, last_error(ptr::null()));
        }
        Ok(Self { ctx, width: textures_width, height: textures_height })
    }

    /// `snapshot: Some(())` makes this frame the new reference (dips_alt/src/lib.rs:222-225, refresh markers :668-670).
    pub fn send_frame<S>(&mut self, frame: &[u8], snapshot: Option<()>, _surface_texture: Option<&S>) -> Vec<u8> {
        if snapshot.is_some() {
            unsafe { sys::dipsb_snapshot(self.ctx) };
        }
        let mut out = vec![0u8; (self.width * self.height * 4) as usize];
        let rc = unsafe {
            sys::dipsb_push_frame(self.ctx, frame.as_ptr(), self.width, self.height, self.width * 4, sys::DIPSB_FMT_RGBX8, out.as_mut_ptr(), ptr::null_mut())
        };
        if rc < 0 {
            panic!("dipsb_push_frame failed ({}): {}", rc, last_error(self.ctx)); // the reference panics on device loss too
        }
        out
    }
}

impl Drop for DiPsCompute {
    fn drop(&mut self) {
        unsafe { sys::dipsb_destroy(self.ctx) };
    }
}

/// dips_alt/src/lib.rs:36
pub const FRAME_COUNT: usize = 2;

/// When the caller loops of `dips_alt` ask for a snapshot: on the frame where `index == FRAME_COUNT` (the third frame, and
/// the third after every refresh marker); `index` saturates one past it, a marker -- a 1-based count of frames processed --
/// resets it (dips_alt/src/lib.rs:222-232 live mode, :560-561 and :662-670 file mode).
#[derive(Debug, Default, Clone)]
pub struct SnapshotSchedule {
    refresh_markers: Vec<usize>,
    index: usize,
    overall_frame: usize,
}

impl SnapshotSchedule {
    pub fn new(refresh_markers: Vec<usize>) -> Self {
        Self { refresh_markers, index: 0, overall_frame: 0 }
    }

    /// Ask before sending the frame.
    pub fn snapshot_now(&self) -> Option<()> {
        if self.index == FRAME_COUNT { Some(()) } else { None }
    }

    /// Call after sending it.
    pub fn frame_sent(&mut self) {
        if self.index <= FRAME_COUNT {
            self.index += 1;
        }
        self.overall_frame += 1;
        if self.refresh_markers.contains(&self.overall_frame) {
            self.index = 0;
        }
    }
}

/// The compute part of `run_dips_on_file` (dips_alt/src/lib.rs:553-690) without the OpenCV capture / writer around it.
pub fn run_dips_on_frames<'a, I, F>(compute: &mut DiPsCompute, frames: I, refresh_markers: Vec<usize>, mut sink: F)
where
    I: IntoIterator<Item = &'a [u8]>,
    F: FnMut(usize, Vec<u8>),
{
    let mut schedule = SnapshotSchedule::new(refresh_markers);
    for (t, frame) in frames.into_iter().enumerate() {
        let out = compute.send_frame(frame, schedule.snapshot_now(), None::<&()>);
        schedule.frame_sent();
        sink(t, out);
    }
}

/// Page-locks a buffer the caller already owns and reuses (a decoder's buffer pool, the `Vec` the output is collected in)
/// in place, for the lifetime of the guard (`dipsb_host_register` / `dipsb_host_unregister`): frames inside it take the
/// direct copy-engine path of `dipsb_push_frame*`.  About a millisecond per 8 MB: once per buffer, not per frame.
pub struct RegisteredFrames<'a> {
    ptr: *mut std::ffi::c_void,
    _buf: std::marker::PhantomData<&'a mut [u8]>,
}

impl<'a> RegisteredFrames<'a> {
    pub fn new(device: i32, buf: &'a mut [u8]) -> anyhow::Result<Self> {
        let p = buf.as_mut_ptr() as *mut std::ffi::c_void;
        let rc = unsafe { sys::dipsb_host_register(device, p, buf.len() as u64) };
        if rc != 0 {
            let msg = unsafe { CStr::from_ptr(sys::dipsb_last_error(ptr::null())) };
            anyhow::bail!("dipsb_host_register failed ({rc}): {}", msg.to_string_lossy());
        }
        Ok(Self { ptr: p, _buf: std::marker::PhantomData })
    }
}

impl Drop for RegisteredFrames<'_> {
    fn drop(&mut self) {
        unsafe { sys::dipsb_host_unregister(self.ptr) };
    }
}

/// Page-locked host buffer (`dipsb_host_alloc`).  A decoder that writes its frames here -- and a caller that receives
/// the difference frame here -- lets `dipsb_push_frame*` skip its two staging copies (the role of the mapped gst buffer
/// in dips/src/frame_extractor.rs:216-226).  Derefs to a byte slice.
pub struct PinnedFrame {
    ptr: *mut u8,
    len: usize,
}

unsafe impl Send for PinnedFrame {}

impl PinnedFrame {
    pub fn new(device: i32, len: usize) -> anyhow::Result<Self> {
        let mut p: *mut std::ffi::c_void = ptr::null_mut();
        let rc = unsafe { sys::dipsb_host_alloc(device, len as u64, &mut p) };
        if rc != 0 {
            let msg = unsafe { CStr::from_ptr(sys::dipsb_last_error(ptr::null())) };
            anyhow::bail!("dipsb_host_alloc failed ({rc}): {}", msg.to_string_lossy());
        }
        Ok(Self { ptr: p as *mut u8, len })
    }
}

impl std::ops::Deref for PinnedFrame {
    type Target = [u8];
    fn deref(&self) -> &[u8] {
        unsafe { std::slice::from_raw_parts(self.ptr, self.len) }
    }
}

impl std::ops::DerefMut for PinnedFrame {
    fn deref_mut(&mut self) -> &mut [u8] {
        unsafe { std::slice::from_raw_parts_mut(self.ptr, self.len) }
    }
}

impl Drop for PinnedFrame {
    fn drop(&mut self) {
        unsafe { sys::dipsb_host_free(self.ptr as *mut std::ffi::c_void) };
    }
}
