//! Raw bindings to `libdips_b200.so`.  One declaration per entry point of `include/dips_b200.h`, same order.
//! NOTE: this crate is shipped as source only -- the build image has no Rust toolchain;
//! it is kept in lock-step with the header by `tests/test_abi.py::test_rust_sys_crate_declares_every_symbol`.
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_void};

#[repr(C)]
pub struct dipsb_ctx {
    _private: [u8; 0],
}

#[repr(C)]
pub struct dipsb_group {
    _private: [u8; 0],
}

pub const DIPSB_UNIQUE_ID_BYTES: usize = 128;
pub const DIPSB_REDUCE_AUTO: i32 = 0;
pub const DIPSB_REDUCE_P2P: i32 = 1;
pub const DIPSB_REDUCE_NCCL: i32 = 2;

pub const DIPSB_OK: i32 = 0;
pub const DIPSB_NOT_READY: i32 = 1;
pub const DIPSB_ERR_INVALID: i32 = -1;
pub const DIPSB_ERR_CUDA: i32 = -2;
pub const DIPSB_ERR_NOMEM: i32 = -3;
pub const DIPSB_ERR_STATE: i32 = -4;

pub const DIPSB_FMT_RGB8: i32 = 0;
pub const DIPSB_FMT_RGBX8: i32 = 1;
pub const DIPSB_FMT_BGR8: i32 = 2;
pub const DIPSB_FMT_BGRX8: i32 = 3;
pub const DIPSB_MODE_OVERALL: i32 = 0;
pub const DIPSB_MODE_PERFRAME: i32 = 1;
pub const DIPSB_FLAVOR_FRAME0: i32 = 0;
pub const DIPSB_FLAVOR_DIPS_RING4: i32 = 1;
pub const DIPSB_FLAVOR_ALT_RING2: i32 = 2;
pub const DIPSB_FLAVOR_ALT_RING2_MEDIAN: i32 = 3;
pub const DIPSB_FILTER_SIGMOID: i32 = 0;
pub const DIPSB_FILTER_INV_SIGMOID: i32 = 1;
pub const DIPSB_FILTER_NONE: i32 = 255;

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct dipsb_config {
    pub struct_size: u32,
    pub device: i32,
    pub width: u32,
    pub height: u32,
    pub format: i32,
    pub mode: i32,
    pub chroma: i32,
    pub threshold: u32,
    pub colorize: i32,
    pub filter: i32,
    pub sigmoid_scalar: f32,
    pub spatial_window: i32,
    pub flavor: i32,
    pub reserved: [u32; 3],
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct dipsb_frame_stats {
    pub frame_index: u64,
    pub sad: u64,
    pub count: u64,
}

extern "C" {
    pub fn dipsb_abi_version() -> i32;
    pub fn dipsb_default_config(cfg: *mut dipsb_config);
    pub fn dipsb_create(cfg: *const dipsb_config, out: *mut *mut dipsb_ctx) -> i32;
    pub fn dipsb_destroy(ctx: *mut dipsb_ctx);
    pub fn dipsb_last_error(ctx: *const dipsb_ctx) -> *const c_char;
    pub fn dipsb_reset(ctx: *mut dipsb_ctx) -> i32;
    pub fn dipsb_set_threshold(ctx: *mut dipsb_ctx, threshold: u32) -> i32;
    pub fn dipsb_set_stream(ctx: *mut dipsb_ctx, stream: *mut c_void) -> i32;
    pub fn dipsb_adopt_stream(ctx: *mut dipsb_ctx, stream: *mut c_void) -> i32;
    pub fn dipsb_use_private_stream(ctx: *mut dipsb_ctx) -> i32;
    pub fn dipsb_synchronize(ctx: *mut dipsb_ctx) -> i32;
    pub fn dipsb_prime_device(ctx: *mut dipsb_ctx, d_frame: *const c_void) -> i32;
    pub fn dipsb_prime_median4_device(ctx: *mut dipsb_ctx, d_frames: *const c_void, frame_stride_bytes: u64) -> i32;
    pub fn dipsb_prime_host(ctx: *mut dipsb_ctx, frame: *const u8) -> i32;
    pub fn dipsb_state_plane_device(ctx: *mut dipsb_ctx, d_state: *mut *mut c_void) -> i32;
    pub fn dipsb_mark_state_valid(ctx: *mut dipsb_ctx, valid: i32) -> i32;
    pub fn dipsb_get_state_plane(ctx: *mut dipsb_ctx, out: *mut u16) -> i32;
    pub fn dipsb_run_clip_device(ctx: *mut dipsb_ctx, d_frames: *const c_void, n_frames: u64, frame_stride_bytes: u64, first_frame_index: u64) -> i32;
    pub fn dipsb_run_clip_host(ctx: *mut dipsb_ctx, frames: *const u8, n_frames: u64, frame_stride_bytes: u64, first_frame_index: u64) -> i32;
    pub fn dipsb_push_frame(ctx: *mut dipsb_ctx, px: *const u8, width: u32, height: u32, stride: u32, format: i32, out_rgba: *mut u8, stats: *mut dipsb_frame_stats) -> i32;
    pub fn dipsb_push_frame_pipelined(ctx: *mut dipsb_ctx, px: *const u8, width: u32, height: u32, stride: u32, format: i32, out_rgba_prev: *mut u8, stats_prev: *mut dipsb_frame_stats) -> i32;
    pub fn dipsb_flush_frame(ctx: *mut dipsb_ctx, out_rgba: *mut u8, stats: *mut dipsb_frame_stats) -> i32;
    pub fn dipsb_stage_frame(ctx: *mut dipsb_ctx, px: *const u8, width: u32, height: u32, stride: u32, format: i32) -> i32;
    pub fn dipsb_dispatch_staged(ctx: *mut dipsb_ctx, out_rgba: *mut u8, stats: *mut dipsb_frame_stats) -> i32;
    pub fn dipsb_snapshot(ctx: *mut dipsb_ctx) -> i32;
    pub fn dipsb_frames_processed(ctx: *const dipsb_ctx) -> u64;
    pub fn dipsb_get_accumulators(ctx: *mut dipsb_ctx, acc_sum: *mut u32, acc_cnt: *mut u32) -> i32;
    pub fn dipsb_set_accumulators(ctx: *mut dipsb_ctx, acc_sum: *const u32, acc_cnt: *const u32) -> i32;
    pub fn dipsb_accumulators_device(ctx: *mut dipsb_ctx, d_acc: *mut *mut c_void, n_elems: *mut u64) -> i32;
    pub fn dipsb_pack_accumulators_device(ctx: *mut dipsb_ctx, total_frames: u64, d_packed: *mut *mut c_void, n_words: *mut u64) -> i32;
    pub fn dipsb_unpack_accumulators_device(ctx: *mut dipsb_ctx) -> i32;
    pub fn dipsb_get_scalars(ctx: *mut dipsb_ctx, first: u64, n: u64, sad: *mut u64, cnt: *mut u64) -> i32;
    pub fn dipsb_get_intensity_map(ctx: *mut dipsb_ctx, n_eff: u64, out: *mut f32) -> i32;
    pub fn dipsb_get_frame_means(ctx: *mut dipsb_ctx, first: u64, n: u64, out: *mut f32) -> i32;
    pub fn dipsb_host_alloc(device: i32, bytes: u64, out: *mut *mut c_void) -> i32;
    pub fn dipsb_host_free(p: *mut c_void) -> i32;
    pub fn dipsb_host_register(device: i32, p: *mut c_void, bytes: u64) -> i32;
    pub fn dipsb_host_unregister(p: *mut c_void) -> i32;
    pub fn dipsb_host_copy2d(dst: *mut c_void, dpitch: u64, src: *const c_void, spitch: u64, row_bytes: u64, rows: u64) -> i32;
    pub fn dipsb_host_copy_threads() -> u32;
    pub fn dipsb_synth_fill_device(device: i32, d_dst: *mut c_void, first_frame: u64, n_frames: u64, width: u32, height: u32, format: i32, seed: u64, profile: i32, stream: *mut c_void) -> i32;
    pub fn dipsb_launch_count() -> u64;
    pub fn dipsb_last_plan(ctx: *const dipsb_ctx, out: *mut u32) -> i32;
    pub fn dipsb_enable_timing(ctx: *mut dipsb_ctx, on: i32) -> i32;
    pub fn dipsb_clip_kernel_time(ctx: *mut dipsb_ctx, total_ms: *mut f64, launches: *mut u64) -> i32;
    pub fn dipsb_stream_probe(ctx: *mut dipsb_ctx, d_frames: *const c_void, n_frames: u64, frame_stride_bytes: u64, reps: u32, ms: *mut f32) -> i32;
    pub fn dipsb_set_kernel(ctx: *mut dipsb_ctx, kernel: i32) -> i32;
    pub fn dipsb_plan_query(width: u32, height: u32, format: i32, num_sms: u32, out: *mut u32) -> i32;
    pub fn dipsb_set_tuning(ctx: *mut dipsb_ctx, stages: u32, tile_px: u32, segments: u32, regs: u32) -> i32;
    // several GPUs: frame-range shards of one clip (include/dips_b200.h, "several GPUs")
    pub fn dipsb_shard_range(total_frames: u64, nranks: u32, rank: u32, first: *mut u64, count: *mut u64);
    pub fn dipsb_xchg_plan_query(total_frames: u64, nranks: u32, n_elems: u64, out: *mut u64) -> i32;
    pub fn dipsb_comm_unique_id(id128: *mut c_void) -> i32;
    pub fn dipsb_comm_init_rank(ctx: *mut dipsb_ctx, nranks: u32, rank: u32, id128: *const c_void) -> i32;
    pub fn dipsb_comm_destroy(ctx: *mut dipsb_ctx) -> i32;
    pub fn dipsb_comm_info(ctx: *const dipsb_ctx, out: *mut u32) -> i32;
    pub fn dipsb_comm_set_reduce(ctx: *mut dipsb_ctx, path: i32) -> i32;
    pub fn dipsb_comm_check(ctx: *mut dipsb_ctx) -> i32;
    pub fn dipsb_run_clip_sharded_device(ctx: *mut dipsb_ctx, d_frames: *const c_void, n_frames: u64, frame_stride_bytes: u64, first_frame_index: u64, total_frames: u64) -> i32;
    pub fn dipsb_run_clip_sharded_host(ctx: *mut dipsb_ctx, frames: *const u8, n_frames: u64, frame_stride_bytes: u64, first_frame_index: u64, total_frames: u64) -> i32;
    pub fn dipsb_comm_phase_times(ctx: *mut dipsb_ctx, out_ms: *mut f64, passes: *mut u64) -> i32;
    pub fn dipsb_comm_probe(ctx: *mut dipsb_ctx, what: i32, total_frames: u64, reps: u32, ms: *mut f32) -> i32;
    pub fn dipsb_gather_accumulators(ctx: *mut dipsb_ctx) -> i32;
    pub fn dipsb_create_group(cfg: *const dipsb_config, ndev: u32, devices: *const i32, out: *mut *mut dipsb_group) -> i32;
    pub fn dipsb_destroy_group(grp: *mut dipsb_group);
    pub fn dipsb_group_size(grp: *const dipsb_group) -> u32;
    pub fn dipsb_group_ctx(grp: *mut dipsb_group, rank: u32) -> *mut dipsb_ctx;
    pub fn dipsb_group_last_error(grp: *const dipsb_group) -> *const c_char;
    pub fn dipsb_group_reset(grp: *mut dipsb_group) -> i32;
    pub fn dipsb_group_run_clip_device(grp: *mut dipsb_group, d_frames: *const *const c_void, n_frames: *const u64, frame_stride_bytes: u64) -> i32;
    pub fn dipsb_group_gather_accumulators(grp: *mut dipsb_group) -> i32;
    pub fn dipsb_group_synchronize(grp: *mut dipsb_group) -> i32;
    pub fn dipsb_group_get_accumulators(grp: *mut dipsb_group, acc_sum: *mut u32, acc_cnt: *mut u32) -> i32;
    pub fn dipsb_group_get_scalars(grp: *mut dipsb_group, first: u64, n: u64, sad: *mut u64, cnt: *mut u64) -> i32;
}
