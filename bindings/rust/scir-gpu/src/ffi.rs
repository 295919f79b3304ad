//! Raw FFI of libscir_b200.so -- GENERATED from include/scir_b200.h by tools/gen_rust_ffi.py; do not edit.
//! Replaces the hand-declared `extern "C"` block over libcuda in the reference crate
//! (crates/scir-gpu/src/lib.rs:549-581): the crate no longer talks to the driver, only to this C ABI.
#![allow(non_camel_case_types, dead_code, missing_docs)]

use std::os::raw::{c_char, c_int, c_void};

/// Opaque handle: one device + one stream + scratch (`scir_b200_ctx`).
#[repr(C)]
pub struct ScirB200Ctx {
    _private: [u8; 0],
}
/// Opaque handle: several ctxs, rows sharded across them (`scir_b200_mg`).
#[repr(C)]
pub struct ScirB200Mg {
    _private: [u8; 0],
}
/// Integer plan of `resample_poly` (`scir_b200_resample_plan`).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]
pub struct ScirB200ResamplePlan {
    pub up: i64,
    pub down: i64,
    pub n_out: i64,
    pub half_len: i64,
    pub n_pre_pad: i64,
    pub n_post_pad: i64,
    pub n_pre_remove: i64,
    pub len_h_padded: i64,
    pub upfirdn_len: i64,
}

pub const SCIR_B200_OK: c_int = 0;
pub const SCIR_B200_ERR_INVALID_ARG: c_int = -1;
pub const SCIR_B200_ERR_NO_DEVICE: c_int = -2;
pub const SCIR_B200_ERR_OOM: c_int = -3;
pub const SCIR_B200_ERR_LAUNCH: c_int = -4;
pub const SCIR_B200_ERR_SHAPE: c_int = -5;
pub const SCIR_B200_ERR_UNSUPPORTED: c_int = -6;
pub const SCIR_B200_MAX_TAPS: usize = 7936;
pub const SCIR_B200_TAPS_SCIR: c_int = 0;
pub const SCIR_B200_TAPS_LFILTER: c_int = 1;
pub const SCIR_B200_PAD_ZERO_STATE: c_int = 0;
pub const SCIR_B200_PAD_ODD: c_int = 1;
pub const SCIR_B200_PAD_EVEN: c_int = 2;
pub const SCIR_B200_PAD_CONSTANT: c_int = 3;
pub const SCIR_B200_PAD_SCIPY_NONE: c_int = 4;
pub const SCIR_B200_EXT_CONSTANT: c_int = 0;
pub const SCIR_B200_EXT_SYMMETRIC: c_int = 1;
pub const SCIR_B200_EXT_EDGE: c_int = 2;
pub const SCIR_B200_EXT_SMOOTH: c_int = 3;
pub const SCIR_B200_EXT_PERIODIC: c_int = 4;
pub const SCIR_B200_EXT_REFLECT: c_int = 5;
pub const SCIR_B200_EXT_ANTISYMMETRIC: c_int = 6;
pub const SCIR_B200_EXT_ANTIREFLECT: c_int = 7;
pub const SCIR_B200_EXT_LINE: c_int = 8;
pub const SCIR_B200_PAD_STAT_MEAN: c_int = 16;
pub const SCIR_B200_PAD_STAT_MEDIAN: c_int = 17;
pub const SCIR_B200_PAD_STAT_MINIMUM: c_int = 18;
pub const SCIR_B200_PAD_STAT_MAXIMUM: c_int = 19;

#[link(name = "scir_b200")]
extern "C" {
    pub fn scir_b200_version() -> *const c_char;
    pub fn scir_b200_last_error() -> *const c_char;
    pub fn scir_b200_device_count(count: *mut c_int) -> c_int;
    pub fn scir_b200_ctx_create(device: c_int, ctx: *mut *mut ScirB200Ctx) -> c_int;
    pub fn scir_b200_ctx_create_on_stream(device: c_int, cuda_stream: *mut c_void, ctx: *mut *mut ScirB200Ctx) -> c_int;
    pub fn scir_b200_ctx_destroy(ctx: *mut ScirB200Ctx) -> c_int;
    pub fn scir_b200_ctx_sync(ctx: *mut ScirB200Ctx) -> c_int;
    pub fn scir_b200_ctx_device(ctx: *const ScirB200Ctx, device: *mut c_int) -> c_int;
    pub fn scir_b200_ctx_stream(ctx: *const ScirB200Ctx, cuda_stream: *mut *mut c_void) -> c_int;
    pub fn scir_b200_ctx_set_option(ctx: *mut ScirB200Ctx, key: *const c_char, value: i64) -> c_int;
    pub fn scir_b200_ctx_get_option(ctx: *const ScirB200Ctx, key: *const c_char, value: *mut i64) -> c_int;
    pub fn scir_b200_ctx_launch_count(ctx: *const ScirB200Ctx, count: *mut u64) -> c_int;
    pub fn scir_b200_malloc(ctx: *mut ScirB200Ctx, bytes: usize, d_ptr: *mut *mut c_void) -> c_int;
    pub fn scir_b200_free(ctx: *mut ScirB200Ctx, d_ptr: *mut c_void) -> c_int;
    pub fn scir_b200_memcpy_h2d(ctx: *mut ScirB200Ctx, d_dst: *mut c_void, h_src: *const c_void, bytes: usize) -> c_int;
    pub fn scir_b200_memcpy_d2h(ctx: *mut ScirB200Ctx, h_dst: *mut c_void, d_src: *const c_void, bytes: usize) -> c_int;
    pub fn scir_b200_host_alloc(bytes: usize, h_ptr: *mut *mut c_void) -> c_int;
    pub fn scir_b200_host_free(h_ptr: *mut c_void) -> c_int;
    pub fn scir_b200_host_register(h_ptr: *mut c_void, bytes: usize) -> c_int;
    pub fn scir_b200_host_unregister(h_ptr: *mut c_void) -> c_int;
    pub fn scir_b200_host_is_pinned(h_ptr: *const c_void, bytes: usize, pinned: *mut c_int) -> c_int;
    pub fn scir_b200_current_device(device: *mut c_int) -> c_int;
    pub fn scir_b200_fir1d_batched_f32(ctx: *mut ScirB200Ctx, d_x: *const f32, ld_x: i64, taps: *const f32, k: i64, tap_order: c_int, d_y: *mut f32, ld_y: i64, batch: i64, n: i64) -> c_int;
    pub fn scir_b200_fir1d_batched_f32_host(ctx: *mut ScirB200Ctx, h_x: *const f32, ld_x: i64, taps: *const f32, k: i64, tap_order: c_int, h_y: *mut f32, ld_y: i64, batch: i64, n: i64) -> c_int;
    pub fn scir_b200_fir1d_batched_f64(ctx: *mut ScirB200Ctx, d_x: *const f64, ld_x: i64, taps: *const f64, k: i64, tap_order: c_int, d_y: *mut f64, ld_y: i64, batch: i64, n: i64) -> c_int;
    pub fn scir_b200_lfilter_fir_f32(ctx: *mut ScirB200Ctx, b: *const f32, k: i64, a0: f32, d_x: *const f32, ld_x: i64, d_zi: *const f32, d_zf: *mut f32, d_y: *mut f32, ld_y: i64, batch: i64, n: i64) -> c_int;
    pub fn scir_b200_upfirdn_out_len(len_h: i64, in_len: i64, up: i64, down: i64) -> i64;
    pub fn scir_b200_upfirdn_f32(ctx: *mut ScirB200Ctx, h: *const f32, len_h: i64, up: i64, down: i64, d_x: *const f32, ld_x: i64, batch: i64, n_in: i64, d_y: *mut f32, ld_y: i64, m_begin: i64, m_count: i64) -> c_int;
    pub fn scir_b200_upfirdn_mode_f32(ctx: *mut ScirB200Ctx, h: *const f32, len_h: i64, up: i64, down: i64, mode: c_int, cval: f32, d_x: *const f32, ld_x: i64, batch: i64, n_in: i64, d_y: *mut f32, ld_y: i64, m_begin: i64, m_count: i64) -> c_int;
    pub fn scir_b200_resample_poly_plan(n_in: i64, len_h: i64, up: i64, down: i64, plan: *mut ScirB200ResamplePlan) -> c_int;
    pub fn scir_b200_resample_poly_f32(ctx: *mut ScirB200Ctx, window: *const f32, len_h: i64, up: i64, down: i64, d_x: *const f32, ld_x: i64, batch: i64, n_in: i64, d_y: *mut f32, ld_y: i64) -> c_int;
    pub fn scir_b200_resample_poly_pad_f32(ctx: *mut ScirB200Ctx, window: *const f32, len_h: i64, up: i64, down: i64, padtype: c_int, cval: f32, d_x: *const f32, ld_x: i64, batch: i64, n_in: i64, d_y: *mut f32, ld_y: i64) -> c_int;
    pub fn scir_b200_resample_poly_f32_host(ctx: *mut ScirB200Ctx, window: *const f32, len_h: i64, up: i64, down: i64, h_x: *const f32, ld_x: i64, batch: i64, n_in: i64, h_y: *mut f32, ld_y: i64) -> c_int;
    pub fn scir_b200_filtfilt_fir_f32(ctx: *mut ScirB200Ctx, b: *const f32, k: i64, pad_mode: c_int, padlen: i64, d_x: *const f32, ld_x: i64, d_y: *mut f32, ld_y: i64, batch: i64, n: i64) -> c_int;
    pub fn scir_b200_filtfilt_fir_f32_host(ctx: *mut ScirB200Ctx, b: *const f32, k: i64, pad_mode: c_int, padlen: i64, h_x: *const f32, ld_x: i64, h_y: *mut f32, ld_y: i64, batch: i64, n: i64) -> c_int;
    pub fn scir_b200_upfirdn_mode_f64(ctx: *mut ScirB200Ctx, h: *const f64, len_h: i64, up: i64, down: i64, mode: c_int, cval: f64, d_x: *const f64, ld_x: i64, batch: i64, n_in: i64, d_y: *mut f64, ld_y: i64, m_begin: i64, m_count: i64) -> c_int;
    pub fn scir_b200_resample_poly_pad_f64(ctx: *mut ScirB200Ctx, window: *const f64, len_h: i64, up: i64, down: i64, padtype: c_int, cval: f64, d_x: *const f64, ld_x: i64, batch: i64, n_in: i64, d_y: *mut f64, ld_y: i64) -> c_int;
    pub fn scir_b200_filtfilt_fir_f64(ctx: *mut ScirB200Ctx, b: *const f64, k: i64, pad_mode: c_int, padlen: i64, d_x: *const f64, ld_x: i64, d_y: *mut f64, ld_y: i64, batch: i64, n: i64) -> c_int;
    pub fn scir_b200_add_scalar_f32(ctx: *mut ScirB200Ctx, d_a: *const f32, alpha: f32, d_y: *mut f32, n: i64) -> c_int;
    pub fn scir_b200_mul_scalar_f32(ctx: *mut ScirB200Ctx, d_a: *const f32, alpha: f32, d_y: *mut f32, n: i64) -> c_int;
    pub fn scir_b200_add_f32(ctx: *mut ScirB200Ctx, d_a: *const f32, d_b: *const f32, d_y: *mut f32, n: i64) -> c_int;
    pub fn scir_b200_mg_create(devices: *const c_int, n_devices: c_int, mg: *mut *mut ScirB200Mg) -> c_int;
    pub fn scir_b200_mg_destroy(mg: *mut ScirB200Mg) -> c_int;
    pub fn scir_b200_mg_device_count(mg: *const ScirB200Mg, n_devices: *mut c_int) -> c_int;
    pub fn scir_b200_mg_ctx(mg: *const ScirB200Mg, shard: c_int, ctx: *mut *mut ScirB200Ctx) -> c_int;
    pub fn scir_b200_mg_sync(mg: *mut ScirB200Mg) -> c_int;
    pub fn scir_b200_shard_rows(batch: i64, world: c_int, rank: c_int, row_begin: *mut i64, row_end: *mut i64) -> c_int;
    pub fn scir_b200_mg_fir1d_batched_f32_host(mg: *mut ScirB200Mg, h_x: *const f32, ld_x: i64, taps: *const f32, k: i64, tap_order: c_int, h_y: *mut f32, ld_y: i64, batch: i64, n: i64) -> c_int;
    pub fn scir_b200_mg_resample_poly_f32_host(mg: *mut ScirB200Mg, window: *const f32, len_h: i64, up: i64, down: i64, h_x: *const f32, ld_x: i64, batch: i64, n_in: i64, h_y: *mut f32, ld_y: i64) -> c_int;
    pub fn scir_b200_mg_filtfilt_fir_f32_host(mg: *mut ScirB200Mg, b: *const f32, k: i64, pad_mode: c_int, padlen: i64, h_x: *const f32, ld_x: i64, h_y: *mut f32, ld_y: i64, batch: i64, n: i64) -> c_int;
    pub fn scir_b200_mg_fir1d_batched_f32(mg: *mut ScirB200Mg, d_x: *const *const f32, ld_x: *const i64, taps: *const f32, k: i64, tap_order: c_int, d_y: *const *mut f32, ld_y: *const i64, batch: i64, n: i64) -> c_int;
    pub fn scir_b200_mg_gather_rows_f32(mg: *mut ScirB200Mg, d_shards: *const *const f32, ld_shards: *const i64, dst_shard: c_int, d_dst: *mut f32, ld_dst: i64, batch: i64, n: i64) -> c_int;
    pub fn scir_b200_microbench_ffma(ctx: *mut ScirB200Ctx, iters: c_int, tflops: *mut f64) -> c_int;
    pub fn scir_b200_microbench_ffma2(ctx: *mut ScirB200Ctx, iters: c_int, mix: c_int, tflops: *mut f64) -> c_int;
    pub fn scir_b200_microbench_copy(ctx: *mut ScirB200Ctx, bytes: usize, iters: c_int, gbps: *mut f64) -> c_int;
    pub fn scir_b200_microbench_pcie(ctx: *mut ScirB200Ctx, bytes: usize, iters: c_int, h2d_gbs: *mut f64, d2h_gbs: *mut f64, duplex_each_gbs: *mut f64) -> c_int;
}
