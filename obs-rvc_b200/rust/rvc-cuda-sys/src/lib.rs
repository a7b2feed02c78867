//! Raw FFI of `librvc_b200.so` - a mechanical mirror of `include/rvc_b200.h`.
//! NOT compiled in the build image (no Rust toolchain); kept thin on purpose.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_float, c_int, c_void};

#[repr(C)]
pub struct rvc_ctx { _private: [u8; 0] }

pub const RVC_OK: c_int = 0;
pub const RVC_ERR_MODEL_NOT_LOADED: c_int = 1;
pub const RVC_ERR_CONTENTVEC_NOT_LOADED: c_int = 2;
pub const RVC_ERR_F0_NOT_LOADED: c_int = 3;
pub const RVC_ERR_CUDA: c_int = 4;
pub const RVC_ERR_BAD_SHAPE: c_int = 5;
pub const RVC_ERR_IO: c_int = 6;
pub const RVC_ERR_INVALID_ARG: c_int = 7;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct rvc_config {
    pub device: i32,
    pub noise_mode: i32,
    pub noise_seed: u64,
    pub index_k: i32,
    pub upstream_pitch_shift: i32,
    pub upstream_cents_window: i32,
    pub use_cuda_graph: i32,
    pub debug_keep: i32,
    pub reserved: [i32; 7],
}

extern "C" {
    pub fn rvc_config_default(cfg: *mut rvc_config);
    pub fn rvc_create(data_path: *const c_char, cfg: *const rvc_config, out: *mut *mut rvc_ctx) -> c_int;
    pub fn rvc_destroy(ctx: *mut rvc_ctx);
    pub fn rvc_last_error(ctx: *const rvc_ctx) -> *const c_char;
    pub fn rvc_last_create_error() -> *const c_char;
    pub fn rvc_load_contentvec(ctx: *mut rvc_ctx, model_version: i32) -> c_int;
    pub fn rvc_load_f0(ctx: *mut rvc_ctx, pitch_algorithm: i32) -> c_int;
    pub fn rvc_load_model(ctx: *mut rvc_ctx, model_path: *const c_char) -> c_int;
    pub fn rvc_unload_model(ctx: *mut rvc_ctx) -> c_int;
    pub fn rvc_load_index(ctx: *mut rvc_ctx, index_path: *const c_char, index_rate: c_float) -> c_int;
    pub fn rvc_set_index(ctx: *mut rvc_ctx, rows: *const c_float, n: usize, c: usize, index_rate: c_float) -> c_int;
    pub fn rvc_set_index_rate(ctx: *mut rvc_ctx, index_rate: c_float) -> c_int;
    pub fn rvc_hubert(ctx: *mut rvc_ctx, pcm: *const c_float, n: usize, out: *mut c_float, cap: usize,
                      out_c: *mut usize, out_t: *mut usize) -> c_int;
    pub fn rvc_extract_feature(ctx: *mut rvc_ctx, pcm: *const c_float, n: usize, out: *mut c_float, cap: usize,
                               out_frames: *mut usize, out_c: *mut usize) -> c_int;
    pub fn rvc_pitch(ctx: *mut rvc_ctx, pcm: *const c_float, n: usize, pitch_shift: i32, sample_frame_16k_size: usize,
                     out: *mut c_float, cap: usize, out_len: *mut usize) -> c_int;
    pub fn rvc_infer(ctx: *mut rvc_ctx, pcm: *const c_float, n: usize, sample_frame_16k_size: u32, pitch_shift: i32,
                     skip_head: u32, return_length: u32, out: *mut c_float, cap: usize, out_len: *mut usize) -> c_int;
    pub fn rvc_reset_state(ctx: *mut rvc_ctx) -> c_int;
    pub fn rvc_cuda_stream(ctx: *mut rvc_ctx) -> *mut c_void;
}
