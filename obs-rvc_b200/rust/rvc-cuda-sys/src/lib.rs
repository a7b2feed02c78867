//! Raw FFI of `librvc_b200.so` - GENERATED from `include/rvc_b200.h` by tools/gen_rust_sys.py; do not edit.
//! NOT compiled in the build image (no Rust toolchain); the crate a maintainer adds to the reference workspace.
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
pub const RVC_MODEL_V1: c_int = 1;
pub const RVC_MODEL_V2: c_int = 2;
pub const RVC_PITCH_RMVPE: c_int = 1;

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

#[repr(C)]
#[derive(Clone, Copy)]
pub struct rvc_stream_config {
    pub sample_rate: u32,
    pub pitch_shift: i32,
    pub sample_length: f64,
    pub crossfade_length: f64,
    pub extra_inference_time: f64,
    pub rms_mix_rate: f64,
    pub skip_inference: i32,
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
    pub fn rvc_hubert(ctx: *mut rvc_ctx, pcm: *const c_float, n: usize, out: *mut c_float, cap: usize, out_c: *mut usize, out_t: *mut usize) -> c_int;
    pub fn rvc_extract_feature(ctx: *mut rvc_ctx, pcm: *const c_float, n: usize, out: *mut c_float, cap: usize, out_frames: *mut usize, out_c: *mut usize) -> c_int;
    pub fn rvc_pitch(ctx: *mut rvc_ctx, pcm: *const c_float, n: usize, pitch_shift: i32, sample_frame_16k_size: usize, out: *mut c_float, cap: usize, out_len: *mut usize) -> c_int;
    pub fn rvc_infer(ctx: *mut rvc_ctx, pcm: *const c_float, n: usize, sample_frame_16k_size: u32, pitch_shift: i32, skip_head: u32, return_length: u32, out: *mut c_float, cap: usize, out_len: *mut usize) -> c_int;
    pub fn rvc_infer_dev(ctx: *mut rvc_ctx, pcm_dev: *const c_float, n: usize, sample_frame_16k_size: u32, pitch_shift: i32, skip_head: u32, return_length: u32, out_dev: *mut c_float, cap: usize, out_len: *mut usize) -> c_int;
    pub fn rvc_infer_batch(ctxs: *const *mut rvc_ctx, n_ctx: usize, pcm: *const *const c_float, n: usize, sample_frame_16k_size: u32, pitch_shift: i32, skip_head: u32, return_length: u32, out: *const *mut c_float, cap: usize, out_len: *mut usize) -> c_int;
    pub fn rvc_infer_batch_dev(ctxs: *const *mut rvc_ctx, n_ctx: usize, pcm_dev: *const *const c_float, n: usize, sample_frame_16k_size: u32, pitch_shift: i32, skip_head: u32, return_length: u32, out_dev: *const *mut c_float, cap: usize, out_len: *mut usize) -> c_int;
    pub fn rvc_infer_windows(ctx: *mut rvc_ctx, pcm: *const c_float, n_pcm: usize, n: usize, sample_frame_16k_size: u32, n_windows: usize, pitch_shift: i32, skip_head: u32, return_length: u32, out: *mut c_float, cap: usize, audio_len: *mut usize, max_batch: i32) -> c_int;
    pub fn rvc_infer_windows_dev(ctx: *mut rvc_ctx, pcm_dev: *const c_float, n_pcm: usize, n: usize, sample_frame_16k_size: u32, n_windows: usize, pitch_shift: i32, skip_head: u32, return_length: u32, out_dev: *mut c_float, cap: usize, audio_len: *mut usize, max_batch: i32) -> c_int;
    pub fn rvc_mel_extract(ctx: *mut rvc_ctx, pcm: *const c_float, n: usize, out: *mut c_float, cap: usize, out_frames: *mut usize) -> c_int;
    pub fn rvc_decode_salience(ctx: *mut rvc_ctx, salience: *const c_float, t_frames: usize, f0_out: *mut c_float, argmax_out: *mut i32) -> c_int;
    pub fn rvc_knn_search(ctx: *mut rvc_ctx, queries: *const c_float, q: usize, c: usize, k: i32, d2: *mut c_float, idx: *mut i32) -> c_int;
    pub fn rvc_knn_fallbacks(ctx: *mut rvc_ctx, total: *mut u64) -> c_int;
    pub fn rvc_stream_config_default(cfg: *mut rvc_stream_config);
    pub fn rvc_stream_open(ctx: *mut rvc_ctx, cfg: *const rvc_stream_config, sample_frame_size: *mut u32) -> c_int;
    pub fn rvc_stream_close(ctx: *mut rvc_ctx) -> c_int;
    pub fn rvc_stream_set(ctx: *mut rvc_ctx, pitch_shift: i32, rms_mix_rate: f64) -> c_int;
    pub fn rvc_stream_info(ctx: *mut rvc_ctx, out: *mut c_char, cap_bytes: usize, out_bytes: *mut usize) -> c_int;
    pub fn rvc_process_frame(ctx: *mut rvc_ctx, input: *const c_float, output: *mut c_float, sola_offset: *mut u32) -> c_int;
    pub fn rvc_resample_chunk(ctx: *mut rvc_ctx, fs_in: u32, fs_out: u32, r#in: *const c_float, n_in: usize, overlap_inout: *mut c_float, out: *mut c_float, cap: usize, n_out: *mut usize) -> c_int;
    pub fn rvc_envelop_mixing(ctx: *mut rvc_ctx, input: *const c_float, n_in: usize, output: *mut c_float, n_out: usize, sample_rate: u32, mix_rate: f64, rms1: *mut c_float, rms2: *mut c_float) -> c_int;
    pub fn rvc_sola_offset(ctx: *mut rvc_ctx, input_buffer: *const c_float, n: usize, sola_buffer: *const c_float, buffer_frame_size: u32, search_frame_size: u32, offset: *mut u32) -> c_int;
    pub fn rvc_sola_crossfade(ctx: *mut rvc_ctx, infer_out: *const c_float, n: usize, sola_buffer: *mut c_float, buffer_frame_size: u32, search_frame_size: u32, sample_frame_size: u32, block_out: *mut c_float, offset: *mut u32) -> c_int;
    pub fn rvc_get_last(ctx: *mut rvc_ctx, name: *const c_char, out: *mut c_void, cap_bytes: usize, out_bytes: *mut usize) -> c_int;
    pub fn rvc_get_last_window(ctx: *mut rvc_ctx, window: i32, name: *const c_char, out: *mut c_void, cap_bytes: usize, out_bytes: *mut usize) -> c_int;
    pub fn rvc_debug_tensor(ctx: *mut rvc_ctx, name: *const c_char, out: *mut c_float, cap: usize, out_len: *mut usize) -> c_int;
    pub fn rvc_debug_list(ctx: *mut rvc_ctx, out: *mut c_char, cap_bytes: usize, out_bytes: *mut usize) -> c_int;
    pub fn rvc_reset_state(ctx: *mut rvc_ctx) -> c_int;
    pub fn rvc_sync(ctx: *mut rvc_ctx) -> c_int;
    pub fn rvc_cuda_stream(ctx: *mut rvc_ctx) -> *mut c_void;
    pub fn rvc_kernel_launches(ctx: *mut rvc_ctx, total: *mut u64) -> c_int;
    pub fn rvc_plan_info(ctx: *mut rvc_ctx, out: *mut c_char, cap_bytes: usize, out_bytes: *mut usize) -> c_int;
    pub fn rvc_event_record(ctx: *mut rvc_ctx, slot: c_int) -> c_int;
    pub fn rvc_event_elapsed_ms(ctx: *mut rvc_ctx, slot_a: c_int, slot_b: c_int, ms: *mut c_float) -> c_int;
    pub fn rvc_profile_ops(ctx: *mut rvc_ctx, iters: c_int, out: *mut c_char, cap_bytes: usize, out_bytes: *mut usize) -> c_int;
    pub fn rvc_profile_timeline(ctx: *mut rvc_ctx, out: *mut c_char, cap_bytes: usize, out_bytes: *mut usize) -> c_int;
    pub fn rvc_profile_chains(ctx: *mut rvc_ctx, out: *mut c_char, cap_bytes: usize, out_bytes: *mut usize) -> c_int;
    pub fn rvc_version() -> *const c_char;
}
