//! Seam 2: drop-in replacement of `obs-rvc/src/rvcadapter.rs` (reference lines 8-126).
//!
//! The reference adapter spawns `rvc-rpc.exe` and speaks a length-prefixed little-endian protocol over its pipes
//! (rvcadapter.rs:33-58 spawn, :60-120 infer, :122-126 kill on drop).  This one keeps the type names, the constructor
//! and `infer` signatures and the error enum the OBS filter uses (`obs-rvc/src/lib.rs:247-254, 701-727`), but calls the
//! B200 engine in-process through `rvc-cuda-sys` - no child process, no pipe copies.  `binary_path` is accepted and
//! ignored so that `lib.rs` compiles unchanged.  NOT compiled in the build image (no Rust toolchain).
use std::ffi::{CStr, CString};
use std::path::PathBuf;

use ndarray::Array1;
use rvc_common::{enums::{PitchAlgorithm, RvcModelVersion}, errors::RvcInferError};
use rvc_cuda_sys as sys;

pub struct RvcInfer {
    ctx: *mut sys::rvc_ctx,
    frame: Option<u32>,      // sample_frame_size of the open device-resident stream (process_frame)
}
// the OBS filter moves the engine into its worker thread (lib.rs:585-600); the context is used by one thread at a time
unsafe impl Send for RvcInfer {}

#[derive(Debug)]
pub enum RvcAdapterError {
    RvcInferError(RvcInferError),
    IoError(std::io::Error),
}
impl From<RvcInferError> for RvcAdapterError { fn from(e: RvcInferError) -> Self { RvcAdapterError::RvcInferError(e) } }
impl From<std::io::Error> for RvcAdapterError { fn from(e: std::io::Error) -> Self { RvcAdapterError::IoError(e) } }

fn map_err(ctx: *mut sys::rvc_ctx, rc: i32) -> RvcAdapterError {
    let msg = unsafe { CStr::from_ptr(sys::rvc_last_error(ctx)) }.to_string_lossy().into_owned();
    match rc {
        sys::RVC_ERR_MODEL_NOT_LOADED => RvcInferError::ModelNotLoaded.into(),
        sys::RVC_ERR_CONTENTVEC_NOT_LOADED => RvcInferError::ContentvecNotLoaded.into(),
        sys::RVC_ERR_F0_NOT_LOADED => RvcInferError::F0NotLoaded.into(),
        // the reference restarts the engine on IoError (lib.rs:716-720): a CUDA failure gets the same treatment
        sys::RVC_ERR_CUDA | sys::RVC_ERR_IO => std::io::Error::new(std::io::ErrorKind::Other, msg).into(),
        // shape / argument errors were panics inside rvc-rpc, i.e. a broken pipe on this side
        _ => std::io::Error::new(std::io::ErrorKind::InvalidInput, msg).into(),
    }
}

impl RvcInfer {
    /// rvcadapter.rs:34-58 + rvc-rpc/src/main.rs:33-54 (the child created the engine and loaded the three models).
    pub fn new(_binary_path: PathBuf, model_version: RvcModelVersion, pitch_algorithm: PitchAlgorithm, model_path: PathBuf,
               data_path: PathBuf) -> Self {
        let data = CString::new(data_path.to_string_lossy().as_bytes()).unwrap();
        let model = CString::new(model_path.to_string_lossy().as_bytes()).unwrap();
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { sys::rvc_create(data.as_ptr(), std::ptr::null(), &mut ctx) };
        assert!(rc == sys::RVC_OK, "rvc_create: {}", unsafe { CStr::from_ptr(sys::rvc_last_create_error()) }.to_string_lossy());
        // rvc-rpc panicked (`.unwrap()`) when a model failed to load: keep that behaviour, it is what `lib.rs` expects
        for (what, rc) in [
            ("load_contentvec", unsafe { sys::rvc_load_contentvec(ctx, i64::from(model_version) as i32) }),
            ("load_f0", unsafe { sys::rvc_load_f0(ctx, i64::from(pitch_algorithm) as i32) }),
            ("load_model", unsafe { sys::rvc_load_model(ctx, model.as_ptr()) }),
        ] {
            assert!(rc == sys::RVC_OK, "{what}: {}", unsafe { CStr::from_ptr(sys::rvc_last_error(ctx)) }.to_string_lossy());
        }
        RvcInfer { ctx, frame: None }
    }

    /// Index settings the reference stores but never uses (lib.rs:78,81,264; TODO at rvc.rs:159).
    pub fn load_index(&mut self, index_path: PathBuf, index_rate: f32) -> Result<(), RvcAdapterError> {
        let p = CString::new(index_path.to_string_lossy().as_bytes()).unwrap();
        let rc = unsafe { sys::rvc_load_index(self.ctx, p.as_ptr(), index_rate) };
        if rc == sys::RVC_OK { Ok(()) } else { Err(map_err(self.ctx, rc)) }
    }

    /// rvcadapter.rs:60-120: same arguments, same result; the wire format is gone.
    pub fn infer(&mut self, input: ndarray::ArrayView1<f32>, sample_frame_16k_size: usize, pitch_shift: i32, skip_head: u32,
                 return_length: u32) -> Result<ndarray::Array1<f32>, RvcAdapterError> {
        let x = input.as_standard_layout();
        let mut out = vec![0f32; return_length as usize * 480 + 16];
        let mut n = 0usize;
        let rc = unsafe {
            sys::rvc_infer(self.ctx, x.as_ptr(), x.len(), sample_frame_16k_size as u32, pitch_shift, skip_head, return_length,
                           out.as_mut_ptr(), out.len(), &mut n)
        };
        if rc != sys::RVC_OK { return Err(map_err(self.ctx, rc)); }
        out.truncate(n);
        Ok(Array1::from_vec(out))
    }

    /// Optional fast path for `process_one_frame` (lib.rs:659-795): the whole frame - ring buffers, both rubato
    /// resamplers, infer, envelope mixing, SOLA, cross-fade - runs on the device; one copy in, one copy out.
    /// `lib.rs` would call `open_stream` where it builds `RvcInferenceState` (:186-300) and `process_frame` in place
    /// of the body of `process_one_frame`.
    pub fn open_stream(&mut self, sample_rate: u32, sample_length: f64, crossfade_length: f64, extra_inference_time: f64,
                       pitch_shift: i32, rms_mix_rate: f64, skip_inference: bool) -> Result<usize, RvcAdapterError> {
        let mut cfg = unsafe { std::mem::zeroed::<sys::rvc_stream_config>() };
        unsafe { sys::rvc_stream_config_default(&mut cfg) };
        cfg.sample_rate = sample_rate; cfg.sample_length = sample_length; cfg.crossfade_length = crossfade_length;
        cfg.extra_inference_time = extra_inference_time; cfg.pitch_shift = pitch_shift; cfg.rms_mix_rate = rms_mix_rate;
        cfg.skip_inference = skip_inference as i32;
        let mut frame = 0u32;
        let rc = unsafe { sys::rvc_stream_open(self.ctx, &cfg, &mut frame) };
        if rc != sys::RVC_OK { return Err(map_err(self.ctx, rc)); }
        self.frame = Some(frame);
        Ok(frame as usize)
    }

    pub fn process_frame(&mut self, input_sample: &[f32]) -> Result<Array1<f32>, RvcAdapterError> {
        let frame = self.frame.expect("open_stream first") as usize;
        assert_eq!(input_sample.len(), frame);
        let mut out = vec![0f32; frame];
        let rc = unsafe { sys::rvc_process_frame(self.ctx, input_sample.as_ptr(), out.as_mut_ptr(), std::ptr::null_mut()) };
        if rc != sys::RVC_OK { return Err(map_err(self.ctx, rc)); }
        Ok(Array1::from_vec(out))
    }
}

impl Drop for RvcInfer {
    /// rvcadapter.rs:122-126 killed the child; here the context is destroyed.
    fn drop(&mut self) { unsafe { sys::rvc_destroy(self.ctx) } }
}
