//! `rvc::RvcInfer` with the reference's method set (rvc/src/rvc.rs:18-220), backed by the B200 engine.
//! NOT compiled in the build image (no Rust toolchain); see INTEGRATION.md.
use ndarray::{Array1, Array3, ArrayView1};
use rvc_common::enums::{PitchAlgorithm, RvcModelVersion};
use rvc_cuda_sys as sys;
use std::ffi::{CStr, CString};
use std::path::PathBuf;

/// Replaces `rvc_common::errors::RvcInferError` (its `Ort(ort::Error)` variant cannot survive).
#[derive(Debug)]
pub enum RvcInferError {
    ModelNotLoaded,
    ContentvecNotLoaded,
    F0NotLoaded,
    Cuda(String),
    BadShape(String),
    Io(String),
    InvalidArg(String),
}

pub struct RvcInfer { ctx: *mut sys::rvc_ctx }
unsafe impl Send for RvcInfer {}

impl RvcInfer {
    fn err(&self, rc: i32) -> RvcInferError {
        let msg = unsafe { CStr::from_ptr(sys::rvc_last_error(self.ctx)) }.to_string_lossy().into_owned();
        match rc {
            sys::RVC_ERR_MODEL_NOT_LOADED => RvcInferError::ModelNotLoaded,
            sys::RVC_ERR_CONTENTVEC_NOT_LOADED => RvcInferError::ContentvecNotLoaded,
            sys::RVC_ERR_F0_NOT_LOADED => RvcInferError::F0NotLoaded,
            sys::RVC_ERR_BAD_SHAPE => RvcInferError::BadShape(msg),
            sys::RVC_ERR_IO => RvcInferError::Io(msg),
            sys::RVC_ERR_INVALID_ARG => RvcInferError::InvalidArg(msg),
            _ => RvcInferError::Cuda(msg),
        }
    }
    fn chk(&self, rc: i32) -> Result<(), RvcInferError> { if rc == sys::RVC_OK { Ok(()) } else { Err(self.err(rc)) } }

    pub fn new(data_path: PathBuf) -> Self {
        let p = CString::new(data_path.to_string_lossy().as_bytes()).unwrap();
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { sys::rvc_create(p.as_ptr(), std::ptr::null(), &mut ctx) };
        assert!(rc == sys::RVC_OK, "rvc_create failed: {}", unsafe { CStr::from_ptr(sys::rvc_last_create_error()) }.to_string_lossy());
        RvcInfer { ctx }
    }
    pub fn load_contentvec(&mut self, v: RvcModelVersion) -> Result<(), RvcInferError> {
        self.chk(unsafe { sys::rvc_load_contentvec(self.ctx, i64::from(v) as i32) })
    }
    pub fn load_model(&mut self, model_path: PathBuf) -> Result<(), RvcInferError> {
        let p = CString::new(model_path.to_string_lossy().as_bytes()).unwrap();
        self.chk(unsafe { sys::rvc_load_model(self.ctx, p.as_ptr()) })
    }
    pub fn load_f0(&mut self, a: PitchAlgorithm) -> Result<(), RvcInferError> {
        self.chk(unsafe { sys::rvc_load_f0(self.ctx, i64::from(a) as i32) })
    }
    pub fn unload_model(&mut self) { unsafe { sys::rvc_unload_model(self.ctx) }; }

    pub fn hubert(&self, input: ArrayView1<f32>) -> Result<Array3<f32>, RvcInferError> {
        let x = input.to_owned();
        let mut out = vec![0f32; 1024 * (x.len() / 320 + 2)];
        let (mut c, mut t) = (0usize, 0usize);
        self.chk(unsafe { sys::rvc_hubert(self.ctx, x.as_ptr(), x.len(), out.as_mut_ptr(), out.len(), &mut c, &mut t) })?;
        out.truncate(c * t);
        Ok(Array3::from_shape_vec((1, c, t), out).unwrap())
    }
    pub fn extract_feature(&self, input: ArrayView1<f32>) -> Result<Array3<f32>, RvcInferError> {
        let x = input.to_owned();
        let mut out = vec![0f32; 1024 * (2 * (x.len() / 320 + 2) + 1)];
        let (mut f, mut c) = (0usize, 0usize);
        self.chk(unsafe { sys::rvc_extract_feature(self.ctx, x.as_ptr(), x.len(), out.as_mut_ptr(), out.len(), &mut f, &mut c) })?;
        out.truncate(f * c);
        Ok(Array3::from_shape_vec((1, f, c), out).unwrap())
    }
    pub fn pitch(&mut self, input: ArrayView1<f32>, pitch_shift: i32, sample_frame_16k_size: usize) -> Result<Array1<f32>, RvcInferError> {
        let x = input.to_owned();
        let mut out = vec![0f32; 4096];
        let mut n = 0usize;
        self.chk(unsafe { sys::rvc_pitch(self.ctx, x.as_ptr(), x.len(), pitch_shift, sample_frame_16k_size, out.as_mut_ptr(), out.len(), &mut n) })?;
        out.truncate(n);
        Ok(Array1::from_vec(out))
    }
    pub fn infer(&mut self, input: ArrayView1<f32>, sample_frame_16k_size: usize, pitch_shift: Option<i32>, skip_head: u32,
                 return_length: u32) -> Result<Array1<f32>, RvcInferError> {
        let x = input.to_owned();
        let mut out = vec![0f32; return_length as usize * 480 + 16];
        let mut n = 0usize;
        self.chk(unsafe {
            sys::rvc_infer(self.ctx, x.as_ptr(), x.len(), sample_frame_16k_size as u32, pitch_shift.unwrap_or(0), skip_head, return_length,
                           out.as_mut_ptr(), out.len(), &mut n)
        })?;
        out.truncate(n);
        Ok(Array1::from_vec(out))
    }
}

impl RvcInfer {
    /// Retrieval (the reference keeps `index_path` / `index_rate` in its settings but never uses them: lib.rs:78,81,264,
    /// TODO at rvc.rs:159): exact top-k over the index rows blended into the features at `index_rate`.
    pub fn load_index(&mut self, index_path: PathBuf, index_rate: f32) -> Result<(), RvcInferError> {
        let p = CString::new(index_path.to_string_lossy().as_bytes()).unwrap();
        self.chk(unsafe { sys::rvc_load_index(self.ctx, p.as_ptr(), index_rate) })
    }
    pub fn set_index_rate(&mut self, index_rate: f32) -> Result<(), RvcInferError> {
        self.chk(unsafe { sys::rvc_set_index_rate(self.ctx, index_rate) })
    }
    /// Offline conversion: `n_windows` consecutive windows of this stream, up to 32 per launch (BASELINE configs[2]).
    pub fn infer_windows(&mut self, pcm: ArrayView1<f32>, n: usize, sample_frame_16k_size: usize, n_windows: usize, pitch_shift: i32,
                         skip_head: u32, return_length: u32) -> Result<Array1<f32>, RvcInferError> {
        let x = pcm.to_owned();
        let mut out = vec![0f32; n_windows * (return_length as usize * 480 + 16)];
        let mut audio_len = 0usize;
        self.chk(unsafe {
            sys::rvc_infer_windows(self.ctx, x.as_ptr(), x.len(), n, sample_frame_16k_size as u32, n_windows, pitch_shift, skip_head,
                                   return_length, out.as_mut_ptr(), out.len(), &mut audio_len, 0)
        })?;
        out.truncate(n_windows * audio_len);
        Ok(Array1::from_vec(out))
    }
}

impl Drop for RvcInfer { fn drop(&mut self) { unsafe { sys::rvc_destroy(self.ctx) } } }
