/*
 * rvc_b200.h - C ABI of the B200-native RVC inference engine (librvc_b200.so).
 *
 * Drop-in boundary for the per-audio-window hot path of RVC-Project/obs-rvc.  Every entry
 * point replaces one item of the reference's `rvc` crate public API (`pub use rvc::*`,
 * rvc/src/lib.rs:5) - the interface `rvc-rpc` (rvc-rpc/src/main.rs:33-54,93) and, through the
 * adapter (obs-rvc/src/rvcadapter.rs:34,60-67), the OBS audio-filter worker thread
 * (obs-rvc/src/lib.rs:701-707) call.  Citations are relative to the reference tree.
 *
 * Conventions
 *   - plain pointers and sizes; caller-allocated outputs (`cap` in elements) + `*out_len`;
 *   - `int` status, 0 = OK; no exceptions or aborts cross the ABI (the reference panics /
 *     unwraps: rvc-rpc/src/main.rs:66-100);
 *   - a context is NOT thread-safe (`infer`/`pitch` take `&mut self`, rvc.rs:111,133);
 *     distinct contexts are independent: each owns a CUDA stream and its pitch cache;
 *   - all PCM is 16 kHz mono float32 host memory unless the name ends in `_dev`.
 */
#ifndef RVC_B200_H
#define RVC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rvc_ctx rvc_ctx;

/* rvc-common/src/errors.rs:1-20 `RvcInferError` (+ the cases where the reference panics) */
enum rvc_status {
    RVC_OK = 0,
    RVC_ERR_MODEL_NOT_LOADED = 1,      /* RvcInferError::ModelNotLoaded      (rvc.rs:141-143) */
    RVC_ERR_CONTENTVEC_NOT_LOADED = 2, /* RvcInferError::ContentvecNotLoaded (rvc.rs:85-88)   */
    RVC_ERR_F0_NOT_LOADED = 3,         /* RvcInferError::F0NotLoaded (rvc.rs:127 unreachable!) */
    RVC_ERR_CUDA = 4,                  /* replaces RvcInferError::Ort(ort::Error)             */
    RVC_ERR_BAD_SHAPE = 5,             /* RvcInferError::NdarrayShapeError / slice panics     */
    RVC_ERR_IO = 6,                    /* model file missing / malformed (ort::Error on load) */
    RVC_ERR_INVALID_ARG = 7
};

/* rvc-common/src/enums.rs:3-28: integer values as `impl From<_> for i64` (enums.rs:32-39,97-103) */
enum rvc_model_version { RVC_MODEL_V1 = 1, RVC_MODEL_V2 = 2 };
enum rvc_pitch_algorithm { RVC_PITCH_RMVPE = 1 };

/* Plain config struct passed at create (replaces ORT session options, rvc/src/models.rs:7-46). */
typedef struct rvc_config {
    int32_t  device;                /* CUDA device ordinal                                      */
    int32_t  noise_mode;            /* 1 = counter-based Gaussian noise (default), 0 = zeros     */
    uint64_t noise_seed;            /* seed of the synthesizer noise (see oracle/noise.py)       */
    int32_t  index_k;               /* neighbours per query for retrieval (default 8)            */
    int32_t  upstream_pitch_shift;  /* 0 = reference-literal integer octaves (rvc.rs:121)        */
    int32_t  upstream_cents_window; /* 0 = reference-literal window (rmvpe.rs:119-125)           */
    int32_t  use_cuda_graph;        /* 1 = replay a captured graph per geometry (default)        */
    int32_t  debug_keep;            /* 1 = keep every intermediate buffer (rvc_debug_tensor)     */
    int32_t  reserved[7];
} rvc_config;

void rvc_config_default(rvc_config* cfg);

/* RvcInfer::new(data_path) - rvc.rs:30-44.  `cfg` may be NULL. */
int rvc_create(const char* data_path, const rvc_config* cfg, rvc_ctx** out);
/* Drop for RvcInfer (sessions freed) */
void rvc_destroy(rvc_ctx* ctx);
/* human-readable text of the last error on this context (never NULL) */
const char* rvc_last_error(const rvc_ctx* ctx);
/* same, for failures of rvc_create (thread-local) */
const char* rvc_last_create_error(void);

/* RvcInfer::load_contentvec(RvcModelVersion) - rvc.rs:46-54; file
 * <data>/contentvec/vec-{256|768}-layer-{9|12}.rvcw (models.rs:58-61 with .rvcw for .onnx) */
int rvc_load_contentvec(rvc_ctx* ctx, int32_t model_version);
/* RvcInfer::load_f0(PitchAlgorithm) - rvc.rs:62-75; file <data>/f0/rmvpe.rvcw (models.rs:71-75) */
int rvc_load_f0(rvc_ctx* ctx, int32_t pitch_algorithm);
/* RvcInfer::load_model(PathBuf) - rvc.rs:56-60 */
int rvc_load_model(rvc_ctx* ctx, const char* model_path);
/* RvcInfer::unload_model() - rvc.rs:77-79 */
int rvc_unload_model(rvc_ctx* ctx);

/* Retrieval index: no reference counterpart - the reference stores `index_path`/`index_rate`
 * (obs-rvc/src/lib.rs:78,81,264) and leaves `// TODO: index search` (rvc.rs:159).
 * `rows` is N x C float32 row-major (FAISS `big_npy`); C must equal the ContentVec width. */
int rvc_load_index(rvc_ctx* ctx, const char* index_path, float index_rate);
int rvc_set_index(rvc_ctx* ctx, const float* rows, size_t n, size_t c, float index_rate);
int rvc_set_index_rate(rvc_ctx* ctx, float index_rate);

/* RvcInfer::hubert(ArrayView1<f32>) -> Array3 (1,C,T) - rvc.rs:81-97.
 * out receives C*T floats in (C,T) row-major order. */
int rvc_hubert(rvc_ctx* ctx, const float* pcm, size_t n, float* out, size_t cap,
               size_t* out_c, size_t* out_t);
/* RvcInfer::extract_feature -> (1, 2T+1, C) - rvc.rs:99-109. out: (2T+1, C) row-major. */
int rvc_extract_feature(rvc_ctx* ctx, const float* pcm, size_t n, float* out, size_t cap,
                        size_t* out_frames, size_t* out_c);
/* RvcInfer::pitch(input, pitch_shift, sample_frame_16k_size) -> f0[T] Hz - rvc.rs:111-131 */
int rvc_pitch(rvc_ctx* ctx, const float* pcm, size_t n, int32_t pitch_shift,
              size_t sample_frame_16k_size, float* out, size_t cap, size_t* out_len);
/* RvcInfer::infer(input, sample_frame_16k_size, pitch_shift, skip_head, return_length)
 * -> audio - rvc.rs:133-220; wire order of rvcadapter.rs:60-67.  Advances the pitch cache. */
int rvc_infer(rvc_ctx* ctx, const float* pcm, size_t n, uint32_t sample_frame_16k_size,
              int32_t pitch_shift, uint32_t skip_head, uint32_t return_length, float* out,
              size_t cap, size_t* out_len);
/* Same call with PCM and audio already resident in device memory of ctx's GPU (no copies;
 * asynchronous on the context stream - call rvc_sync before reading `out_dev`). */
int rvc_infer_dev(rvc_ctx* ctx, const float* pcm_dev, size_t n, uint32_t sample_frame_16k_size,
                  int32_t pitch_shift, uint32_t skip_head, uint32_t return_length,
                  float* out_dev, size_t cap, size_t* out_len);
/* Independent live streams in one call (the path shards by stream; BASELINE configs[3]:
 * 8 streams per GPU).  When the contexts sit on one device and share their models / index (weights are
 * shared by path) the windows run as ONE batched plan - every kernel processes all streams, weights are
 * read once per round - with per-stream state (pitch cache rvc.rs:26,168-179, call counter, noise seed)
 * kept in each context; otherwise each window is enqueued on its own context stream.  Results are
 * identical to n_ctx separate rvc_infer calls. */
int rvc_infer_batch(rvc_ctx* const* ctxs, size_t n_ctx, const float* const* pcm, size_t n,
                    uint32_t sample_frame_16k_size, int32_t pitch_shift, uint32_t skip_head,
                    uint32_t return_length, float* const* out, size_t cap, size_t* out_len);
/* Same with device-resident PCM / audio pointers; asynchronous on ctxs[0]'s stream (rvc_sync(ctxs[0])).
 * Fails with RVC_ERR_INVALID_ARG when the streams cannot share one batched plan. */
int rvc_infer_batch_dev(rvc_ctx* const* ctxs, size_t n_ctx, const float* const* pcm_dev, size_t n,
                        uint32_t sample_frame_16k_size, int32_t pitch_shift, uint32_t skip_head,
                        uint32_t return_length, float* const* out_dev, size_t cap, size_t* out_len);
/* Offline conversion of ONE stream (BASELINE configs[2]: batch = 32): `n_windows` consecutive windows,
 * window w = pcm[w * sample_frame_16k_size, w * sample_frame_16k_size + n) - exactly what n_windows
 * successive RvcInfer::infer calls of the reference's streaming loop see (obs-rvc/src/lib.rs:659-707) -
 * processed `max_batch` (<= 32, 0 = 32) windows per launch: each kernel runs over the whole group, the
 * weights are read once per group, the pitch cache is updated window by window inside the plan.
 * out: n_windows x audio_len samples; results identical to the n_windows single calls. */
int rvc_infer_windows(rvc_ctx* ctx, const float* pcm, size_t n_pcm, size_t n, uint32_t sample_frame_16k_size,
                      size_t n_windows, int32_t pitch_shift, uint32_t skip_head, uint32_t return_length,
                      float* out, size_t cap, size_t* audio_len, int32_t max_batch);
int rvc_infer_windows_dev(rvc_ctx* ctx, const float* pcm_dev, size_t n_pcm, size_t n, uint32_t sample_frame_16k_size,
                          size_t n_windows, int32_t pitch_shift, uint32_t skip_head, uint32_t return_length,
                          float* out_dev, size_t cap, size_t* audio_len, int32_t max_batch);

/* MelSpectrogram::mel_extract - rmvpe.rs:159-205. out: (128, T) row-major, T = 1 + n/160. */
int rvc_mel_extract(rvc_ctx* ctx, const float* pcm, size_t n, float* out, size_t cap,
                    size_t* out_frames);
/* Rmvpe::decode + to_local_average_cents (rmvpe.rs:118-133, 243-248) on caller-supplied salience rows (t_frames, 360):
 * the decode stage of `pitch` without the network (argmax per frame, 9-tap cents window in the configured convention,
 * 0.03 threshold, f0 = 10 * 2^(cents / 1200), unvoiced -> 0).  f0_out / argmax_out: t_frames entries. */
int rvc_decode_salience(rvc_ctx* ctx, const float* salience, size_t t_frames, float* f0_out, int32_t* argmax_out);
/* Exact brute-force L2 top-k on the loaded index (stress config 5). d2/idx: (q, k). */
int rvc_knn_search(rvc_ctx* ctx, const float* queries, size_t q, size_t c, int32_t k,
                   float* d2, int32_t* idx);
/* Index widths C % 64 == 0, C <= 256 (v1 features, configs[4]) with k <= 8 run a tcgen05 candidate pass followed by an exact
 * fp32 re-rank whose guard proves the top-k; a query whose guard fails is recomputed by an exact scan of all rows.
 * Counts those recomputations since the index was set (0 on every test and bench input). */
int rvc_knn_fallbacks(rvc_ctx* ctx, uint64_t* total);

/* ---- the streaming loop around the call, device-resident (obs-rvc/src/lib.rs:186-300 state, :659-795 process_one_frame) ----
 * rvc_stream_open derives RvcInferenceState's sizes from the OBS settings (lib.rs:200-226), builds the two rubato
 * FftFixedInOut resamplers (lib.rs:236-242; OBS rate -> 16 kHz, model rate -> OBS rate) and the ring buffers on the
 * device.  rvc_process_frame is process_one_frame: append the new block, down-sample, RvcInfer::infer, up-sample,
 * envelope mixing (rms_mix_rate < 1), SOLA offset + sin^2 cross-fade - ONE host-to-device copy of sample_frame_size
 * samples in, ONE device-to-host copy of sample_frame_size samples out.  One open stream per context. */
typedef struct rvc_stream_config {
    uint32_t sample_rate;            /* OBS audio rate (lib.rs:186) */
    int32_t pitch_shift;             /* lib.rs:262 */
    double sample_length;            /* seconds, lib.rs:188 (default 0.30) */
    double crossfade_length;         /* lib.rs:189 (0.07) */
    double extra_inference_time;     /* lib.rs:190 (2.00) */
    double rms_mix_rate;             /* lib.rs:265 (0.00): envelope mixing runs while < 1 */
    int32_t skip_inference;          /* lib.rs:198: the 16 kHz tail is passed through instead of the model */
    int32_t reserved[7];
} rvc_stream_config;
void rvc_stream_config_default(rvc_stream_config* cfg);
int rvc_stream_open(rvc_ctx* ctx, const rvc_stream_config* cfg, uint32_t* sample_frame_size);
int rvc_stream_close(rvc_ctx* ctx);
int rvc_stream_set(rvc_ctx* ctx, int32_t pitch_shift, double rms_mix_rate);      /* settings that change without a rebuild */
int rvc_stream_info(rvc_ctx* ctx, char* out, size_t cap_bytes, size_t* out_bytes); /* JSON of the derived sizes */
int rvc_process_frame(rvc_ctx* ctx, const float* input, float* output, uint32_t* sola_offset);
/* One chunk through the rubato 0.15.0 FftFixedInOut equivalent (n_in must be a multiple of fs_in / gcd): out = n_in * fs_out /
 * fs_in samples; overlap_inout ([n_out]) carries the overlap-add state between calls (zeros at the start). */
int rvc_resample_chunk(rvc_ctx* ctx, uint32_t fs_in, uint32_t fs_out, const float* in, size_t n_in, float* overlap_inout,
                       float* out, size_t cap, size_t* n_out);

/* ---- streaming glue around the call ("next" row) ---------------------------------
 * rt_utils::envelop_mixing(input, output, sample_rate, mix_rate) - obs-rvc/src/rt_utils.rs:119-132.
 * `output` (n_out samples) is modified in place; `input` must hold at least n_out samples.
 * rms1/rms2 (optional, n_out each) receive the interpolated envelopes the reference's test checks. */
int rvc_envelop_mixing(rvc_ctx* ctx, const float* input, size_t n_in, float* output, size_t n_out, uint32_t sample_rate,
                       double mix_rate, float* rms1, float* rms2);
/* rt_utils::get_sola_offset(input_buffer, sola_buffer, buffer_frame_size, search_frame_size)
 * - obs-rvc/src/rt_utils.rs:60-90 (normalised cross-correlation, last maximum wins). */
int rvc_sola_offset(rvc_ctx* ctx, const float* input_buffer, size_t n, const float* sola_buffer, uint32_t buffer_frame_size,
                    uint32_t search_frame_size, uint32_t* offset);
/* SOLA tail of process_one_frame - obs-rvc/src/lib.rs:768-794: picks the offset, cross-fades
 * `infer_out` with `sola_buffer` (sin^2 windows, lib.rs:231-233), updates `sola_buffer` in place and
 * returns the `sample_frame_size` samples of this block. */
int rvc_sola_crossfade(rvc_ctx* ctx, const float* infer_out, size_t n, float* sola_buffer, uint32_t buffer_frame_size,
                       uint32_t search_frame_size, uint32_t sample_frame_size, float* block_out, uint32_t* offset);

/* Results of the last call kept on the device and copied on demand: "f0" (f32[T]),
 * "f0_argmax" (i32[T]), "salience" (f32[T*360]), "pitch" (i32[R]), "pitchf" (f32[R]),
 * "phone" (f32[R*C]), "knn_idx" (i32[Q*k]), "knn_d2" (f32[Q*k]), "mel" (f32[T*128], (T,128)). */
int rvc_get_last(rvc_ctx* ctx, const char* name, void* out, size_t cap_bytes, size_t* out_bytes);
/* Same for window `window` of the last batched call (rvc_infer_windows / batched rvc_infer_batch). */
int rvc_get_last_window(rvc_ctx* ctx, int32_t window, const char* name, void* out, size_t cap_bytes, size_t* out_bytes);
/* Any intermediate buffer by plan name (needs debug_keep=1; testing only). */
int rvc_debug_tensor(rvc_ctx* ctx, const char* name, float* out, size_t cap, size_t* out_len);
int rvc_debug_list(rvc_ctx* ctx, char* out, size_t cap_bytes, size_t* out_bytes);

/* state / plumbing */
int rvc_reset_state(rvc_ctx* ctx);                 /* zero the pitch cache + window counter  */
int rvc_sync(rvc_ctx* ctx);                        /* cudaStreamSynchronize(ctx stream)       */
void* rvc_cuda_stream(rvc_ctx* ctx);               /* cudaStream_t of the context             */
int rvc_kernel_launches(rvc_ctx* ctx, uint64_t* total); /* kernels launched by this context  */
int rvc_plan_info(rvc_ctx* ctx, char* out, size_t cap_bytes, size_t* out_bytes);
/* device-side timing on the context stream (CUDA events; 8 slots) */
int rvc_event_record(rvc_ctx* ctx, int slot);
int rvc_event_elapsed_ms(rvc_ctx* ctx, int slot_a, int slot_b, float* ms);
/* Replays every op of the last plan `iters` times back to back between CUDA events and returns a
 * JSON array [{"name","kind","us","flops","wbytes","iobytes","grid"}] (measurement aid: per-op device
 * time + algorithmic work; inputs are whatever the last run left in the work arena). */
int rvc_profile_ops(rvc_ctx* ctx, int iters, char* out, size_t cap_bytes, size_t* out_bytes);
/* End time (us since graph start) of every op of the last plan inside one CUDA-graph replay with all
 * lanes running concurrently: JSON [{"name","lane","end_us"}] (critical-path analysis aid). */
int rvc_profile_timeline(rvc_ctx* ctx, char* out, size_t cap_bytes, size_t* out_bytes);
/* Per-phase device time of every persistent chain kernel of the last plan (most recent run):
 * JSON [{"chain","lane","grid","phases":[{"ops","us"}]}]. */
int rvc_profile_chains(rvc_ctx* ctx, char* out, size_t cap_bytes, size_t* out_bytes);
const char* rvc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RVC_B200_H */
