// launch.h - device launchers for the op descriptors of ops.h (implemented in kernels_*.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "ops.h"

namespace rvc {

struct DeviceBases {
    uint8_t* b[SP_COUNT] = {nullptr};
    // byte distance from the fp32 weights of a space to their tf32 "hi" copy (and again to "lo"); 0 = none
    int64_t hilo_stride[SP_COUNT] = {0};
    // fp16 split planes of a weight space: byte offset of the hi plane from the fp32 arena, byte size of one plane
    // (element i of the arena sits at hi16_off + 2 i, its scaled residual one plane further); 0 = none
    int64_t hilo16_off[SP_COUNT] = {0}, hilo16_plane[SP_COUNT] = {0};
    // Batched plans (several windows per launch - offline conversion, or several live streams of one GPU): every
    // kernel processes `nb` windows; window w finds its activations / state `bstride[space]` bytes after window
    // w-1's (work arena and state block are replicated per window, weights and index are shared: stride 0).
    int32_t nb = 1;
    int64_t bstride[SP_COUNT] = {0};
    template <typename T> T* p(const Ref& r) const {
        return r.null() ? nullptr : reinterpret_cast<T*>(b[r.space] + r.off);
    }
    // window stride of a buffer in 4-byte elements (every plan buffer is f32 / i32)
    long long ws(const Ref& r) const { return r.null() ? 0 : (long long)(bstride[r.space] / 4); }
};

// Each launcher enqueues exactly the kernels of one op on `stream` and returns how many kernels
// it launched (0 for memset-only ops).  Errors are reported through cudaGetLastError by the caller.
int launch_gemm(const GemmOp& g, const DeviceBases& B, cudaStream_t stream);
// lane stamps (pdl.cuh lane_stamp): {STFT start, F0 decode start, pitch cache start, sine source start}, {retrieval gather start, conv_post end, RMVPE pool 0..4 start, GRU start}
void dsp_read_stamps(unsigned long long* out4);
void misc_read_stamps(unsigned long long* out8);
// fused RMVPE residual block (kernels_cbr.cu)
int launch_cbr(const CbrOp& o, const DeviceBases& B, cudaStream_t stream);
void init_cbr_attributes();
int launch_layernorm(const LayerNormOp& o, const DeviceBases& B, cudaStream_t stream);
int launch_attn(const AttnOp& o, const DeviceBases& B, cudaStream_t stream);
int launch_relattn(const RelAttnOp& o, const DeviceBases& B, cudaStream_t stream);
int launch_conv0_stats(const Conv0StatsOp& o, const DeviceBases& B, cudaStream_t stream);
int launch_conv0_apply(const Conv0ApplyOp& o, const DeviceBases& B, cudaStream_t stream);
int launch_stftmel(const StftMelOp& o, const DeviceBases& B, cudaStream_t stream);
int launch_avgpool(const AvgPoolOp& o, const DeviceBases& B, cudaStream_t stream);
int launch_gru(const GruOp& o, const DeviceBases& B, cudaStream_t stream);
int launch_f0decode(const F0DecodeOp& o, const DeviceBases& B, cudaStream_t stream);
int launch_f0post(const F0PostOp& o, const DeviceBases& B, cudaStream_t stream);
int launch_embed(const EmbedOp& o, const DeviceBases& B, cudaStream_t stream);
int launch_zp(const ZpOp& o, const DeviceBases& B, cudaStream_t stream);
int launch_sinegen(const SineGenOp& o, const DeviceBases& B, cudaStream_t stream);
int launch_avg3(const Avg3Op& o, const DeviceBases& B, cudaStream_t stream);
int launch_convpost(const ConvPostOp& o, const DeviceBases& B, cudaStream_t stream);
int launch_knn_scan(const KnnScanOp& o, const DeviceBases& B, cudaStream_t stream);
int launch_knn_select(const KnnSelectOp& o, const DeviceBases& B, cudaStream_t stream);
int launch_knn_blend(const KnnBlendOp& o, const DeviceBases& B, cudaStream_t stream);
// tensor-core candidate pass + exact re-rank (kernels_knn_umma.cu); scan returns -1 when the tensor maps cannot be built
int launch_knn_scan_umma(const KnnScanOp& o, const DeviceBases& B, cudaStream_t stream);
int launch_knn_rerank(const KnnSelectOp& o, const DeviceBases& B, cudaStream_t stream);
// builds [y_hi | y_lo' | |y|^2 | counters] at `planes` (knn_umma_planes_bytes) from N x C fp32 rows; returns max |y|^2 (synchronises)
void knn_umma_debug_read(long long* out);   // [8 events][32 tiles] clock64 stamps of CTA 0 (build with -DRVC_KU_STAMPS)
float launch_knn_build_planes(const float* index, int N, int C, uint8_t* planes, cudaStream_t stream);
int launch_gather_rows(const GatherRowsOp& o, const DeviceBases& B, cudaStream_t stream);

// streaming glue ("next" row #1): obs-rvc/src/rt_utils.rs + lib.rs:779-791 on the device
void launch_rms(const float* y, int n, int frame_length, int hop, float* out, int n_frames, cudaStream_t s);
void launch_envelop_mix(float* out, int n_out, const float* r1, const float* r2, int nr, float power, float* dbg1, float* dbg2,
                        cudaStream_t s);
void launch_sola(const float* x, const float* sola, int buf, int search, float* cor, int* offset, cudaStream_t s);
void launch_sola_crossfade(float* out, const int* offset, float* sola_buffer, int buf, int frame, float* block_out, cudaStream_t s);
// rubato FftFixedInOut in direct polyphase form (resample.h): one chunk x[nin] -> out[nout] (+ previous overlap), new overlap[nout]
void launch_resample(const float* x, const float* kappa, const float* ov_old, float* ov_new, float* out, int a, int b, int period, int nin,
                     int nout, cudaStream_t s);
void launch_shift_append(float* dst, const float* src, int len, int shift, const float* tail, int n_tail, cudaStream_t s);

// Function attributes (dynamic shared memory opt-in) belong to a device: `mask` remembers the devices a call site has
// already configured; returns true the first time it is reached with the calling thread's current device.
inline bool first_time_on_device(unsigned long long& mask) {
    int d = 0;
    cudaGetDevice(&d);
    if (d < 0 || d > 63) return true;
    if ((mask >> d) & 1ull) return false;
    mask |= 1ull << d;
    return true;
}

// per-device kernel attribute setup (dynamic shared memory opt-in)
void init_kernel_attributes();
void init_gemm_v2_attributes();
void init_umma_attributes();
void umma_debug_read(long long* out);
void umma_debug_read2(long long* out);  // [5 events][16 k-blocks] of the last tcgen05 GEMM CTA (0,0,0)
void v2_debug_read(long long* out);  // phase timestamps of the last tcgen05 GEMM CTA (0,0,0)
// tcgen05 path; returns 0 if the TMA descriptors cannot be built (caller falls back)
int launch_gemm_umma(const GemmOp& g, const DeviceBases& B, cudaStream_t stream);
// dst_hi[i] = src[i] with the 13 low mantissa bits cleared, dst_lo[i] = src[i] - dst_hi[i]
void launch_split_hilo(const float* src, float* dst_hi, float* dst_lo, size_t n, cudaStream_t stream);
// dst_hi[i] = half(src[i]), dst_lo[i] = half((src[i] - float(dst_hi[i])) * 2048)  (raw 16-bit storage)
void launch_split_hilo16(const float* src, unsigned short* dst_hi, unsigned short* dst_lo, size_t n, cudaStream_t stream);

}  // namespace rvc
