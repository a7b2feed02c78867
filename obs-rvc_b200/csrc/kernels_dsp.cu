// kernels_dsp.cu - DSP front/back-end ops: fused STFT->mel->log, F0 decode, pitch cache / coarse
// quantisation, NSF sine source.  Reference: rvc/src/f0/rmvpe.rs, rvc/src/f0/mod.rs,
// rvc/src/rvc.rs:111-181 (reproduced literally, quirks included - SURVEY.md Appendix B).
#include <cfloat>

#include "launch.h"
#include "pdl.cuh"
#include "noise.h"

namespace rvc {

namespace {

// ------------------------------------------------------------------------------------------
// stft_mel_log: one CTA per frame.  reflect-pad gather + periodic Hann + 1024-point radix-2 FFT
// in shared memory + magnitude + sparse mel filterbank + ln(max(., clamp))
// (rmvpe.rs:47-68 pad_reflect, 80-116 stft, 159-205 mel_extract).  HBM traffic per frame is the
// 1024 input samples and 128 outputs; the filterbank (1010 non-zeros) stays in L2/L1.
// ------------------------------------------------------------------------------------------
__device__ unsigned long long g_dsp_stamps[4];   // [0] STFT start, [1] F0 decode start, [2] pitch cache start, [3] sine source start

__global__ void __launch_bounds__(256)
stft_mel_log_kernel(const float* __restrict__ pcm, int L, const float* __restrict__ window,
                    const int* __restrict__ band_start, const int* __restrict__ band_count,
                    const int* __restrict__ band_off, const float* __restrict__ band_w, float* __restrict__ mel,
                    float* __restrict__ out2, long long out2_pitch, float scale, float shift, float clamp,
                    long long wPcm, long long wMel, long long wOut2) {
    pdl_enter();
    lane_stamp(&g_dsp_stamps[0]);
    pcm += blockIdx.z * wPcm; mel += blockIdx.z * wMel;
    if (out2) out2 += blockIdx.z * wOut2;
    __shared__ float2 buf[1024];
    __shared__ float2 tw[512];
    __shared__ float mag[513];
    const int t = blockIdx.x, tid = threadIdx.x;
    for (int k = tid; k < 512; k += 256) {
        float s, c;
        sincospif(-float(k) / 512.0f, &s, &c);  // exp(-2 pi i k / 1024)
        tw[k] = make_float2(c, s);
    }
    for (int j = tid; j < 1024; j += 256) {
        int p = t * 160 + j - 512;
        if (p < 0) p = -p;
        if (p >= L) p = 2 * (L - 1) - p;
        float v = pcm[p] * window[j];
        buf[__brev((unsigned)j) >> 22] = make_float2(v, 0.f);
    }
    __syncthreads();
#pragma unroll 1
    for (int s = 0; s < 10; ++s) {
        const int half = 1 << s;
        for (int b = tid; b < 512; b += 256) {
            const int pos = b & (half - 1), i0 = ((b >> s) << (s + 1)) + pos, i1 = i0 + half;
            const float2 w = tw[pos << (9 - s)];
            const float2 u = buf[i0], x = buf[i1];
            const float2 v = make_float2(x.x * w.x - x.y * w.y, x.x * w.y + x.y * w.x);
            buf[i0] = make_float2(u.x + v.x, u.y + v.y);
            buf[i1] = make_float2(u.x - v.x, u.y - v.y);
        }
        __syncthreads();
    }
    for (int k = tid; k < 513; k += 256) mag[k] = sqrtf(buf[k].x * buf[k].x + buf[k].y * buf[k].y);
    __syncthreads();
    if (tid < 128) {
        const int s0 = band_start[tid], n = band_count[tid];
        const float* w = band_w + band_off[tid];
        float a = 0.f;
        for (int j = 0; j < n; ++j) a = fmaf(w[j], mag[s0 + j], a);
        const float v = logf(fmaxf(a, clamp));
        mel[t * 128 + tid] = v;
        if (out2) out2[(long long)t * out2_pitch + tid] = v * scale + shift;
    }
}

// ------------------------------------------------------------------------------------------
// F0 decode (rmvpe.rs:118-133 to_local_average_cents, 243-248 decode; rvc.rs:121 uppower)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
f0_decode_kernel(const float* __restrict__ sal, float* __restrict__ f0, int* __restrict__ argmax,
                 const RunParams* __restrict__ rp, int bins, float threshold, int upstream_window,
                 long long wSal, long long wF0, long long wArg, long long wRp) {
    pdl_enter();
    lane_stamp(&g_dsp_stamps[1]);
    sal += blockIdx.z * wSal; f0 += blockIdx.z * wF0; argmax += blockIdx.z * wArg;
    rp = reinterpret_cast<const RunParams*>(reinterpret_cast<const float*>(rp) + blockIdx.z * wRp);
    const int t = blockIdx.x, tid = threadIdx.x;
    const float* s = sal + (long long)t * bins;
    float bv = -FLT_MAX; int bi = 0x7fffffff;
    for (int i = tid; i < bins; i += 128) { float v = s[i]; if (v > bv) { bv = v; bi = i; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    __shared__ float sv[4]; __shared__ int si[4];
    if ((tid & 31) == 0) { sv[tid >> 5] = bv; si[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 4; ++w) if (sv[w] > bv || (sv[w] == bv && si[w] < bi)) { bv = sv[w]; bi = si[w]; }
        int c = bi;
        if (!(bv > 0.0f)) c = -4;  // argmax of the zero-padded row falls on the padding
        argmax[t] = c;
        float ps = 0.f, ws = 0.f;
        for (int y = 0; y < 9; ++y) {
            int b, ci;
            if (upstream_window) { b = c - 4 + y; ci = c + y; } else { b = c + 4 + y; ci = c + 4 + y; }
            if (b < 0 || b >= bins) continue;
            const float cents = (float(ci) - 4.0f) * 20.0f + 1997.3794084376191f;
            ps = __fadd_rn(ps, __fmul_rn(s[b], cents));
            ws = __fadd_rn(ws, s[b]);
        }
        float cents = ws != 0.f ? __fdiv_rn(ps, ws) : 0.f;
        if (!(bv > threshold)) cents = 0.f;
        float f = 10.0f * exp2f(cents / 1200.0f);
        if (f == 10.0f) f = 0.f;
        f0[t] = f * rp->uppower;
    }
}

// ------------------------------------------------------------------------------------------
// pitch cache roll / write / slice + coarse quantisation (rvc.rs:167-181, f0/mod.rs:7-12)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
f0_post_kernel(const float* __restrict__ f0, float* __restrict__ cache, int* __restrict__ pitch,
               float* __restrict__ pitchf, int pitch_len, int shift, int hubert_length, int skip_head,
               int return_length, int n, float mel_min, float mel_max,
               int seq_windows, long long wF0, long long wCache, long long wPitch, long long wPitchf) {
    pdl_enter();
    lane_stamp(&g_dsp_stamps[2]);
    // Batched plans: independent streams (grid.z = windows, one cache each) or `seq_windows` consecutive windows of
    // ONE stream (grid.z = 1): they update the single cache of window 0 one after the other, as the reference would.
    cache += blockIdx.z * wCache;
    const int i = threadIdx.x;
    const int w0 = seq_windows > 0 ? 0 : blockIdx.z, w1 = seq_windows > 0 ? seq_windows : blockIdx.z + 1;
    for (int w = w0; w < w1; ++w) {
        const float* f0w = f0 + w * wF0;
        int* pitchw = pitch + w * wPitch;
        float* pitchfw = pitchf + w * wPitchf;
        float keep = 0.f;
        if (i + shift < n) keep = cache[i + shift];
        __syncthreads();
        if (i + shift < n) cache[i] = keep;  // copy_within(shift.., 0); the tail keeps its old values
        __syncthreads();
        const int start = n + 4 - pitch_len;
        for (int j = i; j < pitch_len - 1; j += blockDim.x)
            if (j >= 3) cache[start + j - 3] = f0w[j];
        __syncthreads();
        const int a = n - hubert_length + skip_head;
        for (int r = i; r < return_length; r += blockDim.x) {
            const float f = cache[a + r];
            float m = logf(f / 700.0f + 1.0f) * 1127.0f;
            if (!(m <= 0.f)) m = __fadd_rn(__fdiv_rn(__fmul_rn(m - mel_min, 254.0f), mel_max - mel_min), 1.0f);
            m = fminf(fmaxf(m, 1.0f), 255.0f);
            pitchw[r] = int(floor(double(m) + 0.5));  // Rust round(): half away from zero
            pitchfw[r] = f;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// NSF sine source: SineGen(harmonic_num=0) + tanh(linear).  The phase
// accumulation over R*upp samples is a double-precision block scan (torch's CPU cumsum also
// accumulates in double), so parallel order does not change the rounded phase.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
sinegen_kernel(const float* __restrict__ f0, float* __restrict__ out, float* __restrict__ dbg,
               const RunParams* __restrict__ rp, int T, int upp, float sr, float lin_w, float lin_b,
               long long wF0, long long wOut, long long wDbg, long long wRp) {
    pdl_enter();
    lane_stamp(&g_dsp_stamps[3]);
    f0 += blockIdx.z * wF0; out += blockIdx.z * wOut;
    if (dbg) dbg += blockIdx.z * wDbg;
    rp = reinterpret_cast<const RunParams*>(reinterpret_cast<const float*>(rp) + blockIdx.z * wRp);
    extern __shared__ __align__(16) unsigned char smraw[];
    double* part = reinterpret_cast<double*>(smraw);        // [1024] chunk sums -> exclusive offsets
    float* rad = reinterpret_cast<float*>(part + 1024);     // [T]
    float* cum = rad + T;                                   // [T]
    const int tid = threadIdx.x, L = T * upp;
    if (tid == 0) {
        double acc = 0.0;
        for (int t = 0; t < T; ++t) {
            float r = fmodf(f0[t] / sr, 1.0f);
            rad[t] = r; acc += double(r); cum[t] = float(acc) * float(upp);
        }
    }
    __syncthreads();
    const float scale = L > 1 ? float(T - 1) / float(L - 1) : 0.f;
    auto tmp_at = [&](int i) {
        float src = scale * float(i);
        int i0 = min(int(src), T - 1), i1 = i0 + (i0 < T - 1 ? 1 : 0);
        float l1 = fminf(fmaxf(src - float(i0), 0.f), 1.f), l0 = 1.f - l1;
        return fmodf(__fadd_rn(__fmul_rn(l0, cum[i0]), __fmul_rn(l1, cum[i1])), 1.0f);
    };
    const int per = (L + 1023) / 1024, start = tid * per, end = min(L, start + per);
    double sum = 0.0;
    {
        float prev = start > 0 && start < L ? tmp_at(start - 1) : 0.f;
        for (int i = start; i < end; ++i) {
            float cur = tmp_at(i);
            float sh = (i > 0 && (cur - prev) < 0.f) ? -1.f : 0.f;
            sum += double(rad[i / upp] + sh);
            prev = cur;
        }
    }
    // block exclusive scan of the 1024 chunk sums
    const int lane = tid & 31, warp = tid >> 5;
    double incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { double v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    __shared__ double wtot[32];
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        double w = wtot[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { double v = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += v; }
        wtot[lane] = wi - w;  // exclusive
    }
    __syncthreads();
    double ph = wtot[warp] + (incl - sum);
    const uint64_t key = noise_key(rp->noise_seed, rp->window, NOISE_KIND_SINE);
    const int noise_mode = rp->noise_mode;
    float prev = start > 0 && start < L ? tmp_at(start - 1) : 0.f;
    for (int i = start; i < end; ++i) {
        float cur = tmp_at(i);
        float sh = (i > 0 && (cur - prev) < 0.f) ? -1.f : 0.f;
        prev = cur;
        ph += double(rad[i / upp] + sh);
        float s = sinf(float(ph) * 2.0f * 3.14159265358979323846f) * 0.1f;
        float uv = f0[i / upp] > 0.f ? 1.f : 0.f;
        float amp = uv * 0.003f + (1.f - uv) * 0.1f / 3.f;
        float nz = noise_mode ? noise_gauss(key, (uint64_t)i) : 0.f;
        float v = s * uv + amp * nz;
        if (dbg) dbg[i] = v;
        out[i] = tanhf(v * lin_w + lin_b);
    }
}

}  // namespace

int launch_stftmel(const StftMelOp& o, const DeviceBases& B, cudaStream_t s) {
    launch_k(stft_mel_log_kernel, dim3(o.T, 1, B.nb), dim3(256), size_t(0), s, B.p<float>(o.pcm), o.L, B.p<float>(o.window), B.p<int>(o.band_start),
                                            B.p<int>(o.band_count), B.p<int>(o.band_off), B.p<float>(o.band_w), B.p<float>(o.mel),
                                            B.p<float>(o.out2), o.out2_pitch, o.scale, o.shift, o.clamp, B.ws(o.pcm), B.ws(o.mel), B.ws(o.out2));
    return 1;
}

int launch_f0decode(const F0DecodeOp& o, const DeviceBases& B, cudaStream_t s) {
    launch_k(f0_decode_kernel, dim3(o.T, 1, B.nb), dim3(128), size_t(0), s, B.p<float>(o.salience), B.p<float>(o.f0), B.p<int>(o.argmax), B.p<RunParams>(o.params),
                                         o.bins, o.threshold, o.upstream_window, B.ws(o.salience), B.ws(o.f0), B.ws(o.argmax), B.ws(o.params));
    return 1;
}

int launch_f0post(const F0PostOp& o, const DeviceBases& B, cudaStream_t s) {
    const bool seq = o.sequential != 0 && B.nb > 1;
    launch_k(f0_post_kernel, dim3(1, 1, seq ? 1 : B.nb), dim3(1024), size_t(0), s, B.p<float>(o.f0), B.p<float>(o.cache), B.p<int>(o.pitch), B.p<float>(o.pitchf), o.pitch_len,
                                      o.shift, o.hubert_length, o.skip_head, o.return_length, o.cache_len, o.mel_min, o.mel_max,
                                      seq ? B.nb : 0, B.ws(o.f0), B.ws(o.cache), B.ws(o.pitch), B.ws(o.pitchf));
    return 1;
}

int launch_sinegen(const SineGenOp& o, const DeviceBases& B, cudaStream_t s) {
    size_t smem = sizeof(double) * 1024 + sizeof(float) * 2 * o.R;
    launch_k(sinegen_kernel, dim3(1, 1, B.nb), dim3(1024), size_t(smem), s, B.p<float>(o.pitchf), B.p<float>(o.out), B.p<float>(o.sine_dbg), B.p<RunParams>(o.params),
                                         o.R, o.upp, o.sr, o.lin_w, o.lin_b, B.ws(o.pitchf), B.ws(o.out), B.ws(o.sine_dbg), B.ws(o.params));
    return 1;
}

void dsp_read_stamps(unsigned long long* out4) { cudaMemcpyFromSymbol(out4, g_dsp_stamps, 4 * sizeof(unsigned long long)); }

}  // namespace rvc
