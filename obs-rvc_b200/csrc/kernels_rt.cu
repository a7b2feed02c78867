// kernels_rt.cu - streaming glue around the inference call ("next" row #1):
// obs-rvc/src/rt_utils.rs `rms` (93-102), `linear_interpolate_align_corners` (104-117),
// `envelop_mixing` (119-132) and `get_sola_offset` (60-90), plus the SOLA cross-fade of
// obs-rvc/src/lib.rs:779-791.  Moving these next to the engine removes the last host round trip of
// a window.  Parity is PINNED by the reference's own goldens (obs-rvc/src/tests/*.npy):
// SOLA offset == 321, envelope rms1/rms2/mixed within 1e-6.
#include <cfloat>

#include "launch.h"
#include "pdl.cuh"

namespace rvc {

namespace {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// rms(y, frame_length, hop_length): zero padding frame_length/2 both sides, one warp per frame
__global__ void rms_kernel(const float* __restrict__ y, int n, int frame_length, int hop, float* __restrict__ out, int n_frames) {
    pdl_enter();
    const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (f >= n_frames) return;
    const int start = f * hop - frame_length / 2;
    float s = 0.f;
    for (int j = lane; j < frame_length; j += 32) {
        const int p = start + j;
        const float v = (p >= 0 && p < n) ? y[p] : 0.f;
        s = fmaf(v, v, s);
    }
    s = warp_sum_f(s);
    if (lane == 0) out[f] = sqrtf(s / float(frame_length));
}

__device__ __forceinline__ float interp_ac(const float* __restrict__ x, int nx, float step, int i) {
    // rt_utils.rs:108-114 (align_corners): idx = i*step; floor/ceil clamped; linear blend
    const float idx = float(i) * step;
    int fl = int(floorf(idx)), ce = int(ceilf(idx));
    fl = min(max(fl, 0), nx - 1); ce = min(max(ce, 0), nx - 1);
    const float fr = idx - float(fl);
    return __fadd_rn(__fmul_rn(x[fl], 1.0f - fr), __fmul_rn(x[ce], fr));
}

// out[i] *= (rms1[i] / max(rms2[i], 1e-3)) ^ (1 - mix_rate), rms interpolated to n_out+1 points
__global__ void envelop_mix_kernel(float* __restrict__ out, int n_out, const float* __restrict__ r1, const float* __restrict__ r2,
                                   int nr, float power, float* __restrict__ dbg1, float* __restrict__ dbg2) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    const float step = float(nr - 1) / float(n_out);  // (len-1)/(size-1) with size = n_out+1
    const float a = interp_ac(r1, nr, step, i);
    const float b = fmaxf(interp_ac(r2, nr, step, i), 1e-3f);
    if (dbg1) { dbg1[i] = a; dbg2[i] = b; }
    out[i] = out[i] * powf(a / b, power);
}

// normalised cross-correlation for every candidate offset, one CTA per offset
__global__ void __launch_bounds__(256)
sola_corr_kernel(const float* __restrict__ x, const float* __restrict__ sola, int buf, float* __restrict__ cor) {
    pdl_enter();
    const int o = blockIdx.x;
    float nom = 0.f, den = 0.f;
    for (int j = threadIdx.x; j < buf; j += 256) {
        const float v = x[o + j];
        nom = fmaf(v, sola[j], nom);
        den = fmaf(v, v, den);
    }
    __shared__ float sn[8], sd[8];
    nom = warp_sum_f(nom); den = warp_sum_f(den);
    if ((threadIdx.x & 31) == 0) { sn[threadIdx.x >> 5] = nom; sd[threadIdx.x >> 5] = den; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < 8; ++w) { a += sn[w]; b += sd[w]; }
        cor[o] = a / sqrtf(b + 1e-8f);
    }
}

// argmax with LAST maximum winning ties (rt_utils.rs:82-88), then the sin^2 cross-fade and the
// sola_buffer update of lib.rs:779-791 on the shifted output
__global__ void __launch_bounds__(256)
sola_pick_kernel(const float* __restrict__ cor, int n_off, int* __restrict__ offset_out) {
    pdl_enter();
    __shared__ float sv[256]; __shared__ int si[256];
    float bv = -FLT_MAX; int bi = 0;
    for (int i = threadIdx.x; i < n_off; i += 256) { const float v = cor[i]; if (!(bv > v)) { bv = v; bi = i; } }
    sv[threadIdx.x] = bv; si[threadIdx.x] = bi;
    __syncthreads();
    if (threadIdx.x == 0) {
        float best = sv[0]; int idx = si[0];
        for (int t = 0; t < 256; ++t)
            if (sv[t] > best || (sv[t] == best && si[t] > idx)) { best = sv[t]; idx = si[t]; }
        *offset_out = idx;
    }
}

__global__ void sola_fade_kernel(float* __restrict__ out, const int* __restrict__ offset, float* __restrict__ sola_buffer,
                                 int buf, int frame, float* __restrict__ block_out) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float* o = out + *offset;
    if (i < buf) {
        // fade_in = sin(0.5*pi*x)^2 over linspace(0,1,buf), fade_out = 1 - fade_in (lib.rs:231-233)
        const float xx = buf > 1 ? float(i) / float(buf - 1) : 0.f;
        const float s = sinf(xx * 0.5f * 3.14159265358979323846f);
        const float fi = s * s;
        o[i] = o[i] * fi + sola_buffer[i] * (1.0f - fi);
    }
    __syncthreads();
    (void)frame; (void)block_out;
}

__global__ void sola_finish_kernel(const float* __restrict__ out, const int* __restrict__ offset, float* __restrict__ sola_buffer,
                                   int buf, int frame, float* __restrict__ block_out) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float* o = out + *offset;
    if (i < buf) sola_buffer[i] = o[frame + i];
    if (i < frame) block_out[i] = o[i];
}


// rubato FftFixedInOut as a direct-form polyphase filter (resample.h): output m = b i + c of a chunk is
//   y[m] = sum_n x[n] kappa_ph[r][(a i + qc - n) mod P],   r = (a c) mod b, qc = (a c) div b.
// A CTA takes 32 consecutive i of one residue c: their kappa windows overlap (shifted by a), so the window and the
// chunk are staged in shared memory once; a warp produces four outputs per pass (every x element feeds four sums),
// lanes stride over n; fixed summation order (lane partial sums, then a butterfly).  First half + previous overlap
// -> out, second half -> the new overlap (rubato's overlap-add, synchro.rs resample_unit).
constexpr int RS_OUT = 32;
__global__ void __launch_bounds__(256)
resample_kernel(const float* __restrict__ x, const float* __restrict__ kappa, const float* __restrict__ ov_old, float* __restrict__ ov_new,
                float* __restrict__ out, int a, int b, int P, int nin, int nout) {
    pdl_enter();
    extern __shared__ __align__(16) float rs_sm[];
    float* xs = rs_sm;                 // [nin]
    float* W = rs_sm + nin;            // [nin + (RS_OUT - 1) a + 1]
    const int c = blockIdx.y, i0 = blockIdx.x * RS_OUT;
    const int r = (a * c) % b, qc = (a * c) / b;
    const int wlen = nin + (RS_OUT - 1) * a + 1;
    const long long qbase = (long long)a * i0 + qc - (nin - 1);
    const float* kp = kappa + (long long)r * P;
    for (int t = threadIdx.x; t < nin; t += 256) xs[t] = x[t];
    for (int t = threadIdx.x; t < wlen; t += 256) {
        long long q = (qbase + t) % P;
        if (q < 0) q += P;
        W[t] = kp[q];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per_c = (2 * nout) / b;   // outputs of one residue
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int li0 = warp * 4;
    const float* w0 = W + (nin - 1) + a * li0;
    for (int n = lane; n < nin; n += 32) {
        const float xv = xs[n];
        acc[0] = fmaf(xv, w0[-n], acc[0]);
        acc[1] = fmaf(xv, w0[a - n], acc[1]);
        acc[2] = fmaf(xv, w0[2 * a - n], acc[2]);
        acc[3] = fmaf(xv, w0[3 * a - n], acc[3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float y = warp_sum_f(acc[j]);
        const int i = i0 + li0 + j;
        if (lane == 0 && i < per_c) {
            const int m = b * i + c;
            if (m < nout) out[m] = y + ov_old[m];
            else ov_new[m - nout] = y;
        }
    }
}

// dst = [src[shift:], tail]: the left-rotating ring buffers of the streaming loop (lib.rs:661-669) as a ping-pong copy
__global__ void shift_append_kernel(float* __restrict__ dst, const float* __restrict__ src, int len, int shift, const float* __restrict__ tail, int n_tail) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    if (i < len - shift) dst[i] = src[i + shift];
    else if (tail && i - (len - shift) < n_tail) dst[i] = tail[i - (len - shift)];
}

}  // namespace

void launch_resample(const float* x, const float* kappa, const float* ov_old, float* ov_new, float* out, int a, int b, int period, int nin,
                     int nout, cudaStream_t s) {
    const size_t smem = sizeof(float) * size_t(2 * nin + (RS_OUT - 1) * a + 1);
    static unsigned long long attr = 0;
    if (first_time_on_device(attr)) cudaFuncSetAttribute(resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    const int per_c = (2 * nout) / b;
    launch_k(resample_kernel, dim3((per_c + RS_OUT - 1) / RS_OUT, b), dim3(256), smem, s, x, kappa, ov_old, ov_new, out, a, b, period, nin, nout);
}
void launch_shift_append(float* dst, const float* src, int len, int shift, const float* tail, int n_tail, cudaStream_t s) {
    launch_k(shift_append_kernel, dim3((len + 255) / 256), dim3(256), size_t(0), s, dst, src, len, shift, tail, n_tail);
}

void launch_rms(const float* y, int n, int frame_length, int hop, float* out, int n_frames, cudaStream_t s) {
    launch_k(rms_kernel, dim3((n_frames * 32 + 255) / 256), dim3(256), size_t(0), s, y, n, frame_length, hop, out, n_frames);
}
void launch_envelop_mix(float* out, int n_out, const float* r1, const float* r2, int nr, float power, float* dbg1, float* dbg2,
                        cudaStream_t s) {
    launch_k(envelop_mix_kernel, dim3((n_out + 255) / 256), dim3(256), size_t(0), s, out, n_out, r1, r2, nr, power, dbg1, dbg2);
}
void launch_sola(const float* x, const float* sola, int buf, int search, float* cor, int* offset, cudaStream_t s) {
    launch_k(sola_corr_kernel, dim3(search + 1), dim3(256), size_t(0), s, x, sola, buf, cor);
    launch_k(sola_pick_kernel, dim3(1), dim3(256), size_t(0), s, (const float*)cor, search + 1, offset);
}
void launch_sola_crossfade(float* out, const int* offset, float* sola_buffer, int buf, int frame, float* block_out, cudaStream_t s) {
    launch_k(sola_fade_kernel, dim3((buf + 255) / 256), dim3(256), size_t(0), s, out, offset, sola_buffer, buf, frame, block_out);
    const int n = buf > frame ? buf : frame;
    launch_k(sola_finish_kernel, dim3((n + 255) / 256), dim3(256), size_t(0), s, (const float*)out, offset, sola_buffer, buf, frame, block_out);
}

}  // namespace rvc
