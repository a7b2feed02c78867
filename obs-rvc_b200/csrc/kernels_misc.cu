// kernels_misc.cu - the non-GEMM device ops of the plan (normalisation, attention, GRU, glue).
// All arithmetic is fp32 with the same formulas as the reference path's oracle; reductions use
// warp shuffles; every kernel is latency-sized for one audio window.
#include <cooperative_groups.h>

#include <cfloat>
#include <cuda_fp16.h>

#include "chain.h"
#include "cvstack.h"
#include "launch.h"
#include "pdl.cuh"
#include "noise.h"

namespace cg = cooperative_groups;

namespace rvc {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float gelu_f(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }
__device__ __forceinline__ float sigmoid_f(float v) { return 1.0f / (1.0f + expf(-v)); }

// ------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row held in registers (cols <= 32*MAXV)
// ------------------------------------------------------------------------------------------
constexpr int LN_MAXV = 32;  // up to 1024 columns

__global__ void layernorm_kernel(const float* __restrict__ X, long long ldx, float* __restrict__ Y, long long ldy,
                                 const float* __restrict__ gamma, const float* __restrict__ beta, int rows, int cols,
                                 float eps, long long wX, long long wY) {
    pdl_enter();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows) return;
    X += blockIdx.z * wX; Y += blockIdx.z * wY;   // window of a batched plan
    const float* x = X + (long long)warp * ldx;
    float v[LN_MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        int c = lane + i * 32;
        v[i] = c < cols ? x[c] : 0.f;
        s += v[i];
    }
    const float mean = warp_sum(s) / float(cols);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        int c = lane + i * 32;
        float d = c < cols ? v[i] - mean : 0.f;
        q += d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) / float(cols) + eps);
    float* y = Y + (long long)warp * ldy;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        int c = lane + i * 32;
        if (c < cols) y[c] = (v[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
    }
}

// ------------------------------------------------------------------------------------------
// Multi-head attention (ContentVec, head dim 64): CTA = (head, 8 query rows), one warp per row.
// K is staged transposed ([d][T], odd pitch) so the score pass reads consecutive keys per lane,
// V row-major for the P.V pass, q lives in registers.
// ------------------------------------------------------------------------------------------
constexpr int ATT_WARPS = 8, ATT_D = 64;

__global__ void __launch_bounds__(ATT_WARPS * 32)
attn_kernel(const float* __restrict__ qkv, long long ld, float* __restrict__ out, long long ldo, int T, int heads,
            long long wQkv, long long wOut, int rows_per_cta) {
    pdl_enter();
    qkv += blockIdx.z * wQkv; out += blockIdx.z * wOut;
    extern __shared__ __align__(16) float sm[];
    constexpr int D = ATT_D;
    const int h = blockIdx.x, qbase = blockIdx.y * rows_per_cta;
    const int HD = heads * D, Tp = (T | 1) + 2;          // odd pitch
    float* Vs = sm;                                       // [T][D]
    float* Kt = Vs + (size_t)T * D;                       // [D][Tp]
    float* Ps = Kt + (size_t)D * Tp;                      // [ATT_WARPS][Tp]
    for (int i = threadIdx.x; i < T * (D / 4); i += blockDim.x) {
        const int t = i / (D / 4), d4 = i - t * (D / 4);
        const float4 k = __ldg(reinterpret_cast<const float4*>(qkv + (long long)t * ld + HD + h * D) + d4);
        const float4 v = __ldg(reinterpret_cast<const float4*>(qkv + (long long)t * ld + 2 * HD + h * D) + d4);
        Kt[(d4 * 4 + 0) * Tp + t] = k.x; Kt[(d4 * 4 + 1) * Tp + t] = k.y;
        Kt[(d4 * 4 + 2) * Tp + t] = k.z; Kt[(d4 * 4 + 3) * Tp + t] = k.w;
        reinterpret_cast<float4*>(Vs + (size_t)t * D)[d4] = v;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    // K / V of the head are staged once per CTA; a warp takes query rows qbase + warp, + 8, ... (batched plans give a
    // CTA many rows so the staging is amortised; a single window keeps 8 rows per CTA for the widest grid)
    const int qend = min(T, qbase + rows_per_cta);
    if (rows_per_cta > ATT_WARPS) {
        // batched plans: a warp takes TWO query rows per pass (qi, qi + 8): every K / V element read from shared memory
        // feeds both rows (the single-row loop is bound by one LDS per FMA).  Same per-row arithmetic and summation order.
        float* Ps2 = Ps + (size_t)ATT_WARPS * Tp;
        for (int qi = qbase + warp; qi < qend; qi += 2 * ATT_WARPS) {
            const int qj = qi + ATT_WARPS;
            const bool two = qj < qend;
            float qa[D], qb[D];
#pragma unroll
            for (int d4 = 0; d4 < D / 4; ++d4) {
                const float4 t4 = __ldg(reinterpret_cast<const float4*>(qkv + (long long)qi * ld + h * D) + d4);
                qa[d4 * 4] = t4.x; qa[d4 * 4 + 1] = t4.y; qa[d4 * 4 + 2] = t4.z; qa[d4 * 4 + 3] = t4.w;
                const float4 u4 = two ? __ldg(reinterpret_cast<const float4*>(qkv + (long long)qj * ld + h * D) + d4) : make_float4(0.f, 0.f, 0.f, 0.f);
                qb[d4 * 4] = u4.x; qb[d4 * 4 + 1] = u4.y; qb[d4 * 4 + 2] = u4.z; qb[d4 * 4 + 3] = u4.w;
            }
            float* pa = Ps + warp * Tp;
            float* pb = Ps2 + warp * Tp;
            float mxa = -FLT_MAX, mxb = -FLT_MAX;
            for (int j = lane; j < T; j += 32) {
                float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
                for (int d = 0; d < D; d += 2) {
                    const float k0 = Kt[d * Tp + j], k1 = Kt[(d + 1) * Tp + j];
                    a0 = fmaf(qa[d], k0, a0); a1 = fmaf(qa[d + 1], k1, a1);
                    b0 = fmaf(qb[d], k0, b0); b1 = fmaf(qb[d + 1], k1, b1);
                }
                const float sa = a0 + a1, sb = b0 + b1;
                pa[j] = sa; pb[j] = sb;
                mxa = fmaxf(mxa, sa); mxb = fmaxf(mxb, sb);
            }
            mxa = warp_max(mxa); mxb = warp_max(mxb);
            float suma = 0.f, sumb = 0.f;
            for (int j = lane; j < T; j += 32) {
                const float ea = expf(pa[j] - mxa), eb = expf(pb[j] - mxb);
                pa[j] = ea; pb[j] = eb; suma += ea; sumb += eb;
            }
            suma = warp_sum(suma); sumb = warp_sum(sumb);
            __syncwarp();
            float oa0 = 0.f, oa1 = 0.f, ob0 = 0.f, ob1 = 0.f;
            for (int j = 0; j < T; ++j) {
                const float v0 = Vs[j * D + lane], v1 = Vs[j * D + lane + 32];
                const float wa = pa[j], wb = pb[j];
                oa0 = fmaf(wa, v0, oa0); oa1 = fmaf(wa, v1, oa1);
                ob0 = fmaf(wb, v0, ob0); ob1 = fmaf(wb, v1, ob1);
            }
            const float inva = 1.0f / suma, invb = 1.0f / sumb;
            out[(long long)qi * ldo + h * D + lane] = oa0 * inva;
            out[(long long)qi * ldo + h * D + lane + 32] = oa1 * inva;
            if (two) {
                out[(long long)qj * ldo + h * D + lane] = ob0 * invb;
                out[(long long)qj * ldo + h * D + lane + 32] = ob1 * invb;
            }
            __syncwarp();
        }
        return;
    }
    for (int qi = qbase + warp; qi < qend; qi += ATT_WARPS) {
    float q[D];
#pragma unroll
    for (int d4 = 0; d4 < D / 4; ++d4) {
        const float4 t4 = __ldg(reinterpret_cast<const float4*>(qkv + (long long)qi * ld + h * D) + d4);
        q[d4 * 4] = t4.x; q[d4 * 4 + 1] = t4.y; q[d4 * 4 + 2] = t4.z; q[d4 * 4 + 3] = t4.w;
    }
    float* ps = Ps + warp * Tp;
    float mx = -FLT_MAX;
    for (int j = lane; j < T; j += 32) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int d = 0; d < D; d += 2) {
            a0 = fmaf(q[d], Kt[d * Tp + j], a0);
            a1 = fmaf(q[d + 1], Kt[(d + 1) * Tp + j], a1);
        }
        const float a = a0 + a1;
        ps[j] = a;
        mx = fmaxf(mx, a);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < T; j += 32) { const float e = expf(ps[j] - mx); ps[j] = e; sum += e; }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.0f / sum;
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < T; ++j) {
        const float pj = ps[j];
        o0 = fmaf(pj, Vs[j * D + lane], o0);
        o1 = fmaf(pj, Vs[j * D + lane + 32], o1);
    }
    out[(long long)qi * ldo + h * D + lane] = o0 * inv;
    out[(long long)qi * ldo + h * D + lane + 32] = o1 * inv;
    __syncwarp();
    }
}

// VITS windowed relative-position attention (enc_p): one CTA per (head, query row), 128 threads.
__global__ void __launch_bounds__(128)
relattn_kernel(const float* __restrict__ qkv, long long ld, float* __restrict__ out, long long ldo,
               const float* __restrict__ rel_k, const float* __restrict__ rel_v, int T, int heads, int dim, int window,
               long long wQkv, long long wOut) {
    pdl_enter();
    qkv += blockIdx.z * wQkv; out += blockIdx.z * wOut;
    extern __shared__ float sm[];
    const int h = blockIdx.x / T, i = blockIdx.x - h * T;
    const int HD = heads * dim, nrel = 2 * window + 1;
    float* qs = sm;            // [dim]
    float* ps = qs + dim;      // [T]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    for (int d = tid; d < dim; d += blockDim.x) qs[d] = qkv[(long long)i * ld + h * dim + d];
    __syncthreads();
    for (int j = warp; j < T; j += nw) {   // one key per warp pass, lanes over the head dimension
        const float* k = qkv + (long long)j * ld + HD + h * dim;
        const int rel = j - i + window;
        const float* rk = (rel >= 0 && rel < nrel) ? rel_k + rel * dim : nullptr;
        float a = 0.f;
        for (int d = lane; d < dim; d += 32) a = fmaf(qs[d], k[d] + (rk ? rk[d] : 0.f), a);
        a = warp_sum(a);
        if (lane == 0) ps[j] = a;
    }
    __syncthreads();
    if (warp == 0) {
        float mx = -FLT_MAX;
        for (int j = lane; j < T; j += 32) mx = fmaxf(mx, ps[j]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = lane; j < T; j += 32) { const float e = expf(ps[j] - mx); ps[j] = e; sum += e; }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        for (int j = lane; j < T; j += 32) ps[j] *= inv;
    }
    __syncthreads();
    for (int d = tid; d < dim; d += blockDim.x) {
        float a = 0.f;
        for (int j = 0; j < T; ++j) {
            const int rel = j - i + window;
            float v = qkv[(long long)j * ld + 2 * HD + h * dim + d];
            if (rel >= 0 && rel < nrel) v += rel_v[rel * dim + d];
            a = fmaf(ps[j], v, a);
        }
        out[(long long)i * ldo + h * dim + d] = a;
    }
}

// ------------------------------------------------------------------------------------------
// ContentVec conv0 (1 -> 512, k=10, stride 5) + GroupNorm(512 groups) + GELU
// ------------------------------------------------------------------------------------------
constexpr int C0_CH = 4;  // channels per CTA in the stats pass

__global__ void __launch_bounds__(256)
conv0_stats_kernel(const float* __restrict__ pcm, const float* __restrict__ w, float* __restrict__ stats, int T, int C,
                   int k, int stride, float eps, long long wPcm, long long wStats) {
    pdl_enter();
    pcm += blockIdx.z * wPcm; stats += blockIdx.z * wStats;
    const int c0 = blockIdx.x * C0_CH;
    float wr[C0_CH][10];
#pragma unroll
    for (int c = 0; c < C0_CH; ++c)
#pragma unroll
        for (int j = 0; j < 10; ++j) wr[c][j] = (j < k && c0 + c < C) ? w[(c0 + c) * k + j] : 0.f;
    double s[C0_CH], ss[C0_CH];
#pragma unroll
    for (int c = 0; c < C0_CH; ++c) { s[c] = 0.0; ss[c] = 0.0; }
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        float x[10];
#pragma unroll
        for (int j = 0; j < 10; ++j) x[j] = j < k ? __ldg(pcm + t * stride + j) : 0.f;
#pragma unroll
        for (int c = 0; c < C0_CH; ++c) {
            float a = 0.f;
#pragma unroll
            for (int j = 0; j < 10; ++j) a = fmaf(x[j], wr[c][j], a);
            s[c] += double(a);
            ss[c] += double(a) * double(a);
        }
    }
    __shared__ double red[8][C0_CH][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < C0_CH; ++c) {
        double a = warp_sum_d(s[c]), b = warp_sum_d(ss[c]);
        if (lane == 0) { red[warp][c][0] = a; red[warp][c][1] = b; }
    }
    __syncthreads();
    if (threadIdx.x < C0_CH && c0 + threadIdx.x < C) {
        double a = 0, b = 0;
        for (int i = 0; i < 8; ++i) { a += red[i][threadIdx.x][0]; b += red[i][threadIdx.x][1]; }
        double mean = a / T, var = b / T - mean * mean;
        stats[2 * (c0 + threadIdx.x)] = float(mean);
        stats[2 * (c0 + threadIdx.x) + 1] = float(1.0 / sqrt(var + double(eps)));
    }
}

constexpr int C0_TB = 16;  // time steps per CTA in the apply pass

__global__ void __launch_bounds__(256)
conv0_apply_kernel(const float* __restrict__ pcm, const float* __restrict__ w, const float* __restrict__ stats,
                   const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ Y, int T, int C,
                   int k, int stride, long long wPcm, long long wStats, long long wY) {
    pdl_enter();
    pcm += blockIdx.z * wPcm; stats += blockIdx.z * wStats; Y += blockIdx.z * wY;
    __shared__ float xs[C0_TB * 5 + 16];
    const int t0 = blockIdx.x * C0_TB;
    const int nx = (min(C0_TB, T - t0) - 1) * stride + k;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) xs[i] = pcm[t0 * stride + i];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float wr[10];
#pragma unroll
        for (int j = 0; j < 10; ++j) wr[j] = j < k ? w[c * k + j] : 0.f;
        const float mean = stats[2 * c], rstd = stats[2 * c + 1], g = gamma[c], b = beta[c];
        for (int tt = 0; tt < C0_TB && t0 + tt < T; ++tt) {
            float a = 0.f;
#pragma unroll
            for (int j = 0; j < 10; ++j) a = fmaf(xs[tt * stride + j], wr[j], a);
            Y[(long long)(t0 + tt) * C + c] = gelu_f((a - mean) * rstd * g + b);
        }
    }
}

// ------------------------------------------------------------------------------------------
// 2x2 average pool on halo-padded NHWC
// ------------------------------------------------------------------------------------------
__device__ unsigned long long g_misc_stamps[8];   // [0] retrieval gather ("phone") start, [1] conv_post end (CTA 0), [2..6] RMVPE pool 0..4 start, [7] GRU start
__global__ void avgpool_kernel(const float* __restrict__ in, long long ldin, float* __restrict__ out, int T, int F, int C,
                               long long wIn, long long wOut) {
    pdl_enter();
    lane_stamp(&g_misc_stamps[2 + min(4, max(0, 31 - __clz(C >> 4)))]);
    in += blockIdx.z * wIn; out += blockIdx.z * wOut;
    const int To = T / 2, Fo = F / 2;
    const long long n = (long long)To * Fo * C;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        int c = int(e % C);
        long long r = e / C;
        int f = int(r % Fo), t = int(r / Fo);
        const float* p = in + ((long long)(2 * t + 1) * (F + 2) + 2 * f + 1) * ldin + c;
        float v = (p[0] + p[ldin] + p[(long long)(F + 2) * ldin] + p[(long long)(F + 3) * ldin]) * 0.25f;
        out[((long long)(t + 1) * (Fo + 2) + f + 1) * C + c] = v;
    }
}

// ------------------------------------------------------------------------------------------
// Bidirectional GRU (H=256).  The recurrence is latency-bound (T sequential steps), so W_hh is
// made register-resident: one thread-block cluster of 8 CTAs per direction, CTA r owns hidden
// units [32r, 32r+32) = 96 gate rows x 256, 32 weights per thread (768 threads).  Every step each
// CTA computes its 32 new h values and pushes them into all 8 peers' shared memory over DSMEM;
// one cluster barrier per step replaces a trip through L2.
// ------------------------------------------------------------------------------------------
constexpr int GRU_CL = 8, GRU_H = 256, GRU_UNITS = GRU_H / GRU_CL;  // 32 hidden units per CTA

__global__ void __cluster_dims__(GRU_CL, 1, 1) __launch_bounds__(768)
gru_cluster_kernel(const float* __restrict__ gi, const float* __restrict__ whh_t, const float* __restrict__ bhh,
                   float* __restrict__ out, int T, long long wGi, long long wOut) {
    pdl_enter();
    lane_stamp(&g_misc_stamps[7]);
    gi += blockIdx.z * wGi; out += blockIdx.z * wOut;
    cg::cluster_group cluster = cg::this_cluster();
    constexpr int H = GRU_H, G = 3 * GRU_H;
    __shared__ __align__(16) float hbuf[2][H];
    __shared__ float gh[3 * GRU_UNITS];
    const int rank = int(cluster.block_rank()), d = blockIdx.x / GRU_CL, tid = threadIdx.x;
    const int row_local = tid >> 3, part = tid & 7;            // 96 rows x 8 k-slices
    const int gate = row_local / GRU_UNITS, ju = row_local % GRU_UNITS;
    const int g = gate * H + rank * GRU_UNITS + ju;            // global gate row
    float w[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) w[i] = whh_t[((long long)d * H + (i * 8 + part)) * G + g];
    const float bg = bhh[d * G + g];
    if (tid < H) { hbuf[0][tid] = 0.f; hbuf[1][tid] = 0.f; }
    cluster.sync();
    for (int s = 0; s < T; ++s) {
        const int t = d == 0 ? s : T - 1 - s, cur = s & 1;
        float xr = 0.f, xz = 0.f, xn = 0.f;
        if (tid < GRU_UNITS) {
            const float* x = gi + (long long)t * 2 * G + d * G + rank * GRU_UNITS + tid;
            xr = x[0]; xz = x[H]; xn = x[2 * H];
        }
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            a0 = fmaf(w[i], hbuf[cur][i * 8 + part], a0);
            a1 = fmaf(w[i + 1], hbuf[cur][(i + 1) * 8 + part], a1);
        }
        float a = a0 + a1;
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        if (part == 0) gh[row_local] = a + bg;
        __syncthreads();
        if (tid < GRU_UNITS) {
            const int j = rank * GRU_UNITS + tid;
            const float r = sigmoid_f(xr + gh[tid]);
            const float z = sigmoid_f(xz + gh[GRU_UNITS + tid]);
            const float n = tanhf(xn + r * gh[2 * GRU_UNITS + tid]);
            const float hn = (1.0f - z) * n + z * hbuf[cur][j];
            out[(long long)t * 2 * H + d * H + j] = hn;
#pragma unroll
            for (int peer = 0; peer < GRU_CL; ++peer) cluster.map_shared_rank(&hbuf[cur ^ 1][0], peer)[j] = hn;
        }
        cluster.sync();
    }
}

// ------------------------------------------------------------------------------------------
// enc_p embedding: lrelu_0.1((phone Wp^T + bp + emb_pitch[pitch]) * sqrt(H)); CTA per row
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embed_kernel(const float* __restrict__ phone, const int* __restrict__ pitch, const float* __restrict__ wp,
             const float* __restrict__ bp, const float* __restrict__ emb_pitch, float* __restrict__ out, long long ldo,
             int Cin, int H, long long wPhone, long long wPitch, long long wOut) {
    pdl_enter();
    phone += blockIdx.z * wPhone; pitch += blockIdx.z * wPitch; out += blockIdx.z * wOut;
    extern __shared__ float xs[];
    const int r = blockIdx.x;
    for (int i = threadIdx.x; i < Cin; i += blockDim.x) xs[i] = phone[(long long)r * Cin + i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int pi = pitch[r];
    const float sc = sqrtf(float(H));
    // 4 output channels per warp pass: their weight rows stream concurrently (latency-bound otherwise)
    for (int h0 = warp * 4; h0 < H; h0 += nw * 4) {
        float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 6
        for (int k = lane; k < Cin; k += 32) {
            const float x = xs[k];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (h0 + u < H) a[u] = fmaf(x, __ldg(wp + (long long)(h0 + u) * Cin + k), a[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float t = warp_sum(a[u]);
            if (lane == 0 && h0 + u < H) {
                float v = (t + bp[h0 + u] + emb_pitch[pi * H + h0 + u]) * sc;
                out[(long long)r * ldo + h0 + u] = v > 0.f ? v : 0.1f * v;
            }
        }
    }
}

__global__ void zp_kernel(const float* __restrict__ stats, float* __restrict__ out, long long ldo,
                          const RunParams* __restrict__ rp, int R, int H, long long wStats, long long wOut, long long wRp) {
    pdl_enter();
    stats += blockIdx.z * wStats; out += blockIdx.z * wOut;
    rp = reinterpret_cast<const RunParams*>(reinterpret_cast<const float*>(rp) + blockIdx.z * wRp);
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= R * H) return;
    const int r = e / H, c = e - r * H;
    const float m = stats[(long long)r * 2 * H + c], lg = stats[(long long)r * 2 * H + H + c];
    float nz = 0.f;
    if (rp->noise_mode) nz = noise_gauss(noise_key(rp->noise_seed, rp->window, NOISE_KIND_Z), (uint64_t)e);
    out[(long long)r * ldo + c] = m + expf(lg) * nz * 0.66666f;
}

__global__ void avg3_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                            long long ld, float* __restrict__ out, long long ldo, float* __restrict__ raw, long long ldraw,
                            int T, int C, float slope, long long wIn, long long wOut, long long wRaw) {
    pdl_enter();
    a += blockIdx.z * wIn; b += blockIdx.z * wIn; c += blockIdx.z * wIn; out += blockIdx.z * wOut;
    if (raw) raw += blockIdx.z * wRaw;
    const long long n = (long long)T * C;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        int ch = int(e % C);
        long long t = e / C;
        long long i = t * ld + ch;
        float s = (a[i] + b[i] + c[i]) / 3.0f;
        if (raw) raw[t * ldraw + ch] = s;
        out[t * ldo + ch] = s > 0.f ? s : slope * s;
    }
}


// conv_post: tanh(conv1d(C -> 1, k)); thread per output sample, weights in smem
__global__ void __launch_bounds__(256)
convpost_kernel(const float* __restrict__ in, const float* __restrict__ w, float* __restrict__ out, int T, int C, int k,
                long long wIn, long long wOut) {
    pdl_enter();
    in += blockIdx.z * wIn; out += blockIdx.z * wOut;
    extern __shared__ float ws[];
    const int n = k * C;
    for (int i = threadIdx.x; i < n; i += blockDim.x) ws[i] = w[i];
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const float4* p = reinterpret_cast<const float4*>(in + (long long)t * C);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int j = 0; j < n / 4; ++j) {
        float4 v = __ldg(p + j);
        a0 = fmaf(v.x, ws[4 * j + 0], a0); a1 = fmaf(v.y, ws[4 * j + 1], a1);
        a2 = fmaf(v.z, ws[4 * j + 2], a2); a3 = fmaf(v.w, ws[4 * j + 3], a3);
    }
    out[t] = tanhf((a0 + a1) + (a2 + a3));
    lane_stamp(&g_misc_stamps[1]);
}

__global__ void gather_rows_kernel(const float* __restrict__ src, long long lds, float* __restrict__ out, int T, int C,
                                   int skip, int R, int row0, long long wSrc, long long wOut) {
    pdl_enter();
    lane_stamp(&g_misc_stamps[0]);
    src += blockIdx.z * wSrc; out += blockIdx.z * wOut;
    const long long n = (long long)R * C;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        int c = int(e % C), r = int(e / C);
        int s = min((skip + r) / 2, T - 1) - row0;
        out[e] = src[(long long)s * lds + c];
    }
}

__global__ void split_hilo_kernel(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo, size_t n) {
    pdl_enter();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = src[i];
        const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        hi[i] = h; lo[i] = v - h;
    }
}

__global__ void split_hilo16_kernel(const float* __restrict__ src, unsigned short* __restrict__ hi, unsigned short* __restrict__ lo, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = src[i];
        const __half h = __float2half_rn(v);
        hi[i] = __half_as_ushort(h);
        lo[i] = __half_as_ushort(__float2half_rn((v - __half2float(h)) * 2048.0f));
    }
}

inline int grid_for(long long n, int block, int cap = 148 * 8) {
    long long g = (n + block - 1) / block;
    return int(g < 1 ? 1 : (g > cap ? cap : g));
}

size_t attn_smem(int T) { const int Tp = (T | 1) + 2; return sizeof(float) * (size_t(T) * ATT_D + size_t(ATT_D) * Tp + size_t(2 * ATT_WARPS) * Tp); }   // two score rows per warp (batched plans)
size_t relattn_smem(int T, int dim) { return sizeof(float) * (size_t(dim) + size_t(T)); }

}  // namespace

void misc_read_stamps(unsigned long long* out8) { cudaMemcpyFromSymbol(out8, g_misc_stamps, 8 * sizeof(unsigned long long)); }

void init_kernel_attributes() {
    static unsigned long long done = 0;
    if (!first_time_on_device(done)) return;
    cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(relattn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    init_gemm_v2_attributes();
    init_umma_attributes();
    init_chain_attributes();
    init_cvstack_attributes();
    init_cbr_attributes();
    init_wstream_attributes();
}

void launch_split_hilo16(const float* src, unsigned short* dst_hi, unsigned short* dst_lo, size_t n, cudaStream_t stream) {
    split_hilo16_kernel<<<148 * 8, 256, 0, stream>>>(src, dst_hi, dst_lo, n);
}

void launch_split_hilo(const float* src, float* dst_hi, float* dst_lo, size_t n, cudaStream_t stream) {
    launch_k(split_hilo_kernel, dim3(148 * 8), dim3(256), size_t(0), stream, src, dst_hi, dst_lo, n);
}

int launch_layernorm(const LayerNormOp& o, const DeviceBases& B, cudaStream_t s) {
    const int wpb = 4;
    launch_k(layernorm_kernel, dim3((o.rows + wpb - 1) / wpb, 1, B.nb), dim3(wpb * 32), size_t(0), s, B.p<float>(o.X), o.ldx, B.p<float>(o.Y), o.ldy,
                                                                  B.p<float>(o.gamma), B.p<float>(o.beta), o.rows, o.cols, o.eps, B.ws(o.X), B.ws(o.Y));
    return 1;
}

int launch_attn(const AttnOp& o, const DeviceBases& B, cudaStream_t s) {
    // head dim is 64 for every ContentVec variant (validated by the plan builder)
    const int rows = B.nb >= 8 ? 8 * ATT_WARPS : (B.nb > 1 ? 2 * ATT_WARPS : ATT_WARPS);
    dim3 grid(o.heads, (o.T + rows - 1) / rows, B.nb);
    launch_k(attn_kernel, grid, dim3(ATT_WARPS * 32), attn_smem(o.T), s, B.p<float>(o.qkv), o.ldqkv, B.p<float>(o.out), o.ldo, o.T, o.heads,
             B.ws(o.qkv), B.ws(o.out), rows);
    return 1;
}

int launch_relattn(const RelAttnOp& o, const DeviceBases& B, cudaStream_t s) {
    launch_k(relattn_kernel, dim3(o.heads * o.T, 1, B.nb), dim3(128), relattn_smem(o.T, o.dim), s, B.p<float>(o.qkv), o.ldqkv, B.p<float>(o.out), o.ldo,
             B.p<float>(o.rel_k), B.p<float>(o.rel_v), o.T, o.heads, o.dim, o.window, B.ws(o.qkv), B.ws(o.out));
    return 1;
}

int launch_conv0_stats(const Conv0StatsOp& o, const DeviceBases& B, cudaStream_t s) {
    launch_k(conv0_stats_kernel, dim3((o.C + C0_CH - 1) / C0_CH, 1, B.nb), dim3(256), size_t(0), s, B.p<float>(o.pcm), B.p<float>(o.w), B.p<float>(o.stats), o.T, o.C,
                                                                 o.k, o.stride, o.eps, B.ws(o.pcm), B.ws(o.stats));
    return 1;
}

int launch_conv0_apply(const Conv0ApplyOp& o, const DeviceBases& B, cudaStream_t s) {
    launch_k(conv0_apply_kernel, dim3((o.T + C0_TB - 1) / C0_TB, 1, B.nb), dim3(256), size_t(0), s, B.p<float>(o.pcm), B.p<float>(o.w), B.p<float>(o.stats),
                                                                 B.p<float>(o.gamma), B.p<float>(o.beta), B.p<float>(o.Y), o.T, o.C,
                                                                 o.k, o.stride, B.ws(o.pcm), B.ws(o.stats), B.ws(o.Y));
    return 1;
}

int launch_avgpool(const AvgPoolOp& o, const DeviceBases& B, cudaStream_t s) {
    long long n = (long long)(o.T / 2) * (o.F / 2) * o.C;
    launch_k(avgpool_kernel, dim3(grid_for(n, 256), 1, B.nb), dim3(256), size_t(0), s, B.p<float>(o.in), o.ldin, B.p<float>(o.out), o.T, o.F, o.C,
             B.ws(o.in), B.ws(o.out));
    return 1;
}

int launch_gru(const GruOp& o, const DeviceBases& B, cudaStream_t s) {
    // H is fixed by the RMVPE architecture (BiGRU(384, 256)); validated when the model is packed
    launch_k(gru_cluster_kernel, dim3(2 * GRU_CL, 1, B.nb), dim3(768), size_t(0), s, B.p<float>(o.gi), B.p<float>(o.whh_t), B.p<float>(o.bhh), B.p<float>(o.out), o.T,
             B.ws(o.gi), B.ws(o.out));
    return 1;
}

// the pitch half of the embedding when the phone projection was computed ahead of the join (EmbedOp::pre)
__global__ void __launch_bounds__(256)
embed_add_kernel(const float* __restrict__ pre, const int* __restrict__ pitch, const float* __restrict__ emb_pitch, float* __restrict__ out,
                 long long ldo, int R, int H, long long wPre, long long wPitch, long long wOut) {
    pdl_enter();
    pre += blockIdx.z * wPre; pitch += blockIdx.z * wPitch; out += blockIdx.z * wOut;
    const float sc = sqrtf(float(H));
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < R * H; e += gridDim.x * blockDim.x) {
        const int r = e / H, h = e - r * H;
        const float v = (pre[e] + emb_pitch[pitch[r] * H + h]) * sc;
        out[(long long)r * ldo + h] = v > 0.f ? v : 0.1f * v;
    }
}

int launch_embed(const EmbedOp& o, const DeviceBases& B, cudaStream_t s) {
    if (!o.pre.null()) {
        launch_k(embed_add_kernel, dim3((o.R * o.H + 255) / 256, 1, B.nb), dim3(256), size_t(0), s, B.p<float>(o.pre), B.p<int>(o.pitch), B.p<float>(o.emb_pitch),
                 B.p<float>(o.out), o.ldo, o.R, o.H, B.ws(o.pre), B.ws(o.pitch), B.ws(o.out));
        return 1;
    }
    launch_k(embed_kernel, dim3(o.R, 1, B.nb), dim3(256), size_t(sizeof(float) * o.Cin), s, B.p<float>(o.phone), B.p<int>(o.pitch), B.p<float>(o.wp), B.p<float>(o.bp),
                                                         B.p<float>(o.emb_pitch), B.p<float>(o.out), o.ldo, o.Cin, o.H, B.ws(o.phone), B.ws(o.pitch), B.ws(o.out));
    return 1;
}

int launch_zp(const ZpOp& o, const DeviceBases& B, cudaStream_t s) {
    launch_k(zp_kernel, dim3((o.R * o.H + 255) / 256, 1, B.nb), dim3(256), size_t(0), s, B.p<float>(o.stats), B.p<float>(o.out), o.ldo, B.p<RunParams>(o.params), o.R, o.H,
             B.ws(o.stats), B.ws(o.out), B.ws(o.params));
    return 1;
}

int launch_avg3(const Avg3Op& o, const DeviceBases& B, cudaStream_t s) {
    launch_k(avg3_kernel, dim3(grid_for((long long)o.T * o.C, 256), 1, B.nb), dim3(256), size_t(0), s, B.p<float>(o.a), B.p<float>(o.b), B.p<float>(o.c), o.ld,
                                                                    B.p<float>(o.out), o.ldo, B.p<float>(o.raw), o.ldraw, o.T, o.C,
                                                                    o.slope, B.ws(o.a), B.ws(o.out), B.ws(o.raw));
    return 1;
}

int launch_convpost(const ConvPostOp& o, const DeviceBases& B, cudaStream_t s) {
    launch_k(convpost_kernel, dim3((o.T + 255) / 256, 1, B.nb), dim3(256), size_t(sizeof(float) * o.k * o.C), s, B.p<float>(o.in), B.p<float>(o.w), B.p<float>(o.out), o.T,
                                                                              o.C, o.k, B.ws(o.in), B.ws(o.out));
    return 1;
}

int launch_gather_rows(const GatherRowsOp& o, const DeviceBases& B, cudaStream_t s) {
    launch_k(gather_rows_kernel, dim3(grid_for((long long)o.R * o.C, 256), 1, B.nb), dim3(256), size_t(0), s, B.p<float>(o.src), o.lds, B.p<float>(o.out), o.T, o.C,
                                                                          o.skip, o.R, o.row0, B.ws(o.src), B.ws(o.out));
    return 1;
}

}  // namespace rvc
