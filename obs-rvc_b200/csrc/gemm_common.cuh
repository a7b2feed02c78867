// gemm_common.cuh - parameter block + epilogue helpers shared by the GEMM kernels.
#pragma once
#include "launch.h"

namespace rvc {
namespace gemmk {

struct GemmParams {
    const float* A; long long lda; int seg_len; long long seg_stride;
    const float* W; long long ldw;
    const float* bias;
    float* C; long long ldc;
    float* C2; long long ldc2; int act2;
    const float* R; long long ldr;
    int M, N, K, act;
    float alpha;
    int mask_period, mask_valid;
    long long sA, sW, sBias, sC, sR;
    int out_mode, om_a, om_b, om_c, om_d;
    int vec_store;
    // split-K (v2 kernel): partial tiles in scratch[z][M][N] per batch, one counter per output tile
    float* scratch; unsigned int* counters; int splitk, kt_per_split;
    // batched plans: the grid's batch index runs over windows x groups; `batch` = groups per window (the op's own
    // batch count, e.g. the 16 groups of ContentVec's pos-conv), w* = per-window element strides (0 for weights)
    int batch;
    long long wA, wC, wC2, wR, wScratch, wCounters;
};

// batch index -> (window, group) and the operand bases of that pair
struct GemmBases { const float* A; const float* W; const float* bias; float* C; float* C2; const float* R; float* scratch; unsigned int* counters; int grp; };
__device__ __forceinline__ GemmBases gemm_bases(const GemmParams& p, int bz) {
    const int win = bz / p.batch, grp = bz - win * p.batch;
    GemmBases b;
    b.A = p.A + grp * p.sA + win * p.wA;
    b.W = p.W + grp * p.sW;
    b.bias = p.bias ? p.bias + grp * p.sBias : nullptr;
    b.C = p.C + grp * p.sC + win * p.wC;
    b.C2 = p.C2 ? p.C2 + grp * p.sC + win * p.wC2 : nullptr;
    b.R = p.R ? p.R + grp * p.sR + win * p.wR : nullptr;
    b.scratch = p.scratch ? p.scratch + win * p.wScratch : nullptr;
    b.counters = p.counters ? p.counters + win * p.wCounters : nullptr;
    b.grp = grp;
    return b;
}

__device__ __forceinline__ float gelu_f(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }
__device__ __forceinline__ float sigmoid_f(float v) { return 1.0f / (1.0f + expf(-v)); }

__device__ __forceinline__ float apply_act(int act, float v) {
    switch (act) {
        case ACT_GELU: return gelu_f(v);
        case ACT_RELU: return fmaxf(v, 0.0f);
        case ACT_LRELU01: return v > 0.0f ? v : 0.1f * v;
        case ACT_LRELU001: return v > 0.0f ? v : 0.01f * v;
        case ACT_SIGMOID: return sigmoid_f(v);
        case ACT_TANH: return tanhf(v);
        default: return v;
    }
}

// final accumulator -> output element (bias, activation, residual, masks, scatter modes); mirrors the
// scalar interpreter in oracle/plan_exec/cpu_exec.cpp
__device__ __forceinline__ void epilogue_elem(const GemmParams& p, const float* __restrict__ bias, float* C, float* C2,
                                              const float* R, int m, int n, float acc, float acc_partner) {
    const bool masked = p.mask_period > 0 && (m % p.mask_period) >= p.mask_valid;
    if (p.act == ACT_GATE) {
        if (n & 1) return;
        const float v0 = fmaf(p.alpha, acc, bias ? __ldg(bias + n) : 0.f);
        const float v1 = fmaf(p.alpha, acc_partner, bias ? __ldg(bias + n + 1) : 0.f);
        float g = tanhf(v0) * sigmoid_f(v1);
        const int col = n >> 1;
        if (R) g += R[(long long)m * p.ldr + col];
        if (masked) g = 0.f;
        C[(long long)m * p.ldc + col] = g;
        if (C2) C2[(long long)m * p.ldc2 + col] = masked ? 0.f : apply_act(p.act2, g);
        return;
    }
    float v = apply_act(p.act, fmaf(p.alpha, acc, bias ? __ldg(bias + n) : 0.f));
    if (p.out_mode == OUT_PLAIN) {
        if (R) v += R[(long long)m * p.ldr + n];
        if (masked) v = 0.f;
        C[(long long)m * p.ldc + n] = v;
        if (C2) C2[(long long)m * p.ldc2 + n] = masked ? 0.f : apply_act(p.act2, v);
    } else if (p.out_mode == OUT_PIXSHUF2) {
        const int qt = m / p.om_a, qf = m - qt * p.om_a;
        if (qf >= p.om_a - 2) return;
        const int ph = n / p.om_b, co = n - ph * p.om_b, rt = ph >> 1, rf = ph & 1;
        C[((long long)(2 * qt + rt) * p.om_c + (2 * qf + rf)) * p.ldc + co] = masked ? 0.f : v;
    } else {  // OUT_CONVT1D
        const int r = n / p.om_b, co = n - r * p.om_b;
        const int o = m * p.om_a + r - p.om_c;
        if (o < 0 || o >= p.om_d) return;
        C[(long long)o * p.ldc + co] = v;
        if (C2) C2[(long long)o * p.ldc2 + co] = apply_act(p.act2, v);
    }
}

inline GemmParams make_params(const GemmOp& g, const DeviceBases& B) {
    GemmParams p;
    p.A = B.p<float>(g.A); p.lda = g.lda; p.seg_len = g.seg_len; p.seg_stride = g.seg_stride;
    p.W = B.p<float>(g.W); p.ldw = g.ldw; p.bias = B.p<float>(g.bias);
    p.C = B.p<float>(g.C); p.ldc = g.ldc; p.C2 = B.p<float>(g.C2); p.ldc2 = g.ldc2; p.act2 = g.act2;
    p.R = B.p<float>(g.R); p.ldr = g.ldr;
    p.M = g.M; p.N = g.N; p.K = g.K; p.act = g.act; p.alpha = g.alpha;
    p.mask_period = g.mask_period; p.mask_valid = g.mask_valid;
    p.sA = g.sA; p.sW = g.sW; p.sBias = g.sBias; p.sC = g.sC; p.sR = g.sR;
    p.out_mode = g.out_mode; p.om_a = g.om_a; p.om_b = g.om_b; p.om_c = g.om_c; p.om_d = g.om_d;
    p.vec_store = 0;
    p.scratch = B.p<float>(g.scratch); p.counters = B.p<unsigned int>(g.counters); p.splitk = g.splitk; p.kt_per_split = 0;
    p.batch = g.batch;
    p.wA = B.ws(g.A); p.wC = B.ws(g.C); p.wC2 = B.ws(g.C2); p.wR = B.ws(g.R); p.wScratch = B.ws(g.scratch); p.wCounters = B.ws(g.counters);
    return p;
}

}  // namespace gemmk
}  // namespace rvc
