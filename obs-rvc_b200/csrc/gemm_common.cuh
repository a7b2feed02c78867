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
};

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
    return p;
}

}  // namespace gemmk
}  // namespace rvc
