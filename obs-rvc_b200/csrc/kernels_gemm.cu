// kernels_gemm.cu - fp32 implicit-GEMM kernel family (CUDA cores) for sm_100a.
//
// One contraction serves every conv1d / conv2d-3x3 / transposed conv / linear layer of the three
// networks (ops.h GemmOp): A rows are "segmented" views of channels-last halo-padded activations,
// W is [N,K] K-major.  This file is the exact-fp32 path: it is what the small, skinny and
// oddly-shaped contractions run on and what parity is established with.  The dense, large-M
// contractions are additionally served by the tcgen05/TMA kernel in kernels_umma.cu.
//
// Tiling: BMxBN output tile per CTA, BK=16, 256 threads as a 16x16 grid of TMxTN micro-tiles,
// double-buffered shared memory with register prefetch (one __syncthreads per k-tile).  Global
// loads are float4 along K (each lane a different row -> transposed, conflict-free STS); the
// k-major shared layout gives LDS.128 operand fetches.
#include <algorithm>
#include <cstdio>

#include "gemm_common.cuh"
#include "launch.h"
#include "pdl.cuh"

namespace rvc {

namespace {

constexpr int BK = 16;

using namespace gemmk;

// loads 4 consecutive k of one row (zero beyond the row / K range)
template <bool VEC>
__device__ __forceinline__ float4 load_a4(const GemmParams& p, const float* __restrict__ A, int m, int kk) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m >= p.M || kk >= p.K) return v;
    if constexpr (VEC) {
        int seg = kk / p.seg_len, within = kk - seg * p.seg_len;
        return __ldg(reinterpret_cast<const float4*>(A + (long long)m * p.lda + (long long)seg * p.seg_stride + within));
    }
    float t[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int k = kk + j;
        if (k < p.K) {
            int seg = k / p.seg_len, within = k - seg * p.seg_len;
            t[j] = __ldg(A + (long long)m * p.lda + (long long)seg * p.seg_stride + within);
        } else {
            t[j] = 0.f;
        }
    }
    return make_float4(t[0], t[1], t[2], t[3]);
}

template <bool VEC>
__device__ __forceinline__ float4 load_w4(const GemmParams& p, const float* __restrict__ W, int n, int kk) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n >= p.N || kk >= p.K) return v;
    if constexpr (VEC) return __ldg(reinterpret_cast<const float4*>(W + (long long)n * p.ldw + kk));
    float t[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) t[j] = (kk + j < p.K) ? __ldg(W + (long long)n * p.ldw + kk + j) : 0.f;
    return make_float4(t[0], t[1], t[2], t[3]);
}

template <int BM, int BN, int TM, int TN, bool VEC>
__global__ void __launch_bounds__(256) gemm_f32_kernel(GemmParams p) {
    static_assert((BM / TM) * (BN / TN) == 256, "256 threads");
    static_assert(BN / TN == 16, "16 thread columns");
    constexpr int A_F4 = BM * BK / 4, W_F4 = BN * BK / 4;  // float4 loads per k-tile
    constexpr int A_PER = (A_F4 + 255) / 256, W_PER = (W_F4 + 255) / 256;
    __shared__ __align__(16) float As[2][BK][BM];
    __shared__ __align__(16) float Ws[2][BK][BN];

    pdl_enter();
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int bz = blockIdx.z;
    const GemmBases gb = gemm_bases(p, bz);
    const float* __restrict__ A = gb.A;
    const float* __restrict__ W = gb.W;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float4 ra[A_PER], rw[W_PER];
    auto gload = [&](int k0) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            int f = tid + i * 256;
            if (A_F4 >= 256 || f < A_F4) ra[i] = load_a4<VEC>(p, A, m0 + (f % BM), k0 + (f / BM) * 4);
        }
#pragma unroll
        for (int i = 0; i < W_PER; ++i) {
            int f = tid + i * 256;
            if (W_F4 >= 256 || f < W_F4) rw[i] = load_w4<VEC>(p, W, n0 + (f % BN), k0 + (f / BN) * 4);
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            int f = tid + i * 256;
            if (A_F4 >= 256 || f < A_F4) {
                int r = f % BM, kq = (f / BM) * 4;
                As[buf][kq + 0][r] = ra[i].x; As[buf][kq + 1][r] = ra[i].y;
                As[buf][kq + 2][r] = ra[i].z; As[buf][kq + 3][r] = ra[i].w;
            }
        }
#pragma unroll
        for (int i = 0; i < W_PER; ++i) {
            int f = tid + i * 256;
            if (W_F4 >= 256 || f < W_F4) {
                int r = f % BN, kq = (f / BN) * 4;
                Ws[buf][kq + 0][r] = rw[i].x; Ws[buf][kq + 1][r] = rw[i].y;
                Ws[buf][kq + 2][r] = rw[i].z; Ws[buf][kq + 3][r] = rw[i].w;
            }
        }
    };

    const int nkt = (p.K + BK - 1) / BK;
    gload(0);
    sstore(0);
    __syncthreads();
    for (int kt = 0; kt < nkt; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nkt) gload((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
            if (TM % 4 == 0) {
#pragma unroll
                for (int i = 0; i < TM; i += 4) {
                    float4 v = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM + i]);
                    a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < TM; ++i) a[i] = As[buf][k][ty * TM + i];
            }
#pragma unroll
            for (int j = 0; j < TN; j += 4) {
                float4 v = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * TN + j]);
                b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nkt) sstore(buf ^ 1);
        __syncthreads();
    }

    // ---- epilogue -------------------------------------------------------------------------
    const float* __restrict__ bias = gb.bias;
    // C / C2 / R may alias (in-place accumulation: R == C): plain loads, no __restrict__
    float* C = gb.C;
    float* C2 = gb.C2;
    const float* R = gb.R;
    const int nbase = n0 + tx * TN;
    float bv[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) bv[j] = (bias && nbase + j < p.N) ? __ldg(bias + nbase + j) : 0.f;

#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i;
        if (m >= p.M) continue;
        const bool masked = p.mask_period > 0 && (m % p.mask_period) >= p.mask_valid;
        float v[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) v[j] = fmaf(p.alpha, acc[i][j], bv[j]);
        if (p.act == ACT_GATE) {
#pragma unroll
            for (int j = 0; j < TN; j += 2) {
                const int n = nbase + j;
                if (n + 1 < p.N) {
                    float g = tanhf(v[j]) * sigmoid_f(v[j + 1]);
                    const int col = n >> 1;
                    if (R) g += R[(long long)m * p.ldr + col];
                    if (masked) g = 0.f;
                    C[(long long)m * p.ldc + col] = g;
                    if (C2) C2[(long long)m * p.ldc2 + col] = masked ? 0.f : apply_act(p.act2, g);
                }
            }
            continue;
        }
#pragma unroll
        for (int j = 0; j < TN; ++j) v[j] = apply_act(p.act, v[j]);
        if (p.out_mode == OUT_PLAIN) {
            if (R) {
#pragma unroll
                for (int j = 0; j < TN; ++j)
                    if (nbase + j < p.N) v[j] += R[(long long)m * p.ldr + nbase + j];
            }
            if (masked) {
#pragma unroll
                for (int j = 0; j < TN; ++j) v[j] = 0.f;
            }
            if (p.vec_store && nbase + TN <= p.N) {
#pragma unroll
                for (int j = 0; j < TN; j += 4)
                    *reinterpret_cast<float4*>(C + (long long)m * p.ldc + nbase + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                if (C2) {
#pragma unroll
                    for (int j = 0; j < TN; j += 4)
                        *reinterpret_cast<float4*>(C2 + (long long)m * p.ldc2 + nbase + j) =
                            make_float4(masked ? 0.f : apply_act(p.act2, v[j]), masked ? 0.f : apply_act(p.act2, v[j + 1]),
                                        masked ? 0.f : apply_act(p.act2, v[j + 2]), masked ? 0.f : apply_act(p.act2, v[j + 3]));
                }
            } else {
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    if (nbase + j >= p.N) continue;
                    C[(long long)m * p.ldc + nbase + j] = v[j];
                    if (C2) C2[(long long)m * p.ldc2 + nbase + j] = masked ? 0.f : apply_act(p.act2, v[j]);
                }
            }
        } else if (p.out_mode == OUT_PIXSHUF2) {
            const int qt = m / p.om_a, qf = m - qt * p.om_a;
            if (qf >= p.om_a - 2) continue;
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int n = nbase + j;
                if (n >= p.N) continue;
                const int ph = n / p.om_b, co = n - ph * p.om_b, rt = ph >> 1, rf = ph & 1;
                const long long idx = ((long long)(2 * qt + rt) * p.om_c + (2 * qf + rf)) * p.ldc + co;
                C[idx] = masked ? 0.f : v[j];
            }
        } else {  // OUT_CONVT1D
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int n = nbase + j;
                if (n >= p.N) continue;
                const int r = n / p.om_b, co = n - r * p.om_b;
                const int o = m * p.om_a + r - p.om_c;
                if (o < 0 || o >= p.om_d) continue;
                C[(long long)o * p.ldc + co] = v[j];
                if (C2) C2[(long long)o * p.ldc2 + co] = apply_act(p.act2, v[j]);
            }
        }
    }
}

// K <= 32 contractions (1x1 / first-layer convs, Conv1d(1,C,k)): one thread per output element, the
// whole K loop in registers; coalesced along N.  These are outer-product sized - a tiled GEMM only
// adds latency.
__global__ void __launch_bounds__(256) gemm_smallk_kernel(GemmParams p) {
    pdl_enter();
    const int ncols = p.N;
    const long long total = (long long)p.M * ncols;
    const int bz = blockIdx.y;
    const GemmBases gb = gemm_bases(p, bz);
    const float* __restrict__ A = gb.A;
    const float* __restrict__ W = gb.W;
    const float* __restrict__ bias = gb.bias;
    float* C = gb.C;
    float* C2 = gb.C2;
    const float* R = gb.R;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int m = int(e / ncols), n = int(e - (long long)m * ncols);
        float acc = 0.f, accp = 0.f;
        for (int k = 0; k < p.K; ++k) {
            const int seg = k / p.seg_len, within = k - seg * p.seg_len;
            const float a = __ldg(A + (long long)m * p.lda + (long long)seg * p.seg_stride + within);
            acc = fmaf(a, __ldg(W + (long long)n * p.ldw + k), acc);
            if (p.act == ACT_GATE) accp = fmaf(a, __ldg(W + (long long)(n ^ 1) * p.ldw + k), accp);
        }
        epilogue_elem(p, bias, C, C2, R, m, n, acc, accp);
    }
}

template <int BM, int BN, int TM, int TN>
void launch_cfg(const GemmParams& p, int batch, bool vec, cudaStream_t s) {
    dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, batch);
    if (vec) launch_k(gemm_f32_kernel<BM, BN, TM, TN, true>, grid, dim3(256), size_t(0), s, p);
    else launch_k(gemm_f32_kernel<BM, BN, TM, TN, false>, grid, dim3(256), size_t(0), s, p);
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

int launch_gemm_v2(const GemmOp& g, const DeviceBases& B, cudaStream_t stream);  // kernels_gemm2.cu

int launch_gemm(const GemmOp& g, const DeviceBases& B, cudaStream_t stream) {
    if (g.sched_variant >= 5) {
        if (launch_gemm_umma(g, B, stream)) return 1;
        // tensor maps unavailable: exact CUDA-core kernel below (no split-K scratch needed)
    } else if (g.sched_variant > 0) {
        return launch_gemm_v2(g, B, stream);
    }
    GemmParams p = make_params(g, B);
    if (g.K <= 32) {
        const long long total = (long long)g.M * g.N;
        const int blocks = int(std::min<long long>((total + 255) / 256, 148 * 16));
        launch_k(gemm_smallk_kernel, dim3(blocks, g.batch * B.nb), dim3(256), size_t(0), stream, p);
        return 1;
    }
    const bool vec = al16(p.A) && al16(p.W) && g.lda % 4 == 0 && g.seg_len % 4 == 0 && g.seg_stride % 4 == 0 &&
                     g.K % 4 == 0 && g.ldw % 4 == 0 && g.sA % 4 == 0 && g.sW % 4 == 0;
    p.vec_store = (g.out_mode == OUT_PLAIN && g.act != ACT_GATE && al16(p.C) && g.ldc % 4 == 0 && g.sC % 4 == 0 &&
                   (!p.C2 || (al16(p.C2) && g.ldc2 % 4 == 0))) ? 1 : 0;
    if (g.M >= 1024) launch_cfg<128, 64, 8, 4>(p, g.batch * B.nb, vec, stream);
    else if (g.M > 32) launch_cfg<64, 64, 4, 4>(p, g.batch * B.nb, vec, stream);
    else launch_cfg<16, 64, 1, 4>(p, g.batch * B.nb, vec, stream);
    return 1;
}

}  // namespace rvc
