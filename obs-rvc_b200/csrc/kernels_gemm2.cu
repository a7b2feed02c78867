// kernels_gemm2.cu - fp32 implicit GEMM, v2: cp.async multi-stage pipeline + split-K.
//
// Why: at batch 1 the path's contractions are 4..111 rows against MB-sized weight matrices, i.e.
// bound by streaming W from HBM.  This kernel keeps 3-4 k-tiles (BK=32) of 16-byte cp.async loads
// in flight per CTA and splits K across CTAs so that the launch covers ~2 CTAs per SM; partial
// tiles are written to a per-op scratch area and the last CTA to arrive on an output tile (one
// atomic ticket per tile) reduces them in fixed z order and runs the epilogue - one launch,
// deterministic summation order.
//
// Thread layout: lane <-> output column (consecutive W rows), each thread accumulates TM rows x
// TN columns; operands are read from shared memory as float4 along K: the A fragment is a
// broadcast (all lanes, same address), the W fragment is conflict-free because rows are padded to
// 36 words (quarter-warp of 8 lanes x 16 B covers all 32 banks).  Same arithmetic (fp32 FMA,
// k ascending) as the v1 kernel; only the order of the split-K partial sums differs.
#include "gemm_common.cuh"
#include "gemm_sched.h"
#include "launch.h"

namespace rvc {

namespace {

using namespace gemmk;

constexpr int BK2 = GEMM2_BK;   // 32 floats = 128 B per row per k-tile
constexpr int LDS2 = 36;        // padded row stride (words)

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, int src_bytes) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int TM, int WM, int WN, int TN, int STAGES>
__global__ void __launch_bounds__(256) gemm_v2_kernel(GemmParams p) {
    constexpr int BM = TM * WM, BN = 32 * TN * WN;
    static_assert(WM * WN == 8, "8 warps");
    constexpr int STAGE_FLOATS = (BM + BN) * LDS2;
    constexpr int CHUNKS = (BM + BN) * (BK2 / 4);          // 16-byte chunks per stage
    constexpr int CPT = (CHUNKS + 255) / 256;              // chunks per thread
    extern __shared__ __align__(16) float smem[];
    __shared__ int s_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WN, wn = warp % WN;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int z = blockIdx.z % p.splitk, bz = blockIdx.z / p.splitk;
    const float* __restrict__ A = p.A + bz * p.sA;
    const float* __restrict__ W = p.W + bz * p.sW;
    const int nkt_total = (p.K + BK2 - 1) / BK2;
    const int kt0 = z * p.kt_per_split, kt1 = min(nkt_total, kt0 + p.kt_per_split);
    const int nkt = max(0, kt1 - kt0);

    // per-thread chunk descriptors (row base pointers are k-invariant)
    const float* cbase[CPT]; int ckc[CPT]; int cdst[CPT]; bool cisA[CPT], cvalid[CPT];
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
        const int c = tid + i * 256;
        cvalid[i] = c < CHUNKS;
        const int row = c / (BK2 / 4), kc = c % (BK2 / 4);
        ckc[i] = kc * 4;
        cisA[i] = row < BM;
        cdst[i] = row * LDS2 + kc * 4;
        if (cisA[i]) {
            const int m = m0 + row;
            cbase[i] = (cvalid[i] && m < p.M) ? A + (long long)m * p.lda : nullptr;
        } else {
            const int n = n0 + row - BM;
            cbase[i] = (cvalid[i] && n < p.N) ? W + (long long)n * p.ldw : nullptr;
        }
    }
    auto load_stage = [&](int slot, int kt) {
        float* st = smem + slot * STAGE_FLOATS;
        const int k0 = kt * BK2;
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
            if (!cvalid[i]) continue;
            const int kk = k0 + ckc[i];
            const float* src = p.W;  // any valid address when zero-filling
            int bytes = 0;
            if (cbase[i] && kk < p.K) {
                if (cisA[i]) {
                    const int seg = kk / p.seg_len, within = kk - seg * p.seg_len;
                    src = cbase[i] + (long long)seg * p.seg_stride + within;
                } else {
                    src = cbase[i] + kk;
                }
                bytes = 16;
            }
            cp_async16(st + cdst[i], src, bytes);
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nkt) load_stage(s, kt0 + s);
        cp_async_commit();
    }
    for (int kt = 0; kt < nkt; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (kt + STAGES - 1 < nkt) load_stage((kt + STAGES - 1) % STAGES, kt0 + kt + STAGES - 1);
        cp_async_commit();
        const float* As = smem + (kt % STAGES) * STAGE_FLOATS + (wm * TM) * LDS2;
        const float* Ws = smem + (kt % STAGES) * STAGE_FLOATS + (BM + wn * 32 * TN + lane) * LDS2;
#pragma unroll
        for (int kq = 0; kq < BK2 / 4; ++kq) {
            float4 b[TN];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = *reinterpret_cast<const float4*>(Ws + j * 32 * LDS2 + kq * 4);
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                const float4 a = *reinterpret_cast<const float4*>(As + i * LDS2 + kq * 4);
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    acc[i][j] = fmaf(a.x, b[j].x, acc[i][j]);
                    acc[i][j] = fmaf(a.y, b[j].y, acc[i][j]);
                    acc[i][j] = fmaf(a.z, b[j].z, acc[i][j]);
                    acc[i][j] = fmaf(a.w, b[j].w, acc[i][j]);
                }
            }
        }
    }
    cp_async_wait<0>();

    const float* __restrict__ bias = p.bias ? p.bias + bz * p.sBias : nullptr;
    float* C = p.C + bz * p.sC;
    float* C2 = p.C2 ? p.C2 + bz * p.sC : nullptr;
    const float* R = p.R ? p.R + bz * p.sR : nullptr;
    const int mbase = m0 + wm * TM, nbase = n0 + wn * 32 * TN + lane;

    if (p.splitk > 1) {
        float* part = p.scratch + ((long long)(bz * p.splitk + z) * p.M) * p.N;
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int m = mbase + i;
            if (m >= p.M) continue;
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int n = nbase + j * 32;
                if (n < p.N) __stcg(part + (long long)m * p.N + n, acc[i][j]);
            }
        }
        __threadfence();
        __syncthreads();
        const int tile = (bz * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        if (tid == 0) {
            const unsigned ticket = atomicAdd(p.counters + tile, 1u);
            s_last = (ticket == unsigned(p.splitk - 1)) ? 1 : 0;
            if (s_last) p.counters[tile] = 0;  // self-cleaning for the next launch / graph replay
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        const float* base = p.scratch + ((long long)(bz * p.splitk) * p.M) * p.N;
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int m = mbase + i;
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int n = nbase + j * 32;
                float s = 0.f;
                if (m < p.M && n < p.N)
                    for (int zz = 0; zz < p.splitk; ++zz) s += __ldcg(base + ((long long)zz * p.M + m) * p.N + n);
                acc[i][j] = s;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = mbase + i;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = nbase + j * 32;
            const float partner = __shfl_xor_sync(0xffffffffu, acc[i][j], 1);
            if (m < p.M && n < p.N) epilogue_elem(p, bias, C, C2, R, m, n, acc[i][j], partner);
        }
    }
}

template <int TM, int WM, int WN, int TN, int STAGES>
void launch_v2(const GemmOp& g, GemmParams& p, cudaStream_t s) {
    constexpr int BM = TM * WM, BN = 32 * TN * WN;
    constexpr size_t smem = sizeof(float) * size_t(STAGES) * (BM + BN) * LDS2;
    auto kern = gemm_v2_kernel<TM, WM, WN, TN, STAGES>;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)); attr = true; }
    const int nkt = (g.K + BK2 - 1) / BK2;
    p.kt_per_split = (nkt + g.splitk - 1) / g.splitk;
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, g.batch * g.splitk);
    kern<<<grid, 256, smem, s>>>(p);
}

}  // namespace

int launch_gemm_v2(const GemmOp& g, const DeviceBases& B, cudaStream_t stream) {
    GemmParams p = gemmk::make_params(g, B);
    switch (g.sched_variant) {
        case 1: launch_v2<8, 1, 8, 1, 3>(g, p, stream); break;
        case 2: launch_v2<8, 2, 4, 1, 4>(g, p, stream); break;
        case 3: launch_v2<8, 4, 2, 1, 4>(g, p, stream); break;
        default: launch_v2<16, 4, 2, 2, 3>(g, p, stream); break;
    }
    return 1;
}

// instantiates every variant's attribute once, outside stream capture
void init_gemm_v2_attributes() {
    cudaFuncSetAttribute(gemm_v2_kernel<8, 1, 8, 1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(float) * 3 * (8 + 256) * LDS2));
    cudaFuncSetAttribute(gemm_v2_kernel<8, 2, 4, 1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(float) * 4 * (16 + 128) * LDS2));
    cudaFuncSetAttribute(gemm_v2_kernel<8, 4, 2, 1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(float) * 4 * (32 + 64) * LDS2));
    cudaFuncSetAttribute(gemm_v2_kernel<16, 4, 2, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(float) * 3 * (64 + 128) * LDS2));
}

}  // namespace rvc
