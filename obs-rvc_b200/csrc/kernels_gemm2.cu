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
#include "pdl.cuh"

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

__device__ long long g_v2_dbg[16];
#define V2_DBG(i) do { if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) g_v2_dbg[i] = clock64(); } while (0)

// BM x BN output tile, 8 warps.  Inside a k-tile (BK = 32 = 8 k-quads) warp w owns k-quad w and
// accumulates the WHOLE tile for it (lane <-> column, BM rows x BN/32 columns = 64 accumulators):
// per k-tile a warp issues BM broadcast LDS.128 + BN/32 conflict-free LDS.128 for 4*64 FMAs, i.e. the
// tile runs at the FMA rate instead of the shared-memory wavefront rate even at one CTA per SM.
// The eight per-warp partial tiles are summed through shared memory in fixed warp order.
constexpr int V2_THREADS = 512;  // 16 warps: 8 k-quads x 2 row halves (latency hiding at one CTA per SM)

template <int BM, int BN, int STAGES>
__global__ void __launch_bounds__(V2_THREADS, 1) gemm_v2_kernel(GemmParams p) {
    constexpr int TN = BN / 32;
    constexpr int HM = BM / 2;  // rows per warp
    static_assert(HM * TN <= 32 && HM >= 1 && TN >= 1, "at most 32 accumulators per thread");
    constexpr int STAGE_FLOATS = (BM + BN) * LDS2;
    constexpr int CHUNKS = (BM + BN) * (BK2 / 4);          // 16-byte chunks per stage
    constexpr int CPT = (CHUNKS + V2_THREADS - 1) / V2_THREADS;  // chunks per thread
    constexpr int RS = V2_THREADS / BN;                    // row stride of the final thread->output map
    constexpr int OPT = BM * BN / V2_THREADS;              // outputs per thread (4)
    extern __shared__ __align__(16) float smem[];
    __shared__ int s_last;

    pdl_launch_dependents();
    V2_DBG(0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int z = blockIdx.z % p.splitk, bz = blockIdx.z / p.splitk;
    const GemmBases gb = gemm_bases(p, bz);
    const float* __restrict__ A = gb.A;
    const float* __restrict__ W = gb.W;
    const int nkt_total = (p.K + BK2 - 1) / BK2;
    const int kt0 = z * p.kt_per_split, kt1 = min(nkt_total, kt0 + p.kt_per_split);
    const int nkt = max(0, kt1 - kt0);
    const bool contiguous = p.seg_len >= p.K;

    // per-thread chunk descriptors (row base pointers are k-invariant)
    const float* cbase[CPT]; int ckc[CPT]; int cdst[CPT]; bool cisA[CPT], cvalid[CPT];
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
        const int c = tid + i * V2_THREADS;
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
                if (cisA[i] && !contiguous) {
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

    pdl_wait();  // everything above is index math on kernel parameters only
    V2_DBG(1);
    const int kq = warp & 7, rh = warp >> 3;
    float acc[HM][TN];
#pragma unroll
    for (int i = 0; i < HM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nkt) load_stage(s, kt0 + s);
        cp_async_commit();
    }
    V2_DBG(2);
    for (int kt = 0; kt < nkt; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (kt == 0) V2_DBG(3);
        if (kt == 1) V2_DBG(4);
        if (kt + STAGES - 1 < nkt) load_stage((kt + STAGES - 1) % STAGES, kt0 + kt + STAGES - 1);
        cp_async_commit();
        const float* As = smem + (kt % STAGES) * STAGE_FLOATS + (rh * HM) * LDS2 + kq * 4;
        const float* Ws = smem + (kt % STAGES) * STAGE_FLOATS + (BM + lane) * LDS2 + kq * 4;
        float4 b[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = *reinterpret_cast<const float4*>(Ws + j * 32 * LDS2);
        // rows in groups of RG: the RG broadcast loads are issued together and the FMAs are ordered
        // so that consecutive instructions hit different accumulators (FMA latency 4 is covered)
        constexpr int RG = HM >= 8 ? 8 : HM;
#pragma unroll
        for (int i0 = 0; i0 < HM; i0 += RG) {
            float4 a[RG];
#pragma unroll
            for (int r = 0; r < RG; ++r) a[r] = *reinterpret_cast<const float4*>(As + (i0 + r) * LDS2);
#pragma unroll
            for (int j = 0; j < TN; ++j) {
#pragma unroll
                for (int r = 0; r < RG; ++r) acc[i0 + r][j] = fmaf(a[r].x, b[j].x, acc[i0 + r][j]);
#pragma unroll
                for (int r = 0; r < RG; ++r) acc[i0 + r][j] = fmaf(a[r].y, b[j].y, acc[i0 + r][j]);
#pragma unroll
                for (int r = 0; r < RG; ++r) acc[i0 + r][j] = fmaf(a[r].z, b[j].z, acc[i0 + r][j]);
#pragma unroll
                for (int r = 0; r < RG; ++r) acc[i0 + r][j] = fmaf(a[r].w, b[j].w, acc[i0 + r][j]);
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();  // every warp is done with the stage buffers: reuse them for the partial tiles
    V2_DBG(5);
    float* red = smem;  // [8 k-quads][BM][BN]
#pragma unroll
    for (int i = 0; i < HM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) red[(kq * BM + rh * HM + i) * BN + j * 32 + lane] = acc[i][j];
    __syncthreads();
    // final map: thread t owns column t % BN and OPT rows (stride RS): coalesced along N
    const int col = tid % BN, r0 = tid / BN;
    float v[OPT];
#pragma unroll
    for (int i = 0; i < OPT; ++i) {
        const int row = r0 + i * RS;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[(w * BM + row) * BN + col];
        v[i] = s;
    }

    const float* __restrict__ bias = gb.bias;
    float* C = gb.C;
    float* C2 = gb.C2;
    const float* R = gb.R;
    const int n = n0 + col;

    if (p.splitk > 1) {
        float* part = gb.scratch + ((long long)(gb.grp * p.splitk + z) * p.M) * p.N;
#pragma unroll
        for (int i = 0; i < OPT; ++i) {
            const int m = m0 + r0 + i * RS;
            if (m < p.M && n < p.N) __stcg(part + (long long)m * p.N + n, v[i]);
        }
        __threadfence();
        __syncthreads();
        const int tile = (gb.grp * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        if (tid == 0) {
            const unsigned ticket = atomicAdd(gb.counters + tile, 1u);
            s_last = (ticket == unsigned(p.splitk - 1)) ? 1 : 0;
            if (s_last) gb.counters[tile] = 0;  // self-cleaning for the next launch / graph replay
        }
        __syncthreads();
        V2_DBG(6);
        if (!s_last) return;
        __threadfence();
        const float* base = gb.scratch + ((long long)(gb.grp * p.splitk) * p.M) * p.N;
#pragma unroll
        for (int i = 0; i < OPT; ++i) v[i] = 0.f;
        // partials are summed in z order; 8 splits x 4 rows = 32 independent L2 loads in flight per thread
        for (int zz0 = 0; zz0 < p.splitk; zz0 += 8) {
            float t[8][OPT];
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int i = 0; i < OPT; ++i) {
                    const int m = m0 + r0 + i * RS;
                    t[u][i] = (zz0 + u < p.splitk && m < p.M && n < p.N)
                                  ? __ldcg(base + ((long long)(zz0 + u) * p.M + m) * p.N + n) : 0.f;
                }
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int i = 0; i < OPT; ++i) v[i] += t[u][i];
        }
    }
    // epilogue: residual values are fetched up front (R may alias C for in-place accumulation, which
    // would otherwise serialise every load behind the previous store)
    const bool plain = p.out_mode == OUT_PLAIN;
    const bool gate = p.act == ACT_GATE;
    float rv[OPT];
#pragma unroll
    for (int i = 0; i < OPT; ++i) {
        const int m = m0 + r0 + i * RS;
        const int rc = gate ? (n >> 1) : n;
        rv[i] = (R && plain && m < p.M && n < p.N && !(gate && (n & 1))) ? R[(long long)m * p.ldr + rc] : 0.f;
    }
    const float bn_ = (bias && n < p.N) ? __ldg(bias + n) : 0.f;
#pragma unroll
    for (int i = 0; i < OPT; ++i) {
        const int m = m0 + r0 + i * RS;
        const float partner = __shfl_xor_sync(0xffffffffu, v[i], 1);
        if (m >= p.M || n >= p.N) continue;
        if (plain && !gate) {
            float o = apply_act(p.act, fmaf(p.alpha, v[i], bn_)) + rv[i];
            const bool masked = p.mask_period > 0 && (m % p.mask_period) >= p.mask_valid;
            if (masked) o = 0.f;
            C[(long long)m * p.ldc + n] = o;
            if (C2) C2[(long long)m * p.ldc2 + n] = masked ? 0.f : apply_act(p.act2, o);
        } else if (plain && gate) {
            if (n & 1) continue;
            const float v1 = fmaf(p.alpha, partner, bias ? __ldg(bias + n + 1) : 0.f);
            float o = tanhf(fmaf(p.alpha, v[i], bn_)) * sigmoid_f(v1) + rv[i];
            const bool masked = p.mask_period > 0 && (m % p.mask_period) >= p.mask_valid;
            if (masked) o = 0.f;
            C[(long long)m * p.ldc + (n >> 1)] = o;
            if (C2) C2[(long long)m * p.ldc2 + (n >> 1)] = masked ? 0.f : apply_act(p.act2, o);
        } else {
            epilogue_elem(p, bias, C, C2, nullptr, m, n, v[i], partner);
        }
    }
    V2_DBG(7);
}

template <int BM, int BN, int STAGES>
void launch_v2(const GemmOp& g, GemmParams& p, int nb, cudaStream_t s) {
    constexpr size_t stage_bytes = sizeof(float) * size_t(STAGES) * (BM + BN) * LDS2;
    constexpr size_t red_bytes = sizeof(float) * 8 * BM * BN;
    constexpr size_t smem = stage_bytes > red_bytes ? stage_bytes : red_bytes;
    auto kern = gemm_v2_kernel<BM, BN, STAGES>;
    static unsigned long long attr = 0;
    if (first_time_on_device(attr)) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    const int nkt = (g.K + BK2 - 1) / BK2;
    p.kt_per_split = (nkt + g.splitk - 1) / g.splitk;
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, nb * g.batch * g.splitk);
    launch_k(kern, grid, dim3(V2_THREADS), smem, s, p);
}

}  // namespace

int launch_gemm_v2(const GemmOp& g, const DeviceBases& B, cudaStream_t stream) {
    GemmParams p = gemmk::make_params(g, B);
    switch (g.sched_variant) {
        // (deeper rings were measured slower: profiles/README.md)
        case 1: launch_v2<8, 256, 3>(g, p, B.nb, stream); break;
        case 2: launch_v2<16, 128, 4>(g, p, B.nb, stream); break;
        case 4:
            // a short split-K slice (<= 5 k-tiles): the six-stage ring has every k-tile in flight at once - one HBM round trip
            if (g.splitk > 1 && ((g.K + BK2 - 1) / BK2 + g.splitk - 1) / g.splitk <= 5) launch_v2<32, 32, 6>(g, p, B.nb, stream);
            else launch_v2<32, 32, 4>(g, p, B.nb, stream);
            break;
        default: launch_v2<32, 64, 4>(g, p, B.nb, stream); break;
    }
    return 1;
}

void v2_debug_read(long long* out) { cudaMemcpyFromSymbol(out, g_v2_dbg, sizeof(long long) * 16); }

// instantiates every variant's attribute once, outside stream capture
void init_gemm_v2_attributes() {
    cudaFuncSetAttribute(gemm_v2_kernel<8, 256, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(float) * 3 * (8 + 256) * LDS2));
    cudaFuncSetAttribute(gemm_v2_kernel<16, 128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(float) * 4 * (16 + 128) * LDS2));
    cudaFuncSetAttribute(gemm_v2_kernel<32, 64, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(gemm_v2_kernel<32, 32, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(gemm_v2_kernel<32, 32, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
}

}  // namespace rvc
