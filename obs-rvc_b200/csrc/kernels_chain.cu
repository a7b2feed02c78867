// kernels_chain.cu - the persistent chain kernel (chain.h): op table interpreter + grid barrier.
//
// One cooperative launch of G CTAs x 256 threads walks a table of phases.  A phase is a set of
// mutually independent ops (plan.cpp decides that at buffer granularity); its work items (GEMM
// tiles x split-K slices, pooling / normalisation row groups) are dealt round-robin to the CTAs;
// a grid barrier (one L2 atomic + acquire spin per CTA) separates phases.  Activations written in
// one phase are read in the next through L2 only (cp.async.cg / ld.global.cg): the L1s are not
// coherent across SMs and there is no kernel boundary to invalidate them.
//
// GEMM tile (exact fp32, same contraction as ops.h GemmOp): BM x BN = 1024 outputs, BK = 32,
// cp.async ring; warp w owns k-quad w of every k-tile and accumulates the whole tile for it
// (32 accumulators per thread: lanes along N, rows in registers), the 8 per-warp partial tiles
// are summed through shared memory in fixed order; split-K slices go to a scratch area and the
// last CTA to arrive on a tile (atomic ticket) adds them in z order - deterministic sums.
#include <cuda.h>

#include <cstdio>
#include <cstring>

#include "chain.h"
#include "launch.h"
#include "pdl.cuh"

namespace rvc {

namespace {

using namespace gemmk;

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, int src_bytes) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

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

// activation out of line: the transcendental bodies (erff, tanhf, expf) would otherwise be replicated in every
// tile variant and blow the kernel past the instruction cache
__device__ __noinline__ float chain_act(int act, float v) { return apply_act(act, v); }
__device__ __forceinline__ float chain_act_fast(int act, float v) {
    if (act == ACT_NONE) return v;
    if (act == ACT_RELU) return fmaxf(v, 0.0f);
    return chain_act(act, v);
}

// final accumulator -> output element; identical arithmetic to gemmk::epilogue_elem, residual read through L2
__device__ __noinline__ void chain_epilogue(const GemmParams& p, const float* __restrict__ bias, float* C, float* C2, const float* R,
                                               int m, int n, float acc, float acc_partner, float r_pre = 0.f) {
    const bool masked = p.mask_period > 0 && (m % p.mask_period) >= p.mask_valid;
    if (p.act == ACT_GATE) {
        if (n & 1) return;
        const float v0 = fmaf(p.alpha, acc, bias ? __ldg(bias + n) : 0.f);
        const float v1 = fmaf(p.alpha, acc_partner, bias ? __ldg(bias + n + 1) : 0.f);
        float g = tanhf(v0) * sigmoid_f(v1);
        const int col = n >> 1;
        if (R) g += __ldcg(R + (long long)m * p.ldr + col);
        if (masked) g = 0.f;
        C[(long long)m * p.ldc + col] = g;
        if (C2) C2[(long long)m * p.ldc2 + col] = masked ? 0.f : apply_act(p.act2, g);
        return;
    }
    float v = apply_act(p.act, fmaf(p.alpha, acc, bias ? __ldg(bias + n) : 0.f));
    if (p.out_mode == OUT_PLAIN) {
        v += r_pre;   // residual fetched ahead of the contraction (0 when R is read here)
        if (R) v += __ldcg(R + (long long)m * p.ldr + n);
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

__device__ long long g_chain_stamp2[256 * 2];  // [phase]: spin start, spin end (CTA 0)
__device__ long long g_chain_stamp[256 * 8];   // [phase][event] clock64 of CTA 0 thread 0, first item of the phase
#define CH_STAMP(e) do { if (blockIdx.x == 0 && threadIdx.x == 0 && stamp_row >= 0) g_chain_stamp[stamp_row * 8 + (e)] = clock64(); } while (0)

// Per-shape constants of one tile variant.  Lanes = LC columns x LK k-quads (LC * LK = 32), every thread
// keeps all BM rows of its CT columns in registers (BM * CT <= 32 accumulators): the A fragment of a k-quad is
// one 16-byte broadcast per row, the W fragment one conflict-free LDS.128 per column (rows padded by 4 words:
// row r starts at bank 4r).  A k-tile holds BK = 32 * LK * QPW floats = 8 * LK * QPW k-quads, warp w / lane
// group lk owning quads (w * LK + lk) + j * 8 * LK.  Narrow tiles (BN = 8, 16) therefore keep the full
// register reuse of the 32-wide tile: weight-streaming ops (M <= 32) get one tile per CTA without split-K.
template <int BM, int BN, int LK, int QPW, int STAGES>
struct TileCfg {
    static constexpr int LK_ = LK, QPW_ = QPW;
    static constexpr int BK = 32 * LK * QPW;
    static constexpr int LDS = BK + 4;
    static constexpr int LC = 32 / LK;             // lanes along N
    static constexpr int CT = BN / LC;             // columns per thread (stride LC)
    static constexpr int NP = 8 * LK;              // partial tiles to sum at the end
    static constexpr int STAGE_FLOATS = (BM + BN) * LDS;
    static constexpr int CHUNKS_A = BM * (BK / 4), CHUNKS_W = BN * (BK / 4);
    static constexpr int CPT_A = (CHUNKS_A + CHAIN_THREADS - 1) / CHAIN_THREADS;
    static constexpr int CPT_W = (CHUNKS_W + CHAIN_THREADS - 1) / CHAIN_THREADS;
    static constexpr int OUTS = BM * BN;
    static constexpr int OPT = (OUTS + CHAIN_THREADS - 1) / CHAIN_THREADS;
    static_assert(CT >= 1 && CT * LC == BN && BM * CT <= 32, "tile shape");
    static_assert(STAGES * STAGE_FLOATS * 4 <= CHAIN_SMEM_BYTES, "stage ring exceeds the chain's shared memory");
    static_assert(NP * OUTS * 4 <= CHAIN_SMEM_BYTES, "reduction buffer exceeds the chain's shared memory");
};

// W half of the stage loads of k-tiles [kt_first, kt_end): weights do not depend on earlier phases, so the
// caller may issue these BEFORE the grid barrier (no commit here: they join the first A group)
template <class T, int BM>
__device__ __forceinline__ void gemm_issue_w(const GemmParams& p, int n0, int bz, int kt_first, int kt_end, int slot0, float* smem) {
    const float* __restrict__ W = p.W + bz * p.sW;
#pragma unroll 1
    for (int i = 0; i < T::CPT_W; ++i) {
        const int c = threadIdx.x + i * CHAIN_THREADS;
        if (c >= T::CHUNKS_W) break;
        const int row = c / (T::BK / 4), kc = (c % (T::BK / 4)) * 4;
        const int n = n0 + row;
        const float* base = n < p.N ? W + (long long)n * p.ldw : nullptr;
#pragma unroll 1
        for (int kt = kt_first, slot = slot0; kt < kt_end; ++kt, ++slot) {
            const int kk = kt * T::BK + kc;
            const bool ok = base && kk < p.K;
            cp_async16(smem + slot * T::STAGE_FLOATS + (BM + row) * T::LDS + kc, ok ? base + kk : p.W, ok ? 16 : 0);
        }
    }
}

// L2 prefetch of this tile's own weight rows beyond the stages already requested
template <class T>
__device__ __forceinline__ void gemm_prefetch_w(const GemmParams& p, int n0, int bz, int k_from, int k_to) {
    if (p.bias && threadIdx.x == CHAIN_THREADS - 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.bias + bz * p.sBias + n0));
    if (k_to <= k_from) return;
    const char* W = reinterpret_cast<const char*>(p.W + bz * p.sW);
    const int lpr = ((k_to - k_from) * 4 + 127) >> 7;
    const int rows = min(T::CT * T::LC, p.N - n0);
    for (int l = threadIdx.x; l < rows * lpr; l += CHAIN_THREADS) {
        const int r = l / lpr, c = l - r * lpr;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(W + ((long long)(n0 + r) * p.ldw + k_from) * 4 + c * 128));
    }
}

template <class T, int BM, int BN, int STAGES>
__device__ __forceinline__ void gemm_tile(const GemmParams& p, int m0, int n0, int z, int bz, int tile_id, float* smem, int* s_last,
                                          bool w_preloaded, int stamp_row) {
    CH_STAMP(1);
    constexpr int LC = T::LC, LK = T::LK_, CT = T::CT, LDS = T::LDS, BK = T::BK, OUTS = T::OUTS, OPT = T::OPT;
    constexpr int CPT_A = T::CPT_A, CPT_W = T::CPT_W;
    const int tid = threadIdx.x, lane = tid & 31, wq = tid >> 5;
    const int lc = lane % LC, lk = lane / LC;
    const int K = p.K, seg_len = p.seg_len >= p.K ? (1 << 30) : p.seg_len;
    const long long seg_wrap = p.seg_stride - p.seg_len;   // pointer bump when a chunk walks into the next segment
    const int nkt_total = (K + BK - 1) / BK;
    const int kt0 = z * p.kt_per_split, kt1 = min(nkt_total, kt0 + p.kt_per_split);
    const int nkt = max(0, kt1 - kt0);

    // Per-thread chunk cursors (16-byte pieces of the stage): source pointer + position inside the A segment,
    // advanced by BK after every issue - no divisions, no descriptor reads inside the k loop.
    const float* aptr[CPT_A]; int awithin[CPT_A]; int adst[CPT_A];
    const float* wptr[CPT_W]; int wdst[CPT_W];
    int a_k, w_k;   // k index of chunk column 0 of the next k-tile to issue (per operand)
    {
        const float* __restrict__ A = p.A + bz * p.sA;
        const float* __restrict__ W = p.W + bz * p.sW;
        const int kbase = kt0 * BK;
#pragma unroll
        for (int i = 0; i < CPT_A; ++i) {
            const int c = tid + i * CHAIN_THREADS;
            const int row = c / (BK / 4), kc = (c % (BK / 4)) * 4;
            const int m = m0 + row, kk = kbase + kc;
            adst[i] = (c < T::CHUNKS_A) ? row * LDS + kc : -1;
            const int seg = kk / seg_len;
            awithin[i] = kk - seg * seg_len;
            aptr[i] = (m < p.M) ? A + (long long)m * p.lda + (long long)seg * p.seg_stride + awithin[i] : nullptr;
        }
        const int wskip = w_preloaded ? min(nkt, STAGES - 1) : 0;   // W stages already requested before the barrier
#pragma unroll
        for (int i = 0; i < CPT_W; ++i) {
            const int c = tid + i * CHAIN_THREADS;
            const int row = c / (BK / 4), kc = (c % (BK / 4)) * 4;
            const int n = n0 + row;
            wdst[i] = (c < T::CHUNKS_W) ? (BM + row) * LDS + kc : -1;
            wptr[i] = (n < p.N) ? W + (long long)n * p.ldw + kbase + wskip * BK + kc : nullptr;
        }
        a_k = kbase; w_k = kbase + wskip * BK;
    }
    auto issue_a = [&](int slot) {
        float* st = smem + slot * T::STAGE_FLOATS;
#pragma unroll
        for (int i = 0; i < CPT_A; ++i) {
            if (adst[i] >= 0) {
                const int kc = adst[i] % LDS;
                const bool ok = aptr[i] && (a_k + kc) < K;
                cp_async16(st + adst[i], ok ? aptr[i] : p.W, ok ? 16 : 0);
                if (aptr[i]) {
                    aptr[i] += BK; awithin[i] += BK;
                    while (awithin[i] >= seg_len) { awithin[i] -= seg_len; aptr[i] += seg_wrap; }
                }
            }
        }
        a_k += BK;
    };
    auto issue_w = [&](int slot) {
        float* st = smem + slot * T::STAGE_FLOATS;
#pragma unroll
        for (int i = 0; i < CPT_W; ++i) {
            if (wdst[i] >= 0) {
                const int kc = (wdst[i] - BM * LDS) % LDS;
                const bool ok = wptr[i] && (w_k + kc) < K;
                cp_async16(st + wdst[i], ok ? wptr[i] : p.W, ok ? 16 : 0);
                if (wptr[i]) wptr[i] += BK;
            }
        }
        w_k += BK;
    };

    float acc[BM][CT];
#pragma unroll
    for (int i = 0; i < BM; ++i)
#pragma unroll
        for (int j = 0; j < CT; ++j) acc[i][j] = 0.f;
    // residual operand of the epilogue: requested now so its L2 round trip overlaps the contraction
    const float* Rb = p.R ? p.R + bz * p.sR : nullptr;
    const bool r_plain = Rb && p.out_mode == OUT_PLAIN && p.act != ACT_GATE && (p.splitk == 1);
    float rpre[OPT], bpre[OPT];   // bias too: parameter vectors are cold in L2 every window (HBM round trip)
    const float* __restrict__ bias = p.bias ? p.bias + bz * p.sBias : nullptr;
#pragma unroll
    for (int i = 0; i < OPT; ++i) {
        const int idx = tid + i * CHAIN_THREADS;
        const int m = m0 + idx / BN, n = n0 + idx % BN;
        rpre[i] = (r_plain && idx < OUTS && m < p.M && n < p.N) ? __ldcg(Rb + (long long)m * p.ldr + n) : 0.f;
        bpre[i] = (bias && idx < OUTS && n < p.N) ? __ldg(bias + n) : 0.f;
    }

#pragma unroll 1
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nkt) { if (!w_preloaded) issue_w(s); issue_a(s); }
        cp_async_commit();
    }
    CH_STAMP(2);
#pragma unroll 1
    for (int kt = 0; kt < nkt; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (kt == 0) CH_STAMP(3);
        if (kt == 1) CH_STAMP(4);
        if (kt + STAGES - 1 < nkt) {
            const int slot = (kt + STAGES - 1) % STAGES;
            issue_w(slot); issue_a(slot);
        }
        cp_async_commit();
        const float* st = smem + (kt % STAGES) * T::STAGE_FLOATS;
#pragma unroll
        for (int qq = 0; qq < T::QPW_; ++qq) {
            const int kq = wq * LK + lk + qq * 8 * LK;
            const float* As = st + kq * 4;
            const float* Ws = st + (BM + lc) * LDS + kq * 4;
            float4 b[CT];
#pragma unroll
            for (int j = 0; j < CT; ++j) b[j] = *reinterpret_cast<const float4*>(Ws + j * LC * LDS);
            constexpr int RG = BM >= 8 ? 8 : BM;
#pragma unroll
            for (int i0 = 0; i0 < BM; i0 += RG) {
                float4 a[RG];
#pragma unroll
                for (int r = 0; r < RG; ++r) a[r] = *reinterpret_cast<const float4*>(As + (i0 + r) * LDS);
#pragma unroll
                for (int j = 0; j < CT; ++j) {
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
    }
    cp_async_wait<0>();
    __syncthreads();  // stage buffers are free: reuse them for the per-(warp, lane group) partial tiles
    CH_STAMP(5);
    float* red = smem;  // [8 * LK][BM][BN]
    {
        const int pw = wq * LK + lk;
#pragma unroll
        for (int i = 0; i < BM; ++i)
#pragma unroll
            for (int j = 0; j < CT; ++j) red[(pw * BM + i) * BN + j * LC + lc] = acc[i][j];
    }
    __syncthreads();
    float* C = p.C + bz * p.sC;
    float* C2 = p.C2 ? p.C2 + bz * p.sC : nullptr;
    const float* R = r_plain ? nullptr : Rb;
    float v[OPT];
#pragma unroll
    for (int i = 0; i < OPT; ++i) {
        const int idx = tid + i * CHAIN_THREADS;
        float s = 0.f;
        if (idx < OUTS) {
#pragma unroll
            for (int w = 0; w < T::NP; ++w) s += red[w * OUTS + idx];
        }
        v[i] = s;
    }
    if (p.splitk > 1) {
        float* part = p.scratch + ((long long)(bz * p.splitk + z) * p.M) * p.N;
#pragma unroll
        for (int i = 0; i < OPT; ++i) {
            const int idx = tid + i * CHAIN_THREADS;
            const int m = m0 + idx / BN, n = n0 + idx % BN;
            if (idx < OUTS && m < p.M && n < p.N) __stcg(part + (long long)m * p.N + n, v[i]);
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const unsigned ticket = atomicAdd(p.counters + tile_id, 1u);
            *s_last = (ticket == unsigned(p.splitk - 1)) ? 1 : 0;
            if (*s_last) p.counters[tile_id] = 0;  // self-cleaning for the next replay
        }
        __syncthreads();
        const int last = *s_last;
        if (!last) return;   // (uniform per CTA) the caller syncs before the next item touches smem
        __threadfence();
        const float* base = p.scratch + ((long long)(bz * p.splitk) * p.M) * p.N;
#pragma unroll
        for (int i = 0; i < OPT; ++i) v[i] = 0.f;
#pragma unroll 1
        for (int zz = 0; zz < p.splitk; ++zz) {
#pragma unroll
            for (int i = 0; i < OPT; ++i) {
                const int idx = tid + i * CHAIN_THREADS;
                const int m = m0 + idx / BN, n = n0 + idx % BN;
                if (idx < OUTS && m < p.M && n < p.N) v[i] += __ldcg(base + ((long long)zz * p.M + m) * p.N + n);
            }
        }
    }
    if (p.out_mode == OUT_PLAIN && p.act != ACT_GATE) {
        // common case inline, parameters in registers (same arithmetic as chain_epilogue)
        const int act = p.act, act2 = p.act2, mask_period = p.mask_period, mask_valid = p.mask_valid;
        const float alpha = p.alpha;
        const long long ldc = p.ldc, ldc2 = p.ldc2, ldr = p.ldr;
#pragma unroll
        for (int i = 0; i < OPT; ++i) {
            const int idx = tid + i * CHAIN_THREADS;
            const int m = m0 + idx / BN, n = n0 + idx % BN;
            if (idx >= OUTS || m >= p.M || n >= p.N) continue;
            float o = chain_act_fast(act, fmaf(alpha, v[i], bpre[i])) + rpre[i];
            if (R) o += __ldcg(R + (long long)m * ldr + n);
            const bool masked = mask_period > 0 && (m % mask_period) >= mask_valid;
            if (masked) o = 0.f;
            C[(long long)m * ldc + n] = o;
            if (C2) C2[(long long)m * ldc2 + n] = masked ? 0.f : chain_act_fast(act2, o);
        }
    } else {
#pragma unroll 1
        for (int i = 0; i < OPT; ++i) {
            const int idx = tid + i * CHAIN_THREADS;
            const int m = m0 + idx / BN, n = n0 + idx % BN;
            const float partner = __shfl_xor_sync(0xffffffffu, v[i], 1);  // BN is even: the gate partner column is a lane neighbour
            if (idx >= OUTS || m >= p.M || n >= p.N) continue;
            chain_epilogue(p, bias, C, C2, R, m, n, v[i], partner, rpre[i]);
        }
    }
    CH_STAMP(6);
}

// dispatch over the tile variants (chain.h ChainTile); `what` 0 = run the tile, 1 = prefetch the tile's weight rows
// and bias into L2 (issued between the barrier arrival and the wait: weights do not depend on the previous phase.
// Requesting the first cp.async stages there as well was measured slower: the extra instructions delay the poll.)
__device__ __forceinline__ void gemm_dispatch(const ChainOpDev& o, int local, float* smem, int* s_last, bool w_preloaded, int what,
                                              int stamp_row = -1) {
    const GemmParams& p = o.g;
    const int z = local % o.splitk;
    const int tile = local / o.splitk;
    const int tn = tile % o.tiles_n;
    const int tm = (tile / o.tiles_n) % o.tiles_m;
    const int bz = tile / (o.tiles_n * o.tiles_m);
#define RVC_TILE_CASE(ID, BM, BN, LK, QPW, ST)                                                                               \
    case ID: {                                                                                                               \
        using T = TileCfg<BM, BN, LK, QPW, ST>;                                                                              \
        if (what == 0) gemm_tile<T, BM, BN, ST>(p, tm * BM, tn * BN, z, bz, tile, smem, s_last, w_preloaded, stamp_row);     \
        else if (what == 2) {                                                                                                \
            const int nkt_total = (p.K + T::BK - 1) / T::BK, kt0 = z * p.kt_per_split, kt1 = min(nkt_total, kt0 + p.kt_per_split); \
            gemm_issue_w<T, BM>(p, tn * BN, bz, kt0, min(kt1, kt0 + ST - 1), 0, smem);                                       \
        } else {                                                                                                             \
            const int nkt_total = (p.K + T::BK - 1) / T::BK, kt0 = z * p.kt_per_split, kt1 = min(nkt_total, kt0 + p.kt_per_split); \
            gemm_prefetch_w<T>(p, tn * BN, bz, kt0 * T::BK, min(p.K, kt1 * T::BK));                                          \
        }                                                                                                                    \
        break;                                                                                                               \
    }
    switch (o.variant) {
        // (BM, BN, LK, QPW, stages): every ring keeps 75-90 KB of loads in flight per CTA - at ~0.5-0.8 us of L2
        // latency that is what one SM needs to pull ~100 GB/s
        RVC_TILE_CASE(CT_32x32, 32, 32, 1, 2, 5)
        RVC_TILE_CASE(CT_16x64, 16, 64, 1, 2, 4)
        RVC_TILE_CASE(CT_8x128, 8, 128, 1, 1, 4)
        RVC_TILE_CASE(CT_32x16, 32, 16, 2, 1, 6)
        RVC_TILE_CASE(CT_32x8, 32, 8, 4, 1, 4)
        RVC_TILE_CASE(CT_8x32, 8, 32, 1, 4, 4)
        RVC_TILE_CASE(CT_8x16, 8, 16, 2, 2, 6)
        RVC_TILE_CASE(CT_16x16, 16, 16, 2, 2, 5)
        default: break;
    }
#undef RVC_TILE_CASE
}

// tiny / unaligned contractions (K < 64 or rows that are not 16-byte aligned): one output per thread-slot
__device__ __forceinline__ void gemm_direct_item(const GemmParams& p, int local) {
    const long long total = (long long)p.M * p.N;
    const bool contiguous = p.seg_len >= p.K;
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
        const long long o = (long long)local * 1024 + i * CHAIN_THREADS + threadIdx.x;
        const bool ok = o < total;
        const int m = ok ? int(o / p.N) : 0, n = ok ? int(o - (long long)m * p.N) : 0;
        const float* a = p.A + (long long)m * p.lda;
        const float* w = p.W + (long long)n * p.ldw;
        float acc = 0.f;
        if (ok) {
            if (contiguous) {
                for (int k = 0; k < p.K; ++k) acc = fmaf(__ldcg(a + k), __ldg(w + k), acc);
            } else {
                for (int k = 0; k < p.K; ++k) {
                    const int seg = k / p.seg_len, within = k - seg * p.seg_len;
                    acc = fmaf(__ldcg(a + (long long)seg * p.seg_stride + within), __ldg(w + k), acc);
                }
            }
        }
        const float partner = __shfl_xor_sync(0xffffffffu, acc, 1);
        if (ok) chain_epilogue(p, p.bias, p.C, p.C2, p.R, m, n, acc, partner);
    }
}

__device__ __forceinline__ void avgpool_item(const ChainOpDev& o, int local) {
    const int T = o.i0, F = o.i1, C = o.i2, To = T / 2, Fo = F / 2;
    const long long n = (long long)To * Fo * C, ldin = o.ld0;
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
        const long long e = (long long)local * 1024 + i * CHAIN_THREADS + threadIdx.x;
        if (e >= n) continue;
        const int c = int(e % C);
        const long long r = e / C;
        const int f = int(r % Fo), t = int(r / Fo);
        const float* q = o.x0 + ((long long)(2 * t + 1) * (F + 2) + 2 * f + 1) * ldin + c;
        const float v = (__ldcg(q) + __ldcg(q + ldin) + __ldcg(q + (long long)(F + 2) * ldin) + __ldcg(q + (long long)(F + 3) * ldin)) * 0.25f;
        o.y0[((long long)(t + 1) * (Fo + 2) + f + 1) * C + c] = v;
    }
}

// one warp per row, 8 rows per item (cols <= 1024)
__device__ __forceinline__ void layernorm_item(const ChainOpDev& o, int local) {
    const int row = local * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int rows = o.i0, cols = o.i1;
    if (row >= rows) return;
    const float* x = o.x0 + (long long)row * o.ld0;
    float v[32];
    float s = 0.f;
    const int nv = (cols + 31) >> 5;   // columns per lane actually present (<= 32)
    if (nv <= 8) {
        // narrow rows (the synthesizer's 192 channels): everything in flight at once, gamma / beta included -
        // parameter vectors are cold every window
        float gg[8], bb[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = lane + i * 32;
            v[i] = c < cols ? __ldcg(x + c) : 0.f;
            gg[i] = c < cols ? __ldg(o.x1 + c) : 0.f;
            bb[i] = c < cols ? __ldg(o.x2 + c) : 0.f;
            s += v[i];
        }
        const float mean = warp_sum(s) / float(cols);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = lane + i * 32;
            const float d = c < cols ? v[i] - mean : 0.f;
            q += d * d;
        }
        const float rstd = rsqrtf(warp_sum(q) / float(cols) + o.f0);
        float* y = o.y0 + (long long)row * o.ld1;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = lane + i * 32;
            if (c < cols) y[c] = (v[i] - mean) * rstd * gg[i] + bb[i];
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int c = lane + i * 32;
        v[i] = c < cols ? __ldcg(x + c) : 0.f;
        s += v[i];
    }
    const float mean = warp_sum(s) / float(cols);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int c = lane + i * 32;
        const float d = c < cols ? v[i] - mean : 0.f;
        q += d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) / float(cols) + o.f0);
    float* y = o.y0 + (long long)row * o.ld1;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int c = lane + i * 32;
        if (c < cols) y[c] = (v[i] - mean) * rstd * __ldg(o.x1 + c) + __ldg(o.x2 + c);
    }
}

// VITS windowed relative-position attention: one item = one (head, query row).  K, V and the two relative
// tables of the head are staged in shared memory with every load in flight at once (a dependent chain of L2 /
// cold-HBM round trips otherwise); the reductions keep the order of relattn_kernel (kernels_misc.cu), so the
// chain and the stand-alone kernel agree bit for bit.
// CTA-wide sync of the chain kernel (256 threads) / named barrier of the slab kernel's 256 compute threads (its
// producer warp does not take part)
template <bool NAMED> __device__ __forceinline__ void chain_sync() {
    if (NAMED) asm volatile("bar.sync 2, 256;" ::: "memory"); else __syncthreads();
}
template <bool NAMED>
__device__ __forceinline__ void relattn_item_t(const ChainOpDev& o, int local, float* sm) {
    const int T = o.i0, heads = o.i1, dim = o.i2, window = o.i3;
    const int h = local / T, i = local - h * T;
    const int HD = heads * dim, nrel = 2 * window + 1;
    const long long ld = o.ld0;
    const float* qkv = o.x0;
    float* qs = sm;                  // [dim]
    float* ps = qs + dim;            // [T]
    float* ks = ps + ((T + 3) & ~3); // [T][dim]
    float* vs = ks + T * dim;        // [T][dim]
    float* rks = vs + T * dim;       // [nrel][dim]
    float* rvs = rks + nrel * dim;   // [nrel][dim]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = CHAIN_THREADS >> 5;
    for (int d = tid; d < dim; d += CHAIN_THREADS) qs[d] = __ldcg(qkv + (long long)i * ld + h * dim + d);
    for (int e = tid; e < T * dim; e += CHAIN_THREADS) {
        const int j = e / dim, d = e - j * dim;
        ks[e] = __ldcg(qkv + (long long)j * ld + HD + h * dim + d);
        vs[e] = __ldcg(qkv + (long long)j * ld + 2 * HD + h * dim + d);
    }
    for (int e = tid; e < nrel * dim; e += CHAIN_THREADS) { rks[e] = __ldg(o.x1 + e); rvs[e] = __ldg(o.x2 + e); }
    chain_sync<NAMED>();
    for (int j = warp; j < T; j += nw) {
        const int rel = j - i + window;
        const bool inw = rel >= 0 && rel < nrel;
        float a = 0.f;
        for (int d = lane; d < dim; d += 32) a = fmaf(qs[d], ks[j * dim + d] + (inw ? rks[rel * dim + d] : 0.f), a);
        a = warp_sum(a);
        if (lane == 0) ps[j] = a;
    }
    chain_sync<NAMED>();
    if (warp == 0) {
        float mx = -3.402823466e+38f;
        for (int j = lane; j < T; j += 32) mx = fmaxf(mx, ps[j]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = lane; j < T; j += 32) { const float e = expf(ps[j] - mx); ps[j] = e; sum += e; }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        for (int j = lane; j < T; j += 32) ps[j] *= inv;
    }
    chain_sync<NAMED>();
    for (int d = tid; d < dim; d += CHAIN_THREADS) {
        float a = 0.f;
        for (int j = 0; j < T; ++j) {
            const int rel = j - i + window;
            float v = vs[j * dim + d];
            if (rel >= 0 && rel < nrel) v += rvs[rel * dim + d];
            a = fmaf(ps[j], v, a);
        }
        o.y0[(long long)i * o.ld1 + h * dim + d] = a;
    }
}

__device__ __forceinline__ void relattn_item(const ChainOpDev& o, int local, float* sm) { relattn_item_t<false>(o, local, sm); }
__device__ __forceinline__ void relattn_item_sync2(const ChainOpDev& o, int local, float* sm) { relattn_item_t<true>(o, local, sm); chain_sync<true>(); }


__device__ int g_chain_wpre = 0;     // RVC_CHAIN_WPRE=1: first weight stages of the next tile requested before the barrier wait
__device__ int g_chain_poll = 0;     // RVC_CHAIN_POLL=1: waiters poll the arrival counter itself (one hop less than the published phase word)
constexpr int CHAIN_MAX_OPS = 256;   // table entries kept in shared memory (plan.cpp splits longer runs)

__global__ void __launch_bounds__(CHAIN_THREADS, 1)
chain_kernel(const ChainOpDev* __restrict__ ops, const ChainPhaseDev* __restrict__ phases, int n_ops, int n_phases, unsigned int* bar,
             unsigned long long* dbg, int cluster_mode) {
    extern __shared__ __align__(16) float smem[];
    __shared__ ChainOpDev s_op;
    __shared__ ChainPhaseDev s_phase[CHAIN_MAX_OPS];
    __shared__ int s_item_end[CHAIN_MAX_OPS];   // item0 + items of every op (within its phase)
    __shared__ int s_last;
    const int tid = threadIdx.x;
    const unsigned int G = gridDim.x;
    for (int i = tid; i < n_phases; i += CHAIN_THREADS) s_phase[i] = phases[i];
    for (int i = tid; i < n_ops; i += CHAIN_THREADS) s_item_end[i] = ops[i].item0 + ops[i].items;
    __syncthreads();
    int cur = -1;            // op whose descriptor sits in s_op
    bool pre = false;        // s_op of this CTA's first item of the phase is already loaded
    bool wpre = false;       // ... and the first W stages of that item are in flight (RVC_CHAIN_WPRE)
    auto fetch_op = [&](const ChainPhaseDev& P, int item) {   // all threads; ends with a barrier
        int o = P.op0;
        while (o + 1 < P.op1 && item >= s_item_end[o]) ++o;
        if (o != cur) {
            const int* src = reinterpret_cast<const int*>(ops + o);
            int* dst = reinterpret_cast<int*>(&s_op);
            for (int i = tid; i < int(sizeof(ChainOpDev) / 4); i += CHAIN_THREADS) dst[i] = src[i];
            cur = o;
        }
        __syncthreads();
    };
    constexpr int OP_WORDS = int(sizeof(ChainOpDev) / 4);
    static_assert(OP_WORDS <= CHAIN_THREADS, "descriptor is copied one word per thread");
    for (int ph = 0; ph < n_phases; ++ph) {
        const ChainPhaseDev P = s_phase[ph];
        if (dbg && blockIdx.x == 0 && tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); dbg[ph] = t; }
        // descriptor of this CTA's first item of the NEXT phase: requested now (one word per thread, in
        // registers) so that it has long arrived when the phase ends
        int nxt_o = -1, nxt_word = 0;
        if (ph + 1 < n_phases) {
            const ChainPhaseDev Q = s_phase[ph + 1];
            if (int(blockIdx.x) < Q.items) {
                int o = Q.op0;
                while (o + 1 < Q.op1 && int(blockIdx.x) >= s_item_end[o]) ++o;
                nxt_o = o;
                if (tid < OP_WORDS) nxt_word = __ldg(reinterpret_cast<const int*>(ops + o) + tid);
            }
        }
        bool first = true;
        for (int item = blockIdx.x; item < P.items; item += G) {
            const bool preloaded = first && pre;
            if (!preloaded) {
                __syncthreads();  // previous item is done with smem / s_op
                fetch_op(P, item);
            }
            const int stamp_row = (first && ph < 256) ? ph : -1;
            CH_STAMP(0);
            first = false;
            const int local = item - s_op.item0;
            switch (s_op.kind) {
                case CH_GEMM: gemm_dispatch(s_op, local, smem, &s_last, preloaded && wpre, 0, stamp_row); break;
                case CH_GEMM_DIRECT: gemm_direct_item(s_op.g, local); break;
                case CH_AVGPOOL: avgpool_item(s_op, local); break;
                case CH_LAYERNORM: layernorm_item(s_op, local); break;
                case CH_RELATTN: relattn_item(s_op, local, smem); break;
                default: break;
            }
        }
        pre = false; wpre = false;
        if (ph + 1 < n_phases) {
            // arrive first, then use the wait: descriptor + first weight stages of this CTA's next item
            if (cluster_mode) {
                // the whole grid is ONE thread-block cluster: the hardware cluster barrier (release / acquire at cluster
                // scope, a few hundred cycles) stands where the L2 atomic + polling round trip of the grid barrier stood
                asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
                if (blockIdx.x == 0 && tid == 0 && ph < 256) g_chain_stamp[ph * 8 + 7] = clock64();
            } else {
            __syncthreads();
            if (tid == 0) {
                // arrivals are counted on bar[0]; the last one publishes the phase on bar[32] (its own 128-byte
                // line), so the waiters' polling never queues behind the atomics
                __threadfence();
                const unsigned int ticket = atomicAdd(bar, 1u);
                if (ticket == (unsigned int)(ph + 1) * G - 1) { __threadfence(); asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(bar + 32), "r"((unsigned int)(ph + 1)) : "memory"); }
                if (blockIdx.x == 0 && ph < 256) g_chain_stamp[ph * 8 + 7] = clock64();
            }
            }
            if (nxt_o >= 0) {
                if (cluster_mode) __syncthreads();   // every thread is done with s_op of this phase (the cluster arrive does not wait)
                if (tid < OP_WORDS) reinterpret_cast<int*>(&s_op)[tid] = nxt_word;
                cur = nxt_o;
                __syncthreads();
                if (s_op.kind == CH_GEMM) {
                    gemm_dispatch(s_op, int(blockIdx.x) - s_op.item0, smem, &s_last, false, 1);
                    if (g_chain_wpre) { gemm_dispatch(s_op, int(blockIdx.x) - s_op.item0, smem, &s_last, false, 2); wpre = true; }
                } else if (s_op.kind == CH_LAYERNORM || s_op.kind == CH_RELATTN) {
                    // parameter vectors of the next op towards L2 (they are cold every window)
                    const int bytes = (s_op.kind == CH_LAYERNORM ? s_op.i1 : (2 * s_op.i3 + 1) * s_op.i2) * 4;
                    for (int l = tid * 128; l < bytes; l += CHAIN_THREADS * 128) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(s_op.x1) + l));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(s_op.x2) + l));
                    }
                }
                pre = true;
            }
            if (cluster_mode) {
                if (blockIdx.x == 0 && tid == 0 && ph < 256) g_chain_stamp2[ph * 2] = clock64();
                asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
                if (blockIdx.x == 0 && tid == 0 && ph < 256) g_chain_stamp2[ph * 2 + 1] = clock64();
            } else {
            if (tid == 0) {
                const unsigned int target = (unsigned int)(ph + 1);
                const long long t0 = clock64();
                if (blockIdx.x == 0 && ph < 256) g_chain_stamp2[ph * 2] = t0;
                while (true) {
                    unsigned int v;
                    if (g_chain_poll) {
                        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
                        if (v >= target * G) break;
                    } else {
                    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar + 32) : "memory");   // plain L2 poll ...
                    if (v >= target) break;
                    }
                    if (nxt_o < 0) __nanosleep(200);   // CTAs with nothing to do next phase poll gently
                    if (clock64() - t0 > (6ll << 30)) __trap();  // a CTA that never arrives must fail loudly, not hang the GPU
                }
                __threadfence();   // ... one acquire-side fence once the phase is published
                if (blockIdx.x == 0 && ph < 256) g_chain_stamp2[ph * 2 + 1] = clock64();
            }
            __syncthreads();
            }
        }
    }
    if (dbg && blockIdx.x == 0 && tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); dbg[n_phases] = t; }
    // exit ticket: the last CTA out re-arms the barrier words for the next launch / graph replay
    __syncthreads();
    if (tid == 0 && !cluster_mode) {
        __threadfence();
        if (atomicAdd(bar + 1, 1u) == G - 1) { bar[0] = 0; bar[1] = 0; bar[32] = 0; __threadfence(); }
    }
}


// ================================================================================================================
// Slab chain (chain.h): ONE thread-block cluster, GEMMs split by output columns only, weights streamed from a per-CTA
// contiguous stream of 16 KB chunks (cp.async.bulk + mbarrier) that runs ahead of the phase barriers.
// ================================================================================================================
__device__ __forceinline__ uint32_t sl_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void sl_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sl_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void sl_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sl_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sl_mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = sl_smem_u32(bar);
    uint32_t done = 0;
    long long spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) break;
        if (++spins > (1ll << 27)) __trap();   // a chunk that never lands must fail loudly, not hang the GPU
    }
}
__device__ __forceinline__ void sl_bulk_g2s(float* smem_dst, const float* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(sl_smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(sl_smem_u32(bar)) : "memory");
}

__device__ int g_slab_dbg = 0;   // timing experiments (RVC_SLAB_DBG): 1 = no FMA loop, 2 = no activation staging, 3 = no weight stream waits

__device__ __forceinline__ void sl_tma_2d(const CUtensorMap* map, float* smem_dst, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(sl_smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(sl_smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

struct SlabStream {
    const CUtensorMap* map; // the slab buffer as [rows of 32 floats]: a chunk = a box of 128 rows (tensor-map TMA streams at
                            // ~55 B/clk per SM here, 1-D cp.async.bulk measured ~7 B/clk)
    long long row0;         // first row of this CTA's stream
    const float* base;     // this CTA's weight stream
    const int* bytes;      // bytes of every chunk of the chain (shared by all CTAs)
    int total;             // chunks in the stream
    int issued, consumed;  // uniform across the CTA
};

// Compute threads only count: the producer warp (warp 8) issues the copies (an issuing thread is busy for ~1000 cycles
// per bulk copy, which stalled the whole CTA at its next barrier when thread 0 did it between chunks).  After every
// consumed chunk compute thread 0 publishes the count; chunk i may be requested once chunk i - SLAB_RING is consumed.
__device__ __forceinline__ void slab_top_up(SlabStream& st, float* ring, uint64_t* full, volatile int* s_consumed) {
    (void)ring; (void)full;
    if (threadIdx.x == 0) *s_consumed = st.consumed;
}
__device__ __forceinline__ void slab_produce(SlabStream& st, float* ring, uint64_t* full, volatile int* s_consumed) {
    const int lim = min(st.total, *s_consumed + SLAB_RING);
    while (st.issued < lim) {
        const int slot = st.issued % SLAB_RING;
        if (st.map) {
            sl_mbar_expect_tx(&full[slot], SLAB_CHUNK_FLOATS * 4);   // whole 16 KB boxes (the tail of a chunk is zero padding)
            sl_tma_2d(st.map, ring + slot * SLAB_CHUNK_FLOATS, &full[slot], 0, int(st.row0 + (long long)st.issued * (SLAB_CHUNK_FLOATS / 32)));
        } else {
            const uint32_t nb = uint32_t(st.bytes[st.issued]);
            sl_mbar_expect_tx(&full[slot], nb);
            sl_bulk_g2s(ring + slot * SLAB_CHUNK_FLOATS, st.base + (long long)st.issued * SLAB_CHUNK_FLOATS, nb, &full[slot]);
        }
        ++st.issued;
    }
}

// One GEMM of a slab chain on this CTA's column slice: lanes split k inside a chunk, warps split the columns (CW each),
// every thread carries MR x CW partial sums over its k's through all chunks; a halving butterfly then leaves each lane
// with NV / 32 finished outputs (fixed summation order) for the epilogue.
template <int MR, int CW>
__device__ __forceinline__ void slab_gemm(const ChainOpDev& o, SlabStream& st, float* ring, float* abuf, uint64_t* full, volatile int* s_consumed) {
    const GemmParams& p = o.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = p.M, K = p.K, nc = o.slab_nc, kc = o.slab_kc;
    const int n_base = blockIdx.x * nc;
    const int ncols = max(0, min(nc, p.N - n_base));
    const int seg_len = p.seg_len >= K ? K : p.seg_len;
    const int nseg = K / seg_len;
    // ---- activations of this op: the span the (overlapping / segmented) rows cover, L2 -> shared memory ----
    const long long span = (long long)(M - 1) * p.lda + (long long)(nseg - 1) * p.seg_stride + seg_len;
    {
        const float4* src = reinterpret_cast<const float4*>(p.A);
        float4* dst = reinterpret_cast<float4*>(abuf);
        const int n4 = int(span >> 2);
        if (g_slab_dbg != 2) for (int i = tid; i < n4; i += CHAIN_THREADS) dst[i] = __ldcg(src + i);
    }
    asm volatile("bar.sync 2, 256;" ::: "memory");
    const int col0 = warp * CW;                 // first column (inside the slice) of this warp
    const bool active = col0 < ncols;
    const int lda = int(p.lda), seg_stride = int(p.seg_stride);
    // ---- where this lane's finished outputs will sit after the halving butterfly: their bias / residual are requested
    //      NOW (parameter vectors are cold in L2 every window: an HBM round trip each if fetched in the epilogue) ----
    constexpr int NV = (MR * CW + 63) / 64 * 64;   // PER even: the two columns of a gate pair end up in the same lane
    constexpr int PER = NV / 32;
    int first = 0;
#pragma unroll
    for (int sft = 0; sft < 5; ++sft) if (lane & (16 >> sft)) first += NV >> (sft + 1);
    const bool plain = p.out_mode == OUT_PLAIN && p.act != ACT_GATE;
    float bpre[PER], rpre[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int idx = first + i, r = idx / CW, c = idx - r * CW, n = n_base + col0 + c;
        const bool ok = active && r < M && idx < MR * CW && (col0 + c) < ncols && n < p.N;
        bpre[i] = (ok && p.bias) ? __ldg(p.bias + n) : 0.f;
        rpre[i] = (ok && plain && p.R) ? __ldcg(p.R + (long long)r * p.ldr + n) : 0.f;
    }
    float acc[MR][CW];
#pragma unroll
    for (int r = 0; r < MR; ++r)
#pragma unroll
        for (int c = 0; c < CW; ++c) acc[r][c] = 0.f;
    for (int ch = 0; ch < o.slab_chunks; ++ch) {
        const int slot = st.consumed % SLAB_RING;
        if (g_slab_dbg != 3) sl_mbar_wait(&full[slot], (st.consumed / SLAB_RING) & 1);
        if (active && g_slab_dbg != 1) {
            const float* wch = ring + slot * SLAB_CHUNK_FLOATS + col0 * kc;   // chunk layout [nc][kc]
            const int k0 = ch * kc, kn = min(kc, K - k0);
#pragma unroll 1
            for (int kk = lane; kk < kn; kk += 32) {
                const int k = k0 + kk;
                const int sg = k / seg_len;
                const float* ap = abuf + sg * seg_stride + (k - sg * seg_len);
                float wv[CW], av[MR];
#pragma unroll
                for (int c = 0; c < CW; ++c) wv[c] = wch[c * kc + kk];
#pragma unroll
                for (int r = 0; r < MR; ++r) av[r] = r < M ? ap[r * lda] : 0.f;     // all loads of the step first, then the FMAs
#pragma unroll
                for (int r = 0; r < MR; ++r)
#pragma unroll
                    for (int c = 0; c < CW; ++c) acc[r][c] = fmaf(av[r], wv[c], acc[r][c]);
            }
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");   // every compute warp is done with the chunk: its slot may be refilled
        ++st.consumed;
        slab_top_up(st, ring, full, s_consumed);
    }
    if (!active) return;
    // ---- halving butterfly: NV values per lane -> NV / 32 finished sums per lane ----
    float v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = i < MR * CW ? acc[i / CW][i % CW] : 0.f;
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int half = NV >> (s + 1);
        const int mask = 16 >> s;
        const bool upper = (lane & mask) != 0;
#pragma unroll
        for (int i = 0; i < NV / 2; ++i) {
            if (i < half) {
                const float keep = upper ? v[i + half] : v[i];
                const float send = upper ? v[i] : v[i + half];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
            }
        }
    }
    // ---- epilogue: value index = r * CW + c ----
    const bool gate = p.act == ACT_GATE;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int idx = first + i;
        const int r = idx / CW, c = idx - r * CW;
        const float partner = v[i ^ 1];   // the other column of a gate pair (2j, 2j + 1): same lane, neighbouring slot
        const int n = n_base + col0 + c;
        if (!(r < M && idx < MR * CW && (col0 + c) < ncols && n < p.N)) continue;
        if (plain) {
            // common case inline, bias / residual already in registers (same arithmetic as chain_epilogue)
            float ov = chain_act_fast(p.act, fmaf(p.alpha, v[i], bpre[i])) + rpre[i];
            const bool masked = p.mask_period > 0 && (r % p.mask_period) >= p.mask_valid;
            if (masked) ov = 0.f;
            p.C[(long long)r * p.ldc + n] = ov;
            if (p.C2) p.C2[(long long)r * p.ldc2 + n] = masked ? 0.f : chain_act_fast(p.act2, ov);
        } else if (gate && p.out_mode == OUT_PLAIN) {
            if (n & 1) continue;
            const float v0 = fmaf(p.alpha, v[i], bpre[i]), v1 = fmaf(p.alpha, partner, bpre[i ^ 1]);
            float gv = tanhf(v0) * sigmoid_f(v1);
            const int col = n >> 1;
            if (p.R) gv += __ldcg(p.R + (long long)r * p.ldr + col);
            const bool masked = p.mask_period > 0 && (r % p.mask_period) >= p.mask_valid;
            if (masked) gv = 0.f;
            p.C[(long long)r * p.ldc + col] = gv;
            if (p.C2) p.C2[(long long)r * p.ldc2 + col] = masked ? 0.f : apply_act(p.act2, gv);
        } else {
            chain_epilogue(p, p.bias, p.C, p.C2, p.R, r, n, v[i], partner);
        }
    }
}

__device__ __forceinline__ void slab_gemm_dispatch(const ChainOpDev& o, SlabStream& st, float* ring, float* abuf, uint64_t* full, volatile int* s_consumed) {
    const bool small = o.g.M <= 8;
    switch (o.slab_cw) {
        case 2: if (small) slab_gemm<8, 2>(o, st, ring, abuf, full, s_consumed); else slab_gemm<SLAB_MAX_ROWS, 2>(o, st, ring, abuf, full, s_consumed); break;
        case 4: if (small) slab_gemm<8, 4>(o, st, ring, abuf, full, s_consumed); else slab_gemm<SLAB_MAX_ROWS, 4>(o, st, ring, abuf, full, s_consumed); break;
        case 8: slab_gemm<8, 8>(o, st, ring, abuf, full, s_consumed); break;   // the host only picks 8 columns per warp for <= 8 rows
        default: if (small) slab_gemm<8, 6>(o, st, ring, abuf, full, s_consumed); else slab_gemm<SLAB_MAX_ROWS, 6>(o, st, ring, abuf, full, s_consumed); break;
    }
}

constexpr int SLAB_THREADS = CHAIN_THREADS + 32;   // 8 compute warps + the weight-stream producer warp

__global__ void __launch_bounds__(SLAB_THREADS, 1)
slab_chain_kernel(const __grid_constant__ CUtensorMap slab_map, int use_map, const ChainOpDev* __restrict__ ops, const ChainPhaseDev* __restrict__ phases,
                  int n_ops, int n_phases, const float* __restrict__ slabs, long long stream_floats, const int* __restrict__ chunk_bytes,
                  int total_chunks, unsigned long long* dbg) {
    extern __shared__ __align__(128) float smem[];
    __shared__ ChainOpDev s_op;
    __shared__ __align__(8) uint64_t s_full[SLAB_RING];
    __shared__ volatile int s_consumed, s_phase_done;
    float* ring = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem) + 127) & ~uintptr_t(127));   // TMA destinations: 128-byte aligned
    float* abuf = ring + SLAB_RING * SLAB_CHUNK_FLOATS;
    const int tid = threadIdx.x;
    const int G = gridDim.x;
    if (tid == 0) {
        for (int i = 0; i < SLAB_RING; ++i) sl_mbar_init(&s_full[i], 1);
        s_consumed = 0; s_phase_done = -1;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    SlabStream st;
    st.base = slabs + (long long)blockIdx.x * stream_floats; st.bytes = chunk_bytes; st.total = total_chunks; st.issued = 0; st.consumed = 0;
    st.map = use_map ? &slab_map : nullptr; st.row0 = (long long)blockIdx.x * (stream_floats / 32);
    if (tid >= CHAIN_THREADS) {
        // ===== producer warp: keeps SLAB_RING chunks of this CTA's weight stream in flight, across op and phase boundaries
        // (weights do not depend on activations); joins the cluster barrier of a phase once the compute warps are through it =====
        const int lane = tid & 31;
        for (int ph = 0; ph < n_phases; ++ph) {
            if (lane == 0) {
                long long spins = 0;
                while (true) {
                    slab_produce(st, ring, s_full, &s_consumed);
                    if (s_phase_done >= ph) break;
                    __nanosleep(256);   // a tight poll of shared memory starves the compute warps' own shared-memory loads
                    if (++spins > (1ll << 31)) __trap();
                }
            }
            __syncwarp();
            if (ph + 1 < n_phases) {
                asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
                asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
            }
        }
        return;
    }
    constexpr int OP_WORDS = int(sizeof(ChainOpDev) / 4);
    for (int ph = 0; ph < n_phases; ++ph) {
        const ChainPhaseDev P = phases[ph];
        if (dbg && blockIdx.x == 0 && tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); dbg[ph] = t; }
        for (int oi = P.op0; oi < P.op1; ++oi) {
            asm volatile("bar.sync 2, 256;" ::: "memory");   // the previous op is done with s_op / the staging buffer
            if (tid < OP_WORDS) reinterpret_cast<int*>(&s_op)[tid] = __ldg(reinterpret_cast<const int*>(ops + oi) + tid);
            asm volatile("bar.sync 2, 256;" ::: "memory");
            if (s_op.kind == CH_GEMM) {
                slab_gemm_dispatch(s_op, st, ring, abuf, s_full, &s_consumed);
            } else {
                for (int local = blockIdx.x; local < s_op.items; local += G) {
                    switch (s_op.kind) {
                        case CH_GEMM_DIRECT: gemm_direct_item(s_op.g, local); break;
                        case CH_AVGPOOL: avgpool_item(s_op, local); break;
                        case CH_LAYERNORM: layernorm_item(s_op, local); break;
                        case CH_RELATTN: relattn_item_sync2(s_op, local, abuf); break;
                        default: break;
                    }
                }
            }
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (tid == 0) s_phase_done = ph;     // the producer warp may now join this phase's cluster barrier
        if (ph + 1 < n_phases) {
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        }
    }
    if (dbg && blockIdx.x == 0 && tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); dbg[n_phases] = t; }
}

// W[n][k] (row pitch ldw) -> the per-CTA streams: CTA c, chunk j of this op = [nc][kc] floats of columns [c nc, (c + 1) nc),
// k in [j kc, (j + 1) kc), zero-filled outside N x K
__global__ void slab_pack_kernel(const float* __restrict__ W, long long ldw, int N, int K, int nc, int kc, int chunks, float* __restrict__ slabs,
                                 long long stream_floats, int chunk0) {
    const int c = blockIdx.y, j = blockIdx.z;
    float* dst = slabs + (long long)c * stream_floats + (long long)(chunk0 + j) * SLAB_CHUNK_FLOATS;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nc * kc; e += gridDim.x * blockDim.x) {
        const int col = e / kc, kk = e - col * kc;
        const int n = c * nc + col, k = j * kc + kk;
        dst[e] = (n < N && k < K) ? W[(long long)n * ldw + k] : 0.f;
    }
}

int g_chain_max_ctas = 0;

}  // namespace

void init_chain_attributes() {
    { const char* w = getenv("RVC_CHAIN_WPRE"); int v = (w && w[0] == '1') ? 1 : 0; cudaMemcpyToSymbol(g_chain_wpre, &v, sizeof(int)); }
    { const char* w = getenv("RVC_SLAB_DBG"); int v = w ? atoi(w) : 0; cudaMemcpyToSymbol(g_slab_dbg, &v, sizeof(int)); }
    { const char* w = getenv("RVC_CHAIN_POLL"); int v = (w && w[0] == '1') ? 1 : 0; cudaMemcpyToSymbol(g_chain_poll, &v, sizeof(int)); }
    cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CHAIN_SMEM_BYTES);
    cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);   // single-cluster chains of 16 CTAs
    cudaFuncSetAttribute(slab_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SLAB_SMEM_BYTES);
    cudaFuncSetAttribute(slab_chain_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    int per_sm = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chain_kernel, CHAIN_THREADS, CHAIN_SMEM_BYTES);
    g_chain_max_ctas = per_sm * sms;
}

int chain_max_coresident_ctas() { return g_chain_max_ctas; }

void chain_debug_read(long long* out, int n) { cudaMemcpyFromSymbol(out, g_chain_stamp, sizeof(long long) * size_t(n)); }
void chain_debug_read2(long long* out, int n) { cudaMemcpyFromSymbol(out, g_chain_stamp2, sizeof(long long) * size_t(n)); }

void launch_slab_pack(const float* W, long long ldw, int N, int K, int nc, int kc, int chunks, float* slabs, long long stream_floats,
                      int chunk0, int G, cudaStream_t stream) {
    slab_pack_kernel<<<dim3(4, unsigned(G), unsigned(chunks)), 256, 0, stream>>>(W, ldw, N, K, nc, kc, chunks, slabs, stream_floats, chunk0);
}

int launch_chain(const ChainDev& c, cudaStream_t stream) {
    if (c.wstream) return launch_wstream(c, stream);
    if (c.slab) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(unsigned(c.grid)); cfg.blockDim = dim3(SLAB_THREADS); cfg.dynamicSmemBytes = SLAB_SMEM_BYTES; cfg.stream = stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = unsigned(c.grid); attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributePriority;
        attr[1].val.priority = g_launch_priority;
        cfg.attrs = attr; cfg.numAttrs = g_launch_priority != 0 ? 2 : 1;
        const ChainOpDev* ops = c.d_ops; const ChainPhaseDev* phases = c.d_phases; int no = c.n_ops, n = c.n_phases;
        const float* slabs = c.d_slabs; long long sf = c.slab_stream_floats; const int* cb = c.d_chunk_bytes; int tc = c.slab_total_chunks;
        unsigned long long* dbg = c.d_dbg;
        CUtensorMap map;
        std::memset(&map, 0, sizeof(map));
        int use_map = 0;
        {
            static const bool want = getenv("RVC_SLAB_TMA") && getenv("RVC_SLAB_TMA")[0] == '1';   // measured slower than the 1-D bulk copies (profiles/README.md): opt-in
            typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                         const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
            static EncodeFn fn = nullptr;
            static bool tried = false;
            if (!tried) {
                tried = true;
                void* pfn = nullptr;
                cudaDriverEntryPointQueryResult q;
                if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &pfn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
                    fn = reinterpret_cast<EncodeFn>(pfn);
            }
            if (want && fn) {
                const cuuint64_t rows = cuuint64_t(c.grid) * cuuint64_t(sf / 32);
                cuuint64_t dims[2] = {32, rows};
                cuuint64_t strides[1] = {128};
                cuuint32_t box[2] = {32, cuuint32_t(SLAB_CHUNK_FLOATS / 32)};
                cuuint32_t estr[2] = {1, 1};
                if (fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(slabs), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
                    use_map = 1;
            }
        }
        cudaLaunchKernelEx(&cfg, slab_chain_kernel, map, use_map, ops, phases, no, n, slabs, sf, cb, tc, dbg);
        return 1;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(unsigned(c.grid)); cfg.blockDim = dim3(CHAIN_THREADS); cfg.dynamicSmemBytes = CHAIN_SMEM_BYTES; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    if (c.cluster) {
        // the grid is one thread-block cluster (co-scheduled by the hardware): cluster barriers between phases
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = unsigned(c.grid); attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    } else {
        attr[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident or none: the grid barrier cannot deadlock
        attr[0].val.cooperative = 1;
    }
    attr[1].id = cudaLaunchAttributePriority;
    attr[1].val.priority = g_launch_priority;
    cfg.attrs = attr; cfg.numAttrs = g_launch_priority != 0 ? 2 : 1;
    const ChainOpDev* ops = c.d_ops; const ChainPhaseDev* phases = c.d_phases; int no = c.n_ops, n = c.n_phases; unsigned int* bar = c.d_bar;
    unsigned long long* dbg = c.d_dbg;
    int cm = c.cluster ? 1 : 0;
    cudaLaunchKernelEx(&cfg, chain_kernel, ops, phases, no, n, bar, dbg, cm);
    return 1;
}

// largest single-cluster grid the chain kernel can be launched with on this device (0 = clusters unavailable)
int chain_max_cluster_ctas() {
    int best = 0;
    for (int g = 16; g >= 2; g >>= 1) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(unsigned(g)); cfg.blockDim = dim3(CHAIN_THREADS); cfg.dynamicSmemBytes = CHAIN_SMEM_BYTES;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = unsigned(g); attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, chain_kernel, &cfg) == cudaSuccess && n >= 1) { best = g; break; }
        cudaGetLastError();
    }
    return best;
}

}  // namespace rvc
