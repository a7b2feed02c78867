// kernels_knn_umma.cu - exact L2 top-k retrieval for many queries: tcgen05 candidate pass + exact fp32 re-rank.
//
// The fp32 scan of kernels_knn.cu evaluates sum (x-y)^2 on the CUDA cores: with Q = 128 queries against a
// 1 M x 256 index (BASELINE configs[4]) that is 65 GFLOP of dependent FMAs - 11 ms, seventy times the 157 us it
// takes to stream the index from HBM.  A contraction this dense belongs on the tensor cores:
//
//   pass 1 (knn_umma_scan_kernel, HBM-bound):  s[q,n] = |y_n|^2 - 2 x_q.y_n   for every index row, with the queries
//     as the 128-row A operand (resident in shared memory for the whole scan) and tiles of 32 index rows as B,
//     2-term FP16 split (x = hi + lo' 2^-11; D0 += x_hi.y_hi, D1 += x_hi.y_lo' + x_lo'.y_hi: the dropped term is
//     2^-22 relative, accumulation fp32 in TMEM).  The index planes (y_hi, y_lo', |y|^2) are built once when the
//     index is loaded, tile by tile in exactly the shared-memory image the MMA descriptors expect (K-major
//     SWIZZLE_128B): a tile is ONE contiguous 32 KB cp.async.bulk (no tensor map, one issue per tile - the issue
//     cost of several small tensor loads per tile paced the first version at 11 k cycles per tile) through a
//     3-5 stage ring; two TMEM
//     accumulator buffers let the epilogue of a tile overlap the MMAs of the next.  TMEM lane = query: every epilogue
//     thread owns one query and keeps its KC = 16 best (score, row) pairs in registers - no Q x N matrix is ever
//     written.  One candidate list per CTA and query leaves the SM.
//   pass 2 (knn_rerank_kernel):  per query, the KC globally best candidates by approximate score are re-evaluated
//     EXACTLY - fp32 sum (x-y)^2 in the summation order of kernels_knn.cu (lane l accumulates the float4 chunks
//     l, l+32, ... with fmaf in ascending order, then the xor-butterfly 16, 8, 4, 2, 1) - and the k nearest by
//     (distance, row) are returned.  A guard proves the result: every row outside the candidate set has an
//     approximate distance >= a_KC, so its exact distance is >= a_KC - eps with eps the rigorous error bound of pass
//     1 (2 |x| max|y| (2^-21 + C 2^-24), doubled); if a_KC - eps > d_k the candidate set provably contains the true top-k.
//     Otherwise (never seen on the test data; possible for pathological near-duplicate indices) the same CTA
//     falls back to the exact scan of all rows for its query.  The result is therefore ALWAYS the exact fp32 top-k.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cfloat>
#include <climits>
#include <cstdio>
#include <cstdlib>

#include "launch.h"
#include "pdl.cuh"

namespace rvc {

namespace {

constexpr int KU_BN = KNN_UMMA_BN;        // index rows per tile
constexpr int KU_KC = KNN_UMMA_KC;        // candidates kept per query and CTA
constexpr int KU_THREADS = 192;           // warps 0-3: epilogue (TMEM lane quadrants), warp 4: TMA, warp 5: MMA
constexpr int KU_YN_SLOTS = 12;           // |y|^2 ring (>= stages + 3, see the producer)
constexpr int KU_TMEM_COLS = 512;         // [0,128) x_hi | [128,256) x_lo' | [256,384) two accumulator buffers (D0 | D1)
constexpr int KU_ACC_COL = 256;
constexpr float KU_LO_SCALE = 2048.0f;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    long long spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) break;
        if (++spins > (1ll << 28)) __trap();  // a broken pipeline must fail loudly, never hang the GPU
    }
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (rows of 64 halves = 128 B, 8-row atoms of 1024 B)
__device__ __forceinline__ uint64_t umma_desc128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= uint64_t((smem_addr & 0x3FFFFu) >> 4);
    d |= uint64_t(1) << 16;
    d |= uint64_t(1024 >> 4) << 32;
    d |= uint64_t(1) << 46;
    d |= uint64_t(2) << 61;
    return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// A operand from tensor memory (lane = row, 16-bit elements packed two per 32-bit column along K: K = 16 -> 8 columns)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
        "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
          "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
          "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
// one lane of a converged warp (elect.sync): the tcgen05 issue path stays warp-uniform - inside `if (lane == 0)` the
// compiler cannot prove uniformity and wraps every UTCHMMA in an ELECT / BRA.U.ANY loop with R2UR moves
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 columns of this warp's 32 TMEM lanes, twice (main + correction accumulator), one wait
__device__ __forceinline__ void tmem_ld32x2(uint32_t t0, uint32_t t1, float* v, float* c) {
    uint32_t r[32], q[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
        "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
          "=r"(r[30]), "=r"(r[31])
        : "r"(t0));
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
        "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]), "=r"(q[9]),
          "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]), "=r"(q[16]), "=r"(q[17]), "=r"(q[18]), "=r"(q[19]),
          "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]), "=r"(q[25]), "=r"(q[26]), "=r"(q[27]), "=r"(q[28]), "=r"(q[29]),
          "=r"(q[30]), "=r"(q[31])
        : "r"(t1));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) { v[i] = __uint_as_float(r[i]); c[i] = __uint_as_float(q[i]); }
}

__device__ long long g_ku_stamp[8 * 32];   // [event][tile < 32] clock64 of CTA 0: 0 producer empty-ok, 1 producer issued, 2 MMA acc-empty-ok, 3 MMA full-ok,
                                           // 4 MMA issued, 5 epilogue acc-full-ok, 6 epilogue ld done, 7 epilogue tile done
#ifdef RVC_KU_STAMPS
#define KU_STAMP(e, i) do { if (blockIdx.x == 0 && (i) >= 180 && (i) < 212) g_ku_stamp[(e) * 32 + ((i) - 180)] = clock64(); } while (0)   // late tiles
#else
#define KU_STAMP(e, i) do { } while (0)
#endif

template <int KB>
struct KuCfg {
    // the queries (A operand) live in TENSOR memory for the whole scan: an MMA whose A comes from shared memory pays
    // ~100 cycles per instruction for the operand fetch whatever N (measured: t = 100 + N/2 cycles at M = 128, K = 16)
    static constexpr int STAGE_BYTES = KB * 2 * KU_BN * 128;         // y_hi | y_lo' per k-block
    static constexpr int HIT_BYTES = 4 * 32 * 32 * 4;                 // per epilogue warp: the scores of one tile, [element][lane]
    static constexpr int RING = (200 * 1024 - 1024 - KU_YN_SLOTS * KU_BN * 4 - HIT_BYTES) / STAGE_BYTES;
    static constexpr int STAGES = RING > 8 ? 8 : RING;
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + KU_YN_SLOTS * KU_BN * 4 + HIT_BYTES;
    static_assert(STAGES + 3 <= KU_YN_SLOTS, "|y|^2 ring too short");
    static_assert(STAGES >= 2, "ring too small");
};

template <int KB, int KC>
__global__ void __launch_bounds__(KU_THREADS, 1)
knn_umma_scan_kernel(const uint8_t* __restrict__ planes, const float* __restrict__ queries, long long ldq, int Q, int q0, int Qw, long long wQ, int N,
                     float* __restrict__ cand_s, int* __restrict__ cand_i, long long wCandS, long long wCandI) {
    using Cfg = KuCfg<KB>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[STAGES], bar_empty[STAGES], bar_acc_full[2], bar_acc_empty[2];
    __shared__ uint32_t s_tmem;
    pdl_enter();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sB = smem;
    float* sYn = reinterpret_cast<float*>(sB + STAGES * Cfg::STAGE_BYTES);
    float* sHit = sYn + KU_YN_SLOTS * KU_BN;

    const int tiles_total = (N + KU_BN - 1) / KU_BN;
    const int per = (tiles_total + gridDim.x - 1) / gridDim.x;
    const int t_begin = min(tiles_total, int(blockIdx.x) * per), t_end = min(tiles_total, t_begin + per);
    const int ntiles = t_end - t_begin;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&bar_acc_full[a], 1); mbar_init(&bar_acc_empty[a], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(uint32_t(KU_TMEM_COLS)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = s_tmem;
    // queries -> the two fp16 planes of the A operand in tensor memory: TMEM lane = query (rows >= Q are zero), column j of
    // a plane = elements (2j, 2j+1) along K.  Thread q of the epilogue warps converts its own query row.
    if (warp < 4) {
        const int r = tid;
        const uint32_t trow = tmem_base + (uint32_t(warp * 32) << 16);
        const int gq = q0 + r;
        const float* src = queries + (long long)(gq / Qw) * wQ + (long long)(gq % Qw) * ldq;
#pragma unroll 1
        for (int kb = 0; kb < KB; ++kb) {
            uint32_t hw[32], lw[32];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < Q) v = *reinterpret_cast<const float4*>(src + kb * 64 + j * 4);
                const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const __half2 h2 = __floats2half2_rn(x[2 * e], x[2 * e + 1]);
                    const float2 hf = __half22float2(h2);
                    const __half2 l2 = __floats2half2_rn((x[2 * e] - hf.x) * KU_LO_SCALE, (x[2 * e + 1] - hf.y) * KU_LO_SCALE);
                    hw[2 * j + e] = *reinterpret_cast<const uint32_t*>(&h2);
                    lw[2 * j + e] = *reinterpret_cast<const uint32_t*>(&l2);
                }
            }
            tmem_st32(trow + kb * 32, hw);
            tmem_st32(trow + 128 + kb * 32, lw);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (warp == 4) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int i = 0; i < ntiles; ++i) {
                const int s = i % STAGES, ph = (i / STAGES) & 1;
                mbar_wait(&bar_empty[s], ph ^ 1);
                KU_STAMP(0, i);
                mbar_expect_tx(&bar_full[s], Cfg::STAGE_BYTES + KU_BN * 4);
                const uint8_t* tile = planes + (long long)(t_begin + i) * (Cfg::STAGE_BYTES + KU_BN * 4);
                bulk_load(sB + s * Cfg::STAGE_BYTES, tile, Cfg::STAGE_BYTES, &bar_full[s]);
                // |y|^2 of the tile: its slot is reused KU_YN_SLOTS tiles later, when the epilogue of this tile is long
                // done (the producer runs <= STAGES tiles ahead of the MMAs, the MMAs <= 3 tiles ahead of the epilogue)
                bulk_load(sYn + (i % KU_YN_SLOTS) * KU_BN, tile + Cfg::STAGE_BYTES, KU_BN * 4, &bar_full[s]);
                KU_STAMP(1, i);
            }
        }
    } else if (warp == 5) {
        // ===== MMA issuer: the whole warp walks the loop, one elected lane issues =====
        {
            constexpr uint32_t idesc_2n = (1u << 4) | (uint32_t((2 * KU_BN) >> 3) << 17) | (uint32_t(128 >> 4) << 24);
            constexpr uint32_t idesc_n = (1u << 4) | (uint32_t(KU_BN >> 3) << 17) | (uint32_t(128 >> 4) << 24);
            for (int i = 0; i < ntiles; ++i) {
                const int s = i % STAGES, ph = (i / STAGES) & 1, a = i & 1, aph = (i >> 1) & 1;
                mbar_wait(&bar_acc_empty[a], aph ^ 1);    // the epilogue has drained this accumulator buffer
                if (lane == 0) KU_STAMP(2, i);
                mbar_wait(&bar_full[s], ph);
                if (lane == 0) KU_STAMP(3, i);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d0 = tmem_base + KU_ACC_COL + a * (2 * KU_BN);
                const uint32_t bst = smem_u32(sB + s * Cfg::STAGE_BYTES);
                if (elect_one()) {
                // per k-step (16 halves = 8 TMEM columns of A, 32 B along K inside the swizzle row of B):
                // x_hi . [y_hi ; y_lo'] -> D0 | D1 in one instruction (the planes of a k-block are adjacent), x_lo' . y_hi -> D1
#pragma unroll
                for (int kb = 0; kb < KB; ++kb) {
                    const uint32_t b_hi = bst + kb * (2 * KU_BN * 128);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t a_col = tmem_base + (kb * 4 + ks) * 8;
                        umma_f16_ts(d0, a_col, umma_desc128(b_hi + ks * 32), idesc_2n, (kb | ks) ? 1u : 0u);
                        umma_f16_ts(d0 + KU_BN, a_col + 128, umma_desc128(b_hi + ks * 32), idesc_n, 1u);
                    }
                }
                umma_commit(&bar_empty[s]);       // stage reusable once these MMAs have read it
                umma_commit(&bar_acc_full[a]);    // accumulator complete
                KU_STAMP(4, i);
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue warps 0-3: TMEM lane = query; running top-KC by approximate score in registers =====
        // The KC best (score, row) pairs are kept UNSORTED with the position of the worst one: an insertion is KC
        // independent selects plus a log-depth max tree (~40 cycles), taken in a warp-uniform branch only when some lane of
        // the warp has a hit.  (A sorted insertion is a 16-step dependent chain that the compiler predicates: executed for
        // every element it cost ~10 k cycles per tile and paced the whole scan - measured with per-tile stamps.)
        float cs[KC]; int ci[KC];
#pragma unroll
        for (int j = 0; j < KC; ++j) { cs[j] = FLT_MAX; ci[j] = -1; }
        float thr = FLT_MAX; int maxpos = 0;
        const uint32_t trow = tmem_base + (uint32_t(warp * 32) << 16);
        for (int i = 0; i < ntiles; ++i) {
            const int a = i & 1, aph = (i >> 1) & 1;
            mbar_wait(&bar_acc_full[a], aph);
            if (tid == 0) KU_STAMP(5, i);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float v[32], c[32];
            tmem_ld32x2(trow + KU_ACC_COL + a * (2 * KU_BN), trow + KU_ACC_COL + a * (2 * KU_BN) + KU_BN, v, c);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_acc_empty[a]);
            if (tid == 0) KU_STAMP(6, i);
            const float* yn = sYn + (i % KU_YN_SLOTS) * KU_BN;
            const int row0 = (t_begin + i) * KU_BN;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaf(-2.0f, fmaf(c[j], 1.0f / KU_LO_SCALE, v[j]), yn[j]);   // |y|^2 - 2 x.y
            // hits of the whole tile against the threshold at tile start (a superset: thresholds only fall), OR-reduced
            // over the warp in one redux: elements where no lane has a hit are skipped by a warp-uniform branch.  (One vote +
            // branch per element serialised the 32 elements of a tile behind each other: ~100 cycles each, measured.)
            unsigned hm = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) hm |= (v[j] < thr) ? (1u << j) : 0u;
            const unsigned am = __reduce_or_sync(0xffffffffu, hm);
            if (am) {
                // scores of the tile -> this warp's shared-memory strip, then a truly dynamic loop over the set bits only
                // (an unrolled loop of uniform bit tests was predicated by the compiler: all 32 bodies ran whenever am != 0)
                float* strip = sHit + warp * (32 * 32);
#pragma unroll
                for (int j = 0; j < 32; ++j) strip[j * 32 + lane] = v[j];
                __syncwarp();
                unsigned rem = am;
#pragma unroll 1
                while (rem) {
                    const int j = __ffs(rem) - 1;
                    rem &= rem - 1;
                    const float sv = strip[j * 32 + lane];
                    if (sv < thr) {
#pragma unroll
                        for (int u = 0; u < KC; ++u) { const bool w = u == maxpos; cs[u] = w ? sv : cs[u]; ci[u] = w ? row0 + j : ci[u]; }
                        float m[KC]; int mp[KC];
#pragma unroll
                        for (int u = 0; u < KC; ++u) { m[u] = cs[u]; mp[u] = u; }
#pragma unroll
                        for (int st = KC / 2; st >= 1; st >>= 1)
#pragma unroll
                            for (int u = 0; u < st; ++u) { const bool g = m[u + st] > m[u]; m[u] = g ? m[u + st] : m[u]; mp[u] = g ? mp[u + st] : mp[u]; }
                        thr = m[0]; maxpos = mp[0];
                    }
                }
                __syncwarp();
            }
            if (tid == 0) KU_STAMP(7, i);
        }
        const int q = warp * 32 + lane;
        if (q < Q) {
            const int gq = q0 + q, w = gq / Qw, jq = gq - w * Qw;
            // (the candidate buffers are laid out for KU_KC slots per CTA; unused slots stay empty)
            float* ds = cand_s + w * wCandS + ((long long)jq * gridDim.x + blockIdx.x) * KU_KC;
            int* di = cand_i + w * wCandI + ((long long)jq * gridDim.x + blockIdx.x) * KU_KC;
#pragma unroll
            for (int j = 0; j < KU_KC; ++j) { ds[j] = j < KC ? cs[j < KC ? j : 0] : FLT_MAX; di[j] = j < KC ? ci[j < KC ? j : 0] : -1; }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(KU_TMEM_COLS)) : "memory");
}

// exact fp32 squared distance of one index row, in THE summation order of this engine (kernels_knn.cu knn_scan_kernel,
// oracle/knn.py l2_f32_ordered): lane l takes the float4 chunks l, l+32, ... in ascending order, four fmaf per chunk,
// then the xor-butterfly 16, 8, 4, 2, 1.  Every lane returns the total.
__device__ __forceinline__ float exact_d2(const float* __restrict__ x_sm, const float* __restrict__ row, int C4, int lane) {
    float a = 0.f;
    for (int c4 = lane; c4 < C4; c4 += 32) {
        const float4 y = __ldg(reinterpret_cast<const float4*>(row) + c4);
        const float4 x = reinterpret_cast<const float4*>(x_sm)[c4];
        const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
        a = fmaf(d0, d0, a); a = fmaf(d1, d1, a); a = fmaf(d2, d2, a); a = fmaf(d3, d3, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    return a;
}

// One CTA per query: global top-KC by approximate score out of parts x KC candidates, exact re-rank, guard, fallback.
__global__ void __launch_bounds__(256)
knn_rerank_kernel(const float* __restrict__ index, int N, int C, const float* __restrict__ queries, long long ldq,
                  const float* __restrict__ cand_s, const int* __restrict__ cand_i, int parts, int* __restrict__ idx,
                  float* __restrict__ d2, int k, int kc, float ymax2, int* __restrict__ fallbacks,
                  long long wQ, long long wCandS, long long wCandI, long long wIdx, long long wD2) {
    pdl_enter();
    queries += blockIdx.z * wQ; cand_s += blockIdx.z * wCandS; cand_i += blockIdx.z * wCandI; idx += blockIdx.z * wIdx; d2 += blockIdx.z * wD2;
    extern __shared__ __align__(16) float rr_sm[];
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = parts * KU_KC, C4 = C >> 2;
    float* xs = rr_sm;                                   // [C]
    float* sd = xs + C;                                  // [M] approximate scores
    int* si = reinterpret_cast<int*>(sd + M);            // [M]
    __shared__ float sel_s[KU_KC]; __shared__ int sel_i[KU_KC]; __shared__ float sel_d[KU_KC];
    __shared__ float wr_s[8]; __shared__ int wr_i[8]; __shared__ int wr_p[8];
    __shared__ float s_xnorm, s_aout; __shared__ int s_fallback;
    for (int c = tid; c < C; c += 256) xs[c] = queries[(long long)q * ldq + c];
    for (int m = tid; m < M; m += 256) { sd[m] = cand_s[(long long)q * M + m]; si[m] = cand_i[(long long)q * M + m]; }
    __syncthreads();
    if (warp == 0) {
        float a = 0.f;
        for (int c = lane; c < C; c += 32) a = fmaf(xs[c], xs[c], a);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) s_xnorm = a;
    }
    // a_out: every row that is in NO candidate list has an approximate score >= the worst kept score of its CTA's list
    // (a list with an empty slot kept every row its CTA saw: no bound from it)
    {
        float lo = FLT_MAX;
        for (int p = tid; p < parts; p += 256) {
            float mx = -FLT_MAX;
            for (int j = 0; j < kc; ++j) mx = si[p * KU_KC + j] < 0 ? FLT_MAX : fmaxf(mx, sd[p * KU_KC + j]);
            lo = fminf(lo, mx);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        if (lane == 0) wr_s[warp] = lo;
        __syncthreads();
        if (tid == 0) { for (int w = 1; w < 8; ++w) lo = fminf(lo, wr_s[w]); s_aout = lo; }
        __syncthreads();
    }
    // KC rounds of "smallest (score, row) not yet taken"
    int nsel = 0;
    for (int r = 0; r < KU_KC; ++r) {
        float bs = FLT_MAX; int bi = INT_MAX, bp = -1;
        for (int m = tid; m < M; m += 256) {
            const int id = si[m];
            if (id < 0) continue;
            const float s = sd[m];
            if (s < bs || (s == bs && id < bi)) { bs = s; bi = id; bp = m; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, bs, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const int op = __shfl_xor_sync(0xffffffffu, bp, o);
            if (os < bs || (os == bs && oi < bi)) { bs = os; bi = oi; bp = op; }
        }
        if (lane == 0) { wr_s[warp] = bs; wr_i[warp] = bi; wr_p[warp] = bp; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < 8; ++w) if (wr_s[w] < bs || (wr_s[w] == bs && wr_i[w] < bi)) { bs = wr_s[w]; bi = wr_i[w]; bp = wr_p[w]; }
            sel_s[r] = bs; sel_i[r] = bi == INT_MAX ? -1 : bi;
            if (bp >= 0) si[bp] = -1;   // taken
        }
        __syncthreads();
        if (sel_i[r] >= 0) nsel = r + 1;
    }
    // exact distances of the selected rows (one warp per row)
    for (int r = warp; r < nsel; r += 8) {
        const float d = exact_d2(xs, index + (long long)sel_i[r] * C, C4, lane);
        if (lane == 0) sel_d[r] = d;
    }
    __syncthreads();
    if (tid == 0) {
        // insertion sort by (distance, row)
        for (int a = 1; a < nsel; ++a) {
            const float dv = sel_d[a]; const int iv = sel_i[a]; const float sv = sel_s[a];
            int b = a - 1;
            while (b >= 0 && (sel_d[b] > dv || (sel_d[b] == dv && sel_i[b] > iv))) { sel_d[b + 1] = sel_d[b]; sel_i[b + 1] = sel_i[b]; sel_s[b + 1] = sel_s[b]; --b; }
            sel_d[b + 1] = dv; sel_i[b + 1] = iv; sel_s[b + 1] = sv;
        }
        // guard: a row that was not re-evaluated is either in no candidate list (approximate score >= a_out) or a candidate
        // that lost the selection (score >= the largest selected one); + |x|^2 turns scores into approximate distances
        float a_kc = -FLT_MAX;
        for (int r = 0; r < nsel; ++r) a_kc = fmaxf(a_kc, sel_s[r]);
        if (nsel < KU_KC) a_kc = FLT_MAX;      // every candidate was re-evaluated
        a_kc = fminf(a_kc, s_aout);
        const float eps = 4.0f * sqrtf(s_xnorm * ymax2) * (4.76837158203125e-7f + float(C) * 5.9604644775390625e-8f) + 1e-6f * (s_xnorm + ymax2);
        const bool all_rows = a_kc == FLT_MAX;   // every row of the index was a candidate and every candidate was re-evaluated
        const bool ok = all_rows || (nsel >= k && (a_kc + s_xnorm) - eps > sel_d[k - 1]);
        s_fallback = ok ? 0 : 1;
        if (ok) for (int r = 0; r < k; ++r) { idx[q * k + r] = r < nsel ? sel_i[r] : -1; d2[q * k + r] = r < nsel ? sel_d[r] : FLT_MAX; }
        else if (fallbacks) atomicAdd(fallbacks, 1);
    }
    __syncthreads();
    if (!s_fallback) return;
    // ---- fallback: exact scan of every row for this query (same distance order), top-k by (distance, row) ----
    float ld[16]; int li[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { ld[j] = FLT_MAX; li[j] = INT_MAX; }
    for (int n = warp; n < N; n += 8) {
        float cd = exact_d2(xs, index + (long long)n * C, C4, lane);
        int cid = n;
        if (cd < ld[15] || (cd == ld[15] && cid < li[15])) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (cd < ld[j] || (cd == ld[j] && cid < li[j])) { const float td = ld[j]; const int ti = li[j]; ld[j] = cd; li[j] = cid; cd = td; cid = ti; }
        }
    }
    __syncthreads();
    __shared__ float fb_d[8 * 16]; __shared__ int fb_i[8 * 16];
    if (lane == 0) for (int j = 0; j < 16; ++j) { fb_d[warp * 16 + j] = ld[j]; fb_i[warp * 16 + j] = li[j]; }
    __syncthreads();
    if (tid == 0) {
        for (int r = 0; r < k; ++r) {
            float bd = FLT_MAX; int bi = INT_MAX, bp = -1;
            for (int m = 0; m < 128; ++m) if (fb_i[m] != INT_MAX && (fb_d[m] < bd || (fb_d[m] == bd && fb_i[m] < bi))) { bd = fb_d[m]; bi = fb_i[m]; bp = m; }
            idx[q * k + r] = bp >= 0 ? bi : -1; d2[q * k + r] = bd;
            if (bp >= 0) fb_i[bp] = INT_MAX;
        }
    }
}

// index rows -> tile images [k-block][y_hi rows | y_lo' rows] (swizzled as the MMA reads them) + |y|^2 (fp32, fixed
// order: one warp per row, lane-strided fmaf, xor-butterfly) behind each tile
__global__ void __launch_bounds__(256)
knn_planes_kernel(const float* __restrict__ index, int N, int C, uint8_t* __restrict__ planes, int n_pad, unsigned int* __restrict__ ymax_bits) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long row = (long long)blockIdx.x * 8 + warp;
    if (row >= n_pad) return;
    const long long tile_bytes = (long long)(C / 64) * 2 * KU_BN * 128 + KU_BN * 4;
    uint8_t* tile = planes + (row / KU_BN) * tile_bytes;
    const int r = int(row % KU_BN);
    float* yn = reinterpret_cast<float*>(tile + (long long)(C / 64) * 2 * KU_BN * 128) + r;
    if (row >= N) { if (lane == 0) *yn = FLT_MAX; return; }   // padding rows of the last tile (planes zeroed) can never be selected
    float a = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float v = index[row * C + c];
        const __half h = __float2half_rn(v);
        const int kb = c >> 6, ch = (c & 63) >> 3, e = c & 7;
        const int off = kb * (2 * KU_BN * 128) + (r >> 3) * 1024 + (r & 7) * 128 + ((ch ^ (r & 7)) << 4) + e * 2;
        *reinterpret_cast<unsigned short*>(tile + off) = __half_as_ushort(h);
        *reinterpret_cast<unsigned short*>(tile + off + KU_BN * 128) = __half_as_ushort(__float2half_rn((v - __half2float(h)) * KU_LO_SCALE));
        a = fmaf(v, v, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) { *yn = a; atomicMax(ymax_bits, __float_as_uint(a)); }   // non-negative floats order like their bit patterns
}

template <int KB, int KC>
bool scan_umma_launch(const KnnScanOp& o, const DeviceBases& B, int q0, int nq, cudaStream_t s) {
    using Cfg = KuCfg<KB>;
    auto kern = knn_umma_scan_kernel<KB, KC>;
    static unsigned long long attr = 0;
    if (first_time_on_device(attr)) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    const uint8_t* planes = B.b[SP_IDX] + o.planes_off;
    return launch_k(kern, dim3(o.parts), dim3(KU_THREADS), size_t(Cfg::SMEM_BYTES), s, planes, B.p<float>(o.queries), o.ldq, nq, q0, o.Q,
                    B.ws(o.queries), o.N, B.p<float>(o.cand_d), B.p<int>(o.cand_i), B.ws(o.cand_d), B.ws(o.cand_i)) == cudaSuccess;
}

}  // namespace

void knn_umma_debug_read(long long* out) { cudaMemcpyFromSymbol(out, g_ku_stamp, sizeof(long long) * 8 * 32); }

int launch_knn_scan_umma(const KnnScanOp& o, const DeviceBases& B, cudaStream_t s) {
    const int Qall = o.Q * B.nb;
    int launches = 0;
    for (int q0 = 0; q0 < Qall; q0 += 128) {
        const int nq = Qall - q0 < 128 ? Qall - q0 : 128;
        bool ok = false;
        // candidates per query and CTA: twice the requested neighbours (the guard of the re-rank needs slack above rank k)
        const bool kc8 = o.k <= 4;
        switch (o.C / 64) {
            case 1: ok = kc8 ? scan_umma_launch<1, 8>(o, B, q0, nq, s) : scan_umma_launch<1, 16>(o, B, q0, nq, s); break;
            case 2: ok = kc8 ? scan_umma_launch<2, 8>(o, B, q0, nq, s) : scan_umma_launch<2, 16>(o, B, q0, nq, s); break;
            case 3: ok = kc8 ? scan_umma_launch<3, 8>(o, B, q0, nq, s) : scan_umma_launch<3, 16>(o, B, q0, nq, s); break;
            case 4: ok = kc8 ? scan_umma_launch<4, 8>(o, B, q0, nq, s) : scan_umma_launch<4, 16>(o, B, q0, nq, s); break;
            default: break;
        }
        if (!ok) return -1;
        ++launches;
    }
    return launches;
}

int launch_knn_rerank(const KnnSelectOp& o, const DeviceBases& B, cudaStream_t s) {
    static unsigned long long attr = 0;
    if (first_time_on_device(attr)) cudaFuncSetAttribute(knn_rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int M = o.parts * KU_KC;
    const size_t smem = size_t(o.C) * 4 + size_t(M) * 8;
    launch_k(knn_rerank_kernel, dim3(o.Q, 1, B.nb), dim3(256), smem, s, B.p<float>(o.index), o.N, o.C, B.p<float>(o.queries), o.ldq, B.p<float>(o.cand_d),
             B.p<int>(o.cand_i), o.parts, B.p<int>(o.idx), B.p<float>(o.d2), o.k, o.k <= 4 ? 8 : KU_KC, o.ymax2, reinterpret_cast<int*>(B.b[SP_IDX] + o.fallback_off),
             B.ws(o.queries), B.ws(o.cand_d), B.ws(o.cand_i), B.ws(o.idx), B.ws(o.d2));
    return 1;
}

// builds the tile images behind the fp32 rows of an index allocation (ops.h knn_umma_planes_bytes); returns max |y|^2
float launch_knn_build_planes(const float* index, int N, int C, uint8_t* planes, cudaStream_t s) {
    const int n_pad = knn_umma_rows_padded(N);
    unsigned int* tail = reinterpret_cast<unsigned int*>(planes + knn_umma_counters_off(N, C));   // [0] fallback counter, [1] max |y|^2 bits
    cudaMemsetAsync(planes, 0, size_t(knn_umma_planes_bytes(N, C)), s);
    knn_planes_kernel<<<(n_pad + 7) / 8, 256, 0, s>>>(index, N, C, planes, n_pad, tail + 1);
    unsigned int bits = 0;
    cudaMemcpyAsync(&bits, tail + 1, 4, cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    float v; memcpy(&v, &bits, 4);
    return v;
}

}  // namespace rvc
