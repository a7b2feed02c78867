// kernels_umma.cu - tcgen05 / TMEM / TMA implicit GEMM for the dense contractions (sm_100a).
//
// Serves the FLOP-heavy GEMMs of the path (ContentVec conv stem + transformer, pos-conv, HiFiGAN
// convs / transposed convs): CUDA-core fp32 tops out near 25 TFLOP/s on B200, the 5th-gen tensor
// cores do not.  Parity demands fp32-grade results (F0/kNN bit-exact decisions downstream, 1e-3
// waveform RMS), so the kernel runs an error-compensated split:
//   default, PASSES = 16 (2-term FP16 split, kind::f16, fp32 accumulation in TMEM):
//     x = hi + lo' * 2^-11,  hi = half(x),  lo' = half((x - hi) * 2^11)
//     D0 += A_hi.W_hi ;  D1 += A_lo'.W_hi + A_hi.W_lo' ;  D = D0 + 2^-11 D1      (dropped term ~2^-22)
//     A_hi x [W_hi ; W_lo'] is ONE instruction with N = 2 BN (the planes of a stage are adjacent).
//   RVC_UMMA_F16=0, PASSES = 3 (3xTF32): hi = fp32 word (the tensor core ignores 13 mantissa bits), lo = x - hi.
// The weight planes are built once per model and live in HBM next to the fp32 weights.
//
// Structure (one 128 x BN output tile per CTA, 384 threads = 12 warps):
//   warps 0, 11 : TMA producers (alternate k-blocks): cp.async.bulk.tensor 3-D loads of the two weight
//                 planes into a 6-8 stage ring, mbarrier expect_tx;
//   warp 1      : TMEM allocation + single-thread tcgen05.mma issue (M = 128, K = 16 halves per instruction),
//                 tcgen05.commit releases a stage / publishes the accumulator;
//   warps 2-9   : A path (FP16 mode): the segmented / overlapping im2col rows of ops.h are read from L2
//                 (ld.global.cg, three k-blocks ahead in registers), split into the two fp16 planes and written
//                 in the SWIZZLE_64B layout the descriptors expect (fence.proxy.async, then the stage barrier);
//   warps 2-5   : then drain TMEM (tcgen05.ld 32 lanes x 16 columns) into a padded staging tile;
//   all warps   : epilogue - split-K partial tiles of a thread-block cluster are exchanged through an L2 scratch
//                 (one cluster barrier) and summed in z order; bias / activation / residual / masks / scatter modes.
// Batched plans (several windows per launch): grid.z = windows x groups x split-K, weights shared between windows.
#include <cooperative_groups.h>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>

#include "gemm_common.cuh"
#include "gemm_sched.h"
#include "launch.h"
#include "pdl.cuh"

namespace cg = cooperative_groups;

namespace rvc {

namespace {

using namespace gemmk;

constexpr int UM_BM = 128;
constexpr int UM_BK = 32;                 // floats per k-block = 128 B = one swizzle row
constexpr int UM_A_BYTES = UM_BM * 128;   // 16 KB
constexpr int UM_THREADS = 384;           // 12 warps: TMA, MMA, 4 convert/TMEM-drain warps, 6 extra store warps
constexpr int UM_WARPS = UM_THREADS / 32;

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

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, void* smem_dst, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// L2 prefetch of one box of a 3-D tensor map (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, void* smem_dst, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= uint64_t((smem_addr & 0x3FFFFu) >> 4);   // start address            bits [0,14)
    d |= uint64_t(1) << 16;                       // leading byte offset (unused for swizzled K-major)
    d |= uint64_t(1024 >> 4) << 32;               // stride byte offset: 8 rows x 128 B   bits [32,46)
    d |= uint64_t(1) << 46;                       // descriptor version (Blackwell)
    d |= uint64_t(2) << 61;                       // layout type SWIZZLE_128B
    return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Activations: the cheap ones inline, the transcendental bodies (erff / tanhf / expf) out of line - inlined eight times
// per row they made the epilogue loop several thousand instructions long and every pass streamed it through the
// instruction cache (~950 cycles per row, measured)
__device__ __noinline__ float umma_act_slow(int act, float v) { return apply_act(act, v); }
__device__ __forceinline__ float umma_act(int act, float v) {
    if (act == ACT_NONE) return v;
    if (act == ACT_LRELU01) return v > 0.0f ? v : 0.1f * v;
    if (act == ACT_RELU) return fmaxf(v, 0.0f);
    return umma_act_slow(act, v);
}

// two 16-column TMEM loads (main + correction accumulator) behind ONE wait
__device__ __forceinline__ void tmem_ld16x2(uint32_t taddr0, uint32_t taddr1, float* v, float* c) {
    uint32_t r[16], q[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr0));
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]), "=r"(q[9]),
          "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
        : "r"(taddr1));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) { v[i] = __uint_as_float(r[i]); c[i] = __uint_as_float(q[i]); }
}

__device__ long long g_umma_dbg[16];
__device__ long long g_umma_dbg2[5 * 16];   // [event][k-block < 16] of CTA (0,0,0): producer empty-ok / tma-issued, MMA full-ok / conv-ok / issued
#ifdef RVC_UMMA_STAMPS   // per-k-block stamps cost ~100 cycles per iteration of the stamped thread: build with -DRVC_UMMA_STAMPS to use them
#define UMMA_DBG2(e, i) do { if ((i) < 16 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_umma_dbg2[(e) * 16 + (i)] = clock64(); } while (0)
#else
#define UMMA_DBG2(e, i) do { } while (0)
#endif
// 1 = store the truncated A_hi back (explicit); 0 = leave the fp32 tile as delivered by TMA and rely on the tensor
// core ignoring the 13 low mantissa bits of a tf32 operand (saves a third of the converter's shared-memory writes)
// Timing experiments only (results are garbage): RVC_UMMA_DBG_SKIP=1 no A loads, 2 no W loads, 3 no conversion, 4 no
// partial loads in the epilogue, 5 no C stores.  Compiled in only with -DRVC_UMMA_STAMPS; otherwise the constant 0.
#ifdef RVC_UMMA_STAMPS
__device__ int g_dev_dbg_skip_v = 0;
#define g_dev_dbg_skip g_dev_dbg_skip_v
#else
#define g_dev_dbg_skip 0
#endif
__device__ int g_dev_w_prefetch = 0;   // RVC_UMMA_WPREFETCH=1: up-front L2 prefetch of the CTA's weight slice (measured: no gain)
__device__ int g_dev_write_hi = 0;   // RVC_UMMA_WRITE_HI=1 restores the explicit store
#define UMMA_DBG(i) do { if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_umma_dbg[i] = clock64(); } while (0)

// PASSES: 3 = 3xTF32 (A split in smem into tf32 hi / lo fp32 words), 1 = single tf32 pass (debug),
//         16 = 2-term FP16 split (kernel header): same 11+11 mantissa bits per operand as 3xTF32, half the
//              operand bytes in shared memory and in the weight stream.
template <int BN, int PASSES>
struct UmmaCfg {
    static constexpr bool F16 = PASSES == 16;
    static constexpr int W_BYTES = F16 ? BN * 64 : BN * 128;                 // one weight plane of a k-block
    static constexpr int A16_BYTES = UM_BM * 64;                             // one fp16 plane of the A k-block
    // F16: A never lands in shared memory as fp32 - the converter warps read it from L2 into registers (below)
    static constexpr int STAGE_BYTES = F16 ? (2 * A16_BYTES + 2 * W_BYTES)
                                           : (UM_A_BYTES + W_BYTES) * (PASSES == 3 ? 2 : 1);
    static constexpr int MAX_STAGES = F16 ? 8 : 6;
    static constexpr int STAGES = (196 * 1024 / STAGE_BYTES) > MAX_STAGES ? MAX_STAGES : (196 * 1024 / STAGE_BYTES);
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;  // + alignment slack
    static constexpr int ACC_COLS = F16 ? 2 * BN : BN;              // F16: main + correction accumulator
    static constexpr int TMEM_COLS = ACC_COLS < 32 ? 32 : ACC_COLS;
};

// K-major SWIZZLE_64B descriptor (rows of 32 halves = 64 B, 8-row atoms of 512 B)
__device__ __forceinline__ uint64_t umma_desc64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= uint64_t((smem_addr & 0x3FFFFu) >> 4);
    d |= uint64_t(1) << 16;
    d |= uint64_t(512 >> 4) << 32;
    d |= uint64_t(1) << 46;
    d |= uint64_t(4) << 61;                       // layout type SWIZZLE_64B
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
constexpr float F16_LO_SCALE = 2048.0f;   // lo' = (x - float(half(x))) * 2^11 keeps the residual in fp16's normal range

template <int BN, int PASSES>
__global__ void __launch_bounds__(UM_THREADS, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                 const __grid_constant__ CUtensorMap tmWlo, GemmParams p) {
    using Cfg = UmmaCfg<BN, PASSES>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int CT_LD = BN + 4;  // staging tile row stride (words): float4-aligned, conflict-free
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[STAGES], bar_conv[STAGES], bar_empty[STAGES], bar_acc;
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    auto stageA = [&](int s) { return smem + s * Cfg::STAGE_BYTES; };
    auto stageAlo = [&](int s) { return smem + s * Cfg::STAGE_BYTES + UM_A_BYTES; };
    auto stageW = [&](int s) { return smem + s * Cfg::STAGE_BYTES + (Cfg::F16 ? 2 * Cfg::A16_BYTES : UM_A_BYTES * (PASSES == 3 ? 2 : 1)); };
    auto stageWlo = [&](int s) { return stageW(s) + Cfg::W_BYTES; };
    auto stageAhi16 = [&](int s) { return smem + s * Cfg::STAGE_BYTES; };                                 // F16 mode only
    auto stageAlo16 = [&](int s) { return smem + s * Cfg::STAGE_BYTES + Cfg::A16_BYTES; };

    const int m0 = blockIdx.y * UM_BM, n0 = blockIdx.x * BN;
    const int z = blockIdx.z % p.splitk, bzw = blockIdx.z / p.splitk;
    const int win = bzw / p.batch, bz = bzw - win * p.batch;   // window of a batched plan, group (the op's own batch index)
    const int nkb_total = (p.K + UM_BK - 1) / UM_BK;
    const int kb0 = z * p.kt_per_split, kb1 = min(nkb_total, kb0 + p.kt_per_split);
    const int nkb = max(0, kb1 - kb0);

    pdl_launch_dependents();
    if (tid == 0) UMMA_DBG(0);
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
        if (PASSES != 1) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmWlo)) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&bar_full[s], Cfg::F16 ? 9 : 1); mbar_init(&bar_conv[s], 8); mbar_init(&bar_empty[s], 1); }
        mbar_init(&bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(uint32_t(Cfg::TMEM_COLS)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = s_tmem;
    pdl_wait();  // barrier init, TMEM allocation and descriptor prefetch overlapped the previous kernel
    if (tid == 0) UMMA_DBG(1);

    if (warp == 0 || warp == UM_WARPS - 1) {
        // ===== TMA producers: two warps (warp 0 and an otherwise idle store warp) alternate k-blocks; in each, lanes
        // 0 / 1 / 2 issue one of the three loads.  One producer iteration (barrier wait + expect_tx + issue) costs its
        // thread ~1250 cycles whatever it loads (measured), which paced the whole k loop =====
        const int pid = warp == 0 ? 0 : 1;
        for (int i = pid; i < nkb; i += 2) {
            const int s = i % STAGES, ph = (i / STAGES) & 1;
            if (lane == 0) {
                mbar_wait(&bar_empty[s], ph ^ 1);
                UMMA_DBG2(0, i);
                mbar_expect_tx(&bar_full[s], ((Cfg::F16 || g_dev_dbg_skip == 1) ? 0 : UM_A_BYTES) + (g_dev_dbg_skip == 2 ? 0 : Cfg::W_BYTES * (PASSES != 1 ? 2 : 1)));
            }
            __syncwarp();
            const int kk = (kb0 + i) * UM_BK;
            if (lane == 0) {
                const int seg = kk / p.seg_len, within = kk - seg * p.seg_len;
                if (!Cfg::F16 && g_dev_dbg_skip != 1) tma_load_4d(&tmA, stageA(s), &bar_full[s], within, seg, m0, bz);   // (3xTF32 path: nb == 1 only, see launch_gemm_umma)
                UMMA_DBG2(1, i);
            } else if (g_dev_dbg_skip == 2) {
            } else if (lane == 1) {
                tma_load_3d(&tmW, stageW(s), &bar_full[s], kk, n0, bz);
            } else if (lane == 2 && PASSES != 1) {
                tma_load_3d(&tmWlo, stageWlo(s), &bar_full[s], kk, n0, bz);
            }
        }
        if (lane == 0 && pid == 0) UMMA_DBG(2);
    } else if (false) {
        // ===== weight prefetch (a store warp that is idle during the k loop): the weight slice of this CTA is cold in L2
        // every window (850 MB of weights stream through a 126 MB L2) and does not depend on the previous kernel, so it
        // is requested now, one k-block per lane: the stage round trip (TMA -> convert -> MMA -> release) that bounds
        // the k loop then pays an L2 hit instead of an HBM miss =====
        if (g_dev_w_prefetch) {
            for (int i = STAGES + lane; i < nkb; i += 32) {
                const int kk = (kb0 + i) * UM_BK;
                tma_prefetch_3d(&tmW, kk, n0, bz);
                if (PASSES != 1) tma_prefetch_3d(&tmWlo, kk, n0, bz);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(BN >> 3) << 17) | (uint32_t(UM_BM >> 4) << 24);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES, ph = (i / STAGES) & 1;
                mbar_wait(&bar_full[s], ph);
                if (i == 0) UMMA_DBG(3);
                UMMA_DBG2(2, i);
                if (PASSES != 1 && !Cfg::F16) mbar_wait(&bar_conv[s], ph);
                if (i == 0) UMMA_DBG(4);
                UMMA_DBG2(3, i);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if constexpr (Cfg::F16) {
                    // D0 += Ahi.Whi ; D1 += Alo'.Whi + Ahi.Wlo'   (K = 16 halves = 32 B per instruction)
                    const uint32_t idesc16 = (1u << 4) | (uint32_t(BN >> 3) << 17) | (uint32_t(UM_BM >> 4) << 24);
                    const uint32_t a_hi = smem_u32(stageAhi16(s)), a_lo = smem_u32(stageAlo16(s));
                    const uint32_t w_hi = smem_u32(stageW(s)), w_lo = smem_u32(stageWlo(s));
#pragma unroll
                    for (int ks = 0; ks < UM_BK / 16; ++ks) {
                        const uint32_t off = ks * 32;
                        const uint32_t first = (i > 0 || ks > 0) ? 1u : 0u;
                        if constexpr (BN <= 128) {
                            // the two weight planes of a stage are adjacent in shared memory: ONE instruction with N = 2 BN
                            // gives A_hi.W_hi -> D0 (columns 0..BN) and A_hi.W_lo' -> D1 (columns BN..2BN); the issue cost
                            // of a tcgen05.mma is the same whatever N <= 256, so 2 instructions per k-step instead of 3
                            constexpr uint32_t idesc2n = (1u << 4) | (uint32_t((2 * BN) >> 3) << 17) | (uint32_t(UM_BM >> 4) << 24);
                            umma_f16(tmem_base, umma_desc64(a_hi + off), umma_desc64(w_hi + off), idesc2n, first);
                            umma_f16(tmem_base + BN, umma_desc64(a_lo + off), umma_desc64(w_hi + off), idesc16, 1u);
                        } else {
                            umma_f16(tmem_base, umma_desc64(a_hi + off), umma_desc64(w_hi + off), idesc16, first);
                            umma_f16(tmem_base + BN, umma_desc64(a_lo + off), umma_desc64(w_hi + off), idesc16, first);
                            umma_f16(tmem_base + BN, umma_desc64(a_hi + off), umma_desc64(w_lo + off), idesc16, 1u);
                        }
                    }
                    umma_commit(&bar_empty[s]);
                    UMMA_DBG2(4, i);
                    continue;
                }
                const uint32_t a_hi = smem_u32(stageA(s)), a_lo = smem_u32(stageAlo(s));
                const uint32_t w_hi = smem_u32(stageW(s)), w_lo = smem_u32(stageWlo(s));
#pragma unroll
                for (int ks = 0; ks < UM_BK / 8; ++ks) {
                    const uint32_t off = ks * 32;  // 8 tf32 = 32 B along K inside the swizzle atom
                    umma_tf32(tmem_base, umma_desc(a_hi + off), umma_desc(w_hi + off), idesc, (i > 0 || ks > 0) ? 1u : 0u);
                    if (PASSES == 3) {
                        umma_tf32(tmem_base, umma_desc(a_lo + off), umma_desc(w_hi + off), idesc, 1u);
                        umma_tf32(tmem_base, umma_desc(a_hi + off), umma_desc(w_lo + off), idesc, 1u);
                    }
                }
                umma_commit(&bar_empty[s]);  // smem stage reusable once these MMAs have read it
            }
            if (nkb > 0) umma_commit(&bar_acc); else mbar_arrive(&bar_acc);
            UMMA_DBG(5);
        }
    } else if (warp < 10) {
        // ===== converter warps 2-9 (two per stage); warps 2-5 then drain TMEM =====
        const int ct = tid - 64;
        if constexpr (Cfg::F16) {
            // fp32 A rows (L2) -> registers -> two fp16 planes in shared memory (SWIZZLE_64B): hi = half(x),
            // lo' = half((x - hi) * 2^11).  A is NOT staged through shared memory as fp32: the eight converter warps read
            // it straight from L2, one k-block ahead (the loads of k-block i+1 are in flight while i is converted), so a
            // stage holds operands only (16 KB + BN * 128 B -> 8 stages at BN = 64) and the TMA engine moves weights only.
            // One task = one 16-byte chunk of the planes = 8 consecutive k of one row; 512 tasks per k-block, 2 per thread.
            // All 8 warps work on the SAME (oldest) stage: the k loop is bound by the round trip of a stage, not by
            // converter throughput, so the conversion must be short.
            const float* __restrict__ Ab = p.A + bz * p.sA + win * p.wA;
            const float* rowp[2]; int within[2]; int dsto[2];
            const int seg_len = p.seg_len;                      // == K when the rows are contiguous (launch_umma_cfg)
            const long long seg_wrap = p.seg_stride - p.seg_len;
#pragma unroll
            for (int t0 = 0; t0 < 2; ++t0) {
                const int t = t0 * 256 + ct, r = t >> 2, q = t & 3;
                const int m = m0 + r;
                const int kk = kb0 * UM_BK + q * 8;
                const int seg = kk / seg_len;
                within[t0] = kk - seg * seg_len;
                rowp[t0] = (m < p.M) ? Ab + (long long)m * p.lda + (long long)seg * p.seg_stride + within[t0] : nullptr;
                dsto[t0] = r * 64 + ((q ^ ((r >> 1) & 3)) << 4);
            }
            constexpr int PF = 3;                         // k-blocks of A in flight per thread (L2 latency >> conversion time)
            float4 buf[PF][2][2];
            auto fetch = [&](float4 (*dst)[2], int i) {   // loads of k-block i (relative), then advance the cursors
#pragma unroll
                for (int t0 = 0; t0 < 2; ++t0) {
                    const bool ok = rowp[t0] && i < nkb && ((kb0 + i) * UM_BK + (ct & 3) * 8) < p.K;
                    dst[t0][0] = ok ? __ldcg(reinterpret_cast<const float4*>(rowp[t0])) : make_float4(0.f, 0.f, 0.f, 0.f);
                    dst[t0][1] = ok ? __ldcg(reinterpret_cast<const float4*>(rowp[t0]) + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (rowp[t0]) {
                        rowp[t0] += UM_BK; within[t0] += UM_BK;
                        if (within[t0] >= seg_len) { within[t0] -= seg_len; rowp[t0] += seg_wrap; }
                    }
                }
            };
#pragma unroll
            for (int j = 0; j < PF - 1; ++j) fetch(buf[j], j);
#pragma unroll 1
            for (int i0 = 0; i0 < nkb; i0 += PF) {
#pragma unroll
            for (int jj = 0; jj < PF; ++jj) {
                const int i = i0 + jj;
                if (i >= nkb) break;
                float4 (*cur)[2] = buf[jj];
                const int s = i % STAGES, ph = (i / STAGES) & 1;
                fetch(buf[(jj + PF - 1) % PF], i + PF - 1);
                mbar_wait(&bar_empty[s], ph ^ 1);   // the MMAs that read this stage's planes have completed
                uint8_t* h16 = stageAhi16(s);
                uint8_t* l16 = stageAlo16(s);
#pragma unroll
                for (int t0 = 0; t0 < 2; ++t0) {
                    if (g_dev_dbg_skip == 3) break;
                    const float4 v0 = cur[t0][0], v1 = cur[t0][1];
                    const float x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                    uint32_t hw[4], lw[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        // packed conversions (one F2FP per pair)
                        const __half2 h2 = __floats2half2_rn(x[2 * e], x[2 * e + 1]);
                        const float2 hf = __half22float2(h2);
                        const __half2 l2 = __floats2half2_rn((x[2 * e] - hf.x) * F16_LO_SCALE, (x[2 * e + 1] - hf.y) * F16_LO_SCALE);
                        hw[e] = *reinterpret_cast<const uint32_t*>(&h2);
                        lw[e] = *reinterpret_cast<const uint32_t*>(&l2);
                    }
                    *reinterpret_cast<uint4*>(h16 + dsto[t0]) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                    *reinterpret_cast<uint4*>(l16 + dsto[t0]) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> tensor-core reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_full[s]);   // one barrier per stage: weights landed (tx) + planes written (8 arrivals)
            }
            }
        } else if (PASSES == 3) {
            // hi/lo split of the A tile; up to four stages are converted concurrently
            // Ownership is by STAGE (stage s belongs to warp pair s & 3), never by k-block: a parity wait
            // may only ever be one phase ahead of its mbarrier, so every waiter must visit every phase.
            const int cwarp = warp - 2;   // 0..7: every warp converts an eighth (2 KB) of EVERY stage (see the F16 branch)
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES, ph = (i / STAGES) & 1;
                mbar_wait(&bar_full[s], ph);
                float4* a = reinterpret_cast<float4*>(stageA(s)) + cwarp * (UM_A_BYTES / 128);
                float4* lo = reinterpret_cast<float4*>(stageAlo(s)) + cwarp * (UM_A_BYTES / 128);
                {
                    constexpr int b = 0;
                    float4 v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = a[(b + j) * 32 + lane];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4 h;
                        h.x = __uint_as_float(__float_as_uint(v[j].x) & 0xFFFFE000u);
                        h.y = __uint_as_float(__float_as_uint(v[j].y) & 0xFFFFE000u);
                        h.z = __uint_as_float(__float_as_uint(v[j].z) & 0xFFFFE000u);
                        h.w = __uint_as_float(__float_as_uint(v[j].w) & 0xFFFFE000u);
                        if (g_dev_write_hi) a[(b + j) * 32 + lane] = h;
                        lo[(b + j) * 32 + lane] = make_float4(v[j].x - h.x, v[j].y - h.y, v[j].z - h.z, v[j].w - h.w);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> tensor-core reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_conv[s]);
            }
        }
        if (warp < 6) {
        if (ct == 0) UMMA_DBG(6);
        mbar_wait(&bar_acc, 0);
        if (ct == 0) UMMA_DBG(7);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // accumulator: TMEM (one row per lane) -> padded fp32 staging tile in the now idle pipeline smem
        const int q = warp & 3;                       // TMEM lane quadrant this warp may read
        const int row = q * 32 + lane;
        const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16);
        // split-K partial tiles are exchanged through L2 (per-lane scratch, ld/st.cg): pulling them out of the peers'
        // shared memory over DSMEM ran at ~8 B/clk per SM and cost two cluster barriers (measured 12-15 k cycles per GEMM)
        const bool via_l2 = p.splitk > 1 && p.scratch != nullptr;
        const long long tile_lin = ((long long)bz * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        // (the drain itself always goes to the padded smem staging tile: a lane owns a ROW, so writing global memory from
        //  here would be 32 scattered 64-byte pieces per instruction; the coalesced copy to the scratch follows below)
        float* ct_row = reinterpret_cast<float*>(smem) + row * CT_LD;
        (void)tile_lin;
        for (int c0 = 0; c0 < BN; c0 += 16) {
            float v[16];
            if (nkb > 0) {
                if constexpr (Cfg::F16) {
                    float c[16];
                    tmem_ld16x2(trow + c0, trow + BN + c0, v, c);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = fmaf(c[j], 1.0f / F16_LO_SCALE, v[j]);
                } else {
                    tmem_ld16(trow + c0, v);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(ct_row + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        if (ct == 0) UMMA_DBG(11);
        }
    }
    // ---- staged tile -> global: lanes along N (coalesced); split-K partial tiles are reduced across the
    //      cluster through distributed shared memory, each CTA finishing 128/splitk rows ----
    if (tid == 0) UMMA_DBG(8);
    const bool via_l2 = p.splitk > 1 && p.scratch != nullptr;
    // bias of this lane's columns: requested BEFORE the barrier (parameter vectors are cold in L2 every window - an HBM
    // round trip of ~900 cycles that used to sit at the head of the store loop)
    float4 bv_pre = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.vec_store && p.bias) {
        constexpr int LPR0 = BN >= 128 ? 32 : BN / 4;
        const int n_pre = n0 + (lane % LPR0) * 4;
        if (n_pre < p.N) bv_pre = __ldg(reinterpret_cast<const float4*>(p.bias + bz * p.sBias + n_pre));
    }
    if (via_l2) {
        // staged partial tile -> this CTA's slot of the L2 scratch, coalesced (all 12 warps, float4 along the columns)
        __syncthreads();
        const long long tile_lin0 = ((long long)bz * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        float* dst = p.scratch + win * p.wScratch + ((tile_lin0 * p.splitk + z) * UM_BM) * BN;
        const float* src = reinterpret_cast<const float*>(smem);
        const int live_rows = min(UM_BM, p.M - m0);   // rows past M are never read back
        for (int e = tid; e < live_rows * (BN / 4); e += UM_THREADS) {
            const int r = e / (BN / 4), c = (e - r * (BN / 4)) * 4;
            __stcg(reinterpret_cast<float4*>(dst + r * BN + c), *reinterpret_cast<const float4*>(src + r * CT_LD + c));
        }
        __threadfence();   // partial tile visible device-wide before the cluster barrier
    }
    if (p.splitk > 1) cg::this_cluster().sync(); else __syncthreads();
    if (tid == 64) UMMA_DBG(12);
    {
        const float* __restrict__ bias = p.bias ? p.bias + bz * p.sBias : nullptr;
        float* C = p.C + bz * p.sC + win * p.wC;
        float* C2 = p.C2 ? p.C2 + bz * p.sC + win * p.wC2 : nullptr;
        const float* R = p.R ? p.R + bz * p.sR + win * p.wR : nullptr;
        const float* Ct = reinterpret_cast<const float*>(smem);
        const int rows_per = UM_BM / p.splitk;
        const int r_begin = z * rows_per, r_end = r_begin + rows_per;
        const bool gate = p.act == ACT_GATE;
        const float* peers[8];
        const long long tile_lin = ((long long)bz * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        const int ldp = via_l2 ? BN : CT_LD;   // row pitch of a partial tile (global scratch / padded smem staging)
#pragma unroll
        for (int zz = 0; zz < 8; ++zz) {
            if (via_l2) peers[zz] = p.scratch + win * p.wScratch + ((tile_lin * p.splitk + (zz < p.splitk ? zz : 0)) * UM_BM) * BN;
            else peers[zz] = (p.splitk > 1 && zz < p.splitk) ? cg::this_cluster().map_shared_rank(Ct, zz) : Ct;
        }
        if (p.vec_store) {
            // fast path (plain row-major output, no gate): float4 per lane, BN/4 lanes per row
            constexpr int LPR = BN >= 128 ? 32 : BN / 4, RPI = 32 / LPR, NCH = (BN / 4) / LPR;   // BN = 256: two 128-column chunks per row
            const int sub = lane / LPR;
#pragma unroll 1
            for (int ch = 0; ch < NCH; ++ch) {
            const int c4 = (ch * LPR + lane % LPR) * 4, n = n0 + c4;
            const bool ncol = n < p.N;  // N % 4 == 0 on this path
            float4 bv = bv_pre;   // chunk 0 was requested before the barrier
            if (ch > 0 && bias && ncol) bv = __ldg(reinterpret_cast<const float4*>(bias + n));
            if (tid == 64) UMMA_DBG(14);
            // L2 path: the partial sums of up to RB rows of this warp are fetched together (the loop is bound by L2 latency,
            // one row at a time left ~4 loads in flight per warp); pre[k] = sum over z, in z order, of row k of the batch
            constexpr int RB = 3;
            float4 pre[RB];
            int pre_base = -1;   // first row of the batch held in pre[]
            for (int row = r_begin + warp * RPI + sub; row < r_end; row += UM_WARPS * RPI) {
                const int m = m0 + row;
                if (via_l2 && (pre_base < 0 || row >= pre_base + RB * UM_WARPS * RPI)) {
                    pre_base = row;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {       // z = 0..3, then 4..7
                        float4 t[RB][4];
#pragma unroll
                        for (int k = 0; k < RB; ++k) {
                            const int rk = row + k * UM_WARPS * RPI;
                            const bool okr = rk < r_end && (m0 + rk) < p.M && ncol;
#pragma unroll
                            for (int zz = 0; zz < 4; ++zz)
                                t[k][zz] = (okr && (h * 4 + zz) < p.splitk && g_dev_dbg_skip != 4) ? __ldcg(reinterpret_cast<const float4*>(peers[h * 4 + zz] + rk * BN + c4))
                                                                             : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
#pragma unroll
                        for (int k = 0; k < RB; ++k) {
                            if (h == 0) pre[k] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                            for (int zz = 0; zz < 4; ++zz) { pre[k].x += t[k][zz].x; pre[k].y += t[k][zz].y; pre[k].z += t[k][zz].z; pre[k].w += t[k][zz].w; }
                        }
                        if (p.splitk <= 4) break;
                    }
                }
                if (tid == 64 && row == r_begin + warp * RPI + sub && pre[0].x != 12345.678f) UMMA_DBG(15);
                if (m >= p.M || !ncol) continue;
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                if (via_l2) {
                    const int k = (row - pre_base) / (UM_WARPS * RPI);
                    a = k == 0 ? pre[0] : (k == 1 ? pre[1] : pre[2]);
                } else {
#pragma unroll
                    for (int zz = 0; zz < 8; ++zz) {
                        if (zz < p.splitk) {
                            const float4 t = *reinterpret_cast<const float4*>(peers[zz] + row * CT_LD + c4);
                            a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
                        }
                    }
                }
                float4 v;
                if (p.act == ACT_GELU) {   // the FFN's activation: four independent erff chains inline (ILP), not four calls
                    v.x = gelu_f(fmaf(p.alpha, a.x, bv.x)); v.y = gelu_f(fmaf(p.alpha, a.y, bv.y));
                    v.z = gelu_f(fmaf(p.alpha, a.z, bv.z)); v.w = gelu_f(fmaf(p.alpha, a.w, bv.w));
                } else {
                    v.x = umma_act(p.act, fmaf(p.alpha, a.x, bv.x)); v.y = umma_act(p.act, fmaf(p.alpha, a.y, bv.y));
                    v.z = umma_act(p.act, fmaf(p.alpha, a.z, bv.z)); v.w = umma_act(p.act, fmaf(p.alpha, a.w, bv.w));
                }
                if (R) {
                    const float4 r = *reinterpret_cast<const float4*>(R + (long long)m * p.ldr + n);
                    v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
                }
                const bool masked = p.mask_period > 0 && (m % p.mask_period) >= p.mask_valid;
                if (masked) v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (g_dev_dbg_skip != 5) *reinterpret_cast<float4*>(C + (long long)m * p.ldc + n) = v;
                if (C2) {
                    float4 w;
                    w.x = masked ? 0.f : umma_act(p.act2, v.x); w.y = masked ? 0.f : umma_act(p.act2, v.y);
                    w.z = masked ? 0.f : umma_act(p.act2, v.z); w.w = masked ? 0.f : umma_act(p.act2, v.w);
                    *reinterpret_cast<float4*>(C2 + (long long)m * p.ldc2 + n) = w;
                }
            }
            }
        } else {
            for (int row = r_begin + warp; row < r_end; row += UM_WARPS) {
                const int m = m0 + row;
                if (m >= p.M) break;
                for (int col = lane; col < BN; col += 32) {
                    const int n = n0 + col;
                    float v = 0.f, vp = 0.f;
#pragma unroll
                    for (int zz = 0; zz < 8; ++zz) {
                        if (zz < p.splitk) {
                            if (via_l2) {
                                v += __ldcg(peers[zz] + row * ldp + col);
                                if (gate) vp += __ldcg(peers[zz] + row * ldp + (col ^ 1));
                            } else {
                                v += peers[zz][row * ldp + col];
                                if (gate) vp += peers[zz][row * ldp + (col ^ 1)];
                            }
                        }
                    }
                    if (n < p.N) epilogue_elem(p, bias, C, C2, R, m, n, v, vp);
                }
            }
        }
    }
    if (tid == 64) UMMA_DBG(13);
    if (p.splitk > 1 && !via_l2) cg::this_cluster().sync();  // DSMEM path: peers may still be reading this CTA's staging tile
    if (tid == 64) UMMA_DBG(9);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) UMMA_DBG(10);
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(Cfg::TMEM_COLS)) : "memory");
    }
}

// ---- host side ------------------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

bool encode(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box,
            bool f16 = false) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    // fp32 tiles: rows of 32 floats = 128 B (SWIZZLE_128B); fp16 weight planes: rows of 32 halves = 64 B (SWIZZLE_64B)
    CUresult r = fn(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, cuuint32_t(rank), const_cast<void*>(base), dims,
                    strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, f16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <int BN, int PASSES>
bool launch_umma_cfg(const GemmOp& g, GemmParams& p, const void* w_hi, const void* w_lo, int nb, cudaStream_t s) {
    using Cfg = UmmaCfg<BN, PASSES>;
    constexpr int WB = Cfg::F16 ? 2 : 4;   // bytes per weight element of the planes
    CUtensorMap tmA, tmW, tmWlo;
    const int nseg = (g.seg_len >= g.K) ? 1 : g.K / g.seg_len;
    const int seg_len = nseg == 1 ? g.K : g.seg_len;
    const cuuint64_t big = cuuint64_t(1) << 36;  // stride of a size-1 dimension (any multiple of 16)
    {
        cuuint64_t dims[4] = {cuuint64_t(seg_len), cuuint64_t(nseg), cuuint64_t(g.M), cuuint64_t(g.batch)};
        cuuint64_t str[3] = {nseg > 1 ? cuuint64_t(g.seg_stride) * 4 : big, cuuint64_t(g.lda) * 4, g.batch > 1 ? cuuint64_t(g.sA) * 4 : big};
        cuuint32_t box[4] = {UM_BK, 1, UM_BM, 1};
        if (!encode(&tmA, p.A, 4, dims, str, box)) return false;
    }
    {
        cuuint64_t dims[3] = {cuuint64_t(g.K), cuuint64_t(g.N), cuuint64_t(g.batch)};
        cuuint64_t str[2] = {cuuint64_t(g.ldw) * WB, g.batch > 1 ? cuuint64_t(g.sW) * WB : big};
        cuuint32_t box[3] = {UM_BK, cuuint32_t(BN), 1};
        if (!encode(&tmW, w_hi, 3, dims, str, box, Cfg::F16)) return false;
        if (!encode(&tmWlo, w_lo, 3, dims, str, box, Cfg::F16)) return false;
    }
    const int nkb = (g.K + UM_BK - 1) / UM_BK;
    p.kt_per_split = (nkb + g.splitk - 1) / g.splitk;
    p.seg_len = seg_len;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    p.vec_store = (g.out_mode == OUT_PLAIN && g.act != ACT_GATE && g.N % 4 == 0 && al16(p.C) && g.ldc % 4 == 0 && g.sC % 4 == 0 &&
                   (!p.C2 || (al16(p.C2) && g.ldc2 % 4 == 0)) && (!p.R || (al16(p.R) && g.ldr % 4 == 0 && g.sR % 4 == 0)) &&
                   (!p.bias || (al16(p.bias) && g.sBias % 4 == 0))) ? 1 : 0;
    auto kern = umma_gemm_kernel<BN, PASSES>;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((g.N + BN - 1) / BN, (g.M + UM_BM - 1) / UM_BM, nb * g.batch * g.splitk);
    cfg.blockDim = dim3(UM_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;   // split-K group = one thread-block cluster along z
    attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = unsigned(g.splitk);
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = (g_use_pdl || g_pdl_op) ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 2;
    if (cudaLaunchKernelEx(&cfg, kern, tmA, tmW, tmWlo, p) != cudaSuccess) return false;
    return true;
}

int g_umma_passes = 3;
bool g_umma_f16 = true;   // 2-term FP16 split (RVC_UMMA_F16=0: 3xTF32)

}  // namespace

void init_umma_attributes() {
    cudaFuncSetAttribute(umma_gemm_kernel<128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<128, 3>::SMEM_BYTES);
    cudaFuncSetAttribute(umma_gemm_kernel<64, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<64, 3>::SMEM_BYTES);
    cudaFuncSetAttribute(umma_gemm_kernel<32, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<32, 3>::SMEM_BYTES);
    cudaFuncSetAttribute(umma_gemm_kernel<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<128, 1>::SMEM_BYTES);
    cudaFuncSetAttribute(umma_gemm_kernel<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<64, 1>::SMEM_BYTES);
    cudaFuncSetAttribute(umma_gemm_kernel<32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<32, 1>::SMEM_BYTES);
    cudaFuncSetAttribute(umma_gemm_kernel<256, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<256, 16>::SMEM_BYTES);
    cudaFuncSetAttribute(umma_gemm_kernel<128, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<128, 16>::SMEM_BYTES);
    cudaFuncSetAttribute(umma_gemm_kernel<64, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<64, 16>::SMEM_BYTES);
    cudaFuncSetAttribute(umma_gemm_kernel<32, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<32, 16>::SMEM_BYTES);
    cudaFuncSetAttribute(umma_gemm_kernel<16, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaCfg<16, 16>::SMEM_BYTES);
    { const char* f = getenv("RVC_UMMA_F16"); g_umma_f16 = !(f && f[0] == '0'); }
#ifdef RVC_UMMA_STAMPS
    { const char* w = getenv("RVC_UMMA_DBG_SKIP"); int v = w ? atoi(w) : 0; cudaMemcpyToSymbol(g_dev_dbg_skip_v, &v, sizeof(int)); }
#endif
    { const char* w = getenv("RVC_UMMA_WPREFETCH"); int v = (w && w[0] == '1') ? 1 : 0; cudaMemcpyToSymbol(g_dev_w_prefetch, &v, sizeof(int)); }
    { const char* w = getenv("RVC_UMMA_WRITE_HI"); int v = (w && w[0] == '1') ? 1 : 0; cudaMemcpyToSymbol(g_dev_write_hi, &v, sizeof(int)); }
    const char* e = getenv("RVC_UMMA_PASSES");
    if (e && e[0] == '1') g_umma_passes = 1;
}

void umma_debug_read(long long* out) { cudaMemcpyFromSymbol(out, g_umma_dbg, sizeof(long long) * 16); }
void umma_debug_read2(long long* out) { cudaMemcpyFromSymbol(out, g_umma_dbg2, sizeof(long long) * 80); }

// returns 0 when the tensor maps could not be encoded (caller falls back to the CUDA-core kernel)
int launch_gemm_umma(const GemmOp& g, const DeviceBases& B, cudaStream_t stream) {
    const int64_t hl = B.hilo_stride[g.W.space];
    if (hl <= 0) return 0;
    GemmParams p = gemmk::make_params(g, B);
    const float* w_hi = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(p.W) + hl);
    const float* w_lo = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(p.W) + 2 * hl);
    const int bn = g.sched_variant == 8 ? 256 : (g.sched_variant == 5 ? 128 : (g.sched_variant == 6 ? 64 : (g.sched_variant == 9 ? 16 : 32)));
    bool ok;
    const int64_t h16 = B.hilo16_off[g.W.space];
    const int64_t woff = reinterpret_cast<const uint8_t*>(p.W) - B.b[g.W.space];   // byte offset of W inside its fp32 arena
    // fp16 planes: TMA needs 16-byte aligned bases and row pitches -> element offsets / strides that are multiples of 8
    if (g_umma_f16 && g_umma_passes == 3 && h16 > 0 && woff % 32 == 0 && g.ldw % 8 == 0 && g.sW % 8 == 0) {
        const uint8_t* hi16 = B.b[g.W.space] + h16 + woff / 2;
        const uint8_t* lo16 = hi16 + B.hilo16_plane[g.W.space];
        ok = bn == 256 ? launch_umma_cfg<256, 16>(g, p, hi16, lo16, B.nb, stream)
           : bn == 128 ? launch_umma_cfg<128, 16>(g, p, hi16, lo16, B.nb, stream)
           : bn == 64 ? launch_umma_cfg<64, 16>(g, p, hi16, lo16, B.nb, stream)
           : bn == 32 ? launch_umma_cfg<32, 16>(g, p, hi16, lo16, B.nb, stream) : launch_umma_cfg<16, 16>(g, p, hi16, lo16, B.nb, stream);
        return ok ? 1 : 0;
    }
    if (B.nb > 1) return 0;   // the 3xTF32 / single-pass instances move A with a 4-D tensor map: one window per launch only
    if (bn == 256 || bn == 16) return 0;   // only the FP16-split kernel has 256- / 16-wide instances: the caller falls back to CUDA cores
    if (g.seg_len < g.K && g.seg_len % 32 != 0) return 0;   // odd segment lengths: FP16-split kernel only (see gemm_sched.h)
    if (g_umma_passes == 3) {
        ok = bn == 128 ? launch_umma_cfg<128, 3>(g, p, w_hi, w_lo, 1, stream)
           : bn == 64 ? launch_umma_cfg<64, 3>(g, p, w_hi, w_lo, 1, stream) : launch_umma_cfg<32, 3>(g, p, w_hi, w_lo, 1, stream);
    } else {
        ok = bn == 128 ? launch_umma_cfg<128, 1>(g, p, p.W, p.W, 1, stream)
           : bn == 64 ? launch_umma_cfg<64, 1>(g, p, p.W, p.W, 1, stream) : launch_umma_cfg<32, 1>(g, p, p.W, p.W, 1, stream);
    }
    return ok ? 1 : 0;
}

}  // namespace rvc
