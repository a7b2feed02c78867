// kernels_cvstack.cu - the ContentVec transformer stack as ONE persistent tcgen05 kernel (sm_100a).
//
// Reference call site: rvc/src/rvc.rs:81-97 (`hubert`: the encoder layers of the ContentVec graph).  At batch 1 the
// stack is 12 layers x (QKV, attention, out-proj + residual, LayerNorm, FC1 + GELU, FC2 + residual, LayerNorm) on
// T <= 128 rows: 84 dependent steps of ~1 us of work each.  As separate launches every step paid launch + prologue
// (TMEM allocation, barrier init, descriptor fetch) + a cold weight pipeline + a split-K exchange (~13 us per GEMM).
// Here G CTAs stay resident for the whole stack (cooperative launch) and walk a phase table:
//   * a grid barrier (one release atomic per CTA, acquire polling) stands where the kernel boundaries stood;
//   * activations feeding a GEMM live in HBM/L2 as two fp16 planes (hi = half(x), lo' = half((x - hi) * 2^11)) written
//     by the PRODUCING epilogue / LayerNorm / attention, so the A operand is a plain TMA load (SWIZZLE_128B, 64-wide
//     k-blocks) - no fp32 -> fp16 conversion inside the GEMM pipeline;
//   * the TMA producer runs ahead of the barrier: the weight halves of the next GEMM's first stages are already in
//     shared memory when the previous phase ends (weights do not depend on activations);
//   * one 128 x 48 output tile per CTA with the full K (QKV, FC1) or a quarter of it (out-proj, FC2: the four
//     partial tiles are summed - in z order - by the LayerNorm phase that follows anyway): no split-K exchange;
//   * same 2-term FP16 split arithmetic as kernels_umma.cu (D0 += A_hi.W_hi, D1 += A_hi.W_lo' + A_lo'.W_hi in TMEM,
//     fp32 accumulation, D = D0 + 2^-11 D1), attention and LayerNorm in exact fp32 on the CUDA cores.
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2-9 = workers (TMEM drain
// + epilogue, attention, LayerNorm).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cfloat>
#include <cstdio>
#include <vector>

#include <cstdlib>
#include <cstring>

#include "cvstack.h"
#include "gemm_common.cuh"

namespace rvc {

namespace {

using gemmk::gelu_f;

constexpr int CS_BN = CVS_BN;                 // output columns per tile
constexpr int CS_STAGES = CVS_BN > 48 ? 3 : 4;
constexpr int CS_A_BYTES = 128 * 128;         // one fp16 plane of an A k-block: 128 rows x 64 halves
constexpr int CS_W_BYTES = CS_BN * 128;       // one fp16 plane of a W k-block: 48 rows x 64 halves
constexpr int CS_STAGE_BYTES = 2 * CS_A_BYTES + 2 * CS_W_BYTES;   // 45056
constexpr int CS_SCRATCH_BYTES = CVS_SCRATCH_BYTES;   // worker scratch (attention staging: K^T / V 28.9 KB, scores 10.8 KB, q rows 6 KB at T = 111)
constexpr int CS_SMEM_BYTES = CS_STAGES * CS_STAGE_BYTES + CS_SCRATCH_BYTES + 1024;
constexpr int CS_THREADS = 320;
constexpr int CS_WORKERS = 256;
constexpr int CS_TMEM_COLS = CVS_BN > 64 ? 256 : 128;   // 2 x BN accumulator columns, power of two
constexpr float CS_LO_SCALE = 2048.0f;
constexpr int CS_ATT_ROWS = CVS_ATT_ROWS;     // query rows per attention item
constexpr int CS_D = 64;                      // head dim

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
        if (++spins > (1ll << 27)) __trap();   // a broken pipeline must fail loudly, never hang the GPU
    }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, void* smem_dst, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor: rows of 128 B, 8-row atoms of 1024 B
__device__ __forceinline__ uint64_t desc128(uint32_t smem_addr) {
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
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// main + correction accumulator, 8 columns each, behind one wait
__device__ __forceinline__ void tmem_ld8x2(uint32_t t0, uint32_t t1, float* v, float* c) {
    uint32_t r[8], q[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(t0));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]) : "r"(t1));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = __uint_as_float(r[i]); c[i] = __uint_as_float(q[i]); }
}
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

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

// every CTA has finished all phases < ph  <=>  arrivals >= ph * G
__device__ __forceinline__ void wait_grid(const unsigned int* bar, unsigned int target) {
    const long long t0 = clock64();
    while (true) {
        unsigned int v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        if (v >= target) break;
        __nanosleep(32);   // keep the polls off the arrival atomics' L2 line
        if (clock64() - t0 > (6ll << 30)) __trap();
    }
}

// two 16-bit planes of x: hi = half(x), lo' = half((x - hi) * 2^11)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h2 = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn((a - hf.x) * CS_LO_SCALE, (b - hf.y) * CS_LO_SCALE);
    hi = *reinterpret_cast<const uint32_t*>(&h2);
    lo = *reinterpret_cast<const uint32_t*>(&l2);
}

__device__ long long g_cvs_stamp[CVS_MAX_PHASES * 4];   // CTA 0: [phase][worker start, work done, arrived, grid wait over]
__device__ long long g_cvs_stamp2[CVS_MAX_PHASES * 8];  // CTA 0, GEMM phases: producer [wait over, last issue], MMA [first full, last full, last commit], worker [acc full]
#define CVS_ST2(e) do { if (blockIdx.x == 0 && ph < CVS_MAX_PHASES) g_cvs_stamp2[ph * 8 + (e)] = clock64(); } while (0)

__global__ void __launch_bounds__(CS_THREADS, 1)
cvstack_kernel(const CUtensorMap* __restrict__ maps, const CvsPhase* __restrict__ phases, int n_phases, unsigned int* bar, int T) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[CS_STAGES], bar_empty[CS_STAGES], bar_acc_full, bar_acc_empty;
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned int G = gridDim.x;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* scratch = smem + CS_STAGES * CS_STAGE_BYTES;

    if (tid == 0) {
        for (int s = 0; s < CS_STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        mbar_init(&bar_acc_full, 1); mbar_init(&bar_acc_empty, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(uint32_t(CS_TMEM_COLS)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = s_tmem;

    if (warp == 0) {
        // ===== TMA producer: lanes 0-3 of warp 0 issue one of the four loads of a k-block each (A_hi, A_lo', W_hi, W_lo':
        // an issuing thread is busy for a few hundred cycles per TMA instruction, four in a row paced the whole k loop).
        // The warp runs ahead of the grid barrier with the WEIGHT halves of a phase's first stages; the activation
        // halves follow once every CTA has finished the previous phase =====
        {
            unsigned int it = 0;   // k-blocks issued so far (ring position)
            for (int ph = 0; ph < n_phases; ++ph) {
                const CvsPhase& P = phases[ph];
                if (P.kind != CVS_GEMM) continue;
                const CUtensorMap* mymap = maps + ((lane < 2) ? P.a_map + lane : P.w_map + (lane - 2));   // lanes 0-3 only
                const int my_off = lane == 0 ? 0 : (lane == 1 ? CS_A_BYTES : (lane == 2 ? 2 * CS_A_BYTES : 2 * CS_A_BYTES + CS_W_BYTES));
                bool passed = false;
                for (int item = blockIdx.x; item < P.items; item += G) {
                    const int tn = item / P.splitk, z = item - tn * P.splitk;
                    const int n0 = tn * CS_BN, kb0 = z * P.nkb;
                    // every CTA walks the same A planes: each starts at its own k-block (the sum over k is commutative up
                    // to fp32 rounding, fixed per tile) so that the CTAs do not hit the same L2 lines in lockstep
                    const int rot = tn % P.nkb;
                    auto kx = [&](int kb_) { int k = kb_ + rot; if (k >= P.nkb) k -= P.nkb; return (kb0 + k) * 64; };
                    const int my_row = lane < 2 ? 0 : n0;
                    int kb = 0;
                    if (!passed) {
                        const int pre = min(P.nkb, CS_STAGES);
                        const unsigned int it0 = it;
                        for (; kb < pre; ++kb, ++it) {
                            const int s = it % CS_STAGES;
                            if (lane == 0) {
                                mbar_wait(&bar_empty[s], ((it / CS_STAGES) & 1) ^ 1);
                                mbar_expect_tx(&bar_full[s], CS_STAGE_BYTES);
                            }
                            __syncwarp();
                            if (lane == 2 || lane == 3) tma_load_2d(mymap, smem + s * CS_STAGE_BYTES + my_off, &bar_full[s], kx(kb), my_row);
                        }
                        if (ph > 0 && lane == 0) wait_grid(bar, (unsigned int)ph * G);
                        if (lane == 0) CVS_ST2(0);
                        __syncwarp();
                        asm volatile("fence.proxy.async;" ::: "memory");   // peers' generic-proxy stores -> this thread's TMA reads
                        passed = true;
                        if (lane < 2) {
                            for (int j = 0; j < pre; ++j) {
                                const int s = (it0 + j) % CS_STAGES;
                                tma_load_2d(mymap, smem + s * CS_STAGE_BYTES + my_off, &bar_full[s], kx(j), 0);
                            }
                        }
                    }
                    for (; kb < P.nkb; ++kb, ++it) {
                        const int s = it % CS_STAGES;
                        if (lane == 0) {
                            mbar_wait(&bar_empty[s], ((it / CS_STAGES) & 1) ^ 1);
                            mbar_expect_tx(&bar_full[s], CS_STAGE_BYTES);
                        }
                        __syncwarp();
                        if (lane < 4) tma_load_2d(mymap, smem + s * CS_STAGE_BYTES + my_off, &bar_full[s], kx(kb), my_row);
                    }
                    if (lane == 0) CVS_ST2(1);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread): per 64-wide k-block four k-steps of
        //   D[0:96)  (+)= A_hi x [W_hi ; W_lo']   (one N = 96 instruction: main accumulator | correction accumulator)
        //   D[48:96)  += A_lo' x W_hi =====
        if (lane == 0) {
            constexpr uint32_t idesc_2n = (1u << 4) | (uint32_t((2 * CS_BN) >> 3) << 17) | (uint32_t(128 >> 4) << 24);
            constexpr uint32_t idesc_n = (1u << 4) | (uint32_t(CS_BN >> 3) << 17) | (uint32_t(128 >> 4) << 24);
            unsigned int it = 0, tiles = 0;
            for (int ph = 0; ph < n_phases; ++ph) {
                const CvsPhase& P = phases[ph];
                if (P.kind != CVS_GEMM) continue;
                for (int item = blockIdx.x; item < P.items; item += G, ++tiles) {
                    mbar_wait(&bar_acc_empty, (tiles & 1) ^ 1);   // the workers have drained the previous tile
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int kb = 0; kb < P.nkb; ++kb, ++it) {
                        const int s = it % CS_STAGES;
                        mbar_wait(&bar_full[s], (it / CS_STAGES) & 1);
                        if (kb == 0) CVS_ST2(2);
                        if (kb == P.nkb - 1) CVS_ST2(3);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t a_hi = smem_u32(smem + s * CS_STAGE_BYTES), a_lo = a_hi + CS_A_BYTES, w_hi = a_hi + 2 * CS_A_BYTES;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint32_t off = ks * 32;   // 16 halves = 32 B along K inside the swizzle atom
                            umma_f16(tmem_base, desc128(a_hi + off), desc128(w_hi + off), idesc_2n, (kb > 0 || ks > 0) ? 1u : 0u);
                            umma_f16(tmem_base + CS_BN, desc128(a_lo + off), desc128(w_hi + off), idesc_n, 1u);
                        }
                        umma_commit(&bar_empty[s]);   // stage reusable once these MMAs have read it
                    }
                    umma_commit(&bar_acc_full);
                    CVS_ST2(4);
                }
            }
        }
    } else {
        // ===== workers (8 warps) =====
        const int wt = tid - 64, ww = warp - 2;
        const int q = warp & 3, half = ww >> 2;         // TMEM lane quadrant of this warp, column half of the tile
        const int row = q * 32 + lane;
        const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16);
        unsigned int tiles = 0;
        for (int ph = 0; ph < n_phases; ++ph) {
            const CvsPhase& P = phases[ph];
            if (blockIdx.x == 0 && wt == 0 && ph < CVS_MAX_PHASES) g_cvs_stamp[ph * 4] = clock64();
            // Every CTA must have finished phase ph - 1 before this one starts phase ph: attention / LayerNorm read what
            // the previous phase wrote, and the arrival count is only meaningful when nobody arrives for phase ph early
            // (a CTA without items in a GEMM phase would otherwise arrive twice before a slow CTA has arrived once).
            // parameter vectors are cold in L2 every window (850 MB of weights stream through it): requested before the wait
            if (P.kind == CVS_LN) {
                for (int l = wt * 32; l < P.cols; l += CS_WORKERS * 32) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.gamma + l));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.beta + l));
                    if (P.bias) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.bias + l));
                }
            } else if (P.kind == CVS_GEMM && P.epi != CVS_EPI_PARTIAL && int(blockIdx.x) < P.items) {
                if (wt < 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.bias + (blockIdx.x / P.splitk) * CS_BN + wt * 32));
            }
            if (ph > 0) {
                if (wt == 0) wait_grid(bar, (unsigned int)ph * G);
                worker_sync();
            }
            if (blockIdx.x == 0 && wt == 0 && ph < CVS_MAX_PHASES) g_cvs_stamp[ph * 4 + 3] = clock64();
            if (P.kind == CVS_GEMM) {
                for (int item = blockIdx.x; item < P.items; item += G, ++tiles) {
                    const int tn = item / P.splitk, z = item - tn * P.splitk;
                    const int n0 = tn * CS_BN + half * (CS_BN / 2);
                    // bias of this thread's 24 columns: requested before the accumulator is ready (cold parameter vectors)
                    float4 bq[CS_BN / 8];
                    if (P.epi != CVS_EPI_PARTIAL) {
#pragma unroll
                        for (int j = 0; j < CS_BN / 8; ++j) bq[j] = __ldg(reinterpret_cast<const float4*>(P.bias + n0) + j);
                    }
                    mbar_wait(&bar_acc_full, tiles & 1);
                    if (wt == 0) CVS_ST2(5);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                    for (int c0 = 0; c0 < CS_BN / 2; c0 += 8) {
                        float v[8], c[8];
                        tmem_ld8x2(trow + half * (CS_BN / 2) + c0, trow + CS_BN + half * (CS_BN / 2) + c0, v, c);
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = fmaf(c[j], 1.0f / CS_LO_SCALE, v[j]);
                        if (row < T) {
                            const int n = n0 + c0;
                            if (P.epi == CVS_EPI_PARTIAL) {
                                float* dst = P.C + ((long long)z * 128 + row) * P.ldc + n;
                                __stcg(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3]));
                                __stcg(reinterpret_cast<float4*>(dst) + 1, make_float4(v[4], v[5], v[6], v[7]));
                            } else {
                                const float4 b0 = bq[c0 / 4], b1 = bq[c0 / 4 + 1];
                                v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
                                if (P.epi == CVS_EPI_BIAS) {
                                    float* dst = P.C + (long long)row * P.ldc + n;
                                    __stcg(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3]));
                                    __stcg(reinterpret_cast<float4*>(dst) + 1, make_float4(v[4], v[5], v[6], v[7]));
                                } else {   // CVS_EPI_GELU_PLANES
                                    uint32_t hw[4], lw[4];
#pragma unroll
                                    for (int j = 0; j < 4; ++j) split2(gelu_f(v[2 * j]), gelu_f(v[2 * j + 1]), hw[j], lw[j]);
                                    *reinterpret_cast<uint4*>(P.p_hi + (long long)row * P.ldp + n) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                                    *reinterpret_cast<uint4*>(P.p_lo + (long long)row * P.ldp + n) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                                }
                            }
                        }
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_acc_empty);
                }
            } else {
                if (P.kind == CVS_ATTN) {
                    // item = (head, block of CS_ATT_ROWS query rows); K^T staged, scores + softmax for the block's rows,
                    // then V staged over K^T, P.V; every reduction in a fixed order
                    const int Tp = (T | 1) + 2;
                    float* KV = reinterpret_cast<float*>(scratch);              // [64][Tp] (K^T) then [T][64] (V)
                    float* Ps = KV + CS_D * Tp;                                  // [CS_ATT_ROWS][Tp]
                    float* inv_s = Ps + P.att_rows * Tp;                         // [att_rows]
                    const int HD = P.heads * CS_D;
                    const int AR = P.att_rows;   // query rows per item (8, 16 or 24: what the scratch holds at this T)
                    const int nblk = (T + AR - 1) / AR;
                    for (int item = blockIdx.x; item < P.items; item += G) {
                        const int h = item / nblk, qbase = (item - h * nblk) * AR;
                        const int qend = min(T, qbase + AR);
                        const float* qkv = P.qkv;
                        if (wt == 0 && item == int(blockIdx.x)) CVS_ST2(0);
                        constexpr int NLD = 8;   // float4 per thread: T * 16 <= 256 * 8  (T <= 128)
                        float4 kreg[NLD], vreg[NLD];
#pragma unroll
                        for (int u = 0; u < NLD; ++u) {
                            const int i = wt + u * CS_WORKERS;
                            const int t = i / (CS_D / 4), d4 = i - t * (CS_D / 4);
                            kreg[u] = t < T ? __ldcg(reinterpret_cast<const float4*>(qkv + (long long)t * P.ldqkv + HD + h * CS_D) + d4) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        // the block's query rows of this warp (up to CS_ATT_ROWS / 8 = 3), requested together with K
                        float4 qreg[CS_ATT_ROWS / 8][2];
#pragma unroll
                        for (int rr_ = 0; rr_ < CS_ATT_ROWS / 8; ++rr_) {
                            const int qi = qbase + ww + rr_ * 8;
                            // lane l holds q[2l], q[2l+1] ... as one float2 would need shuffles: keep two float4 = dims [8 (l%8) .. +8) of the row
                            const int d8 = (lane & 7) * 8;
                            qreg[rr_][0] = qi < qend ? __ldcg(reinterpret_cast<const float4*>(qkv + (long long)qi * P.ldqkv + h * CS_D + d8)) : make_float4(0.f, 0.f, 0.f, 0.f);
                            qreg[rr_][1] = qi < qend ? __ldcg(reinterpret_cast<const float4*>(qkv + (long long)qi * P.ldqkv + h * CS_D + d8 + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
#pragma unroll
                        for (int u = 0; u < NLD; ++u) {
                            const int i = wt + u * CS_WORKERS;
                            const int t = i / (CS_D / 4), d4 = i - t * (CS_D / 4);
                            vreg[u] = t < T ? __ldcg(reinterpret_cast<const float4*>(qkv + (long long)t * P.ldqkv + 2 * HD + h * CS_D) + d4) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
#pragma unroll
                        for (int u = 0; u < NLD; ++u) {
                            const int i = wt + u * CS_WORKERS;
                            const int t = i / (CS_D / 4), d4 = i - t * (CS_D / 4);
                            if (t < T) {
                                KV[(d4 * 4 + 0) * Tp + t] = kreg[u].x; KV[(d4 * 4 + 1) * Tp + t] = kreg[u].y;
                                KV[(d4 * 4 + 2) * Tp + t] = kreg[u].z; KV[(d4 * 4 + 3) * Tp + t] = kreg[u].w;
                            }
                        }
                        if (wt == 0 && item == int(blockIdx.x)) CVS_ST2(1);
                        // q rows -> shared memory (row-major [rows][64]) so every lane can read the whole row as broadcasts
                        float* Qs = inv_s + AR;   // [att_rows][64]
#pragma unroll
                        for (int rr_ = 0; rr_ < CS_ATT_ROWS / 8; ++rr_) {
                            if (lane < 8 && ww + rr_ * 8 < AR) {
                                float* qd = Qs + (ww + rr_ * 8) * CS_D + lane * 8;
                                *reinterpret_cast<float4*>(qd) = qreg[rr_][0];
                                *reinterpret_cast<float4*>(qd + 4) = qreg[rr_][1];
                            }
                        }
                        worker_sync();
                        if (wt == 0 && item == int(blockIdx.x)) CVS_ST2(2);
                        {
                            // scores of this warp's (up to three) rows qbase + ww + {0, 8, 16} together: a lane owns keys lane, lane + 32,
                            // lane + 64, lane + 96; every K element read from shared memory feeds three rows, every q element (broadcast)
                            // four keys - 12 independent FMA chains per lane instead of one load -> one FMA at a time.  Per (row, key)
                            // the sum runs over d ascending in two interleaved partial sums (even / odd d), as in attn_kernel.
                            constexpr int NR = CS_ATT_ROWS / 8, NK = 4;
                            float a0[NR][NK], a1[NR][NK];
#pragma unroll
                            for (int r_ = 0; r_ < NR; ++r_)
#pragma unroll
                                for (int k_ = 0; k_ < NK; ++k_) { a0[r_][k_] = 0.f; a1[r_][k_] = 0.f; }
                            int jk[NK];
#pragma unroll
                            for (int k_ = 0; k_ < NK; ++k_) jk[k_] = min(lane + 32 * k_, T - 1);   // clamped: the extra keys are never stored
                            const float* qrow[NR];
#pragma unroll
                            for (int r_ = 0; r_ < NR; ++r_) qrow[r_] = Qs + min(ww + r_ * 8, AR - 1) * CS_D;   // clamped: rows past the block are never stored
#pragma unroll 4
                            for (int d = 0; d < CS_D; d += 4) {
                                float4 qv[NR];
#pragma unroll
                                for (int r_ = 0; r_ < NR; ++r_) qv[r_] = *reinterpret_cast<const float4*>(qrow[r_] + d);
                                float kx0[NK], kx1[NK], kx2[NK], kx3[NK];
#pragma unroll
                                for (int k_ = 0; k_ < NK; ++k_) {
                                    kx0[k_] = KV[d * Tp + jk[k_]]; kx1[k_] = KV[(d + 1) * Tp + jk[k_]];
                                    kx2[k_] = KV[(d + 2) * Tp + jk[k_]]; kx3[k_] = KV[(d + 3) * Tp + jk[k_]];
                                }
#pragma unroll
                                for (int r_ = 0; r_ < NR; ++r_)
#pragma unroll
                                    for (int k_ = 0; k_ < NK; ++k_) {
                                        a0[r_][k_] = fmaf(qv[r_].x, kx0[k_], a0[r_][k_]);
                                        a1[r_][k_] = fmaf(qv[r_].y, kx1[k_], a1[r_][k_]);
                                        a0[r_][k_] = fmaf(qv[r_].z, kx2[k_], a0[r_][k_]);
                                        a1[r_][k_] = fmaf(qv[r_].w, kx3[k_], a1[r_][k_]);
                                    }
                            }
#pragma unroll
                            for (int r_ = 0; r_ < NR; ++r_) {
                                const int qi = qbase + ww + r_ * 8;
                                if (qi >= qend || ww + r_ * 8 >= AR) continue;
                                float* ps = Ps + (qi - qbase) * Tp;
                                float sc[NK];
                                float mx = -FLT_MAX;
#pragma unroll
                                for (int k_ = 0; k_ < NK; ++k_) {
                                    sc[k_] = a0[r_][k_] + a1[r_][k_];
                                    if (lane + 32 * k_ < T) mx = fmaxf(mx, sc[k_]);
                                }
                                mx = warp_max(mx);
                                float sum = 0.f;
#pragma unroll
                                for (int k_ = 0; k_ < NK; ++k_) {
                                    if (lane + 32 * k_ < T) { const float e = expf(sc[k_] - mx); ps[lane + 32 * k_] = e; sum += e; }
                                }
                                sum = warp_sum(sum);
                                if (lane == 0) inv_s[qi - qbase] = 1.0f / sum;
                            }
                        }
                        if (wt == 0 && item == int(blockIdx.x)) CVS_ST2(3);
                        worker_sync();   // scores done: K^T may be overwritten by V (already in registers)
                        if (wt == 0 && item == int(blockIdx.x)) CVS_ST2(4);
#pragma unroll
                        for (int u = 0; u < NLD; ++u) {
                            const int i = wt + u * CS_WORKERS;
                            const int t = i / (CS_D / 4), d4 = i - t * (CS_D / 4);
                            if (t < T) reinterpret_cast<float4*>(KV + (size_t)t * CS_D)[d4] = vreg[u];
                        }
                        worker_sync();
                        {
                            // o = sum_j p_j v_j, j ascending, for the warp's three rows together (every V element feeds three rows)
                            constexpr int NR = CS_ATT_ROWS / 8;
                            float o0[NR], o1[NR];
                            const float* psr[NR];
#pragma unroll
                            for (int r_ = 0; r_ < NR; ++r_) { o0[r_] = 0.f; o1[r_] = 0.f; psr[r_] = Ps + min(ww + r_ * 8, AR - 1) * Tp; }
                            int j = 0;
                            for (; j + 4 <= T; j += 4) {
                                float va[4], vb[4];
#pragma unroll
                                for (int u = 0; u < 4; ++u) { va[u] = KV[(j + u) * CS_D + lane]; vb[u] = KV[(j + u) * CS_D + lane + 32]; }
#pragma unroll
                                for (int r_ = 0; r_ < NR; ++r_) {
                                    const float p0 = psr[r_][j], p1 = psr[r_][j + 1], p2 = psr[r_][j + 2], p3 = psr[r_][j + 3];
                                    o0[r_] = fmaf(p0, va[0], o0[r_]); o1[r_] = fmaf(p0, vb[0], o1[r_]);
                                    o0[r_] = fmaf(p1, va[1], o0[r_]); o1[r_] = fmaf(p1, vb[1], o1[r_]);
                                    o0[r_] = fmaf(p2, va[2], o0[r_]); o1[r_] = fmaf(p2, vb[2], o1[r_]);
                                    o0[r_] = fmaf(p3, va[3], o0[r_]); o1[r_] = fmaf(p3, vb[3], o1[r_]);
                                }
                            }
                            for (; j < T; ++j) {
                                const float v0 = KV[j * CS_D + lane], v1 = KV[j * CS_D + lane + 32];
#pragma unroll
                                for (int r_ = 0; r_ < NR; ++r_) { const float pj = psr[r_][j]; o0[r_] = fmaf(pj, v0, o0[r_]); o1[r_] = fmaf(pj, v1, o1[r_]); }
                            }
#pragma unroll
                            for (int r_ = 0; r_ < NR; ++r_) {
                                const int qi = qbase + ww + r_ * 8;
                                if (qi >= qend || ww + r_ * 8 >= AR) continue;
                                const float inv = inv_s[qi - qbase];
                                const float x0 = o0[r_] * inv, x1 = o1[r_] * inv;
                                // planes of the attention output (the A operand of the out-projection)
                                const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
                                const __half l0 = __float2half_rn((x0 - __half2float(h0)) * CS_LO_SCALE), l1 = __float2half_rn((x1 - __half2float(h1)) * CS_LO_SCALE);
                                __half* ph_ = reinterpret_cast<__half*>(P.p_hi) + (long long)qi * P.ldp + h * CS_D;
                                __half* pl_ = reinterpret_cast<__half*>(P.p_lo) + (long long)qi * P.ldp + h * CS_D;
                                ph_[lane] = h0; ph_[lane + 32] = h1; pl_[lane] = l0; pl_[lane + 32] = l1;
                            }
                        }
                        if (wt == 0 && item == int(blockIdx.x)) CVS_ST2(5);
                        worker_sync();   // the next item restages K^T
                    }
                } else if (P.kind == CVS_LN) {
                    // item = 8 rows, one per warp: t = sum_z partial[z] (z ascending) + bias + residual; y = LayerNorm(t)
                    const int nv = P.cols >> 7;   // float4 per lane (cols % 128 == 0, <= 1024)
                    for (int item = blockIdx.x; item < P.items; item += G) {
                        const int r = item * 8 + ww;
                        if (r >= T) continue;
                        // the loads of three float4 columns (partials, residual, bias) in flight at once - a dependent chain of L2
                        // round trips otherwise - then summed in the fixed order  ((p0 + p1 + p2 + p3) + bias) + residual
                        float4 v[8];
                        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int i0 = 0; i0 < 8; i0 += 3) {
                            float4 pz[3][CVS_SPLITK], rr[3], bb[3];
#pragma unroll
                            for (int u = 0; u < 3; ++u) {
                                const int i = i0 + u;
                                if (i >= 8 || i >= nv) break;
                                const int c = i * 128 + lane * 4;
#pragma unroll
                                for (int zz = 0; zz < CVS_SPLITK; ++zz)
                                    pz[u][zz] = zz < P.S ? __ldcg(reinterpret_cast<const float4*>(P.X + (long long)zz * P.slab + (long long)r * P.ldx + c)) : zero4;
                                rr[u] = P.R ? __ldcg(reinterpret_cast<const float4*>(P.R + (long long)r * P.ldr + c)) : zero4;
                                bb[u] = P.bias ? __ldg(reinterpret_cast<const float4*>(P.bias + c)) : zero4;
                            }
#pragma unroll
                            for (int u = 0; u < 3; ++u) {
                                const int i = i0 + u;
                                if (i >= 8 || i >= nv) break;
                                float4 a = pz[u][0];
#pragma unroll
                                for (int zz = 1; zz < CVS_SPLITK; ++zz) {
                                    if (zz < P.S) { a.x += pz[u][zz].x; a.y += pz[u][zz].y; a.z += pz[u][zz].z; a.w += pz[u][zz].w; }
                                }
                                if (P.bias) { a.x += bb[u].x; a.y += bb[u].y; a.z += bb[u].z; a.w += bb[u].w; }
                                if (P.R) { a.x += rr[u].x; a.y += rr[u].y; a.z += rr[u].z; a.w += rr[u].w; }
                                v[i] = a;
                            }
                        }
                        float s = 0.f;
#pragma unroll
                        for (int i = 0; i < 8; ++i) { if (i >= nv) break; s += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
                        const float mean = warp_sum(s) / float(P.cols);
                        float qv = 0.f;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if (i >= nv) break;
                            const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
                            qv += (dx * dx + dy * dy) + (dz * dz + dw * dw);
                        }
                        const float rstd = rsqrtf(warp_sum(qv) / float(P.cols) + P.eps);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if (i >= nv) break;
                            const int c = i * 128 + lane * 4;
                            const float4 g = __ldg(reinterpret_cast<const float4*>(P.gamma + c)), b = __ldg(reinterpret_cast<const float4*>(P.beta + c));
                            float4 y;
                            y.x = (v[i].x - mean) * rstd * g.x + b.x; y.y = (v[i].y - mean) * rstd * g.y + b.y;
                            y.z = (v[i].z - mean) * rstd * g.z + b.z; y.w = (v[i].w - mean) * rstd * g.w + b.w;
                            __stcg(reinterpret_cast<float4*>(P.Y + (long long)r * P.ldy + c), y);
                            uint32_t h0, l0, h1, l1;
                            split2(y.x, y.y, h0, l0); split2(y.z, y.w, h1, l1);
                            *reinterpret_cast<uint2*>(P.p_hi + (long long)r * P.ldp + c) = make_uint2(h0, h1);
                            *reinterpret_cast<uint2*>(P.p_lo + (long long)r * P.ldp + c) = make_uint2(l0, l1);
                        }
                    }
                }
            }
            if (blockIdx.x == 0 && wt == 0 && ph < CVS_MAX_PHASES) g_cvs_stamp[ph * 4 + 1] = clock64();
            // ---- end of phase: this CTA's stores -> visible to every SM (generic and async proxy), then one arrival ----
            asm volatile("fence.proxy.async;" ::: "memory");
            worker_sync();
            if (wt == 0) {
                __threadfence();
                atomicAdd(bar, 1u);
                if (blockIdx.x == 0 && ph < CVS_MAX_PHASES) g_cvs_stamp[ph * 4 + 2] = clock64();
            }
        }
    }
    // ---- teardown ----
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(CS_TMEM_COLS)) : "memory");
    }
    if (tid == 0) {
        // exit ticket: the last CTA out re-arms the barrier words for the next launch / graph replay
        __threadfence();
        if (atomicAdd(bar + 32, 1u) == G - 1) { bar[0] = 0; bar[32] = 0; __threadfence(); }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn cvs_encode_fn() {
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

int g_cvs_max_ctas = 0;

}  // namespace

// fp16 row-major [rows][K] matrix -> 2-D tensor map with a {64, box_rows} SWIZZLE_128B box
bool cvstack_encode_map(void* out128, const void* base, int K, int rows, int box_rows) {
    EncodeTiledFn fn = cvs_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {cuuint64_t(K), cuuint64_t(rows)};
    cuuint64_t strides[1] = {cuuint64_t(K) * 2};
    cuuint32_t box[2] = {64, cuuint32_t(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(reinterpret_cast<CUtensorMap*>(out128), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

void init_cvstack_attributes() {
    cudaFuncSetAttribute(cvstack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM_BYTES);
    int per_sm = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cvstack_kernel, CS_THREADS, CS_SMEM_BYTES);
    g_cvs_max_ctas = per_sm * sms;
}
int cvstack_max_ctas() { return g_cvs_max_ctas; }

// Experiment (RVC_EXP_SPIN=us[,smem]): instead of the stack, a kernel of the same grid that only waits - no memory
// traffic, no tensor pipe.  What the F0 lane then loses beside it is the cost of a co-resident grid as such.
__global__ void __launch_bounds__(CS_THREADS) cvstack_spin_kernel(unsigned long long ns) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do { __nanosleep(1000); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while (t - t0 < ns);
}

int launch_cvstack(const CvsDev& c, cudaStream_t stream) {
    static const char* spin = getenv("RVC_EXP_SPIN");
    if (spin && *spin) {
        const unsigned long long ns = 1000ull * (unsigned long long)atoi(spin);
        const char* comma = strchr(spin, ',');
        const int smem = comma ? atoi(comma + 1) : CS_SMEM_BYTES;
        static bool once = false;
        if (!once) { cudaFuncSetAttribute(cvstack_spin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM_BYTES); once = true; }
        cvstack_spin_kernel<<<c.grid, CS_THREADS, smem, stream>>>(ns);
        return 1;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(unsigned(c.grid)); cfg.blockDim = dim3(CS_THREADS); cfg.dynamicSmemBytes = CS_SMEM_BYTES; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident: the grid barrier cannot deadlock
    attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const CUtensorMap* maps = reinterpret_cast<const CUtensorMap*>(c.d_maps);
    const CvsPhase* phases = c.d_phases; int n = c.n_phases; unsigned int* bar = c.d_bar; int T = c.T;
    cudaLaunchKernelEx(&cfg, cvstack_kernel, maps, phases, n, bar, T);
    return 1;
}

void cvstack_debug_read(long long* out, int n) { cudaMemcpyFromSymbol(out, g_cvs_stamp, sizeof(long long) * size_t(n)); }
void cvstack_debug_read2(long long* out, int n) { cudaMemcpyFromSymbol(out, g_cvs_stamp2, sizeof(long long) * size_t(n)); }

}  // namespace rvc
