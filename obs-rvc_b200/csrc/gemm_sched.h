// gemm_sched.h - plan-time choice of the GEMM kernel variant and split-K factor (host-only).
//
// At batch 1 most contractions of the path are small-M (4..111 rows) against multi-megabyte
// weight matrices: they are bound by how fast the weights stream out of HBM, which needs (a) a
// CTA count that is a multiple of the 148 SMs and (b) several k-tiles of loads in flight per
// CTA.  The v2 kernel (kernels_gemm2.cu) therefore splits K across CTAs; partial tiles go to a
// per-op scratch area and the last CTA to arrive on a tile reduces them in a fixed order and
// applies the epilogue (deterministic, one launch).
#pragma once
#include <algorithm>
#include <cstdlib>

#include "ops.h"

namespace rvc {

struct GemmSched {
    int variant = 0;  // 0: v1 (kernels_gemm.cu); 1: BM8/BN256; 2: BM16/BN128; 3: BM32/BN64; 4: BM32/BN32 (v2; six-stage ring when a split-K slice has <= 5 k-tiles);
                      // 5/6/7/8/9: tcgen05 kernel (kernels_umma.cu) with BN = 128/64/32/256/16
    int bm = 0, bn = 0, splitk = 1, tiles = 0;
};

constexpr int GEMM2_BK = 32;
constexpr int NUM_SMS = 148;

inline int sched_env(const char* name, int dflt) {
    const char* e = std::getenv(name);
    return (e && *e) ? std::atoi(e) : dflt;
}

// nb = windows per launch (batched plans): the grid is nb x larger.  Batched plans also put RMVPE's wide levels and the
// small-M ops (enc_p / flow, 21 rows per window) on the tensor-core kernel: with nb windows per launch they are
// throughput- not latency-bound, and the 2-term FP16 split is fp32-grade (argmax / pitch parity is asserted by the tests).
inline GemmSched gemm_schedule(const GemmOp& g, bool allow_umma, int nb = 1, bool f0_umma = false) {
    GemmSched s;
    static const int kWant = sched_env("RVC_UMMA_WANT", 96), kBn128 = sched_env("RVC_UMMA_BN128_MIN", 1),
                     kKbMin = sched_env("RVC_UMMA_KB_MIN", 4);
    const bool aligned = g.A.off % 16 == 0 && g.W.off % 16 == 0 && g.lda % 4 == 0 && g.seg_len % 4 == 0 &&
                         g.seg_stride % 4 == 0 && g.K % 4 == 0 && g.ldw % 4 == 0 && g.sA % 4 == 0 && g.sW % 4 == 0;
    if (!aligned || g.K < 64) return s;  // tiny / oddly aligned contractions stay on the v1 kernel
    // tensor-core path: dense contractions of ContentVec / synthesizer with at least half a 128-row tile;
    // RMVPE (SP_F0) stays on exact-fp32 CUDA cores (its 360-bin argmax is a bit-exact parity item)
    const bool contiguous = g.seg_len >= g.K;
    const int min_m = nb > 1 ? 16 : 64;
    if (allow_umma && (g.W.space == SP_CV || g.W.space == SP_SYN || (f0_umma && g.W.space == SP_F0)) && g.M >= min_m && g.N % 16 == 0 && g.N >= 16 &&
        (contiguous || (g.seg_len % 32 == 0 && g.K % g.seg_len == 0) ||
         (sched_env("RVC_UMMA_F16", 1) != 0 && g.seg_len % 8 == 0 && g.K % g.seg_len == 0)) && g.K >= 96) {
        // (segments that are only a multiple of 8 long - ContentVec's grouped pos-conv, 48 channels per group - need the
        //  FP16-split kernel, whose A path walks the segmented rows itself; the TMA box of the 3xTF32 path cannot)
        const int tm = (g.M + 127) / 128;
        const int nkb = (g.K + 31) / 32;
        // one CTA per SM (smem-limited) and clusters must pack into GPCs: aim for a single wave of
        // <= ~112 CTAs; prefer 128-wide tiles (less operand traffic per flop) when that still fills it
        const int tiles128 = tm * ((g.N + 127) / 128) * g.batch * nb;
        const int bn128_min = g.cta_budget > 0 ? std::max(1, kBn128 * g.cta_budget / kWant) : kBn128;
        int bn = 128;
        if (g.N <= 16) bn = 16;
        else if (g.N <= 32) bn = 32;
        else if (g.N <= 64 || tiles128 < bn128_min) bn = 64;
        // 256-wide tiles (FP16-split kernel only): one tcgen05.mma costs ~175 cycles of latency on the accumulator chain
        // whatever its N <= 256, so for wide outputs a quarter of the instructions per unit of work - taken when the
        // tile grid x the deepest split-K still gives ~half a wave of CTAs
        static const bool kF16 = sched_env("RVC_UMMA_F16", 1) != 0, k256 = sched_env("RVC_UMMA_BN256", 0) != 0;   // measured: no gain over 128-wide tiles (2.93 vs 2.90 ms/window), off
        const int tiles256 = tm * ((g.N + 255) / 256) * g.batch * nb;
        if (kF16 && k256 && g.N >= 512 && tiles256 * std::min(8, std::max(1, nkb / kKbMin)) >= 48) bn = 256;
        s.variant = bn == 256 ? 8 : (bn == 128 ? 5 : (bn == 64 ? 6 : (bn == 32 ? 7 : 9)));
        s.bm = 128; s.bn = bn;
        s.tiles = tm * ((g.N + bn - 1) / bn) * g.batch;
        const int want = std::max(1, (g.cta_budget > 0 ? g.cta_budget : kWant) / (s.tiles * nb));
        const int maxsplit = std::max(1, nkb / kKbMin);
        // split-K group = one thread-block cluster (partials reduced over DSMEM): power of two <= 8
        s.splitk = 1;
        while (s.splitk * 2 <= std::min(std::min(want, maxsplit), 8)) s.splitk *= 2;
        return s;
    }
    if (g.M <= 8) { s.variant = 1; s.bm = 8; s.bn = 256; }
    else if (g.M <= 16) { s.variant = 2; s.bm = 16; s.bn = 128; }
    else if (g.M <= 256) { s.variant = 3; s.bm = 32; s.bn = 64; }
    else { s.variant = 3; s.bm = 32; s.bn = 64; }
    // tall, narrow outputs (RMVPE's top levels: thousands of pixels x 16 / 32 channels): half the tile width, half the k loop
    if (g.N <= 32 && g.M >= 256) { s.variant = 4; s.bm = 32; s.bn = 32; }
    // weight-streaming convs with a single row tile (RMVPE's deep levels: <= 32 pixels x 256..512 channels x K in the
    // thousands): 32-wide column tiles double the CTAs that pull the weights (RVC_V2_NARROW=0 restores 64-wide tiles)
    static const int kNarrow = sched_env("RVC_V2_NARROW", 1);
    static const int kNarrowM = sched_env("RVC_V2_NARROW_M", 32);
    if (kNarrow && g.M <= kNarrowM && g.N >= 64 && g.K >= 512 && s.variant == 3) { s.variant = 4; s.bm = 32; s.bn = 32; }
    // narrow outputs: do not waste a 256/128-wide tile on a 32..64-column problem
    if (s.variant == 1 && g.N <= 64) { s.variant = 3; s.bm = 32; s.bn = 64; }
    if (s.variant == 2 && g.N <= 64) { s.variant = 3; s.bm = 32; s.bn = 64; }
    const int tm = (g.M + s.bm - 1) / s.bm, tn = (g.N + s.bn - 1) / s.bn;
    s.tiles = tm * tn * g.batch;
    const int nkt = (g.K + GEMM2_BK - 1) / GEMM2_BK;
    static const int kV2Want = sched_env("RVC_V2_WANT", 2 * NUM_SMS);
    int want = (kV2Want + s.tiles * nb - 1) / (s.tiles * nb);   // ~2 CTAs per SM in total
    int maxsplit = std::max(1, nkt / 4);                        // at least 4 k-tiles per split
    // every extra split costs a partial-tile round trip through L2 in the last CTA: keep the group small
    s.splitk = std::max(1, std::min(std::min(want, maxsplit), 8));
    // Weight-streaming convs with one row tile (RMVPE's deep levels): a CTA's k extent is a chain of HBM round trips
    // (ring of 3 k-tiles in flight) - more, shorter slices whose k-tiles are ALL in flight at once (six-stage ring,
    // <= 5 k-tiles per CTA) cut the chain to one round trip.  RVC_V2_DEEP = largest split-K factor (0 = off)
    static const int kDeep = sched_env("RVC_V2_DEEP", 0);   // measured: 2.79 vs 2.72-2.75 ms / window with 15-16 slices (the last CTA's reduction of 15 partial tiles costs more than the shorter load chain saves): off
    if (kDeep > 8 && s.variant == 4 && g.M <= kNarrowM && g.N >= 64 && g.K >= 512 && nb == 1) {
        const int sk = std::min(std::min(kDeep, (nkt + 4) / 5 > 0 ? nkt : 1), std::max(1, (NUM_SMS + s.tiles - 1) / s.tiles));
        int best = s.splitk;
        for (int c = s.splitk; c <= sk; ++c) if ((nkt + c - 1) / c <= 5) { best = c; break; }
        if (best > s.splitk) s.splitk = best;   // launch_gemm_v2 picks the six-stage instance when a slice has <= 5 k-tiles
    }
    if (g.out_mode != OUT_PLAIN && false) s.splitk = 1;
    return s;
}

}  // namespace rvc
