// chain.h - persistent "chain" kernel: a run of small dependent ops executed by ONE cooperative
// launch, with a grid-wide barrier where a kernel boundary used to be.
//
// Why: at batch 1 the RMVPE U-Net (137 exact-fp32 convolutions, reference call site
// rvc/src/f0/rmvpe.rs:235) and the synthesizer's text encoder / flow (rvc/src/rvc.rs:195) are
// chains of ~10 us kernels whose cost is launch + prologue + pipeline-fill latency, not work.  Inside
// one resident grid a dependent op costs a ~1.5 us barrier instead, the next op's weights are
// prefetched into L2 while the current one computes, and the grid size is a fixed SM budget, so
// the F0 lane no longer fights the ContentVec lane for SMs (DESIGN.md "Chains").
#pragma once
#include <cuda_runtime.h>

#include "gemm_common.cuh"

namespace rvc {

enum ChainKind : int { CH_GEMM = 0, CH_GEMM_DIRECT = 1, CH_AVGPOOL = 2, CH_LAYERNORM = 3, CH_RELATTN = 4 };

// tile shapes of the chain GEMM (256 threads; warp w owns k-quads w, w+8, ... of every k-tile); the small
// shapes give weight-streaming ops (M <= 32) one tile per CTA without a split-K round trip
enum ChainTile : int { CT_32x32 = 0, CT_16x64 = 1, CT_8x128 = 2, CT_32x16 = 3, CT_32x8 = 4, CT_8x32 = 5, CT_8x16 = 6, CT_16x16 = 7 };

struct ChainOpDev {
    int kind, variant;
    int tiles_m, tiles_n, splitk, batch;
    int item0, items;          // work items [item0, item0+items) of the op's phase
    gemmk::GemmParams g;       // CH_GEMM / CH_GEMM_DIRECT
    // CH_AVGPOOL: x0=in y0=out ld0=ldin i0=T i1=F i2=C
    // CH_LAYERNORM: x0=X y0=Y x1=gamma x2=beta ld0=ldx ld1=ldy i0=rows i1=cols f0=eps
    // CH_RELATTN: x0=qkv y0=out x1=rel_k x2=rel_v ld0=ldqkv ld1=ldo i0=T i1=heads i2=dim i3=window
    const float* x0; float* y0; const float* x1; const float* x2;
    long long ld0, ld1;
    int i0, i1, i2, i3;
    float f0;
    // weights worth pulling into L2 while the previous phase runs: rows x row_bytes at stride
    const char* pf_base; long long pf_stride; int pf_rows, pf_row_bytes;
    // slab chains (below): columns per CTA / per warp, k per 16 KB chunk, chunks of this op, index of its first chunk in
    // the per-CTA weight stream
    int slab_nc, slab_cw, slab_kc, slab_chunks, slab_chunk0, slab_pad;
};

struct ChainPhaseDev { int op0, op1, items, pad; };

// One step of a weight-streaming chain (kernels_wstream.cu): a skinny GEMM, at most four valid rows, contiguous A rows
struct WsOpDev {
    const float* A; const float* W; const float* bias; const float* R; float* C;
    long long ldw, ldc, ldr;
    int K, a_floats, n_rows, relu;
    int row[4], row_off[4];   // valid rows m and their offsets m * lda inside the activation span
};

struct ChainDev {
    ChainOpDev* d_ops = nullptr;
    ChainPhaseDev* d_phases = nullptr;
    unsigned int* d_bar = nullptr;   // [0] arrivals, [1] exits (self-cleaning)
    unsigned long long* d_dbg = nullptr;  // optional: globaltimer at every phase start (+ end), CTA 0
    int n_ops = 0, n_phases = 0, grid = 0;
    int wstream = 0;   // 1: every op is a skinny GEMM of the same width: the weight-streaming kernel runs the chain (kernels_wstream.cu)
    WsOpDev* d_wsops = nullptr;
    int cluster = 0;   // 1: the grid is ONE thread-block cluster (<= 16 CTAs), phases separated by the hardware cluster barrier
    // Slab chain (slab_chain_kernel): a single cluster whose GEMMs (M <= 24 rows) are split by output columns only - CTA c
    // owns columns [c nc, (c + 1) nc) of EVERY GEMM with the full K, no split-K - and whose weights were re-packed once, per
    // CTA, into one contiguous stream of 16 KB chunks in execution order: a CTA pulls its stream with bulk copies that run
    // ahead of the phase barriers (weights do not depend on activations), so the HBM latency of every op but the first is hidden.
    int slab = 0;
    float* d_slabs = nullptr; long long slab_stream_floats = 0; int slab_total_chunks = 0;
    int* d_chunk_bytes = nullptr;   // bytes of every chunk of the stream (the same for all CTAs)
};

constexpr int CHAIN_THREADS = 256;
constexpr int CHAIN_SMEM_BYTES = 90 * 1024;  // stage rings of 75-90 KB (kernels_chain.cu tile table)
constexpr int SLAB_CHUNK_FLOATS = 4096;      // 16 KB weight chunks
constexpr int SLAB_RING = 6;                 // chunks in flight per CTA (96 KB)
constexpr int SLAB_A_FLOATS = 18432;         // staged activation span of one GEMM (72 KB)
constexpr int SLAB_SMEM_BYTES = (SLAB_RING * SLAB_CHUNK_FLOATS + SLAB_A_FLOATS) * 4 + 256;
constexpr int SLAB_MAX_ROWS = 24;
// columns per CTA / per warp and k per chunk of a GEMM inside a slab chain of G CTAs (host and device agree through ChainOpDev)
inline void slab_shape(int N, int K, int G, bool gate, int& nc, int& cw, int& kc, int& chunks) {
    nc = (N + G - 1) / G;
    if (gate && (nc & 1)) ++nc;
    cw = (nc + 7) / 8;
    if (cw <= 2) cw = 2; else if (cw <= 4) cw = 4; else if (cw <= 6) cw = 6; else cw = 8;   // 8 only with <= 8 rows (registers)
    kc = (SLAB_CHUNK_FLOATS / nc) / 32 * 32;
    const int kpad = (K + 31) / 32 * 32;
    if (kc > kpad) kc = kpad;
    chunks = (K + kc - 1) / kc;
}
// packs the per-CTA weight streams of a slab chain (one launch per GEMM op, at plan build time)
void launch_slab_pack(const float* W, long long ldw, int N, int K, int nc, int kc, int chunks, float* slabs, long long stream_floats,
                      int chunk0, int G, cudaStream_t stream);

int launch_chain(const ChainDev& c, cudaStream_t stream);  // returns kernels launched (1)
int launch_wstream(const ChainDev& c, cudaStream_t stream);
bool wstream_shape_ok(int M, int N, int K, long long lda, int valid_rows);
void init_wstream_attributes();
void init_chain_attributes();
void chain_debug_read(long long* out, int n);  // [256 phases][8] clock64 stamps of the LAST chain launch, CTA 0
void chain_debug_read2(long long* out, int n); // [256 phases][2]: barrier spin start / end of CTA 0
int chain_max_cluster_ctas();     // largest single-cluster grid (16 with the non-portable size allowed), 0 = none
int chain_max_coresident_ctas();  // occupancy-derived upper bound for a cooperative launch

}  // namespace rvc
