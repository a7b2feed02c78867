// ops.h - operation descriptors of the per-window execution plan.
//
// The engine turns one `RvcInfer::infer` call (reference rvc/src/rvc.rs:133-220) into a fixed
// list of device operations ("plan") over channels-last, halo-padded fp32 buffers.  A
// descriptor is plain data: buffer references are (space, byte offset) pairs that are resolved
// against the arenas of a context at launch time, so the same plan can be replayed by the CUDA
// launcher (kernels_*.cu) and, in the test-suite only, by the scalar interpreter in
// oracle/plan_exec/.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace rvc {

enum Space : int32_t { SP_NULL = 0, SP_CV = 1, SP_F0 = 2, SP_SYN = 3, SP_IDX = 4, SP_WORK = 5,
                       SP_STATE = 6,  // persistent per-stream state: params, pitch cache, pcm, audio
                       SP_COUNT = 7 };

struct Ref {
    int32_t space = SP_NULL;
    int64_t off = 0;  // bytes
    bool null() const { return space == SP_NULL; }
    Ref plus(int64_t elems, int64_t elem_size = 4) const { return Ref{space, off + elems * elem_size}; }
};

enum Act : int32_t {
    ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2, ACT_LRELU01 = 3, ACT_LRELU001 = 4, ACT_SIGMOID = 5,
    ACT_TANH = 6,
    ACT_GATE = 7  // columns (2j, 2j+1) -> tanh(v0)*sigmoid(v1) written to column j
};

enum OutMode : int32_t {
    OUT_PLAIN = 0,     // C[m*ldc + col]
    OUT_PIXSHUF2 = 1,  // ConvTranspose2d k3 s2: row m=(qt,qf) on the padded input grid, column
                       // n=(rt,rf,co) -> pixel (2qt+rt, 2qf+rf); om_a = Fin+2, om_b = cout,
                       // om_c = output row pitch in pixels
    OUT_CONVT1D = 2    // ConvTranspose1d: row m=q, column n=(r,co) -> sample o=q*u+r-p, skipped
                       // unless 0<=o<Tout; om_a=u, om_b=cout, om_c=p, om_d=Tout
};

// Implicit GEMM:  C[m,n] = epilogue( alpha * sum_k A(m,k) * W[n,k] + bias[n] ).
// A(m,k) = A[m*lda + (k / seg_len)*seg_stride + (k % seg_len)]  - "segmented rows": with
// channels-last halo-padded activations every conv1d / conv2d-3x3 / transposed conv / linear of
// the three networks is this one contraction (DESIGN.md "Data layout").
struct GemmOp {
    Ref A; int64_t lda = 0; int32_t seg_len = 0; int64_t seg_stride = 0;
    Ref W; int64_t ldw = 0;
    Ref bias;
    Ref C; int64_t ldc = 0;
    Ref C2; int64_t ldc2 = 0; int32_t act2 = ACT_NONE;  // optional second output act2(out)
    Ref R; int64_t ldr = 0;                             // optional residual added AFTER act
    int32_t M = 0, N = 0, K = 0;
    int32_t act = ACT_NONE;
    float alpha = 1.0f;
    int32_t mask_period = 0, mask_valid = 0;  // rows with (m % period) >= valid write 0
    int32_t batch = 1; int64_t sA = 0, sW = 0, sBias = 0, sC = 0, sR = 0;  // element strides
    int32_t out_mode = OUT_PLAIN; int32_t om_a = 0, om_b = 0, om_c = 0, om_d = 0;
    // kernel schedule chosen at plan time (gemm_sched.h): 0 = v1 tile kernel, >0 = v2 cp.async
    // split-K variant; scratch[splitk][batch][M][N] partial tiles + one arrival counter per tile
    int32_t cta_budget = 0;  // >0: CTA target of this op (it shares the GPU with sibling lanes); 0 = scheduler default
    int32_t sched_variant = 0, splitk = 1; Ref scratch, counters;
    // schedule of the same op when it runs inside a persistent chain (chain.h): tile shape
    // (-1 = scalar "direct" path for tiny / unaligned contractions), tile grid and split-K
    int32_t ch_variant = -1, ch_tiles_m = 0, ch_tiles_n = 0, ch_splitk = 1; Ref ch_scratch, ch_counters;
};

struct LayerNormOp {  // y = (x-mean)/sqrt(var+eps)*gamma+beta over `cols`, biased variance
    Ref X; int64_t ldx = 0; Ref Y; int64_t ldy = 0; Ref gamma, beta;
    int32_t rows = 0, cols = 0; float eps = 1e-5f;
};

struct AttnOp {  // softmax(q k^T) v per head; q pre-scaled; qkv = [T, 3*heads*dim]
    Ref qkv; int64_t ldqkv = 0; Ref out; int64_t ldo = 0; int32_t T = 0, heads = 0, dim = 0;
};

struct RelAttnOp {  // + windowed relative-position terms (VITS MultiHeadAttention)
    Ref qkv; int64_t ldqkv = 0; Ref out; int64_t ldo = 0; Ref rel_k, rel_v;  // [2w+1, dim]
    int32_t T = 0, heads = 0, dim = 0, window = 0;
};

struct Conv0StatsOp {  // per-channel mean / rstd over time of conv1d(1->C,k,stride) (GroupNorm)
    Ref pcm; Ref w; Ref stats; int32_t T = 0, C = 0, k = 0, stride = 0; float eps = 1e-5f;
};
struct Conv0ApplyOp {  // y = gelu((conv-mean)*rstd*gamma+beta), channels-last [T,C]
    Ref pcm; Ref w; Ref stats; Ref gamma, beta; Ref Y; int32_t T = 0, C = 0, k = 0, stride = 0;
};

struct StftMelOp {  // reflect-pad + periodic Hann + 1024-pt FFT + |.| + sparse mel + ln(max(.,clamp))
    Ref pcm; int32_t L = 0, T = 0;         // window of L samples, T = 1 + L/160 frames
    Ref window;                            // f32[1024]
    Ref band_start, band_count, band_off;  // i32[128] each: bins [start, start+count), weights at off
    Ref band_w;                            // f32[sum count]
    Ref mel;                               // f32[T,128] raw log-mel (time-major)
    Ref out2; int64_t out2_pitch = 0;      // optional: affine copy into a padded [T+2, 130] plane
    float scale = 1.0f, shift = 0.0f; float clamp = 1e-5f;
};

struct AvgPoolOp {  // 2x2 average on halo-padded NHWC
    Ref in; int64_t ldin = 0; Ref out; int32_t T = 0, F = 0, C = 0;  // input interior T x F
};

struct GruOp {  // bidirectional single-layer GRU, gate order [r,z,n]; whh_t is [dir][H][3H]
    Ref gi; Ref whh_t; Ref bhh; Ref out; int32_t T = 0, H = 0;
};

struct F0DecodeOp {  // rmvpe.rs:118-133,243-248 (+ rvc.rs:121 uppower from the params block)
    Ref salience; Ref f0; Ref argmax; Ref params; int32_t T = 0, bins = 360; float threshold = 0.03f;
    int32_t upstream_window = 0;
};

struct F0PostOp {  // rvc.rs:167-181 + f0/mod.rs:7-12
    Ref f0; Ref cache; Ref pitch; Ref pitchf; int32_t pitch_len = 0, shift = 0, hubert_length = 0,
        skip_head = 0, return_length = 0, cache_len = 1024; float mel_min = 0, mel_max = 0;
    int32_t sequential = 0;  // batched plans: the windows are consecutive windows of ONE stream (single pitch cache, updated in order)
};

struct EmbedOp {  // lrelu_0.1((phone Wp^T + bp + emb_pitch[pitch]) * sqrt(H))
    Ref phone; Ref pitch; Ref wp; Ref bp; Ref emb_pitch; Ref out; int64_t ldo = 0;
    Ref pre;  // non-null: [R, H] = phone Wp^T + bp, computed by an earlier GEMM (it does not need the pitch: off the critical path)
    int32_t R = 0, Cin = 0, H = 0;
};

struct ZpOp {  // z = m + exp(logs) * noise * 0.66666 ; stats = [R, 2H] (m | logs)
    Ref stats; Ref out; int64_t ldo = 0; Ref params; int32_t R = 0, H = 0;
};

struct SineGenOp {  // SineGen(harmonic_num=0) + tanh(linear) -> har (halo-padded, 1 channel)
    Ref pitchf; Ref out; Ref sine_dbg; Ref params; int32_t R = 0, upp = 0; float sr = 0, lin_w = 0,
        lin_b = 0;
};

struct Avg3Op {  // s = (a+b+c)/3 ; out = lrelu_slope(s) ; optional raw s
    Ref a, b, c; int64_t ld = 0; Ref out; int64_t ldo = 0; Ref raw; int64_t ldraw = 0;
    int32_t T = 0, C = 0; float slope = 0.1f;
};

struct ConvPostOp {  // tanh(conv1d(Cin->1, k)) on halo-padded channels-last input
    Ref in; Ref w; Ref out; int32_t T = 0, C = 0, k = 0;
};

struct KnnScanOp {  // per-part top-k of D[q,n] = sum_c (x[q,c]-index[n,c])^2; the parts partition the rows (how is up to the executor)
    Ref index; Ref queries; int64_t ldq = 0; Ref cand_d, cand_i;  // [Q][parts][k]
    int32_t N = 0, C = 0, Q = 0, k = 0, parts = 0;
    // umma = 1: candidate pass on the tensor cores (kernels_knn_umma.cu): cand = [Q][parts][KNN_UMMA_KC] approximate
    // scores |y|^2 - 2 x.y; the fp16 planes of the index sit planes_off bytes behind its fp32 rows
    int32_t umma = 0; int64_t planes_off = 0;
};
struct KnnSelectOp {  // k smallest (d, idx) per query over parts*k candidates, ascending
    Ref cand_d, cand_i; Ref idx; Ref d2; int32_t Q = 0, k = 0, parts = 0;
    // rerank = 1 (after a umma scan): the KNN_UMMA_KC best candidates by approximate score are re-evaluated exactly in
    // fp32 (the scan's summation order), a guard proves the top-k, a failed guard falls back to the exact scan of all rows
    int32_t rerank = 0; Ref index, queries; int64_t ldq = 0; int32_t N = 0, C = 0; float ymax2 = 0.f;
    int64_t planes_off = 0, fallback_off = 0;
};
struct KnnBlendOp {  // out[Q,C] = rate * sum_i w_i index[idx_i] + (1-rate) x ; w = (1/d2)^2 normalised
    Ref index; Ref idx; Ref d2; Ref x; int64_t ldx = 0; Ref out; Ref params; int32_t C = 0, Q = 0, k = 0;
};

struct GatherRowsOp {  // out[r] = src[min((skip+r)/2, T-1) - row0]  (rvc.rs:99-109,155)
    Ref src; int64_t lds = 0; Ref out; int32_t T = 0, C = 0, skip = 0, R = 0, row0 = 0;
};

struct FillOp { Ref dst; int64_t bytes = 0; };
struct WaitOp { int32_t src_lane = 0, dst_lane = 0; };

// RMVPE ConvBlockRes executed by one kernel (kernels_cbr.cu); carried by the block's LAST GEMM op (Op::fuse == 2)
struct CbrOp {
    Ref in; Ref w1, b1, w2, b2, wsc, bsc; Ref dst; int64_t ld_dst = 0;
    int32_t T = 0, F = 0, Cin = 0, C = 0;
};
// shapes the fused kernel is instantiated for: 16 / 32 output channels, strips of 32 / 8 or 16 pixels
inline bool cbr_supported(int C, int Cin, int F) {
    return Cin % 4 == 0 && Cin <= 64 && ((C == 16 && F % 32 == 0) || (C == 32 && F % 16 == 0));
}

enum OpKind : int32_t {
    OP_GEMM, OP_LAYERNORM, OP_ATTN, OP_RELATTN, OP_CONV0_STATS, OP_CONV0_APPLY, OP_STFTMEL,
    OP_AVGPOOL, OP_GRU, OP_F0DECODE, OP_F0POST, OP_EMBED, OP_ZP, OP_SINEGEN, OP_AVG3,
    OP_CONVPOST, OP_KNN_SCAN, OP_KNN_SELECT, OP_KNN_BLEND, OP_GATHER_ROWS, OP_FILL, OP_WAIT
};

struct Op {
    int32_t kind = OP_FILL;
    int32_t lane = 0;
    int32_t chain = -1;  // index into Plan::chains when the op executes inside a persistent chain kernel
    int32_t stack = 0;   // 1: the op is executed by the persistent ContentVec stack kernel (Plan::cvstack, kernels_cvstack.cu)
    int32_t fuse = 0;    // RMVPE residual block run by the fused kernel: 1 = covered by it (no launch), 2 = launches it (Op::cbr)
    std::string name;  // plan-unique; debug lookups + parity tests
    // exactly one of these is meaningful, selected by `kind`
    GemmOp gemm; LayerNormOp ln; AttnOp attn; RelAttnOp relattn; Conv0StatsOp c0s; Conv0ApplyOp c0a;
    StftMelOp stft; AvgPoolOp pool; GruOp gru; F0DecodeOp f0d; F0PostOp f0p; EmbedOp embed; ZpOp zp;
    SineGenOp sine; Avg3Op avg3; ConvPostOp cpost; KnnScanOp kd; KnnSelectOp ks; KnnBlendOp kb;
    GatherRowsOp gather; FillOp fill; WaitOp wait; CbrOp cbr;
};

// Runtime parameters that change per call without changing the plan (device-resident block,
// refreshed by one small H2D copy before the graph launch).
struct RunParams {
    float uppower;        // 2^(pitch_shift/12) (integer division, rvc.rs:121)
    float index_rate;
    uint64_t noise_seed;
    uint64_t window;      // call counter of this stream
    int32_t noise_mode;   // 0 zeros, 1 counter-based Gaussian
    int32_t pad[3];
};

static const int NOISE_KIND_Z = 1;
static const int NOISE_KIND_SINE = 2;
static const int KNN_PARTS = 148;  // one candidate list per CTA (8 warp lists merged on chip), one CTA per SM
// Few queries (the per-window case: 11 x 768): two CTAs per SM without the register double buffer of the index rows -
// the second CTA's warps cover the load latency (scan 91 -> 66 us measured).  Needs the padded query block of a CTA
// to fit twice into shared memory and the small top-k lists (k <= 8).
inline bool knn_two_ctas_per_sm(int Q, int C, int k) {
    const int qn = Q <= 8 ? 8 : (Q <= 16 ? 16 : 32);
    return Q <= 32 && k <= 8 && int64_t(qn) * C * 4 <= 96 * 1024;
}
inline int knn_parts(int Q, int C, int k, int n_rows) {
    const int p = knn_two_ctas_per_sm(Q, C, k) ? 2 * KNN_PARTS : KNN_PARTS;
    return p < n_rows ? p : n_rows;
}

// tensor-core candidate pass of the retrieval (kernels_knn_umma.cu): tile of index rows, candidates per query and CTA
constexpr int KNN_UMMA_BN = 32, KNN_UMMA_KC = 16;
inline bool knn_umma_ok(int C, int k) { return C % 64 == 0 && C >= 64 && C <= 256 && k >= 1 && k <= 8; }
inline int knn_umma_rows_padded(int n_rows) { return (n_rows + KNN_UMMA_BN - 1) / KNN_UMMA_BN * KNN_UMMA_BN; }
inline int knn_umma_parts(int n_rows) { const int t = knn_umma_rows_padded(n_rows) / KNN_UMMA_BN; return t < 148 ? t : 148; }
// The planes are stored tile by tile in exactly the shared-memory image the MMA reads (K-major SWIZZLE_128B, per 64-wide
// k-block the y_hi rows then the y_lo' rows), followed by the tile's |y|^2: one contiguous cp.async.bulk per tile.
inline int64_t knn_umma_tile_bytes(int C) { return int64_t(C / 64) * 2 * KNN_UMMA_BN * 128 + KNN_UMMA_BN * 4; }
// [tiles x tile image | fallback counter, max |y|^2 bits]
inline int64_t knn_umma_counters_off(int n_rows, int C) { return int64_t(knn_umma_rows_padded(n_rows) / KNN_UMMA_BN) * knn_umma_tile_bytes(C); }
inline int64_t knn_umma_planes_bytes(int n_rows, int C) { return knn_umma_counters_off(n_rows, C) + 256; }

// A run of consecutive same-lane ops executed by one persistent cooperative kernel (chain.h).
// phase[i] is the barrier phase of op first+i: ops of one phase touch disjoint buffers.
struct ChainInfo { int32_t first = 0, count = 0, lane = 0, grid = 0, n_phases = 0; std::vector<int32_t> phase; };

// The ContentVec transformer layers as one persistent tcgen05 kernel (cvstack.h): ops [first, first + count) of the
// plan (enc_in LayerNorm, then per layer qkv / attn / o / ln1 / fc1 / fc2 / ln2) plus the buffers only that kernel
// uses: fp16 operand planes [hi | lo'] of the four GEMM inputs and the split-K partial tiles.
struct CvStackInfo {
    int32_t first = -1, count = 0, T = 0, width = 0, ffn = 0;
    Ref planes_x, planes_x1, planes_a, planes_h, partial;
};

// A named buffer of the plan (debug / result lookups).
struct NamedBuf { std::string name; Ref ref; int64_t elems = 0; int32_t is_int = 0; };

}  // namespace rvc
