// cvstack.h - phase table of the persistent ContentVec transformer kernel (kernels_cvstack.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace rvc {

#ifndef RVC_CVS_BN
#define RVC_CVS_BN 48
#endif
constexpr int CVS_BN = RVC_CVS_BN;   // output columns per GEMM tile (N of every stack GEMM is a multiple of it; 48 or 96)
constexpr int CVS_ATT_ROWS = 24;    // most query rows per attention item
constexpr int CVS_SCRATCH_BYTES = 46 * 1024;   // worker scratch of a CTA: K^T / V staging, scores, q rows
// query rows per attention item that fit the scratch at T rows: K^T [64][Tp] + per row (scores [Tp], 1 / sum, q [64])
inline int cvs_att_rows(int T) {
    const int Tp = (T | 1) + 2;
    for (int r = CVS_ATT_ROWS; r >= 8; r -= 8)
        if (64 * Tp * 4 + r * (Tp * 4 + 4 + 256) <= CVS_SCRATCH_BYTES) return r;
    return 0;
}
constexpr int CVS_MAX_PHASES = 128;
constexpr int CVS_SPLITK = 4;       // out-proj / FC2: K split in four, partial tiles summed by the following LayerNorm phase

enum CvsKind : int { CVS_GEMM = 0, CVS_ATTN = 1, CVS_LN = 2 };
enum CvsEpi : int { CVS_EPI_BIAS = 0,          // C[m, n] = acc + bias[n]                      (QKV)
                    CVS_EPI_PARTIAL = 1,       // C[z][m][n] = acc                             (split-K partial, summed by the next LN phase)
                    CVS_EPI_GELU_PLANES = 2 }; // planes(gelu(acc + bias[n]))                  (FC1)

struct CvsPhase {
    int kind, items;
    // CVS_GEMM: A planes = maps[a_map] (hi), maps[a_map + 1] (lo'); W planes = maps[w_map], maps[w_map + 1];
    // items = (N / 48) * splitk, nkb = 64-wide k-blocks per item
    int a_map, w_map, splitk, nkb, epi, pad0;
    const float* bias;
    float* C; long long ldc;
    // fp16 planes written by this phase (GELU epilogue, attention, LayerNorm): element [m * ldp + n]
    unsigned short* p_hi; unsigned short* p_lo; long long ldp;
    // CVS_ATTN: qkv = [T][3 * heads * 64] fp32, q pre-scaled; items = heads * ceil(T / att_rows), att_rows <= CVS_ATT_ROWS
    const float* qkv; long long ldqkv; int heads, att_rows;
    // CVS_LN: t = sum_{z < S} X[z * slab + m * ldx + c] (+ bias[c]) (+ R[m * ldr + c]); Y = LayerNorm(t) * gamma + beta; items = ceil(T / 8)
    const float* X; long long ldx, slab; int S, cols;
    const float* R; long long ldr;
    const float* gamma; const float* beta;
    float* Y; long long ldy;
    float eps; int pad2;
};

struct CvsDev {
    void* d_maps = nullptr;          // CUtensorMap[n_maps] (128 B each)
    CvsPhase* d_phases = nullptr;
    unsigned int* d_bar = nullptr;   // [0] arrivals, [32] exits (self-cleaning)
    int n_phases = 0, grid = 0, T = 0;
};

bool cvstack_encode_map(void* out128, const void* base, int K, int rows, int box_rows);
int launch_cvstack(const CvsDev& c, cudaStream_t stream);   // returns kernels launched (1)
void init_cvstack_attributes();
int cvstack_max_ctas();
void cvstack_debug_read2(long long* out, int n);  // [phase][8] GEMM pipeline stamps of CTA 0
void cvstack_debug_read(long long* out, int n);   // [phase][4] clock64 stamps of CTA 0 of the last launch

}  // namespace rvc
