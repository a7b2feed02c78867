// kernels_knn.cu - exact brute-force L2 retrieval (FAISS IndexFlatL2 semantics, upstream RVC;
// the reference has only `// TODO: index search`, rvc/src/rvc.rs:159).
//
// HBM-bound design: the N x C index is streamed exactly once with coalesced 128-bit loads, one
// index row per warp iteration; the Q query rows live in shared memory; per-lane partial sums of
// squared differences for a group of up to 32 queries are reduced with a butterfly of warp
// shuffles (31 shuffles per 32 queries) that leaves query j's distance in lane j; every lane keeps
// the running top-k of "its" queries in registers, so no N x Q distance matrix ever touches memory.
// Per-warp candidate lists are merged by a tiny second kernel.  Distances are fp32 sum (x-y)^2 in
// a fixed order; ties resolve to the lowest row index.
#include <cfloat>
#include <climits>
#include <cstdlib>

#include "launch.h"
#include "pdl.cuh"

namespace rvc {

namespace {

// RR index rows per warp iteration are held in registers, so every query fragment fetched from shared memory
// is used RR times (the scan is otherwise bound by shared-memory reads of the queries, not by HBM); only the Q
// real queries of a padded group are evaluated.  The 8 per-warp top-k lists of a CTA are merged in shared
// memory before they leave the SM: one candidate list per CTA (`parts` = gridDim.x).
template <int CV, int QN, int G, int KK, int RR, int PF>
__global__ void __launch_bounds__(256, PF ? 1 : 2)
knn_scan_kernel(const float* __restrict__ index, int N, int C, const float* __restrict__ queries, long long ldq, int Q,
                float* __restrict__ cand_d, int* __restrict__ cand_i, int parts, int k,
                int q0, int Qw, long long wQ, long long wCandD, long long wCandI) {
    pdl_enter();
    // Q queries of this launch = global queries q0 .. q0+Q; global query g belongs to window g / Qw of a batched plan
    // (row g % Qw of that window's query block): ONE pass over the index serves every window of the batch
    extern __shared__ __align__(16) float qs[];  // [G*QN][C], zero padded; reused for the CTA merge at the end
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C4 = C >> 2;
    for (int e = tid; e < G * QN * C; e += 256) {
        int q = e / C, c = e - q * C;
        const int gq = q0 + q;
        qs[e] = q < Q ? queries[(long long)(gq / Qw) * wQ + (long long)(gq % Qw) * ldq + c] : 0.f;
    }
    __syncthreads();

    float ld[G][KK]; int li[G][KK];
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
        for (int j = 0; j < KK; ++j) { ld[g][j] = FLT_MAX; li[g][j] = -1; }

    const float4* qs4 = reinterpret_cast<const float4*>(qs);
    float4 y[RR][CV], yn[RR][CV];
    auto load_rows = [&](float4 (*dst)[CV], long long base) {
#pragma unroll
        for (int r = 0; r < RR; ++r) {
            const float4* row = reinterpret_cast<const float4*>(index + (base + r) * C);
            const bool ok = base + r < N;
#pragma unroll
            for (int i = 0; i < CV; ++i) {
                int c4 = lane + i * 32;
                dst[r][i] = (ok && c4 < C4) ? __ldcs(row + c4) : make_float4(0.f, 0.f, 0.f, 0.f);  // streamed once: evict-first
            }
        }
    };
    const long long stride = (long long)parts * 8 * RR;
    long long base = ((long long)blockIdx.x * 8 + warp) * RR;
    if (PF && base < N) load_rows(y, base);
    for (; base < N; base += stride) {
        const bool more = base + stride < N;
        // PF = 1: next rows in flight while these are reduced (one CTA per SM); PF = 0: no register double buffer, two
        // CTAs per SM - the second CTA's warps cover the load latency instead
        if (PF) { if (more) load_rows(yn, base + stride); }
        else load_rows(y, base);
#pragma unroll
        for (int g = 0; g < G; ++g) {
            float v[RR][QN];
#pragma unroll
            for (int qq = 0; qq < QN; ++qq) {
                float a[RR];
#pragma unroll
                for (int r = 0; r < RR; ++r) a[r] = 0.f;
                if (g * QN + qq < Q) {   // warp-uniform: padded queries cost nothing
                    const float4* xq = qs4 + (size_t)(g * QN + qq) * C4;
#pragma unroll
                    for (int i = 0; i < CV; ++i) {
                        int c4 = lane + i * 32;
                        if (c4 < C4) {
                            const float4 x = xq[c4];
#pragma unroll
                            for (int r = 0; r < RR; ++r) {
                                float d0 = x.x - y[r][i].x, d1 = x.y - y[r][i].y, d2 = x.z - y[r][i].z, d3 = x.w - y[r][i].w;
                                a[r] = fmaf(d0, d0, a[r]); a[r] = fmaf(d1, d1, a[r]); a[r] = fmaf(d2, d2, a[r]); a[r] = fmaf(d3, d3, a[r]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < RR; ++r) v[r][qq] = a[r];
            }
#pragma unroll
            for (int r = 0; r < RR; ++r) {
                // full-width steps while fewer than 32 values per lane group
#pragma unroll
                for (int s = 16; s >= QN; s >>= 1)
#pragma unroll
                    for (int i = 0; i < QN; ++i) v[r][i] += __shfl_xor_sync(0xffffffffu, v[r][i], s);
                // halving butterfly: after it, lane l holds the total of query (l & (QN-1))
#pragma unroll
                for (int s = QN / 2; s >= 1; s >>= 1) {
                    const bool up = (lane & s) != 0;
#pragma unroll
                    for (int i = 0; i < s; ++i) {
                        float send = up ? v[r][i] : v[r][i + s];
                        float keep = up ? v[r][i + s] : v[r][i];
                        v[r][i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
                    }
                }
                float cd = v[r][0];
                if (base + r < N && cd < ld[g][KK - 1]) {
                    int ci = int(base + r);
#pragma unroll
                    for (int j = 0; j < KK; ++j) {
                        if (cd < ld[g][j]) {
                            float td = ld[g][j]; int ti = li[g][j];
                            ld[g][j] = cd; li[g][j] = ci; cd = td; ci = ti;
                        }
                    }
                }
            }
        }
        if (PF && more) {
#pragma unroll
            for (int r = 0; r < RR; ++r)
#pragma unroll
                for (int i = 0; i < CV; ++i) y[r][i] = yn[r][i];
        }
    }
    // ---- CTA merge: 8 warp lists -> one list per query, (distance, row) ascending ----
    __syncthreads();   // every warp is done with the queries
    float* md = qs;                                               // [8][G*QN][KK]
    int* mi = reinterpret_cast<int*>(qs + 8 * G * QN * KK);       // [8][G*QN][KK]
    if (lane < QN) {
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int j = 0; j < KK; ++j) {
                md[(warp * G * QN + g * QN + lane) * KK + j] = ld[g][j];
                mi[(warp * G * QN + g * QN + lane) * KK + j] = li[g][j];
            }
    }
    __syncthreads();
    constexpr int NC = 8 * KK;            // candidates per query in the CTA
    constexpr int CPL = (NC + 31) / 32;   // per lane
    for (int q = warp; q < Q; q += 8) {
        float cd_[CPL]; int ci_[CPL];
#pragma unroll
        for (int u = 0; u < CPL; ++u) {
            const int c = lane + u * 32, w = c / KK, j = c - w * KK;
            const bool ok = c < NC;
            const int id = ok ? mi[(w * G * QN + q) * KK + j] : -1;
            cd_[u] = (ok && id >= 0) ? md[(w * G * QN + q) * KK + j] : FLT_MAX;
            ci_[u] = (ok && id >= 0) ? id : INT_MAX;
        }
        for (int r = 0; r < k; ++r) {
            float bd = FLT_MAX; int bi = INT_MAX;
#pragma unroll
            for (int u = 0; u < CPL; ++u)
                if (cd_[u] < bd || (cd_[u] == bd && ci_[u] < bi)) { bd = cd_[u]; bi = ci_[u]; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                float od = __shfl_xor_sync(0xffffffffu, bd, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
            }
#pragma unroll
            for (int u = 0; u < CPL; ++u)
                if (ci_[u] == bi && bi != INT_MAX) { cd_[u] = FLT_MAX; ci_[u] = INT_MAX; }   // row ids are unique: taken
            if (lane == 0) {
                const int gq = q0 + q, w = gq / Qw, j = gq - w * Qw;
                cand_d[w * wCandD + ((long long)j * parts + blockIdx.x) * k + r] = bd;
                cand_i[w * wCandI + ((long long)j * parts + blockIdx.x) * k + r] = bi == INT_MAX ? -1 : bi;
            }
        }
    }
}

// merges parts*k candidates per query: k rounds of "smallest (d, idx) greater than the previous".
// The candidate list is staged once in shared memory when it fits (it does for 2368 parts x k <= 8).
__global__ void __launch_bounds__(256)
knn_select_kernel(const float* __restrict__ cand_d, const int* __restrict__ cand_i, int* __restrict__ idx,
                  float* __restrict__ d2, int M, int k, int staged, long long wCandD, long long wCandI, long long wIdx, long long wD2) {
    pdl_enter();
    cand_d += blockIdx.z * wCandD; cand_i += blockIdx.z * wCandI; idx += blockIdx.z * wIdx; d2 += blockIdx.z * wD2;
    extern __shared__ __align__(16) float sel_sm[];
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* cd = cand_d + (long long)q * M;
    const int* ci = cand_i + (long long)q * M;
    if (staged) {
        float* sd_ = sel_sm; int* si_ = reinterpret_cast<int*>(sel_sm + M);
        for (int m = tid; m < M; m += 256) { sd_[m] = cd[m]; si_[m] = ci[m]; }
        __syncthreads();
        cd = sd_; ci = si_;
    }
    __shared__ float sd[8]; __shared__ int si[8];
    __shared__ float pd_s; __shared__ int pi_s;
    float pd = -FLT_MAX; int pi = -1;
    for (int r = 0; r < k; ++r) {
        float bd = FLT_MAX; int bi = INT_MAX;
        for (int m = tid; m < M; m += 256) {
            float d = cd[m]; int i = ci[m];
            if (i < 0) continue;
            if (!(d > pd || (d == pd && i > pi))) continue;
            if (d < bd || (d == bd && i < bi)) { bd = d; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float od = __shfl_xor_sync(0xffffffffu, bd, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        if (lane == 0) { sd[warp] = bd; si[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < 8; ++w) if (sd[w] < bd || (sd[w] == bd && si[w] < bi)) { bd = sd[w]; bi = si[w]; }
            idx[q * k + r] = bi == INT_MAX ? -1 : bi; d2[q * k + r] = bd;
            pd_s = bd; pi_s = bi;
        }
        __syncthreads();
        pd = pd_s; pi = pi_s;
    }
}

__global__ void __launch_bounds__(256)
knn_blend_kernel(const float* __restrict__ index, const int* __restrict__ idx, const float* __restrict__ d2,
                 const float* __restrict__ x, long long ldx, float* __restrict__ out, const RunParams* __restrict__ rp,
                 int C, int k, long long wIdx, long long wD2, long long wX, long long wOut, long long wRp) {
    pdl_enter();
    idx += blockIdx.z * wIdx; d2 += blockIdx.z * wD2; x += blockIdx.z * wX; out += blockIdx.z * wOut;
    rp = reinterpret_cast<const RunParams*>(reinterpret_cast<const float*>(rp) + blockIdx.z * wRp);
    const int q = blockIdx.x;
    const float rate = rp->index_rate;
    if (rate == 0.0f) {   // retrieval switched off: the features pass through untouched (no NaN can leak in from a degenerate hit)
        for (int c = threadIdx.x; c < C; c += 256) out[(long long)q * C + c] = x[(long long)q * ldx + c];
        return;
    }
    __shared__ float w[32]; __shared__ int id[32];
    if (threadIdx.x == 0) {
        // weights (1/d)^2 normalised (upstream RVC).  An exact hit (d == 0: replayed training audio, duplicate rows) would
        // give inf/inf: the distance is floored at the smallest normal float, which makes such a hit dominate the blend.
        float ws = 0.f;
        for (int i = 0; i < k; ++i) { float r = 1.0f / fmaxf(d2[q * k + i], 1e-30f); r = fminf(r, 1e18f); w[i] = r * r; ws += w[i]; id[i] = idx[q * k + i]; }
        for (int i = 0; i < k; ++i) w[i] /= ws;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float a = 0.f;
        for (int i = 0; i < k; ++i) a = fmaf(w[i], __ldg(index + (long long)id[i] * C + c), a);
        out[(long long)q * C + c] = rate * a + (1.0f - rate) * x[(long long)q * ldx + c];
    }
}

template <int CV, int QN, int G, int KK, int RR>
void scan_launch(const KnnScanOp& o, const DeviceBases& B, int q0, int nq, cudaStream_t s) {
    const long long wQ = B.ws(o.queries), wCD = B.ws(o.cand_d), wCI = B.ws(o.cand_i);
    if (G == 1 && knn_two_ctas_per_sm(nq, o.C, o.k) && (o.parts > KNN_PARTS || o.parts == o.N)) {
        auto kern2 = knn_scan_kernel<CV, QN, G, KK, RR, 0>;
        const size_t q_bytes2 = sizeof(float) * size_t(G) * QN * o.C, m_bytes2 = size_t(8) * G * QN * KK * 8;
        const size_t smem2 = q_bytes2 > m_bytes2 ? q_bytes2 : m_bytes2;
        static unsigned long long attr2 = 0;
        if (first_time_on_device(attr2)) cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        launch_k(kern2, dim3(o.parts), dim3(256), smem2, s, B.p<float>(o.index), o.N, o.C, B.p<float>(o.queries), o.ldq, nq,
                 B.p<float>(o.cand_d), B.p<int>(o.cand_i), o.parts, o.k, q0, o.Q, wQ, wCD, wCI);
        return;
    }
    auto kern = knn_scan_kernel<CV, QN, G, KK, RR, 1>;
    const size_t q_bytes = sizeof(float) * size_t(G) * QN * o.C, m_bytes = size_t(8) * G * QN * KK * 8;
    const size_t smem = q_bytes > m_bytes ? q_bytes : m_bytes;
    static unsigned long long attr = 0;
    if (first_time_on_device(attr)) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    launch_k(kern, dim3(o.parts), dim3(256), smem, s, B.p<float>(o.index), o.N, o.C, B.p<float>(o.queries), o.ldq, nq,
                                              B.p<float>(o.cand_d), B.p<int>(o.cand_i), o.parts, o.k, q0, o.Q, wQ, wCD, wCI);
}

template <int CV, int KK>
int scan_dispatch_q(const KnnScanOp& o, const DeviceBases& B, cudaStream_t s) {
    const int maxq = (KK > 8 || o.C > 384) ? 32 : 128;  // per launch: register (top-k lists) and smem (queries) budget
    int launches = 0;
    const int Qall = o.Q * B.nb;   // every window's queries in the same pass over the index
    for (int q0 = 0; q0 < Qall; q0 += maxq) {
        int nq = Qall - q0 < maxq ? Qall - q0 : maxq;
        if (nq <= 8) scan_launch<CV, 8, 1, KK, (CV <= 2 ? 4 : 2)>(o, B, q0, nq, s);
        else if (nq <= 16) scan_launch<CV, 16, 1, KK, 2>(o, B, q0, nq, s);
        else if (nq <= 32) scan_launch<CV, 32, 1, KK, (CV <= 2 ? 2 : 1)>(o, B, q0, nq, s);
        else scan_launch<CV, 32, 4, 8, (CV <= 2 ? 2 : 1)>(o, B, q0, nq, s);
        ++launches;
    }
    return launches;
}

}  // namespace

int launch_knn_scan(const KnnScanOp& o, const DeviceBases& B, cudaStream_t s) {
    // C must be a multiple of 4 and <= 1024; k <= 16 (validated when the index is set)
    if (o.k <= 8) {
        if (o.C <= 256) return scan_dispatch_q<2, 8>(o, B, s);
        if (o.C <= 768) return scan_dispatch_q<6, 8>(o, B, s);
        return scan_dispatch_q<8, 8>(o, B, s);
    }
    if (o.C <= 256) return scan_dispatch_q<2, 16>(o, B, s);
    if (o.C <= 768) return scan_dispatch_q<6, 16>(o, B, s);
    return scan_dispatch_q<8, 16>(o, B, s);
}

int launch_knn_select(const KnnSelectOp& o, const DeviceBases& B, cudaStream_t s) {
    static unsigned long long attr = 0;
    if (first_time_on_device(attr)) cudaFuncSetAttribute(knn_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int M = o.parts * o.k;
    const size_t need = size_t(M) * 8;
    const int staged = need <= 200 * 1024 ? 1 : 0;
    launch_k(knn_select_kernel, dim3(o.Q, 1, B.nb), dim3(256), staged ? need : size_t(0), s, B.p<float>(o.cand_d), B.p<int>(o.cand_i), B.p<int>(o.idx),
             B.p<float>(o.d2), M, o.k, staged, B.ws(o.cand_d), B.ws(o.cand_i), B.ws(o.idx), B.ws(o.d2));
    return 1;
}

int launch_knn_blend(const KnnBlendOp& o, const DeviceBases& B, cudaStream_t s) {
    launch_k(knn_blend_kernel, dim3(o.Q, 1, B.nb), dim3(256), size_t(0), s, B.p<float>(o.index), B.p<int>(o.idx), B.p<float>(o.d2), B.p<float>(o.x), o.ldx, B.p<float>(o.out),
                                         B.p<RunParams>(o.params), o.C, o.k, B.ws(o.idx), B.ws(o.d2), B.ws(o.x), B.ws(o.out), B.ws(o.params));
    return 1;
}

}  // namespace rvc
