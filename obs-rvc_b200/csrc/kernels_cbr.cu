// kernels_cbr.cu - RMVPE ConvBlockRes at the two full-resolution U-Net levels as ONE kernel.
//
// relu(bn(conv3x3(relu(bn(conv3x3(x)))))) + (conv1x1(x) | x)   (reference: rmvpe.rs ConvBlockRes; BatchNorm folded
// into the weights when the model is packed, model.cpp pack_convblockres).  At 16 / 32 channels the two 3x3 convs
// are a few MFLOP over a 32x128 / 16x64 map: as implicit GEMMs they cost two (three with the shortcut) dependent
// launches of ~6-8 us each, nearly all of it launch + pipeline latency.  Here one CTA owns a strip of output pixels,
// keeps the input strip with a 2-pixel halo, both filter banks and the intermediate strip (1-pixel halo, recomputed
// per CTA) in shared memory and runs both convs back to back in fp32 (same arithmetic type as the GEMM path; the
// summation order over k differs, as it already does between the tile kernels).
//
// Thread mapping of a conv pass: lane = (channel group cg, k-slice ks, pixel slot); a thread accumulates 4 output
// channels {cg + q * C/4} of PXT pixels over its slice of the 9 * Cs / 4 k-quads; k-slices are summed with shuffles.
// Shared-memory layouts are chosen for conflict-free 128-bit loads: activations [pixel][Cs + 4], filters
// [k-quad][C + 1] float4 with the four k of a quad innermost (adjacent lanes read adjacent channels).
#include <cuda_runtime.h>

#include <cstdint>

#include "launch.h"
#include "pdl.cuh"

namespace rvc {

namespace {

constexpr int CBR_THREADS = 512;   // 16 warps: four per scheduler to cover the shared-memory latency

struct CbrParams {
    const float* in;      // halo-padded NHWC map [T + 2][F + 2][Cin]
    float* dst;           // interior pixel (0, 0) of the destination map, pixel stride ld_dst
    const float* w1; const float* b1;   // [C][9 * Cin], [C]
    const float* w2; const float* b2;   // [C][9 * C], [C]
    const float* wsc; const float* bsc; // [C][Cin], [C] or null (identity residual, Cin == C)
    long long ld_dst, wIn, wDst;
    int T, F, Cin;
};

__device__ __forceinline__ void cp16(void* dst, const void* src) {
    unsigned d = unsigned(__cvta_generic_to_shared(dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// filters W[C][K] (K = k-quads * 4) -> shared [K / 4][C + 1] float4
template <int C>
__device__ __forceinline__ void stage_filters(float4* ws, const float* __restrict__ W, int K) {
    const int nq = K >> 2;
    for (int i = threadIdx.x; i < nq * C; i += CBR_THREADS) {
        const int co = i % C, kq = i / C;
        cp16(ws + kq * (C + 1) + co, W + (long long)co * K + kq * 4);
    }
}

// One 3x3 conv over a shared-memory strip.  src: [rows][SW][Cs + 4]; output pixels p in [0, NP), p = (p / PW, p % PW)
// on the strip shifted by one pixel (the 3x3 window of output (r, c) starts at source (r, c)).  The k loop is flat over
// the 9 * Cs / 4 k-quads (filter row = k-quad index) and software-pipelined: the operands of quad i + 1 are requested
// before the FMAs of quad i - with two to four warps per scheduler the shared-memory latency is otherwise exposed
// (ncu, first version: short-scoreboard stall 3.4 per issue, 0.26 instructions / clock / scheduler).
template <int C, int PXT, int KS, typename Emit>
__device__ __forceinline__ void conv_pass(const float* __restrict__ src, int SW, int Cs, const float4* __restrict__ w,
                                          int NP, int PW, Emit&& emit) {
    constexpr int CG = C / 4, NSLOT = CBR_THREADS / (CG * KS);
    static_assert(CG * KS <= 32, "k-slices are summed with warp shuffles");
    const int cg = threadIdx.x % CG, ks = (threadIdx.x / CG) % KS, slot = threadIdx.x / (CG * KS);
    const int Q = Cs >> 2, lq = 31 - __clz(Q), NI = 9 * Q, ps = Cs + 4;
    const int i0 = ks * NI / KS, i1 = (ks + 1) * NI / KS;
    const float4* wc = w + cg;
    for (int base = 0; base < NP; base += NSLOT * PXT) {
        float acc[PXT][4];
        const float* sp[PXT];
#pragma unroll
        for (int i = 0; i < PXT; ++i) {
            const int p = min(base + slot + i * NSLOT, NP - 1);
            sp[i] = src + ((p / PW) * SW + p % PW) * ps;
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[i][q] = 0.f;
        }
        float4 an[PXT], wn[4];
        auto request = [&](int it) {
            const int tap = it >> lq, q4 = it & (Q - 1), kt = (tap * 11) >> 5, kf = tap - 3 * kt;
            const int toff = (kt * SW + kf) * ps + q4 * 4;
#pragma unroll
            for (int i = 0; i < PXT; ++i) an[i] = *reinterpret_cast<const float4*>(sp[i] + toff);
#pragma unroll
            for (int q = 0; q < 4; ++q) wn[q] = wc[it * (C + 1) + q * CG];
        };
        request(i0);
#pragma unroll 2
        for (int it = i0; it < i1; ++it) {
            float4 a[PXT], ww[4];
#pragma unroll
            for (int i = 0; i < PXT; ++i) a[i] = an[i];
#pragma unroll
            for (int q = 0; q < 4; ++q) ww[q] = wn[q];
            if (it + 1 < i1) request(it + 1);
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int i = 0; i < PXT; ++i) {
                    acc[i][q] = fmaf(a[i].x, ww[q].x, acc[i][q]);
                    acc[i][q] = fmaf(a[i].y, ww[q].y, acc[i][q]);
                    acc[i][q] = fmaf(a[i].z, ww[q].z, acc[i][q]);
                    acc[i][q] = fmaf(a[i].w, ww[q].w, acc[i][q]);
                }
        }
        if (KS > 1) {
#pragma unroll
            for (int o = CG; o < CG * KS; o <<= 1)
#pragma unroll
                for (int i = 0; i < PXT; ++i)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[i][q] += __shfl_xor_sync(0xffffffffu, acc[i][q], o);
        }
        if (ks == 0) {
#pragma unroll
            for (int i = 0; i < PXT; ++i) {
                const int p = base + slot + i * NSLOT;
                if (p < NP) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) emit(p, cg + q * CG, acc[i][q]);
                }
            }
        }
    }
}

// grid (F / TC, T, nb): one output strip of 1 x TC pixels per CTA
template <int C, int TC, int PXT1, int KS1, int KS2>
__global__ void __launch_bounds__(CBR_THREADS) cbr_kernel(CbrParams p) {
    extern __shared__ __align__(16) float sm[];
    constexpr int IW = TC + 4, IH = 5, MW = TC + 2, MH = 3;
    const int Cin = p.Cin, ips = Cin + 4;
    float* in_s = sm;
    float* t1_s = in_s + IH * IW * ips;
    float4* w1_s = reinterpret_cast<float4*>(t1_s + MH * MW * (C + 4));
    float4* w2_s = w1_s + 9 * (Cin >> 2) * (C + 1);
    float4* wsc_s = w2_s + 9 * (C >> 2) * (C + 1);
    const int c0 = blockIdx.x * TC, r0 = blockIdx.y;
    const float* in = p.in + blockIdx.z * p.wIn;
    float* dst = p.dst + blockIdx.z * p.wDst;

    pdl_launch_dependents();
    // the filters do not depend on the upstream kernel
    stage_filters<C>(w1_s, p.w1, 9 * Cin);
    if (p.wsc) stage_filters<C>(wsc_s, p.wsc, Cin);
    cp_commit();
    pdl_wait();
    {   // input strip with a 2-pixel halo; pixels outside the padded map are zero
        const int Q = Cin >> 2;
        for (int i = threadIdx.x; i < IH * IW * Q; i += CBR_THREADS) {
            const int q = i % Q, px = i / Q, pr = px / IW, pc = px % IW;
            const int gr = r0 - 1 + pr, gc = c0 - 1 + pc;   // padded-map coordinates
            float* d = in_s + px * ips + q * 4;
            if (gr >= 0 && gr < p.T + 2 && gc >= 0 && gc < p.F + 2) cp16(d, in + ((long long)gr * (p.F + 2) + gc) * Cin + q * 4);
            else *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    cp_commit();
    stage_filters<C>(w2_s, p.w2, 9 * C);
    cp_commit();
    cp_wait<1>();
    __syncthreads();

    // conv 1 over the strip grown by one pixel; positions outside the map are conv 2's zero padding
    conv_pass<C, PXT1, KS1>(in_s, IW, Cin, w1_s, MH * MW, MW, [&](int px, int co, float v) {
        const int r = r0 - 1 + px / MW, c = c0 - 1 + px % MW;
        const bool inside = r >= 0 && r < p.T && c >= 0 && c < p.F;
        t1_s[px * (C + 4) + co] = inside ? fmaxf(v + __ldg(p.b1 + co), 0.f) : 0.f;
    });
    cp_wait<0>();
    __syncthreads();

    conv_pass<C, 1, KS2>(t1_s, MW, C, w2_s, TC, TC, [&](int px, int co, float v) {
        const int c = c0 + px;
        if (c >= p.F) return;
        const float* x = in_s + (2 * IW + px + 2) * ips;   // the block's input at this pixel
        float res;
        if (p.wsc) {
            res = __ldg(p.bsc + co);
            for (int q4 = 0; q4 < (Cin >> 2); ++q4) {
                const float4 a = *reinterpret_cast<const float4*>(x + q4 * 4);
                const float4 ww = wsc_s[q4 * (C + 1) + co];
                res = fmaf(a.x, ww.x, res); res = fmaf(a.y, ww.y, res); res = fmaf(a.z, ww.z, res); res = fmaf(a.w, ww.w, res);
            }
        } else {
            res = x[co];
        }
        dst[((long long)r0 * (p.F + 2) + c) * p.ld_dst + co] = fmaxf(v + __ldg(p.b2 + co), 0.f) + res;
    });
}

template <int C, int TC>
size_t cbr_smem(int Cin, bool sc) {
    return size_t(4) * (size_t(5) * (TC + 4) * (Cin + 4) + size_t(3) * (TC + 2) * (C + 4)) +
           size_t(16) * (size_t(9) * (Cin / 4) + size_t(9) * (C / 4) + (sc ? Cin / 4 : 0)) * (C + 1);
}

constexpr int CBR_TC16 = 32, CBR_TC32 = 8;

}  // namespace

void init_cbr_attributes() {
    cudaFuncSetAttribute(cbr_kernel<16, CBR_TC16, 1, 1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(cbr_smem<16, CBR_TC16>(64, true)));
    cudaFuncSetAttribute(cbr_kernel<32, CBR_TC32, 1, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(cbr_smem<32, CBR_TC32>(64, true)));
}

int launch_cbr(const CbrOp& o, const DeviceBases& B, cudaStream_t s) {
    CbrParams p;
    p.in = B.p<float>(o.in); p.dst = B.p<float>(o.dst);
    p.w1 = B.p<float>(o.w1); p.b1 = B.p<float>(o.b1); p.w2 = B.p<float>(o.w2); p.b2 = B.p<float>(o.b2);
    p.wsc = B.p<float>(o.wsc); p.bsc = B.p<float>(o.bsc);
    p.ld_dst = o.ld_dst; p.wIn = B.ws(o.in); p.wDst = B.ws(o.dst);
    p.T = o.T; p.F = o.F; p.Cin = o.Cin;
    const bool sc = !o.wsc.null();
    if (o.C == 16)
        launch_k(cbr_kernel<16, CBR_TC16, 1, 1, 4>, dim3(o.F / CBR_TC16, o.T, B.nb), dim3(CBR_THREADS), cbr_smem<16, CBR_TC16>(o.Cin, sc), s, p);
    else
        launch_k(cbr_kernel<32, CBR_TC32, 1, 2, 4>, dim3(o.F / CBR_TC32, o.T, B.nb), dim3(CBR_THREADS), cbr_smem<32, CBR_TC32>(o.Cin, sc), s, p);
    return 1;
}

}  // namespace rvc
