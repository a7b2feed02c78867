// kernels_wstream.cu - persistent weight-streaming kernel for a chain of skinny GEMMs (RMVPE's U-Net bottleneck).
//
// The bottleneck of RMVPE (reference: rvc/src/f0/rmvpe.rs, intermediate layers of the U-Net) is 32 dependent 3x3
// convolutions over a 1 x 4 pixel map with 512 channels: as GEMMs, 4 valid rows x 512 columns x K = 1536 each, i.e. 3.1 MB
// of fp32 filters per step and almost no arithmetic.  In the generic chain kernel a step costs 11.7 us (tile prologue,
// cp.async ring of 3 k-tiles, 32 CTAs at ~5 B/clk each).  Here a step is barrier-bound:
//   * G = N / 8 CTAs; CTA c owns output columns [8c, 8c + 8) of EVERY step: 8 filter rows of K floats (6 KB each, contiguous
//     in the packed weights) = 48 KB per step, fetched with 8 `cp.async.bulk` copies into a ring of WS_STAGES stages that
//     runs up to WS_STAGES steps AHEAD of the compute (filters do not depend on activations), completion on an mbarrier;
//   * a step = grid barrier (previous step's outputs visible) -> the <= 16 KB activation span from L2 into shared memory
//     -> warp w computes column 8c + w: lanes split K, four row accumulators, shuffle reduction -> bias / ReLU / residual
//     -> 4 x 8 outputs -> arrival.
// fp32 throughout (same arithmetic type as the GEMM path; summation order over k differs, as between the tile kernels).
#include <cuda_runtime.h>

#include <cstdint>

#include "chain.h"
#include "pdl.cuh"

namespace rvc {

namespace {

constexpr int WS_THREADS = 256;          // 8 warps = the 8 columns of a CTA
constexpr int WS_NB = 8;
constexpr int WS_KMAX = 1536;            // floats per filter row of a step
constexpr int WS_A_FLOATS = 4096;        // activation span: (M - 1) * lda + K
constexpr int WS_STAGES = 3;
constexpr int WS_STAGE_FLOATS = WS_NB * WS_KMAX;
constexpr int WS_SMEM_BYTES = (WS_STAGES * WS_STAGE_FLOATS + WS_A_FLOATS) * 4 + 128;
constexpr int WS_ROWS = 4;               // valid rows of a step (the others are halo pixels)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void wait_grid(const unsigned int* bar, unsigned int target) {
    const long long t0 = clock64();
    while (true) {
        unsigned int v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        if (v >= target) break;
        __nanosleep(32);
        if (clock64() - t0 > (6ll << 30)) __trap();
    }
}
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(WS_THREADS) wstream_kernel(const WsOpDev* __restrict__ ops, int n_ops, unsigned int* bar, unsigned long long* dbg) {
    extern __shared__ __align__(128) float sm[];
    float* wst = sm;                                   // [WS_STAGES][WS_NB][K]
    float* as = sm + WS_STAGES * WS_STAGE_FLOATS;      // activation span
    __shared__ __align__(8) uint64_t full[WS_STAGES];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned int G = gridDim.x;
    const int n0 = blockIdx.x * WS_NB;

    if (tid == 0) {
        for (int s = 0; s < WS_STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // filters of the first WS_STAGES steps: they depend on nothing upstream
    auto request = [&](int step) {
        const WsOpDev& o = ops[step];
        const int s = step % WS_STAGES;
        const uint32_t row_bytes = uint32_t(o.K) * 4u;
        mbar_expect_tx(&full[s], row_bytes * WS_NB);
        for (int r = 0; r < WS_NB; ++r)
            bulk_load(wst + s * WS_STAGE_FLOATS + r * o.K, o.W + (long long)(n0 + r) * o.ldw, row_bytes, &full[s]);
    };
    if (tid == 0)
        for (int st = 0; st < WS_STAGES && st < n_ops; ++st) request(st);
    pdl_launch_dependents();
    pdl_wait();

    for (int step = 0; step < n_ops; ++step) {
        const WsOpDev& o = ops[step];
        if (step > 0) {
            if (tid == 0) wait_grid(bar, (unsigned int)step * G);
            __syncthreads();
        }
        if (blockIdx.x == 0 && tid == 0 && dbg) dbg[step] = gtime();
        // activation span (written by other SMs during the previous steps: L2, not L1)
        for (int i = tid * 4; i < o.a_floats; i += WS_THREADS * 4)
            *reinterpret_cast<float4*>(as + i) = __ldcg(reinterpret_cast<const float4*>(o.A + i));
        __syncthreads();
        mbar_wait(&full[step % WS_STAGES], uint32_t(step / WS_STAGES) & 1u);
        {
            const float* wr = wst + (step % WS_STAGES) * WS_STAGE_FLOATS + warp * o.K;
            float acc[WS_ROWS] = {0.f, 0.f, 0.f, 0.f};
            for (int k = lane * 4; k < o.K; k += 128) {
                const float4 w = *reinterpret_cast<const float4*>(wr + k);
#pragma unroll
                for (int m = 0; m < WS_ROWS; ++m) {
                    const float4 a = *reinterpret_cast<const float4*>(as + o.row_off[m] + k);
                    acc[m] = fmaf(a.x, w.x, acc[m]); acc[m] = fmaf(a.y, w.y, acc[m]);
                    acc[m] = fmaf(a.z, w.z, acc[m]); acc[m] = fmaf(a.w, w.w, acc[m]);
                }
            }
#pragma unroll
            for (int m = 0; m < WS_ROWS; ++m)
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], off);
            if (lane < o.n_rows) {
                const int n = n0 + warp, m = o.row[lane];
                float v = (lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3]) + (o.bias ? __ldg(o.bias + n) : 0.f);
                if (o.relu) v = fmaxf(v, 0.f);
                if (o.R) v += __ldcg(o.R + (long long)m * o.ldr + n);
                __stcg(o.C + (long long)m * o.ldc + n, v);
            }
        }
        __syncthreads();   // every warp is done with this stage and with the activation span
        if (tid == 0) {
            if (step + WS_STAGES < n_ops) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of the stage before the bulk copy overwrites it
                request(step + WS_STAGES);
            }
            if (step + 1 < n_ops) { __threadfence(); atomicAdd(bar, 1u); }
        }
    }
    if (blockIdx.x == 0 && tid == 0 && dbg) dbg[n_ops] = gtime();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(bar + 32, 1u) == G - 1) { bar[0] = 0; bar[32] = 0; __threadfence(); }
    }
}

}  // namespace

bool wstream_shape_ok(int M, int N, int K, long long lda, int valid_rows) {
    return N % WS_NB == 0 && N / WS_NB >= 2 && N / WS_NB <= 96 && K % 128 == 0 && K <= WS_KMAX && lda % 4 == 0 &&
           (long long)(M - 1) * lda + K <= WS_A_FLOATS && valid_rows >= 1 && valid_rows <= WS_ROWS;
}

void init_wstream_attributes() { cudaFuncSetAttribute(wstream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM_BYTES); }

int launch_wstream(const ChainDev& c, cudaStream_t stream) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(unsigned(c.grid)); cfg.blockDim = dim3(WS_THREADS); cfg.dynamicSmemBytes = WS_SMEM_BYTES; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident or none: the grid barrier cannot deadlock
    attr[0].val.cooperative = 1;
    attr[1].id = cudaLaunchAttributePriority;
    attr[1].val.priority = g_launch_priority;
    cfg.attrs = attr; cfg.numAttrs = g_launch_priority != 0 ? 2 : 1;
    const WsOpDev* ops = c.d_wsops; int n = c.n_ops; unsigned int* bar = c.d_bar; unsigned long long* dbg = c.d_dbg;
    cudaLaunchKernelEx(&cfg, wstream_kernel, ops, n, bar, dbg);
    return 1;
}

}  // namespace rvc
