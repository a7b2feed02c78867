// engine.cu - context, weight residency, plan cache, CUDA-graph replay and the C ABI
// (include/rvc_b200.h).  Replaces rvc::RvcInfer (reference rvc/src/rvc.rs:18-220) and the
// rvc-rpc dispatch around it (rvc-rpc/src/main.rs:56-101) with an in-process engine:
//   * weights are packed once (model.cpp) and stay resident in HBM;
//   * per (call kind, geometry) an op list is built (plan.cpp), its work arena allocated and
//     zeroed once (halo invariants), and after one eager run the whole window is captured into a
//     CUDA graph spanning up to three stream lanes (F0 chain || ContentVec, 3 ResBlocks);
//   * per-call variability (pitch shift, window counter, index rate) travels in a 64-byte
//     parameter block written by a one-thread kernel, so graph replays need no re-instantiation;
//   * the pitch cache (rvc.rs:26,168-179) is per-context device state.
// There is no CPU execution path: every op is a kernel launch and a missing device is an error.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sys/stat.h>

#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rvc_b200.h"
#include "chain.h"
#include "cvstack.h"
#include "launch.h"
#include "model.h"
#include "pdl.cuh"
#include "resample.h"

using namespace rvc;

static bool g_sync_each = false;
namespace rvc { bool g_use_pdl = false; thread_local bool g_pdl_op = false; thread_local int g_launch_priority = 0; }  // measured: no gain on this path (profiles/README), opt-in with RVC_PDL=1

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    uint8_t* d = nullptr; size_t bytes = 0;
    void release() { if (d) cudaFree(d); d = nullptr; bytes = 0; }
};

// Immutable weights of one model file, resident in HBM.  Shared (by path) between all contexts of a
// process on the same device: independent streams replicate state, never weights.
struct ModelData {
    Packed packed;  // offsets kept; host copy dropped after upload
    DevBuf dev;
    int64_t hilo_stride = 0;  // bytes from the fp32 arena to its tf32 hi copy (and again to lo); 0 = none
    int64_t hilo16_off = 0, hilo16_plane = 0;  // fp16 split planes (launch.h DeviceBases)
    CvInfo cvi; F0Info f0i; SynInfo syi;
    int rows = 0, cols = 0;   // retrieval index only
    int64_t planes_off = 0; float ymax2 = 0.f;   // retrieval index: fp16 planes behind the fp32 rows (kernels_knn_umma.cu)
    ~ModelData() { dev.release(); }
};

struct Model {
    std::shared_ptr<ModelData> d;
    bool loaded = false;
    Packed& packed_ref() { return d->packed; }
    void unload() { d.reset(); loaded = false; }
};

std::mutex g_model_mu;
std::map<std::string, std::weak_ptr<ModelData>> g_model_cache;

std::string model_key(int device, const std::string& path, bool hilo) {
    struct stat st{};
    long long sz = 0, mt = 0;
    if (stat(path.c_str(), &st) == 0) { sz = (long long)st.st_size; mt = (long long)st.st_mtime; }
    return std::to_string(device) + "|" + path + "|" + std::to_string(sz) + "|" + std::to_string(mt) + (hilo ? "|hilo" : "");
}

std::shared_ptr<ModelData> cache_lookup(const std::string& key) {
    std::lock_guard<std::mutex> lk(g_model_mu);
    auto it = g_model_cache.find(key);
    if (it == g_model_cache.end()) return nullptr;
    return it->second.lock();
}
void cache_store(const std::string& key, const std::shared_ptr<ModelData>& d) {
    std::lock_guard<std::mutex> lk(g_model_mu);
    g_model_cache[key] = d;
}

struct PlanKey {
    int kind; Geometry g; int with_index; int index_rows; int k;
    int chains = 1;   // persistent chains on (single live stream on the device) or off (several streams share it)
    int nb = 1, sequential = 0;   // batched plans: windows per launch; consecutive windows of one stream or independent streams
    bool operator<(const PlanKey& o) const {
        if (kind != o.kind) return kind < o.kind;
        if (chains != o.chains) return chains < o.chains;
        if (nb != o.nb) return nb < o.nb;
        if (sequential != o.sequential) return sequential < o.sequential;
        if (g < o.g) return true;
        if (o.g < g) return false;
        if (with_index != o.with_index) return with_index < o.with_index;
        if (index_rows != o.index_rows) return index_rows < o.index_rows;
        return k < o.k;
    }
};

struct PlanEntry {
    Plan plan;
    DevBuf work;
    DevBuf bstate;   // batched plans: nb state blocks (params | pitch cache | pcm | audio), plan.state_block bytes apart
    std::vector<ChainDev> chains;   // device tables of plan.chains (pointers resolved against this entry's arenas)
    CvsDev cvs;                     // device tables of plan.cvstack (grid == 0: the stack runs as separate kernels)
    cudaGraphExec_t exec = nullptr;
    cudaGraph_t graph = nullptr;
    int runs = 0;
    int launches = 0;
    ~PlanEntry() {
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
        for (ChainDev& c : chains) { cudaFree(c.d_ops); cudaFree(c.d_phases); cudaFree(c.d_bar); cudaFree(c.d_dbg); cudaFree(c.d_slabs); cudaFree(c.d_chunk_bytes); cudaFree(c.d_wsops); }
        cudaFree(cvs.d_maps); cudaFree(cvs.d_phases); cudaFree(cvs.d_bar);
        work.release(); bstate.release();
    }
};

constexpr int MAX_LANES = 4;

__global__ void set_params_kernel(RunParams* p, float uppower, float index_rate, unsigned long long seed,
                                  unsigned long long window, int noise_mode) {
    pdl_enter();
    p->uppower = uppower; p->index_rate = index_rate; p->noise_seed = seed; p->window = window; p->noise_mode = noise_mode;
}

// Batched plans: one parameter block per window.  Windows of one stream (sequential) count the call counter up;
// independent streams bring their own seed and counter.
struct BatchSeeds { unsigned long long seed[64]; unsigned long long window[64]; };
__global__ void set_params_batch_kernel(uint8_t* state, long long block_bytes, int nb, float uppower, float index_rate, int noise_mode,
                                        BatchSeeds bs) {
    pdl_enter();
    const int w = threadIdx.x;
    if (w >= nb) return;
    RunParams* p = reinterpret_cast<RunParams*>(state + w * block_bytes + StateLayout::off_params);
    p->uppower = uppower; p->index_rate = index_rate; p->noise_seed = bs.seed[w]; p->window = bs.window[w]; p->noise_mode = noise_mode;
}

// window w of a batched offline call = src[w * hop, w * hop + n) (the reference's streaming geometry: every call sees the
// last n samples, advanced by hop): copied into the pcm slot of state block w
__global__ void gather_windows_kernel(const float* __restrict__ src, uint8_t* state, long long block_bytes, long long off_pcm, int n, int hop) {
    pdl_enter();
    const int w = blockIdx.y;
    const float* s = src + (long long)w * hop;
    float* d = reinterpret_cast<float*>(state + w * block_bytes + off_pcm);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) d[i] = s[i];
}

// pitch caches of nb independent streams <-> the cache slots of the state blocks (dir 0: stream -> block, 1: block -> stream)
struct CachePtrs { float* p[64]; };
__global__ void move_caches_kernel(CachePtrs cp, uint8_t* state, long long block_bytes, int dir) {
    pdl_enter();
    float* blk = reinterpret_cast<float*>(state + blockIdx.x * block_bytes + StateLayout::off_cache);
    float* own = cp.p[blockIdx.x];
    for (int i = threadIdx.x; i < int(StateLayout::CACHE_LEN); i += blockDim.x) {
        if (dir == 0) blk[i] = own[i]; else own[i] = blk[i];
    }
}

__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
    pdl_enter();
    __shared__ float tile[32][33];
    int c = blockIdx.x * 32 + threadIdx.x, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) if (r0 + i < rows && c < cols) tile[i][threadIdx.x] = in[(long long)(r0 + i) * cols + c];
    __syncthreads();
    int r = r0 + threadIdx.x, c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += 8) if (c0 + i < cols && r < rows) out[(long long)(c0 + i) * rows + r] = tile[threadIdx.x][i];
}

}  // namespace

// Device-resident state of the streaming loop around the call (obs-rvc/src/lib.rs RvcInferenceState, :88-130, sizes
// :200-245): ring buffers at the OBS rate and at 16 kHz, the two rubato resamplers as polyphase tables + overlaps, the
// SOLA buffer.  One per context (one OBS filter instance drives one engine).
struct ResamplerDev {
    rvc::ResampleTable t;          // host copy of the geometry (table moved to the device)
    float* kappa = nullptr;
    float* overlap[2] = {nullptr, nullptr};
    int cur = 0;
    void release() { cudaFree(kappa); cudaFree(overlap[0]); cudaFree(overlap[1]); kappa = nullptr; overlap[0] = overlap[1] = nullptr; }
};
struct StreamState {
    bool open = false;
    rvc_stream_config cfg{};
    int zc = 0, sample_frame_time = 0, sample_frame_size = 0, sample_frame_16k = 0, crossfade = 0, sola_buf = 0, sola_search = 0,
        extra = 0, in_size = 0, in16k_size = 0, ret_len = 0, ret_size = 0, model_sr = 0, skip_head = 0, up_out = 0;
    ResamplerDev down, up;
    float* inbuf[2] = {nullptr, nullptr};
    float* in16k[2] = {nullptr, nullptr};
    int cur = 0;
    float* scratch = nullptr;      // frame in | down out | model out | up out | rms1 | rms2 | cor | offset | sola | block
    size_t scratch_floats = 0;
    uint64_t frames = 0;
    void release() {
        down.release(); up.release();
        for (int i = 0; i < 2; ++i) { cudaFree(inbuf[i]); cudaFree(in16k[i]); inbuf[i] = in16k[i] = nullptr; }
        cudaFree(scratch); scratch = nullptr; open = false;
    }
};

struct rvc_ctx {
    StreamState stream;
    std::map<std::string, ResamplerDev> resamplers;   // rvc_resample_chunk: tables by (fs_in, fs_out, chunk)
    std::string data_path, err;
    rvc_config cfg{};
    cudaStream_t streams[MAX_LANES] = {nullptr};
    std::vector<cudaEvent_t> events;
    Model cv, f0, syn;
    CvInfo cvi; F0Info f0i; SynInfo syi;
    Model index; int index_rows = 0, index_c = 0; float index_rate = 0.f;
    DevBuf state;
    std::map<PlanKey, std::unique_ptr<PlanEntry>> plans;
    PlanEntry* last = nullptr;
    uint64_t window = 0, total_launches = 0;
    cudaEvent_t timers[8] = {nullptr};
    bool allow_umma = true;
    int chain_grid_main = 0, chain_grid_side = 0, chain_side_max_m = 8;
    int f0_priority = 0;        // launch priority of the F0 lanes' kernels (highest stream priority of the device; RVC_F0_PRIO=0 disables)
    bool chain_force = false;   // RVC_CHAIN=2: keep chains even when several contexts share the device
    int cvstack_grid = 64;      // CTAs of the persistent ContentVec stack kernel (RVC_CVSTACK=0 disables, RVC_CVSTACK_G sets the grid)
    bool cvstack_all = false;   // RVC_CVSTACK=1: also for the hubert / feature plans
    bool knn_umma = true;       // RVC_KNN_UMMA=0: retrieval always on the fp32 scan (kernels_knn.cu)

    int fail(int code, const std::string& m) { err = m; return code; }
    int cuda_fail(cudaError_t e, const char* what) {
        err = std::string(what) + ": " + cudaGetErrorString(e);
        return RVC_ERR_CUDA;
    }
    void drop_plans() { sync_all(); plans.clear(); last = nullptr; }
    void sync_all() { for (auto s : streams) if (s) cudaStreamSynchronize(s); }
    DeviceBases bases(const PlanEntry& e) const {
        DeviceBases B;
        B.b[SP_CV] = cv.d ? cv.d->dev.d : nullptr; B.b[SP_F0] = f0.d ? f0.d->dev.d : nullptr;
        B.b[SP_SYN] = syn.d ? syn.d->dev.d : nullptr; B.b[SP_IDX] = index.d ? index.d->dev.d : nullptr;
        B.b[SP_WORK] = e.work.d; B.b[SP_STATE] = state.d;
        if (e.plan.nb > 1) {
            B.nb = e.plan.nb; B.b[SP_STATE] = e.bstate.d;
            B.bstride[SP_WORK] = e.plan.work_bytes; B.bstride[SP_STATE] = e.plan.state_block;
        }
        B.hilo_stride[SP_CV] = cv.d ? cv.d->hilo_stride : 0; B.hilo_stride[SP_SYN] = syn.d ? syn.d->hilo_stride : 0;
        B.hilo_stride[SP_F0] = f0.d ? f0.d->hilo_stride : 0; B.hilo16_off[SP_F0] = f0.d ? f0.d->hilo16_off : 0; B.hilo16_plane[SP_F0] = f0.d ? f0.d->hilo16_plane : 0;
        B.hilo16_off[SP_CV] = cv.d ? cv.d->hilo16_off : 0; B.hilo16_off[SP_SYN] = syn.d ? syn.d->hilo16_off : 0;
        B.hilo16_plane[SP_CV] = cv.d ? cv.d->hilo16_plane : 0; B.hilo16_plane[SP_SYN] = syn.d ? syn.d->hilo16_plane : 0;
        return B;
    }
};

namespace {

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return ctx->cuda_fail(e_, #call); } while (0)

int upload(rvc_ctx* ctx, ModelData& m, bool with_hilo) {
    size_t bytes = (m.packed.host.size() * sizeof(float) + 1023) & ~size_t(1023);
    m.dev.release();
    // layout: fp32 arena | tf32 hi | tf32 lo | fp16 hi plane, fp16 scaled-lo plane (one more arena's worth)
    CK(cudaMalloc(&m.dev.d, bytes * (with_hilo ? 4 : 1) + 256));
    m.dev.bytes = bytes;
    CK(cudaMemcpy(m.dev.d, m.packed.host.data(), m.packed.host.size() * sizeof(float), cudaMemcpyHostToDevice));
    if (with_hilo) {
        // tf32 hi / lo copies of the whole arena for the tcgen05 3xTF32 GEMMs (kernels_umma.cu)
        launch_split_hilo(reinterpret_cast<const float*>(m.dev.d), reinterpret_cast<float*>(m.dev.d + bytes),
                          reinterpret_cast<float*>(m.dev.d + 2 * bytes), m.packed.host.size(), ctx->streams[0]);
        m.hilo_stride = int64_t(bytes);
        // 2-term fp16 split planes for the tcgen05 kind::f16 path
        launch_split_hilo16(reinterpret_cast<const float*>(m.dev.d), reinterpret_cast<unsigned short*>(m.dev.d + 3 * bytes),
                            reinterpret_cast<unsigned short*>(m.dev.d + 3 * bytes + bytes / 2), m.packed.host.size(), ctx->streams[0]);
        m.hilo16_off = int64_t(3 * bytes); m.hilo16_plane = int64_t(bytes / 2);
    }
    CK(cudaDeviceSynchronize());
    std::vector<float>().swap(m.packed.host);
    return RVC_OK;
}

int issue_one(rvc_ctx* ctx, const Op& op, const DeviceBases& B, cudaStream_t s, int* n) {
    // RMVPE residual block run by one kernel: launched where the block's last GEMM stood, the other ops are covered
    if (op.fuse == 1) return RVC_OK;
    if (op.fuse == 2) { *n += launch_cbr(op.cbr, B, s); return RVC_OK; }
    switch (op.kind) {
        case OP_GEMM: *n += launch_gemm(op.gemm, B, s); break;
        case OP_LAYERNORM: *n += launch_layernorm(op.ln, B, s); break;
        case OP_ATTN: *n += launch_attn(op.attn, B, s); break;
        case OP_RELATTN: *n += launch_relattn(op.relattn, B, s); break;
        case OP_CONV0_STATS: *n += launch_conv0_stats(op.c0s, B, s); break;
        case OP_CONV0_APPLY: *n += launch_conv0_apply(op.c0a, B, s); break;
        case OP_STFTMEL: *n += launch_stftmel(op.stft, B, s); break;
        case OP_AVGPOOL: *n += launch_avgpool(op.pool, B, s); break;
        case OP_GRU: *n += launch_gru(op.gru, B, s); break;
        case OP_F0DECODE: *n += launch_f0decode(op.f0d, B, s); break;
        case OP_F0POST: *n += launch_f0post(op.f0p, B, s); break;
        case OP_EMBED: *n += launch_embed(op.embed, B, s); break;
        case OP_ZP: *n += launch_zp(op.zp, B, s); break;
        case OP_SINEGEN: *n += launch_sinegen(op.sine, B, s); break;
        case OP_AVG3: *n += launch_avg3(op.avg3, B, s); break;
        case OP_CONVPOST: *n += launch_convpost(op.cpost, B, s); break;
        case OP_KNN_SCAN:
            if (op.kd.umma) {
                const int l = launch_knn_scan_umma(op.kd, B, s);
                if (l < 0) return ctx->fail(RVC_ERR_CUDA, "kNN tensor maps could not be encoded");
                *n += l;
            } else *n += launch_knn_scan(op.kd, B, s);
            break;
        case OP_KNN_SELECT: *n += op.ks.rerank ? launch_knn_rerank(op.ks, B, s) : launch_knn_select(op.ks, B, s); break;
        case OP_KNN_BLEND: *n += launch_knn_blend(op.kb, B, s); break;
        case OP_GATHER_ROWS: *n += launch_gather_rows(op.gather, B, s); break;
        case OP_FILL: CK(cudaMemsetAsync(B.p<uint8_t>(op.fill.dst), 0, op.fill.bytes, s)); break;
        default: break;
    }
    return RVC_OK;
}

// NVTX ranges per stage of the window - the reference's own timers (rvc.rs:145 hubert, :157 / :184 pitch, :217 synthesizer)
// as profiler ranges (RVC_NVTX=1).  Host-side: they bracket the enqueue of a stage's kernels (eager pass / graph capture) and
// the whole call; device time per stage comes from rvc_profile_timeline (RVC_TL_MARKS).
static bool g_nvtx = false;
static const char* stage_of(const Op& op) {
    const std::string& n = op.name;
    if (n.compare(0, 3, "cv.") == 0) return "contentvec (rvc.rs:81-97)";
    if (n.compare(0, 3, "rm.") == 0 || n == "mel" || n == "f0") return "rmvpe f0 (rmvpe.rs:225-261)";
    if (n.compare(0, 3, "knn") == 0 || n == "phone") return "retrieval (rvc.rs:159)";
    if (n == "pitch") return "pitch cache (rvc.rs:167-181)";
    if (n.compare(0, 3, "sy.") == 0) return "synthesizer (rvc.rs:193-214)";
    return nullptr;
}

// RVC_PDL_OPS=pattern,pattern,...: programmatic dependent launch for the ops whose names match (the kernel may be
// scheduled while its predecessor on the lane still runs and waits at griddepcontrol.wait).  Which ops gain is an
// empirical matter (profiles/README.md, tools/pdl_search.py): a dependent launch costs 1.3-3 us more while ANY other grid
// is resident, PDL hides that for the F0 encoder and the vocoder, and costs elsewhere.
#define RVC_PDL_OPS_DEFAULT "rm.enc,rm.dec0,rm.dec1,rm.dec*c2,sy.,knn_scan,phone"   // tools/pdl_search.py: 2.74 -> 2.66-2.67 ms / window
static std::vector<std::string> pdl_patterns(int nb) {
    std::vector<std::string> v;
    const char* e = getenv("RVC_PDL_OPS");
    // the default set was searched on the single-window plan (tools/pdl_search.py); batched plans (throughput-bound
    // kernels whose early-launched successors take SM slots) want less: 8 streams 1111 -> 1133, 32 offline windows
    // 1378 -> 1388 windows/s with the RMVPE encoder + synthesizer only
    std::string s(e ? e : (nb > 1 ? "rm.enc,sy." : RVC_PDL_OPS_DEFAULT)), t;
    for (char c : s + ",") { if (c == ',') { if (!t.empty()) v.push_back(t); t.clear(); } else t += c; }
    return v;
}
// "prefix" or "prefix*suffix"
static bool pdl_for(const std::vector<std::string>& pre, const std::string& name) {
    for (const std::string& p : pre) {
        const size_t star = p.find('*');
        if (star == std::string::npos) { if (name.compare(0, p.size(), p) == 0) return true; continue; }
        const std::string a = p.substr(0, star), b = p.substr(star + 1);
        if (name.size() >= a.size() + b.size() && name.compare(0, a.size(), a) == 0 && name.compare(name.size() - b.size(), b.size(), b) == 0) return true;
    }
    return false;
}

int issue_ops(rvc_ctx* ctx, PlanEntry& e, int* launches) {
    const DeviceBases B = ctx->bases(e);
    const std::vector<std::string> pdl_pat = pdl_patterns(e.plan.nb);
    size_t ev = 0;
    int n = 0;
    const char* cur_stage = nullptr;
    struct RangeGuard { bool open = false; ~RangeGuard() { if (open) nvtxRangePop(); } } guard;
    for (const Op& op : e.plan.ops) {
        if (g_nvtx && op.kind != OP_WAIT) {
            const char* st = stage_of(op);
            if (st != cur_stage) {
                if (guard.open) { nvtxRangePop(); guard.open = false; }
                if (st) { nvtxRangePushA(st); guard.open = true; }
                cur_stage = st;
            }
        }
        if (op.kind == OP_WAIT) {
            if (ev >= ctx->events.size()) {
                cudaEvent_t x; CK(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
                ctx->events.push_back(x);
            }
            CK(cudaEventRecord(ctx->events[ev], ctx->streams[op.wait.src_lane]));
            CK(cudaStreamWaitEvent(ctx->streams[op.wait.dst_lane], ctx->events[ev], 0));
            ++ev;
            continue;
        }
        rvc::g_launch_priority = (ctx->f0_priority != 0 && (op.lane == 1 || op.lane == 3) && op.name.compare(0, 3, "sy.") != 0) ? ctx->f0_priority : 0;
        rvc::g_pdl_op = pdl_for(pdl_pat, op.name);
        if (op.stack && e.cvs.grid > 0) {
            // ContentVec's transformer layers: one persistent tcgen05 kernel, launched where the first op stood
            if (&op == &e.plan.ops[size_t(e.plan.cvstack.first)]) n += launch_cvstack(e.cvs, ctx->streams[op.lane]);
            continue;
        }
        if (op.chain >= 0 && size_t(op.chain) < e.chains.size()) {
            // the whole run executes in one persistent kernel, launched where its first op stood
            if (&op == &e.plan.ops[size_t(e.plan.chains[size_t(op.chain)].first)]) n += launch_chain(e.chains[size_t(op.chain)], ctx->streams[op.lane]);
            continue;
        }
        int rc = issue_one(ctx, op, B, ctx->streams[op.lane], &n);
        if (rc != RVC_OK) return rc;
        if (g_sync_each) {  // RVC_SYNC_EACH=1: name the op whose kernel faults (debugging aid)
            cudaError_t se = cudaStreamSynchronize(ctx->streams[op.lane]);
            if (se == cudaSuccess) se = cudaGetLastError();
            if (se != cudaSuccess) return ctx->fail(RVC_ERR_CUDA, "op '" + op.name + "' failed: " + cudaGetErrorString(se));
        }
    }
    rvc::g_launch_priority = 0; rvc::g_pdl_op = false;
    CK(cudaGetLastError());
    *launches = n;
    return RVC_OK;
}

// Live contexts per device.  A chain is a cooperative launch that holds its CTAs for hundreds of microseconds:
// the right trade for one latency-critical stream, the wrong one when several streams share the GPU (their
// chains would queue behind each other: measured 450 vs 547 windows/s with 8 streams) - so plans are built
// with chains only while the context is alone on its device.
std::mutex g_live_mu;
std::map<int, int> g_live_ctx;
int live_contexts(int device) { std::lock_guard<std::mutex> lk(g_live_mu); return g_live_ctx[device]; }

// Resolves the ops of every chain of the plan into the device tables the chain kernel walks.
int build_chain_tables(rvc_ctx* ctx, PlanEntry& e) {
    const DeviceBases B = ctx->bases(e);
    for (const ChainInfo& ci : e.plan.chains) {
        std::vector<ChainOpDev> ops(size_t(ci.count));
        std::vector<ChainPhaseDev> phases(size_t(ci.n_phases), ChainPhaseDev{0, 0, 0, 0});
        for (int k = 0; k < ci.count; ++k) {
            const Op& op = e.plan.ops[size_t(ci.first + k)];
            ChainOpDev d;
            std::memset(&d, 0, sizeof(d));
            d.splitk = 1; d.batch = 1;
            switch (op.kind) {
                case OP_GEMM: {
                    const GemmOp& g = op.gemm;
                    d.g = gemmk::make_params(g, B);
                    d.g.splitk = g.ch_splitk; d.g.scratch = B.p<float>(g.ch_scratch); d.g.counters = B.p<unsigned int>(g.ch_counters);
                    d.batch = g.batch;
                    if (g.ch_variant < 0) {
                        d.kind = CH_GEMM_DIRECT; d.g.splitk = 1;
                        d.items = int((int64_t(g.M) * g.N + 1023) / 1024);
                    } else {
                        d.kind = CH_GEMM; d.variant = g.ch_variant; d.tiles_m = g.ch_tiles_m; d.tiles_n = g.ch_tiles_n; d.splitk = g.ch_splitk;
                        const int nkt = (g.K + 31) / 32;
                        d.g.kt_per_split = (nkt + g.ch_splitk - 1) / g.ch_splitk;
                        d.items = g.ch_tiles_m * g.ch_tiles_n * g.batch * g.ch_splitk;
                        if (g.batch == 1 && int64_t(g.N) * g.K * 4 >= 32768) {
                            d.pf_base = reinterpret_cast<const char*>(d.g.W); d.pf_stride = g.ldw * 4; d.pf_rows = g.N; d.pf_row_bytes = g.K * 4;
                        }
                    }
                    break;
                }
                case OP_AVGPOOL:
                    d.kind = CH_AVGPOOL; d.x0 = B.p<float>(op.pool.in); d.y0 = B.p<float>(op.pool.out); d.ld0 = op.pool.ldin;
                    d.i0 = op.pool.T; d.i1 = op.pool.F; d.i2 = op.pool.C;
                    d.items = int((int64_t(op.pool.T / 2) * (op.pool.F / 2) * op.pool.C + 1023) / 1024);
                    break;
                case OP_LAYERNORM:
                    d.kind = CH_LAYERNORM; d.x0 = B.p<float>(op.ln.X); d.y0 = B.p<float>(op.ln.Y); d.x1 = B.p<float>(op.ln.gamma);
                    d.x2 = B.p<float>(op.ln.beta); d.ld0 = op.ln.ldx; d.ld1 = op.ln.ldy; d.i0 = op.ln.rows; d.i1 = op.ln.cols; d.f0 = op.ln.eps;
                    d.items = (op.ln.rows + 7) / 8;
                    break;
                case OP_RELATTN:
                    d.kind = CH_RELATTN; d.x0 = B.p<float>(op.relattn.qkv); d.y0 = B.p<float>(op.relattn.out); d.x1 = B.p<float>(op.relattn.rel_k);
                    d.x2 = B.p<float>(op.relattn.rel_v); d.ld0 = op.relattn.ldqkv; d.ld1 = op.relattn.ldo; d.i0 = op.relattn.T;
                    d.i1 = op.relattn.heads; d.i2 = op.relattn.dim; d.i3 = op.relattn.window;
                    d.items = op.relattn.T * op.relattn.heads;
                    break;
                default: return ctx->fail(RVC_ERR_INVALID_ARG, "op kind cannot run inside a chain: " + op.name);
            }
            ChainPhaseDev& ph = phases[size_t(ci.phase[size_t(k)])];
            if (ph.op1 == 0) ph.op0 = k;
            ph.op1 = k + 1;
            d.item0 = ph.items; ph.items += d.items;
            ops[size_t(k)] = d;
        }
        ChainDev cd;
        cd.n_ops = ci.count; cd.n_phases = ci.n_phases;
        cd.grid = std::max(1, std::min(ci.grid, chain_max_coresident_ctas()));
        {   // slab chain (chain.h): every GEMM of the run has few rows and fits the column-split shape -> one 16-CTA cluster
            static const int max_cluster = chain_max_cluster_ctas();
            const char* se = getenv("RVC_SLAB");
            bool slab = se && se[0] == '1' && max_cluster >= 16;
            const int G = 16;
            std::vector<int> chunk_bytes;
            for (int k = 0; k < ci.count && slab; ++k) {
                ChainOpDev& d = ops[size_t(k)];
                if (d.kind != CH_GEMM) continue;
                const gemmk::GemmParams& g = d.g;
                const int seg_len = g.seg_len >= g.K ? g.K : g.seg_len;
                int nc, cw, kc, chunks;
                slab_shape(g.N, g.K, G, g.act == ACT_GATE, nc, cw, kc, chunks);
                const long long span = (long long)(g.M - 1) * g.lda + (long long)(g.K / seg_len - 1) * g.seg_stride + seg_len;
                if (g.M > SLAB_MAX_ROWS || d.batch != 1 || g.K % seg_len != 0 || span > SLAB_A_FLOATS || span % 4 != 0 || nc > (g.M <= 8 ? 64 : 48) || nc * kc > SLAB_CHUNK_FLOATS ||
                    (reinterpret_cast<uintptr_t>(g.A) & 15) != 0) {
                    if (getenv("RVC_SLAB_VERBOSE")) std::fprintf(stderr, "slab: chain at op %d rejected: M=%d N=%d K=%d seg_len=%d batch=%d span=%lld nc=%d kc=%d align=%d\n", ci.first + k, g.M, g.N, g.K, seg_len, d.batch, span, nc, kc, int(reinterpret_cast<uintptr_t>(g.A) & 15));
                    slab = false; break;
                }
                d.slab_nc = nc; d.slab_cw = cw; d.slab_kc = kc; d.slab_chunks = chunks; d.slab_chunk0 = int(chunk_bytes.size());
                d.g.splitk = 1;
                for (int j = 0; j < chunks; ++j) chunk_bytes.push_back(nc * kc * 4);
            }
            if (slab && !chunk_bytes.empty()) {
                cd.slab = 1; cd.grid = G; cd.slab_total_chunks = int(chunk_bytes.size());
                cd.slab_stream_floats = (long long)chunk_bytes.size() * SLAB_CHUNK_FLOATS;
                CK(cudaMalloc(&cd.d_slabs, size_t(G) * size_t(cd.slab_stream_floats) * sizeof(float)));
                CK(cudaMalloc(&cd.d_chunk_bytes, chunk_bytes.size() * sizeof(int)));
                CK(cudaMemcpyAsync(cd.d_chunk_bytes, chunk_bytes.data(), chunk_bytes.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->streams[0]));
                for (int k = 0; k < ci.count; ++k) {
                    const ChainOpDev& d = ops[size_t(k)];
                    if (d.kind != CH_GEMM) continue;
                    launch_slab_pack(d.g.W, d.g.ldw, d.g.N, d.g.K, d.slab_nc, d.slab_kc, d.slab_chunks, cd.d_slabs, cd.slab_stream_floats, d.slab_chunk0, G,
                                     ctx->streams[0]);
                }
                CK(cudaStreamSynchronize(ctx->streams[0]));
                CK(cudaGetLastError());
            }
        }
        {   // a side-lane chain made only of skinny GEMMs of one width (RMVPE's bottleneck: 4 valid rows x 512 x 1536) runs on
            // the weight-streaming kernel: 8 columns per CTA, filters fetched whole steps ahead (kernels_wstream.cu)
            const char* we = getenv("RVC_WSTREAM");
            bool ws = !(we && we[0] == '0') && !cd.slab && ci.n_phases == ci.count && ci.count >= 2;
            std::vector<WsOpDev> wops;
            int N0 = 0;
            for (int k = 0; ws && k < ci.count; ++k) {
                const Op& op = e.plan.ops[size_t(ci.first + k)];
                if (op.kind != OP_GEMM || ci.phase[size_t(k)] != k) { ws = false; break; }
                const GemmOp& g = op.gemm;
                WsOpDev w; std::memset(&w, 0, sizeof(w));
                int nv = 0;
                for (int m = 0; m < g.M; ++m) {
                    if (g.mask_period > 0 && (m % g.mask_period) >= g.mask_valid) continue;
                    if (nv < 4) { w.row[nv] = m; w.row_off[nv] = int(int64_t(m) * g.lda); }
                    ++nv;
                }
                if (k == 0) N0 = g.N;
                if (g.batch != 1 || g.out_mode != OUT_PLAIN || (g.act != ACT_NONE && g.act != ACT_RELU) || g.alpha != 1.0f || !g.C2.null() || g.seg_len < g.K ||
                    g.N != N0 || nv > 4 || !wstream_shape_ok(g.M, g.N, g.K, g.lda, nv)) { ws = false; break; }
                const gemmk::GemmParams gp = gemmk::make_params(g, B);
                w.A = gp.A; w.W = gp.W; w.bias = gp.bias; w.R = gp.R; w.C = gp.C; w.ldw = g.ldw; w.ldc = g.ldc; w.ldr = g.ldr;
                w.K = g.K; w.n_rows = nv; w.relu = g.act == ACT_RELU ? 1 : 0;
                w.a_floats = w.row_off[nv - 1] + g.K;
                for (int i = nv; i < 4; ++i) { w.row[i] = w.row[nv - 1]; w.row_off[i] = w.row_off[nv - 1]; }
                if ((reinterpret_cast<uintptr_t>(w.A) & 15) || (reinterpret_cast<uintptr_t>(w.W) & 15) || (g.ldw & 3)) { ws = false; break; }
                wops.push_back(w);
            }
            if (ws) {
                cd.wstream = 1; cd.grid = N0 / 8;
                CK(cudaMalloc(&cd.d_wsops, wops.size() * sizeof(WsOpDev)));
                CK(cudaMemcpyAsync(cd.d_wsops, wops.data(), wops.size() * sizeof(WsOpDev), cudaMemcpyHostToDevice, ctx->streams[0]));
                CK(cudaStreamSynchronize(ctx->streams[0]));
            }
        }
        {   // small grids run as ONE thread-block cluster: hardware cluster barrier instead of the L2 grid barrier
            static const int max_cluster = chain_max_cluster_ctas();
            const char* ce = getenv("RVC_CHAIN_CLUSTER");
            cd.cluster = (!cd.slab && !cd.wstream && !(ce && ce[0] == '0') && cd.grid >= 2 && cd.grid <= max_cluster) ? 1 : 0;
        }
        CK(cudaMalloc(&cd.d_ops, ops.size() * sizeof(ChainOpDev)));
        CK(cudaMalloc(&cd.d_phases, phases.size() * sizeof(ChainPhaseDev)));
        CK(cudaMalloc(&cd.d_bar, 256));
        CK(cudaMemcpyAsync(cd.d_ops, ops.data(), ops.size() * sizeof(ChainOpDev), cudaMemcpyHostToDevice, ctx->streams[0]));
        CK(cudaMemcpyAsync(cd.d_phases, phases.data(), phases.size() * sizeof(ChainPhaseDev), cudaMemcpyHostToDevice, ctx->streams[0]));
        CK(cudaMemsetAsync(cd.d_bar, 0, 256, ctx->streams[0]));
        CK(cudaMalloc(&cd.d_dbg, size_t(ci.n_phases + 1) * sizeof(unsigned long long)));
        CK(cudaMemsetAsync(cd.d_dbg, 0, size_t(ci.n_phases + 1) * sizeof(unsigned long long), ctx->streams[0]));
        CK(cudaStreamSynchronize(ctx->streams[0]));  // the host vectors go out of scope
        e.chains.push_back(cd);
    }
    return RVC_OK;
}

// Phase table + tensor maps of the persistent ContentVec stack kernel (cvstack.h) from the plan's own ops.
int build_cvstack_tables(rvc_ctx* ctx, PlanEntry& e) {
    const CvStackInfo& cs = e.plan.cvstack;
    if (cs.first < 0 || cs.count <= 0 || ctx->cvstack_grid <= 0) return RVC_OK;
    const DeviceBases B = ctx->bases(e);
    const int64_t h16 = B.hilo16_off[SP_CV], plane = B.hilo16_plane[SP_CV];
    if (h16 <= 0) return RVC_OK;   // no fp16 weight planes: the ops run as separate kernels
    const int W = cs.width, F = cs.ffn;
    auto unfuse = [&]() { for (int k = 0; k < cs.count; ++k) e.plan.ops[size_t(cs.first + k)].stack = 0; return RVC_OK; };
    if (W % 128 != 0 || W > 1024 || W % CVS_BN != 0 || F % CVS_BN != 0 || (3 * W) % CVS_BN != 0 || (W / 64) % CVS_SPLITK != 0 || (F / 64) % CVS_SPLITK != 0) return unfuse();
    std::vector<CvsPhase> phases;
    std::vector<uint8_t> maps;   // 128 bytes per CUtensorMap
    auto add_map = [&](const void* base, int K, int rows, int box_rows) -> int {
        const int idx = int(maps.size() / 128);
        maps.resize(maps.size() + 128);
        if (!cvstack_encode_map(maps.data() + size_t(idx) * 128, base, K, rows, box_rows)) return -1;
        return idx;
    };
    struct Planes { unsigned short* hi; unsigned short* lo; int K; int map; };
    auto mk_planes = [&](const Ref& r, int K) { Planes p; p.hi = B.p<unsigned short>(r); p.lo = p.hi + size_t(128) * K; p.K = K; p.map = -1; return p; };
    Planes px = mk_planes(cs.planes_x, W), px1 = mk_planes(cs.planes_x1, W), pa = mk_planes(cs.planes_a, W), phh = mk_planes(cs.planes_h, F);
    for (Planes* p : {&px, &px1, &pa, &phh}) {
        p->map = add_map(p->hi, p->K, 128, 128);
        if (p->map < 0 || add_map(p->lo, p->K, 128, 128) < 0) return unfuse();
    }
    float* partial = B.p<float>(cs.partial);
    const GemmOp* pending = nullptr;   // split-K GEMM whose partial tiles the next LayerNorm sums
    int n_ln = 0, n_gemm = 0;
    for (int k = 0; k < cs.count; ++k) {
        const Op& op = e.plan.ops[size_t(cs.first + k)];
        CvsPhase P;
        std::memset(&P, 0, sizeof(P));
        if (op.kind == OP_LAYERNORM) {
            const LayerNormOp& ln = op.ln;
            if (ln.cols != W || ln.rows != cs.T) return unfuse();
            P.kind = CVS_LN; P.items = (cs.T + 7) / 8; P.cols = W; P.eps = ln.eps;
            P.gamma = B.p<float>(ln.gamma); P.beta = B.p<float>(ln.beta); P.Y = B.p<float>(ln.Y); P.ldy = ln.ldy;
            if (pending) {
                P.X = partial; P.ldx = pending->N; P.slab = int64_t(128) * pending->N; P.S = CVS_SPLITK;
                P.bias = B.p<float>(pending->bias); P.R = B.p<float>(pending->R); P.ldr = pending->ldr;
                pending = nullptr;
            } else {
                P.X = B.p<float>(ln.X); P.ldx = ln.ldx; P.slab = 0; P.S = 1;
            }
            const Planes& out = (n_ln % 2 == 0) ? px : px1;   // enc_in / ln2 feed the next QKV, ln1 feeds FC1
            P.p_hi = out.hi; P.p_lo = out.lo; P.ldp = W;
            ++n_ln;
        } else if (op.kind == OP_ATTN) {
            const AttnOp& a = op.attn;
            if (a.dim != 64 || a.T != cs.T || a.heads * 64 != W) return unfuse();
            P.att_rows = cvs_att_rows(cs.T);
            if (P.att_rows <= 0) return unfuse();
            P.kind = CVS_ATTN; P.items = a.heads * ((cs.T + P.att_rows - 1) / P.att_rows);
            P.qkv = B.p<float>(a.qkv); P.ldqkv = a.ldqkv; P.heads = a.heads;
            P.p_hi = pa.hi; P.p_lo = pa.lo; P.ldp = W;
        } else if (op.kind == OP_GEMM) {
            const GemmOp& g = op.gemm;
            const int role = n_gemm % 4;   // qkv, o, fc1, fc2
            ++n_gemm;
            const Planes& in = role == 0 ? px : (role == 1 ? pa : (role == 2 ? px1 : phh));
            const bool split = !g.R.null();
            if (g.K != in.K || g.N % CVS_BN != 0 || g.M != cs.T || g.batch != 1 || g.seg_len < g.K || g.ldw != g.K || g.alpha != 1.0f ||
                (split != (role == 1 || role == 3)) || (role == 2) != (g.act == ACT_GELU) || (role != 2 && g.act != ACT_NONE) || g.bias.null() ||
                g.W.space != SP_CV || g.W.off % 32 != 0)
                return unfuse();
            const uint8_t* w_hi = B.b[SP_CV] + h16 + g.W.off / 2;
            P.kind = CVS_GEMM; P.a_map = in.map;
            P.w_map = add_map(w_hi, g.K, g.N, CVS_BN);
            if (P.w_map < 0 || add_map(w_hi + plane, g.K, g.N, CVS_BN) < 0) return unfuse();
            P.splitk = split ? CVS_SPLITK : 1; P.nkb = g.K / 64 / P.splitk; P.items = (g.N / CVS_BN) * P.splitk;
            P.bias = B.p<float>(g.bias);
            if (split) { P.epi = CVS_EPI_PARTIAL; P.C = partial; P.ldc = g.N; pending = &g; }
            else if (role == 2) { P.epi = CVS_EPI_GELU_PLANES; P.p_hi = phh.hi; P.p_lo = phh.lo; P.ldp = F; }
            else { P.epi = CVS_EPI_BIAS; P.C = B.p<float>(g.C); P.ldc = g.ldc; }
        } else {
            return unfuse();
        }
        phases.push_back(P);
    }
    if (pending || int(phases.size()) > CVS_MAX_PHASES) return unfuse();
    CvsDev& d = e.cvs;
    d.n_phases = int(phases.size()); d.T = cs.T;
    d.grid = std::max(1, std::min(ctx->cvstack_grid, cvstack_max_ctas()));
    CK(cudaMalloc(&d.d_maps, maps.size()));
    CK(cudaMalloc(&d.d_phases, phases.size() * sizeof(CvsPhase)));
    CK(cudaMalloc(&d.d_bar, 256));
    CK(cudaMemcpyAsync(d.d_maps, maps.data(), maps.size(), cudaMemcpyHostToDevice, ctx->streams[0]));
    CK(cudaMemcpyAsync(d.d_phases, phases.data(), phases.size() * sizeof(CvsPhase), cudaMemcpyHostToDevice, ctx->streams[0]));
    CK(cudaMemsetAsync(d.d_bar, 0, 256, ctx->streams[0]));
    CK(cudaStreamSynchronize(ctx->streams[0]));
    return RVC_OK;
}

int get_plan(rvc_ctx* ctx, PlanKind kind, const Geometry& g, PlanEntry** out, int nb = 1, bool sequential = false) {
    PlanKey key{int(kind), g, (ctx->index.loaded && kind == PLAN_INFER) ? 1 : 0, ctx->index_rows, ctx->cfg.index_k};
    key.chains = (nb <= 1 && (ctx->chain_force || live_contexts(ctx->cfg.device) <= 1)) ? 1 : 0;
    key.nb = nb > 1 ? nb : 1; key.sequential = (nb > 1 && sequential) ? 1 : 0;
    auto it = ctx->plans.find(key);
    if (it != ctx->plans.end()) { *out = it->second.get(); return RVC_OK; }
    PlanOptions opt;
    opt.index_k = ctx->cfg.index_k; opt.upstream_cents_window = ctx->cfg.upstream_cents_window;
    opt.with_index = key.with_index || kind == PLAN_KNN; opt.index_rows = ctx->index_rows; opt.multi_lane = true;
    opt.allow_umma = ctx->allow_umma;
    opt.chain_grid_main = key.chains ? ctx->chain_grid_main : 0; opt.chain_grid_side = key.chains ? ctx->chain_grid_side : 0;
    opt.chain_side_max_m = ctx->chain_side_max_m;
    // fused residual blocks of RMVPE's two full-resolution levels (kernels_cbr.cu): 0 off (default), 1 encoder + decoder,
    // 2 decoder only.  Opt-in: with programmatic dependent launch on the block's GEMMs the two launches are cheaper
    // than the fused kernel (F0 branch alone 1273 vs 1327 us, window 2.643 vs 2.657 ms; profiles/README.md)
    const int fuse_cbr = [] { const char* ev = getenv("RVC_CBR"); return ev ? atoi(ev) : 0; }();
    opt.fuse_cbr = key.nb > 1 ? 0 : fuse_cbr;   // batched plans run these levels on the tensor cores (f0_umma)
    opt.cv_stack = key.chains && ctx->cvstack_grid > 0 && ctx->allow_umma && (kind == PLAN_INFER || ctx->cvstack_all);
    opt.nb = key.nb; opt.sequential = key.sequential != 0; opt.index_cols = ctx->index_c;
    if (ctx->index.loaded && ctx->knn_umma) { opt.index_planes_off = ctx->index.d->planes_off; opt.index_ymax2 = ctx->index.d->ymax2; }
    {   // RVC_F0_UMMA: 0 never, 1 always, default = batched plans only
        const char* ev = getenv("RVC_F0_UMMA");
        const bool want = ev ? ev[0] == '1' : key.nb > 1;
        opt.f0_umma = want && ctx->allow_umma && ctx->f0.loaded && ctx->f0.d->hilo16_off > 0;
    }
    auto e = std::make_unique<PlanEntry>();
    std::string err;
    if (!build_plan(kind, g, opt, ctx->cv.loaded ? &ctx->cv.d->packed : nullptr, &ctx->cvi, ctx->f0.loaded ? &ctx->f0.d->packed : nullptr,
                    &ctx->f0i, ctx->syn.loaded ? &ctx->syn.d->packed : nullptr, &ctx->syi, e->plan, err))
        return ctx->fail(RVC_ERR_BAD_SHAPE, err);
    if (e->plan.n_lanes > MAX_LANES) return ctx->fail(RVC_ERR_INVALID_ARG, "too many lanes");
    CK(cudaMalloc(&e->work.d, size_t(e->plan.work_bytes) * size_t(e->plan.nb)));
    e->work.bytes = size_t(e->plan.work_bytes) * size_t(e->plan.nb);
    CK(cudaMemsetAsync(e->work.d, 0, e->work.bytes, ctx->streams[0]));  // establishes the zero halos once
    if (e->plan.nb > 1) {
        CK(cudaMalloc(&e->bstate.d, size_t(e->plan.state_block) * size_t(e->plan.nb)));
        e->bstate.bytes = size_t(e->plan.state_block) * size_t(e->plan.nb);
        CK(cudaMemsetAsync(e->bstate.d, 0, e->bstate.bytes, ctx->streams[0]));
    }
    { int rc = build_chain_tables(ctx, *e); if (rc != RVC_OK) return rc; }
    { int rc = build_cvstack_tables(ctx, *e); if (rc != RVC_OK) return rc; }
    *out = e.get();
    ctx->plans[key] = std::move(e);
    return RVC_OK;
}

// enqueues one execution of the plan on the context streams (lane 0 = ctx->streams[0])
int run_plan(rvc_ctx* ctx, PlanEntry& e) {
    int n = 0;
    if (ctx->cfg.use_cuda_graph && e.exec) {
        CK(cudaGraphLaunch(e.exec, ctx->streams[0]));
        n = e.launches;
    } else if (ctx->cfg.use_cuda_graph && e.runs >= 1) {
        CK(cudaStreamBeginCapture(ctx->streams[0], cudaStreamCaptureModeThreadLocal));
        int rc = issue_ops(ctx, e, &n);
        cudaGraph_t g = nullptr;
        cudaError_t ce = cudaStreamEndCapture(ctx->streams[0], &g);
        if (rc != RVC_OK) { if (g) cudaGraphDestroy(g); return rc; }
        if (ce != cudaSuccess) return ctx->cuda_fail(ce, "cudaStreamEndCapture");
        e.graph = g; e.launches = n;
        CK(cudaGraphInstantiate(&e.exec, g, 0));
        CK(cudaGraphLaunch(e.exec, ctx->streams[0]));
    } else {
        int rc = issue_ops(ctx, e, &n);
        if (rc != RVC_OK) return rc;
        e.launches = n;
    }
    e.runs++;
    ctx->total_launches += uint64_t(n);
    ctx->last = &e;
    return RVC_OK;
}

int set_params(rvc_ctx* ctx, int32_t pitch_shift) {
    const float up = ctx->cfg.upstream_pitch_shift ? std::pow(2.0f, float(pitch_shift) / 12.0f)
                                                   : std::ldexp(1.0f, pitch_shift / 12);  // 2.0f32.powi(pitch_shift / 12): i32 division (rvc.rs:121)
    launch_k(set_params_kernel, dim3(1), dim3(1), size_t(0), ctx->streams[0],
             reinterpret_cast<RunParams*>(ctx->state.d + StateLayout::off_params), up, ctx->index_rate,
             (unsigned long long)ctx->cfg.noise_seed, (unsigned long long)ctx->window, ctx->cfg.noise_mode);
    ctx->total_launches++;
    return RVC_OK;
}

int enter(rvc_ctx* ctx) {
    if (!ctx) return RVC_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->cfg.device));
    return RVC_OK;
}

float* state_pcm(rvc_ctx* ctx) { return reinterpret_cast<float*>(ctx->state.d + StateLayout::off_pcm); }
float* state_audio(rvc_ctx* ctx) { return reinterpret_cast<float*>(ctx->state.d + StateLayout::off_audio); }

int enqueue_infer(rvc_ctx* ctx, const float* pcm, size_t n, bool pcm_on_device, uint32_t sf16k, int32_t shift, uint32_t skip_head,
                  uint32_t return_length, float* out, bool out_on_device, size_t cap, size_t* out_len) {
    if (!ctx->syn.loaded) return ctx->fail(RVC_ERR_MODEL_NOT_LOADED, "ModelNotLoaded");          // rvc.rs:141
    if (!ctx->cv.loaded) return ctx->fail(RVC_ERR_CONTENTVEC_NOT_LOADED, "ContentvecNotLoaded");  // rvc.rs:85
    if (!ctx->f0.loaded) return ctx->fail(RVC_ERR_F0_NOT_LOADED, "F0NotLoaded");
    if (!pcm || !out || n == 0 || n > size_t(StateLayout::PCM_CAP)) return ctx->fail(RVC_ERR_INVALID_ARG, "bad pcm/out");
    Geometry g{int32_t(n), int32_t(sf16k), int32_t(skip_head), int32_t(return_length)};
    struct CallRange { bool on; CallRange(bool o) : on(o) { if (on) nvtxRangePushA("RvcInfer::infer (rvc.rs:133-220)"); } ~CallRange() { if (on) nvtxRangePop(); } } call_range(g_nvtx);
    PlanEntry* e = nullptr;
    int rc = get_plan(ctx, PLAN_INFER, g, &e);
    if (rc != RVC_OK) return rc;
    if (size_t(e->plan.audio_len) > cap) return ctx->fail(RVC_ERR_INVALID_ARG, "output buffer too small");
    cudaStream_t s = ctx->streams[0];
    CK(cudaMemcpyAsync(state_pcm(ctx), pcm, n * sizeof(float), pcm_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
    set_params(ctx, shift);
    rc = run_plan(ctx, *e);
    if (rc != RVC_OK) return rc;
    ctx->window++;
    CK(cudaMemcpyAsync(out, state_audio(ctx), size_t(e->plan.audio_len) * sizeof(float),
                       out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
    if (out_len) *out_len = size_t(e->plan.audio_len);
    return RVC_OK;
}

constexpr int MAX_BATCH = 32;   // windows per launch of a batched plan (BASELINE configs[2]: 32)

float uppower_of(const rvc_ctx* ctx, int32_t pitch_shift) {
    if (ctx->cfg.upstream_pitch_shift) return std::pow(2.0f, float(pitch_shift) / 12.0f);
    return std::ldexp(1.0f, pitch_shift / 12);  // 2.0f32.powi(pitch_shift / 12): i32 division (rvc.rs:121)
}

int check_infer_ready(rvc_ctx* ctx) {
    if (!ctx->syn.loaded) return ctx->fail(RVC_ERR_MODEL_NOT_LOADED, "ModelNotLoaded");          // rvc.rs:141
    if (!ctx->cv.loaded) return ctx->fail(RVC_ERR_CONTENTVEC_NOT_LOADED, "ContentvecNotLoaded");  // rvc.rs:85
    if (!ctx->f0.loaded) return ctx->fail(RVC_ERR_F0_NOT_LOADED, "F0NotLoaded");
    return RVC_OK;
}

// `nb` consecutive windows of THIS stream through one execution of a batched plan: window w is pcm[w * sf16k, w * sf16k + n)
// (what nb successive rvc_infer calls of the reference's streaming loop would see, obs-rvc/src/lib.rs:659-707); the pitch
// cache is updated window by window inside the plan, the call counter advances by nb.  Weights are read once per nb windows.
int enqueue_windows(rvc_ctx* ctx, const float* pcm, bool on_device, size_t n, uint32_t sf16k, int nb, int32_t shift, uint32_t skip_head,
                    uint32_t return_length, float* out, size_t* audio_len) {
    Geometry g{int32_t(n), int32_t(sf16k), int32_t(skip_head), int32_t(return_length)};
    PlanEntry* e = nullptr;
    int rc = get_plan(ctx, PLAN_INFER, g, &e, nb, true);
    if (rc != RVC_OK) return rc;
    cudaStream_t s = ctx->streams[0];
    const size_t span = n + size_t(nb - 1) * sf16k;
    const float* src = pcm;
    if (!on_device) {
        if (span > size_t(StateLayout::PCM_CAP)) return ctx->fail(RVC_ERR_INVALID_ARG, "window span exceeds the staging buffer");
        CK(cudaMemcpyAsync(state_pcm(ctx), pcm, span * sizeof(float), cudaMemcpyHostToDevice, s));
        src = state_pcm(ctx);
    }
    uint8_t* blocks = e->bstate.d;
    const long long blk = e->plan.state_block;
    launch_k(gather_windows_kernel, dim3(8, unsigned(nb)), dim3(256), size_t(0), s, src, blocks, blk, (long long)e->plan.pcm.off, int(n), int(sf16k));
    CK(cudaMemcpyAsync(blocks + StateLayout::off_cache, ctx->state.d + StateLayout::off_cache, StateLayout::CACHE_LEN * 4, cudaMemcpyDeviceToDevice, s));
    BatchSeeds bs{};
    for (int w = 0; w < nb; ++w) { bs.seed[w] = ctx->cfg.noise_seed; bs.window[w] = ctx->window + uint64_t(w); }
    launch_k(set_params_batch_kernel, dim3(1), dim3(64), size_t(0), s, blocks, blk, nb, uppower_of(ctx, shift), ctx->index_rate, ctx->cfg.noise_mode, bs);
    ctx->total_launches += 2;
    rc = run_plan(ctx, *e);
    if (rc != RVC_OK) return rc;
    ctx->window += uint64_t(nb);
    CK(cudaMemcpyAsync(ctx->state.d + StateLayout::off_cache, blocks + StateLayout::off_cache, StateLayout::CACHE_LEN * 4, cudaMemcpyDeviceToDevice, s));
    const size_t al = size_t(e->plan.audio_len);
    CK(cudaMemcpy2DAsync(out, al * 4, blocks + e->plan.audio.off, size_t(blk), al * 4, size_t(nb),
                         on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
    if (audio_len) *audio_len = al;
    return RVC_OK;
}

// whole-buffer driver of enqueue_windows: n_windows windows in groups of up to MAX_BATCH
int infer_windows(rvc_ctx* ctx, const float* pcm, bool on_device, size_t n_pcm, size_t n, uint32_t sf16k, size_t n_windows, int32_t shift,
                  uint32_t skip_head, uint32_t return_length, float* out, size_t cap, size_t* audio_len, int max_batch) {
    int rc = check_infer_ready(ctx); if (rc) return rc;
    if (!pcm || !out || n == 0 || n_windows == 0 || sf16k == 0) return ctx->fail(RVC_ERR_INVALID_ARG, "bad pcm/out");
    if (n + (n_windows - 1) * size_t(sf16k) > n_pcm) return ctx->fail(RVC_ERR_BAD_SHAPE, "pcm shorter than n + (n_windows - 1) * sf16k");
    if (max_batch <= 0 || max_batch > MAX_BATCH) max_batch = MAX_BATCH;
    size_t al = 0, done = 0;
    while (done < n_windows) {
        const int nb = int(std::min<size_t>(size_t(max_batch), n_windows - done));
        if (nb == 1) {
            if (al == 0) {   // geometry check + output length before anything is written
                PlanEntry* e1 = nullptr;
                rc = get_plan(ctx, PLAN_INFER, Geometry{int32_t(n), int32_t(sf16k), int32_t(skip_head), int32_t(return_length)}, &e1); if (rc) return rc;
                al = size_t(e1->plan.audio_len);
            }
            if ((done + 1) * al > cap) return ctx->fail(RVC_ERR_INVALID_ARG, "output buffer too small");
            size_t ol = 0;
            rc = enqueue_infer(ctx, pcm + done * sf16k, n, on_device, sf16k, shift, skip_head, return_length, out + done * al, on_device, al, &ol);
            if (rc) return rc;
        } else {
            if (al == 0) {
                PlanEntry* eb = nullptr;
                rc = get_plan(ctx, PLAN_INFER, Geometry{int32_t(n), int32_t(sf16k), int32_t(skip_head), int32_t(return_length)}, &eb, nb, true); if (rc) return rc;
                al = size_t(eb->plan.audio_len);
            }
            if ((done + size_t(nb)) * al > cap) return ctx->fail(RVC_ERR_INVALID_ARG, "output buffer too small");
            rc = enqueue_windows(ctx, pcm + done * sf16k, on_device, n, sf16k, nb, shift, skip_head, return_length, out + done * al, nullptr);
            if (rc) return rc;
        }
        done += size_t(nb);
    }
    if (audio_len) *audio_len = al;
    return RVC_OK;
}

// several live streams of one GPU through ONE batched plan (BASELINE configs[3]: 8 per GPU): possible when the contexts
// sit on the same device and share every model (weights are shared by path, engine.cu cache_lookup) and the settings
// that shape the plan.  Stream state (pitch cache, call counter, noise seed) stays per context.
bool batchable(rvc_ctx* const* ctxs, size_t n_ctx) {
    if (n_ctx < 2 || n_ctx > size_t(MAX_BATCH)) return false;
    const rvc_ctx* a = ctxs[0];
    for (size_t i = 0; i < n_ctx; ++i) {
        const rvc_ctx* c = ctxs[i];
        if (!c || !c->syn.loaded || !c->cv.loaded || !c->f0.loaded) return false;
        for (size_t j = 0; j < i; ++j) if (ctxs[j] == c) return false;
        if (c->cfg.device != a->cfg.device || c->cv.d != a->cv.d || c->f0.d != a->f0.d || c->syn.d != a->syn.d || c->index.d != a->index.d ||
            c->index_rate != a->index_rate || c->cfg.index_k != a->cfg.index_k || c->cfg.noise_mode != a->cfg.noise_mode ||
            c->cfg.upstream_cents_window != a->cfg.upstream_cents_window || c->cfg.upstream_pitch_shift != a->cfg.upstream_pitch_shift)
            return false;
    }
    return true;
}

int enqueue_streams(rvc_ctx* const* ctxs, int nb, const float* const* pcm, bool on_device, size_t n, uint32_t sf16k, int32_t shift,
                    uint32_t skip_head, uint32_t return_length, float* const* out, size_t cap, size_t* out_len) {
    rvc_ctx* ctx = ctxs[0];
    if (n == 0 || n > size_t(StateLayout::PCM_CAP)) return ctx->fail(RVC_ERR_INVALID_ARG, "bad pcm/out");
    for (int w = 0; w < nb; ++w) if (!pcm[w] || !out[w]) return ctx->fail(RVC_ERR_INVALID_ARG, "bad pcm/out");
    Geometry g{int32_t(n), int32_t(sf16k), int32_t(skip_head), int32_t(return_length)};
    PlanEntry* e = nullptr;
    int rc = get_plan(ctx, PLAN_INFER, g, &e, nb, false);
    if (rc != RVC_OK) return rc;
    const size_t al = size_t(e->plan.audio_len);
    if (al > cap) return ctx->fail(RVC_ERR_INVALID_ARG, "output buffer too small");
    cudaStream_t s = ctx->streams[0];
    for (int w = 1; w < nb; ++w) CK(cudaStreamSynchronize(ctxs[w]->streams[0]));   // their own pending work touches their caches
    uint8_t* blocks = e->bstate.d;
    const long long blk = e->plan.state_block;
    CachePtrs cp{}; BatchSeeds bs{};
    for (int w = 0; w < nb; ++w) {
        cp.p[w] = reinterpret_cast<float*>(ctxs[w]->state.d + StateLayout::off_cache);
        bs.seed[w] = ctxs[w]->cfg.noise_seed; bs.window[w] = ctxs[w]->window;
        CK(cudaMemcpyAsync(blocks + w * blk + e->plan.pcm.off, pcm[w], n * sizeof(float), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
    }
    launch_k(move_caches_kernel, dim3(unsigned(nb)), dim3(256), size_t(0), s, cp, blocks, blk, 0);
    launch_k(set_params_batch_kernel, dim3(1), dim3(64), size_t(0), s, blocks, blk, nb, uppower_of(ctx, shift), ctx->index_rate, ctx->cfg.noise_mode, bs);
    rc = run_plan(ctx, *e);
    if (rc != RVC_OK) return rc;
    launch_k(move_caches_kernel, dim3(unsigned(nb)), dim3(256), size_t(0), s, cp, blocks, blk, 1);
    ctx->total_launches += 3;
    for (int w = 0; w < nb; ++w) {
        ctxs[w]->window++;
        CK(cudaMemcpyAsync(out[w], blocks + w * blk + e->plan.audio.off, al * sizeof(float), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
    }
    if (out_len) *out_len = al;
    return RVC_OK;
}

int copy_named(rvc_ctx* ctx, const PlanEntry& e, const std::string& name, void* out, size_t cap_bytes, size_t* out_bytes, int window = 0) {
    const NamedBuf* nb = e.plan.find(name);
    if (!nb) return ctx->fail(RVC_ERR_INVALID_ARG, "no such buffer: " + name);
    if (window < 0 || window >= e.plan.nb) return ctx->fail(RVC_ERR_INVALID_ARG, "no such window in the last call");
    size_t bytes = size_t(nb->elems) * 4;
    if (out_bytes) *out_bytes = bytes;
    if (!out) return RVC_OK;
    if (bytes > cap_bytes) return ctx->fail(RVC_ERR_INVALID_ARG, "buffer too small for " + name);
    ctx->sync_all();
    const DeviceBases B = ctx->bases(e);
    CK(cudaMemcpy(out, B.b[nb->ref.space] + nb->ref.off + int64_t(window) * B.bstride[nb->ref.space], bytes, cudaMemcpyDeviceToHost));
    return RVC_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

const char* rvc_version(void) { return "rvc_b200 0.1 (sm_100a)"; }

void rvc_config_default(rvc_config* cfg) {
    if (!cfg) return;
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->device = 0; cfg->noise_mode = 1; cfg->noise_seed = 0; cfg->index_k = 8; cfg->use_cuda_graph = 1;
}

const char* rvc_last_create_error(void) { return g_create_error.c_str(); }
const char* rvc_last_error(const rvc_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int rvc_create(const char* data_path, const rvc_config* cfg, rvc_ctx** out) {
    if (!out || !data_path) { g_create_error = "null argument"; return RVC_ERR_INVALID_ARG; }
    *out = nullptr;
    auto ctx = std::make_unique<rvc_ctx>();
    if (cfg) ctx->cfg = *cfg; else rvc_config_default(&ctx->cfg);
    if (ctx->cfg.index_k <= 0 || ctx->cfg.index_k > 16) { g_create_error = "index_k must be in [1,16]"; return RVC_ERR_INVALID_ARG; }
    ctx->data_path = data_path;
    { const char* ev = getenv("RVC_UMMA"); ctx->allow_umma = !(ev && ev[0] == '0'); }
    { const char* ev = getenv("RVC_SYNC_EACH"); g_sync_each = (ev && ev[0] == '1'); }
    { const char* ev = getenv("RVC_PDL"); rvc::g_use_pdl = (ev && ev[0] == '1'); }
    { const char* ev = getenv("RVC_NVTX"); g_nvtx = (ev && ev[0] == '1'); }
    { const char* ev = getenv("RVC_KNN_UMMA"); ctx->knn_umma = !(ev && ev[0] == '0'); }
    {
        const char* ev = getenv("RVC_CVSTACK"); const char* eg = getenv("RVC_CVSTACK_G");
        ctx->cvstack_grid = (ev && ev[0] == '0') ? 0 : (eg ? atoi(eg) : 64);
        ctx->cvstack_all = ev && ev[0] == '1';   // default: infer plans only (ContentVec has ~0.3 ms of slack behind the F0 lane there and
                                                 // the stack keeps 84 SMs free for it: 2.72 vs 2.76 ms / window); rvc_hubert alone is
                                                 // 0.14 ms faster on the separate kernels.  RVC_CVSTACK=1: every plan, =0: never
    }
    {   // persistent chains: CTA budgets (0 = off).  RVC_CHAIN=0 disables both.
        const char* ev = getenv("RVC_CHAIN"); const bool on = !(ev && ev[0] == '0');
        ctx->chain_force = ev && ev[0] == '2';
        const char* em = getenv("RVC_CHAIN_MAIN"); const char* es = getenv("RVC_CHAIN_SIDE");
        // 147, not 148: a chain CTA fills an SM (247 registers x 256 threads), and the cooperative launch needs all its CTAs
        // resident at once - with one SM left over the one-CTA sine-source kernel of lane 1 runs beside enc_p instead of
        // holding its launch up (or being held up until both chains are over): 2.621 -> 2.604 ms / window
        ctx->chain_grid_main = on ? (em ? atoi(em) : 147) : 0;
        ctx->chain_grid_side = on ? (es ? atoi(es) : 32) : 0;
        const char* emm = getenv("RVC_CHAIN_SIDE_MAXM"); if (emm) ctx->chain_side_max_m = atoi(emm);
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this engine has no CPU path)";
        return RVC_ERR_CUDA;
    }
    if (ctx->cfg.device < 0 || ctx->cfg.device >= ndev) { g_create_error = "bad device ordinal"; return RVC_ERR_INVALID_ARG; }
    if ((e = cudaSetDevice(ctx->cfg.device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return RVC_ERR_CUDA; }
    cudaDeviceProp prop{};
    cudaGetDeviceProperties(&prop, ctx->cfg.device);
    if (prop.major != 10) {
        g_create_error = "device is not sm_100 (kernels are built for sm_100a only): " + std::string(prop.name);
        return RVC_ERR_CUDA;
    }
    init_kernel_attributes();
    for (int i = 0; i < MAX_LANES; ++i) {
        // lane 1 carries the F0 chain, the longest branch of the window: its kernels go first when SMs free up
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        { const char* fp = getenv("RVC_F0_PRIO"); ctx->f0_priority = (fp && fp[0] == '0') ? 0 : prio_hi; }
        const char* pe = getenv("RVC_LANE1_PRIO");
        const bool boost = !(pe && pe[0] == '0');
        if ((e = cudaStreamCreateWithPriority(&ctx->streams[i], cudaStreamNonBlocking, (i == 1 && boost) ? prio_hi : prio_lo)) != cudaSuccess) {
            g_create_error = cudaGetErrorString(e); return RVC_ERR_CUDA;
        }
    }
    if ((e = cudaMalloc(&ctx->state.d, size_t(StateLayout::bytes))) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return RVC_ERR_CUDA; }
    ctx->state.bytes = size_t(StateLayout::bytes);
    cudaMemsetAsync(ctx->state.d, 0, ctx->state.bytes, ctx->streams[0]);
    cudaStreamSynchronize(ctx->streams[0]);
    { std::lock_guard<std::mutex> lk(g_live_mu); g_live_ctx[ctx->cfg.device]++; }
    *out = ctx.release();
    return RVC_OK;
}

void rvc_destroy(rvc_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->cfg.device);
    ctx->sync_all();
    ctx->plans.clear();
    ctx->stream.release();
    for (auto& kv : ctx->resamplers) kv.second.release();
    ctx->cv.unload(); ctx->f0.unload(); ctx->syn.unload(); ctx->index.unload(); ctx->state.release();
    for (auto ev : ctx->events) cudaEventDestroy(ev);
    for (auto ev : ctx->timers) if (ev) cudaEventDestroy(ev);
    for (auto s : ctx->streams) if (s) cudaStreamDestroy(s);
    { std::lock_guard<std::mutex> lk(g_live_mu); g_live_ctx[ctx->cfg.device]--; }
    delete ctx;
}

int rvc_load_contentvec(rvc_ctx* ctx, int32_t model_version) {
    int rc = enter(ctx); if (rc) return rc;
    if (model_version != RVC_MODEL_V1 && model_version != RVC_MODEL_V2) model_version = RVC_MODEL_V2;  // enums.rs:42-49
    const int c = model_version == RVC_MODEL_V1 ? 256 : 768, l = model_version == RVC_MODEL_V1 ? 9 : 12;   // enums.rs:9-23
    std::string path = ctx->data_path + "/contentvec/vec-" + std::to_string(c) + "-layer-" + std::to_string(l) + ".rvcw";
    const std::string key = model_key(ctx->cfg.device, path, ctx->allow_umma);
    std::shared_ptr<ModelData> d = cache_lookup(key);
    if (!d) {
        RvcwFile f; std::string err;
        if (!f.load(path, err)) return ctx->fail(RVC_ERR_IO, err);
        d = std::make_shared<ModelData>();
        if (!pack_contentvec(f, d->packed, d->cvi, err)) return ctx->fail(RVC_ERR_IO, err);
        // the file must BE the requested version (enums.rs:9-23): a 12-layer file under the v1 name (e.g. a converted
        // checkpoint without meta tensors) would silently run the wrong network
        if (d->cvi.n_layers != l || d->cvi.out_dim != c)
            return ctx->fail(RVC_ERR_BAD_SHAPE, path + " holds " + std::to_string(d->cvi.n_layers) + " layers / " + std::to_string(d->cvi.out_dim) +
                                                    " output channels, the requested model version needs " + std::to_string(l) + " / " + std::to_string(c));
        rc = upload(ctx, *d, ctx->allow_umma); if (rc) return rc;
        cache_store(key, d);
    }
    ctx->drop_plans(); ctx->cv.unload();
    ctx->cv.d = d; ctx->cv.loaded = true; ctx->cvi = d->cvi;
    return RVC_OK;
}

int rvc_load_f0(rvc_ctx* ctx, int32_t pitch_algorithm) {
    int rc = enter(ctx); if (rc) return rc;
    (void)pitch_algorithm;  // enums.rs:106-113: every value maps to Rmvpe
    std::string path = ctx->data_path + "/f0/rmvpe.rvcw";
    const std::string key = model_key(ctx->cfg.device, path, ctx->allow_umma);
    std::shared_ptr<ModelData> d = cache_lookup(key);
    if (!d) {
        RvcwFile f; std::string err;
        if (!f.load(path, err)) return ctx->fail(RVC_ERR_IO, err);
        d = std::make_shared<ModelData>();
        if (!pack_rmvpe(f, d->packed, d->f0i, err)) return ctx->fail(RVC_ERR_IO, err);
        rc = upload(ctx, *d, ctx->allow_umma); if (rc) return rc;   // weight planes: batched plans run the wide U-Net levels on the tcgen05 kernel
        cache_store(key, d);
    }
    ctx->drop_plans(); ctx->f0.unload();
    ctx->f0.d = d; ctx->f0.loaded = true; ctx->f0i = d->f0i;
    return RVC_OK;
}

int rvc_load_model(rvc_ctx* ctx, const char* model_path) {
    int rc = enter(ctx); if (rc) return rc;
    if (!model_path) return ctx->fail(RVC_ERR_INVALID_ARG, "null path");
    const std::string key = model_key(ctx->cfg.device, model_path, ctx->allow_umma);
    std::shared_ptr<ModelData> d = cache_lookup(key);
    if (!d) {
        RvcwFile f; std::string err;
        if (!f.load(model_path, err)) return ctx->fail(RVC_ERR_IO, err);
        d = std::make_shared<ModelData>();
        if (!pack_synth(f, d->packed, d->syi, err)) return ctx->fail(RVC_ERR_IO, err);
        rc = upload(ctx, *d, ctx->allow_umma); if (rc) return rc;
        cache_store(key, d);
    }
    ctx->drop_plans(); ctx->syn.unload();
    ctx->syn.d = d; ctx->syn.loaded = true; ctx->syi = d->syi;
    return RVC_OK;
}

int rvc_unload_model(rvc_ctx* ctx) {
    int rc = enter(ctx); if (rc) return rc;
    ctx->drop_plans(); ctx->syn.unload();
    return RVC_OK;
}

static int attach_index(rvc_ctx* ctx, const std::shared_ptr<ModelData>& d, float index_rate) {
    // the index may be loaded before the ContentVec: the width is checked again when an infer plan is built
    if (ctx->cv.loaded && d->cols != ctx->cvi.out_dim)
        return ctx->fail(RVC_ERR_BAD_SHAPE, "retrieval index is " + std::to_string(d->cols) + " wide, the loaded ContentVec produces " +
                                                std::to_string(ctx->cvi.out_dim) + "-wide features");
    ctx->drop_plans(); ctx->index.unload();
    ctx->index.d = d; ctx->index.loaded = true; ctx->index_rows = d->rows; ctx->index_c = d->cols; ctx->index_rate = index_rate;
    return RVC_OK;
}

static int upload_index(rvc_ctx* ctx, const float* rows, size_t n, size_t c, std::shared_ptr<ModelData>& out) {
    if (c % 4 != 0 || c > 1024 || n > 0x7fffffffull || n < size_t(ctx->cfg.index_k))
        return ctx->fail(RVC_ERR_BAD_SHAPE, "index must be N x C with C % 4 == 0, C <= 1024, N >= k");
    out = std::make_shared<ModelData>();
    const size_t row_bytes = (n * c * sizeof(float) + 1023) & ~size_t(1023);
    const bool planes = knn_umma_ok(int(c), 1);   // widths the tensor-core candidate pass serves (C % 64 == 0, C <= 256)
    CK(cudaMalloc(&out->dev.d, row_bytes + (planes ? size_t(knn_umma_planes_bytes(int(n), int(c))) : 0)));
    out->dev.bytes = row_bytes;
    CK(cudaMemcpy(out->dev.d, rows, n * c * sizeof(float), cudaMemcpyHostToDevice));
    if (planes) {
        out->planes_off = int64_t(row_bytes);
        out->ymax2 = launch_knn_build_planes(reinterpret_cast<const float*>(out->dev.d), int(n), int(c), out->dev.d + row_bytes, ctx->streams[0]);
    }
    CK(cudaDeviceSynchronize());
    out->rows = int(n); out->cols = int(c);
    return RVC_OK;
}

int rvc_set_index(rvc_ctx* ctx, const float* rows, size_t n, size_t c, float index_rate) {
    int rc = enter(ctx); if (rc) return rc;
    ctx->index_rate = index_rate;
    if (!rows || n == 0) {  // clears the index
        ctx->drop_plans(); ctx->index.unload(); ctx->index_rows = 0; ctx->index_c = 0;
        return RVC_OK;
    }
    std::shared_ptr<ModelData> d;
    rc = upload_index(ctx, rows, n, c, d); if (rc) return rc;
    return attach_index(ctx, d, index_rate);
}

int rvc_load_index(rvc_ctx* ctx, const char* index_path, float index_rate) {
    int rc = enter(ctx); if (rc) return rc;
    if (!index_path) return ctx->fail(RVC_ERR_INVALID_ARG, "null path");
    const std::string key = model_key(ctx->cfg.device, index_path, false) + "|index";
    std::shared_ptr<ModelData> d = cache_lookup(key);
    if (!d) {
        RvcwFile f; std::string err;
        if (!f.load(index_path, err)) return ctx->fail(RVC_ERR_IO, err);
        const HostTensor* t = f.find("big_npy");
        if (!t || t->dtype != 0 || t->shape.size() != 2) return ctx->fail(RVC_ERR_IO, "index file has no big_npy [N,C] tensor");
        rc = upload_index(ctx, t->f(), size_t(t->shape[0]), size_t(t->shape[1]), d); if (rc) return rc;
        cache_store(key, d);
    }
    return attach_index(ctx, d, index_rate);
}

int rvc_set_index_rate(rvc_ctx* ctx, float index_rate) {
    if (!ctx) return RVC_ERR_INVALID_ARG;
    ctx->index_rate = index_rate;
    return RVC_OK;
}

int rvc_infer(rvc_ctx* ctx, const float* pcm, size_t n, uint32_t sf16k, int32_t pitch_shift, uint32_t skip_head,
              uint32_t return_length, float* out, size_t cap, size_t* out_len) {
    int rc = enter(ctx); if (rc) return rc;
    rc = enqueue_infer(ctx, pcm, n, false, sf16k, pitch_shift, skip_head, return_length, out, false, cap, out_len);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->streams[0]));
    return RVC_OK;
}

int rvc_infer_dev(rvc_ctx* ctx, const float* pcm_dev, size_t n, uint32_t sf16k, int32_t pitch_shift, uint32_t skip_head,
                  uint32_t return_length, float* out_dev, size_t cap, size_t* out_len) {
    int rc = enter(ctx); if (rc) return rc;
    return enqueue_infer(ctx, pcm_dev, n, true, sf16k, pitch_shift, skip_head, return_length, out_dev, true, cap, out_len);
}

int rvc_infer_batch(rvc_ctx* const* ctxs, size_t n_ctx, const float* const* pcm, size_t n, uint32_t sf16k, int32_t pitch_shift,
                    uint32_t skip_head, uint32_t return_length, float* const* out, size_t cap, size_t* out_len) {
    if (!ctxs || !pcm || !out) return RVC_ERR_INVALID_ARG;
    if (batchable(ctxs, n_ctx)) {   // one batched plan for all streams (weights read once per round of windows)
        rvc_ctx* ctx = ctxs[0];
        int rc = enter(ctx); if (rc) return rc;
        rc = enqueue_streams(ctxs, int(n_ctx), pcm, false, n, sf16k, pitch_shift, skip_head, return_length, out, cap, out_len);
        if (rc) return rc;
        CK(cudaStreamSynchronize(ctx->streams[0]));
        return RVC_OK;
    }
    for (size_t i = 0; i < n_ctx; ++i) {
        int rc = enter(ctxs[i]); if (rc) return rc;
        rc = enqueue_infer(ctxs[i], pcm[i], n, false, sf16k, pitch_shift, skip_head, return_length, out[i], false, cap, out_len);
        if (rc) return rc;
    }
    for (size_t i = 0; i < n_ctx; ++i) {
        rvc_ctx* ctx = ctxs[i];
        cudaSetDevice(ctx->cfg.device);
        CK(cudaStreamSynchronize(ctx->streams[0]));
    }
    return RVC_OK;
}

int rvc_infer_batch_dev(rvc_ctx* const* ctxs, size_t n_ctx, const float* const* pcm_dev, size_t n, uint32_t sf16k, int32_t pitch_shift,
                        uint32_t skip_head, uint32_t return_length, float* const* out_dev, size_t cap, size_t* out_len) {
    if (!ctxs || !pcm_dev || !out_dev || n_ctx == 0 || !ctxs[0]) return RVC_ERR_INVALID_ARG;
    rvc_ctx* ctx = ctxs[0];
    int rc = enter(ctx); if (rc) return rc;
    if (!batchable(ctxs, n_ctx)) return ctx->fail(RVC_ERR_INVALID_ARG, "streams cannot share one batched plan (different device / models / settings)");
    return enqueue_streams(ctxs, int(n_ctx), pcm_dev, true, n, sf16k, pitch_shift, skip_head, return_length, out_dev, cap, out_len);
}

int rvc_infer_windows(rvc_ctx* ctx, const float* pcm, size_t n_pcm, size_t n, uint32_t sf16k, size_t n_windows, int32_t pitch_shift,
                      uint32_t skip_head, uint32_t return_length, float* out, size_t cap, size_t* audio_len, int32_t max_batch) {
    int rc = enter(ctx); if (rc) return rc;
    rc = infer_windows(ctx, pcm, false, n_pcm, n, sf16k, n_windows, pitch_shift, skip_head, return_length, out, cap, audio_len, max_batch);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->streams[0]));
    return RVC_OK;
}

int rvc_infer_windows_dev(rvc_ctx* ctx, const float* pcm_dev, size_t n_pcm, size_t n, uint32_t sf16k, size_t n_windows, int32_t pitch_shift,
                          uint32_t skip_head, uint32_t return_length, float* out_dev, size_t cap, size_t* audio_len, int32_t max_batch) {
    int rc = enter(ctx); if (rc) return rc;
    return infer_windows(ctx, pcm_dev, true, n_pcm, n, sf16k, n_windows, pitch_shift, skip_head, return_length, out_dev, cap, audio_len, max_batch);
}

int rvc_hubert(rvc_ctx* ctx, const float* pcm, size_t n, float* out, size_t cap, size_t* out_c, size_t* out_t) {
    int rc = enter(ctx); if (rc) return rc;
    if (!ctx->cv.loaded) return ctx->fail(RVC_ERR_CONTENTVEC_NOT_LOADED, "ContentvecNotLoaded");
    if (!pcm || !out || n == 0 || n > size_t(StateLayout::PCM_CAP)) return ctx->fail(RVC_ERR_INVALID_ARG, "bad pcm/out");
    PlanEntry* e = nullptr;
    rc = get_plan(ctx, PLAN_HUBERT, Geometry{int32_t(n), 0, 0, 0}, &e); if (rc) return rc;
    const int T = e->plan.hubert_T, C = e->plan.hubert_C;
    if (size_t(T) * C > cap) return ctx->fail(RVC_ERR_INVALID_ARG, "output buffer too small");
    cudaStream_t s = ctx->streams[0];
    CK(cudaMemcpyAsync(state_pcm(ctx), pcm, n * sizeof(float), cudaMemcpyHostToDevice, s));
    rc = run_plan(ctx, *e); if (rc) return rc;
    const NamedBuf* nb = e->plan.find("cv.out");
    const float* src = reinterpret_cast<const float*>(e->work.d + nb->ref.off);
    dim3 grid((C + 31) / 32, (T + 31) / 32), block(32, 8);
    launch_k(transpose_kernel, grid, block, size_t(0), s, src, state_audio(ctx), T, C);  // (T,C) -> (C,T) as rvc.rs:96
    ctx->total_launches++;
    CK(cudaMemcpyAsync(out, state_audio(ctx), size_t(T) * C * sizeof(float), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (out_c) *out_c = size_t(C);
    if (out_t) *out_t = size_t(T);
    return RVC_OK;
}

int rvc_extract_feature(rvc_ctx* ctx, const float* pcm, size_t n, float* out, size_t cap, size_t* out_frames, size_t* out_c) {
    int rc = enter(ctx); if (rc) return rc;
    if (!ctx->cv.loaded) return ctx->fail(RVC_ERR_CONTENTVEC_NOT_LOADED, "ContentvecNotLoaded");
    if (!pcm || !out || n == 0 || n > size_t(StateLayout::PCM_CAP)) return ctx->fail(RVC_ERR_INVALID_ARG, "bad pcm/out");
    PlanEntry* e = nullptr;
    rc = get_plan(ctx, PLAN_FEATURE, Geometry{int32_t(n), 0, 0, 0}, &e); if (rc) return rc;
    const size_t frames = 2 * size_t(e->plan.hubert_T) + 1, C = size_t(e->plan.hubert_C);
    if (frames * C > cap) return ctx->fail(RVC_ERR_INVALID_ARG, "output buffer too small");
    cudaStream_t s = ctx->streams[0];
    CK(cudaMemcpyAsync(state_pcm(ctx), pcm, n * sizeof(float), cudaMemcpyHostToDevice, s));
    rc = run_plan(ctx, *e); if (rc) return rc;
    CK(cudaMemcpyAsync(out, state_audio(ctx), frames * C * sizeof(float), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (out_frames) *out_frames = frames;
    if (out_c) *out_c = C;
    return RVC_OK;
}

int rvc_pitch(rvc_ctx* ctx, const float* pcm, size_t n, int32_t pitch_shift, size_t sf16k, float* out, size_t cap, size_t* out_len) {
    int rc = enter(ctx); if (rc) return rc;
    if (!ctx->f0.loaded) return ctx->fail(RVC_ERR_F0_NOT_LOADED, "F0NotLoaded");
    if (!pcm || !out || n == 0 || n > size_t(StateLayout::PCM_CAP)) return ctx->fail(RVC_ERR_INVALID_ARG, "bad pcm/out");
    PlanEntry* e = nullptr;
    rc = get_plan(ctx, PLAN_PITCH, Geometry{int32_t(n), int32_t(sf16k), 0, 0}, &e); if (rc) return rc;
    if (size_t(e->plan.f0_T) > cap) return ctx->fail(RVC_ERR_INVALID_ARG, "output buffer too small");
    cudaStream_t s = ctx->streams[0];
    CK(cudaMemcpyAsync(state_pcm(ctx), pcm, n * sizeof(float), cudaMemcpyHostToDevice, s));
    set_params(ctx, pitch_shift);
    rc = run_plan(ctx, *e); if (rc) return rc;
    const NamedBuf* nb = e->plan.find("f0");
    CK(cudaMemcpyAsync(out, e->work.d + nb->ref.off, size_t(e->plan.f0_T) * sizeof(float), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (out_len) *out_len = size_t(e->plan.f0_T);
    return RVC_OK;
}

// Rmvpe::decode (rmvpe.rs:243-248) + to_local_average_cents (rmvpe.rs:118-133) on caller-supplied salience rows
// [T][360]: the decode stage of `pitch` without the network in front of it, so that tests can drive every argmax bin,
// the exact-threshold case and both cents-window conventions.  f0 is returned unshifted (pitch factor 1).
int rvc_decode_salience(rvc_ctx* ctx, const float* salience, size_t t_frames, float* f0_out, int32_t* argmax_out) {
    int rc = enter(ctx); if (rc) return rc;
    if (!salience || !f0_out || !argmax_out || t_frames == 0 || t_frames * 362 > size_t(StateLayout::AUDIO_CAP))
        return ctx->fail(RVC_ERR_INVALID_ARG, "bad salience / outputs");
    cudaStream_t s = ctx->streams[0];
    const int T = int(t_frames);
    float* d_sal = state_audio(ctx);
    CK(cudaMemcpyAsync(d_sal, salience, size_t(T) * 360 * sizeof(float), cudaMemcpyHostToDevice, s));
    set_params(ctx, 0);
    DeviceBases B;
    B.b[SP_STATE] = ctx->state.d;
    F0DecodeOp o;
    o.salience = Ref{SP_STATE, StateLayout::off_audio};
    o.f0 = Ref{SP_STATE, StateLayout::off_audio + int64_t(T) * 360 * 4};
    o.argmax = Ref{SP_STATE, StateLayout::off_audio + int64_t(T) * 361 * 4};
    o.params = Ref{SP_STATE, StateLayout::off_params};
    o.T = T; o.bins = 360; o.threshold = 0.03f; o.upstream_window = ctx->cfg.upstream_cents_window;
    ctx->total_launches += uint64_t(launch_f0decode(o, B, s));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(f0_out, d_sal + size_t(T) * 360, size_t(T) * sizeof(float), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(argmax_out, d_sal + size_t(T) * 361, size_t(T) * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return RVC_OK;
}

int rvc_mel_extract(rvc_ctx* ctx, const float* pcm, size_t n, float* out, size_t cap, size_t* out_frames) {
    int rc = enter(ctx); if (rc) return rc;
    if (!ctx->f0.loaded) return ctx->fail(RVC_ERR_F0_NOT_LOADED, "F0NotLoaded");
    if (!pcm || !out || n == 0 || n > size_t(StateLayout::PCM_CAP)) return ctx->fail(RVC_ERR_INVALID_ARG, "bad pcm/out");
    PlanEntry* e = nullptr;
    rc = get_plan(ctx, PLAN_MEL, Geometry{int32_t(n), 0, 0, 0}, &e); if (rc) return rc;
    const int T = e->plan.f0_T;
    if (size_t(T) * 128 > cap) return ctx->fail(RVC_ERR_INVALID_ARG, "output buffer too small");
    cudaStream_t s = ctx->streams[0];
    CK(cudaMemcpyAsync(state_pcm(ctx), pcm, n * sizeof(float), cudaMemcpyHostToDevice, s));
    rc = run_plan(ctx, *e); if (rc) return rc;
    const NamedBuf* nb = e->plan.find("mel");
    dim3 grid(4, (T + 31) / 32), block(32, 8);
    launch_k(transpose_kernel, grid, block, size_t(0), s, reinterpret_cast<const float*>(e->work.d + nb->ref.off), state_audio(ctx), T, 128);
    ctx->total_launches++;
    CK(cudaMemcpyAsync(out, state_audio(ctx), size_t(T) * 128 * sizeof(float), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (out_frames) *out_frames = size_t(T);
    return RVC_OK;
}

int rvc_knn_search(rvc_ctx* ctx, const float* queries, size_t q, size_t c, int32_t k, float* d2, int32_t* idx) {
    int rc = enter(ctx); if (rc) return rc;
    if (!ctx->index.loaded) return ctx->fail(RVC_ERR_INVALID_ARG, "no index loaded");
    if (!queries || !d2 || !idx || q == 0 || c != size_t(ctx->index_c) || k <= 0 || k > 16 || q * c > size_t(StateLayout::AUDIO_CAP))
        return ctx->fail(RVC_ERR_BAD_SHAPE, "bad kNN query shape");
    PlanEntry* e = nullptr;
    rc = get_plan(ctx, PLAN_KNN, Geometry{int32_t(q), int32_t(c), 0, k}, &e); if (rc) return rc;
    cudaStream_t s = ctx->streams[0];
    CK(cudaMemcpyAsync(state_audio(ctx), queries, q * c * sizeof(float), cudaMemcpyHostToDevice, s));
    rc = run_plan(ctx, *e); if (rc) return rc;
    CK(cudaMemcpyAsync(idx, e->work.d + e->plan.find("knn_idx")->ref.off, q * k * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(d2, e->work.d + e->plan.find("knn_d2")->ref.off, q * k * sizeof(float), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return RVC_OK;
}

int rvc_knn_fallbacks(rvc_ctx* ctx, uint64_t* total) {
    int rc = enter(ctx); if (rc) return rc;
    if (!total) return ctx->fail(RVC_ERR_INVALID_ARG, "null argument");
    *total = 0;
    if (!ctx->index.loaded || ctx->index.d->planes_off == 0) return RVC_OK;
    ctx->sync_all();
    const int64_t off = ctx->index.d->planes_off + knn_umma_counters_off(ctx->index_rows, ctx->index_c);
    unsigned int v = 0;
    CK(cudaMemcpy(&v, ctx->index.d->dev.d + off, 4, cudaMemcpyDeviceToHost));
    *total = v;
    return RVC_OK;
}

int rvc_get_last(rvc_ctx* ctx, const char* name, void* out, size_t cap_bytes, size_t* out_bytes) {
    int rc = enter(ctx); if (rc) return rc;
    if (!name) return ctx->fail(RVC_ERR_INVALID_ARG, "null name");
    if (!ctx->last) return ctx->fail(RVC_ERR_INVALID_ARG, "nothing has run yet");
    std::string nm(name);
    if (nm == "salience") nm = "rm.salience";
    if (nm == "hubert") nm = "cv.out";
    return copy_named(ctx, *ctx->last, nm, out, cap_bytes, out_bytes);
}

int rvc_get_last_window(rvc_ctx* ctx, int32_t window, const char* name, void* out, size_t cap_bytes, size_t* out_bytes) {
    int rc = enter(ctx); if (rc) return rc;
    if (!name) return ctx->fail(RVC_ERR_INVALID_ARG, "null name");
    if (!ctx->last) return ctx->fail(RVC_ERR_INVALID_ARG, "nothing has run yet");
    std::string nm(name);
    if (nm == "salience") nm = "rm.salience";
    if (nm == "hubert") nm = "cv.out";
    return copy_named(ctx, *ctx->last, nm, out, cap_bytes, out_bytes, window);
}

int rvc_debug_tensor(rvc_ctx* ctx, const char* name, float* out, size_t cap, size_t* out_len) {
    size_t bytes = 0;
    int rc = rvc_get_last(ctx, name, out, cap * 4, &bytes);
    if (out_len) *out_len = bytes / 4;
    return rc;
}

int rvc_debug_list(rvc_ctx* ctx, char* out, size_t cap_bytes, size_t* out_bytes) {
    if (!ctx || !ctx->last) return RVC_ERR_INVALID_ARG;
    std::string s;
    for (const auto& b : ctx->last->plan.bufs) { s += b.name; s += '\n'; }
    if (out_bytes) *out_bytes = s.size();
    if (out && cap_bytes > 0) { size_t n = s.size() < cap_bytes - 1 ? s.size() : cap_bytes - 1; std::memcpy(out, s.data(), n); out[n] = 0; }
    return RVC_OK;
}

int rvc_reset_state(rvc_ctx* ctx) {
    int rc = enter(ctx); if (rc) return rc;
    ctx->sync_all();
    CK(cudaMemsetAsync(ctx->state.d + StateLayout::off_cache, 0, StateLayout::CACHE_LEN * 4, ctx->streams[0]));
    CK(cudaStreamSynchronize(ctx->streams[0]));
    ctx->window = 0;
    return RVC_OK;
}

int rvc_sync(rvc_ctx* ctx) {
    int rc = enter(ctx); if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->streams[0]));
    return RVC_OK;
}

void* rvc_cuda_stream(rvc_ctx* ctx) { return ctx ? static_cast<void*>(ctx->streams[0]) : nullptr; }

int rvc_kernel_launches(rvc_ctx* ctx, uint64_t* total) {
    if (!ctx || !total) return RVC_ERR_INVALID_ARG;
    *total = ctx->total_launches;
    return RVC_OK;
}

int rvc_plan_info(rvc_ctx* ctx, char* out, size_t cap_bytes, size_t* out_bytes) {
    if (!ctx || !ctx->last) return RVC_ERR_INVALID_ARG;
    const Plan& p = ctx->last->plan;
    char buf[512];
    int n = std::snprintf(buf, sizeof(buf),
                          "{\"ops\": %zu, \"kernels_per_run\": %d, \"lanes\": %d, \"work_bytes\": %lld, \"hubert_T\": %d, \"hubert_C\": %d, "
                          "\"f0_T\": %d, \"audio_len\": %d, \"knn_q\": %d, \"graph\": %d, \"windows\": %d, \"chains\": %zu, \"cvstack\": %d}",
                          p.ops.size(), ctx->last->launches, p.n_lanes, (long long)p.work_bytes, p.hubert_T, p.hubert_C, p.f0_T,
                          p.audio_len, p.knn_q, ctx->last->exec ? 1 : 0, p.nb, p.chains.size(), ctx->last->cvs.grid > 0 ? 1 : 0);
    if (out_bytes) *out_bytes = size_t(n);
    if (out && cap_bytes > 0) { size_t m = size_t(n) < cap_bytes - 1 ? size_t(n) : cap_bytes - 1; std::memcpy(out, buf, m); out[m] = 0; }
    return RVC_OK;
}

// ---- streaming glue ("next" row #1): scratch lives in the audio staging region of the state arena ----
int rvc_envelop_mixing(rvc_ctx* ctx, const float* input, size_t n_in, float* output, size_t n_out, uint32_t sample_rate,
                       double mix_rate, float* rms1, float* rms2) {
    int rc = enter(ctx); if (rc) return rc;
    const int zc = int(sample_rate / 100);
    if (!input || !output || n_out == 0 || n_in < n_out || zc <= 0 || n_out * 4 + 4096 > size_t(StateLayout::AUDIO_CAP))
        return ctx->fail(RVC_ERR_BAD_SHAPE, "bad envelop_mixing shape");
    cudaStream_t s = ctx->streams[0];
    float* base = state_audio(ctx);
    float* d_in = base; float* d_out = base + n_out; float* d_r1 = d_out + n_out; float* d_r2 = d_r1 + 2048;
    float* d_dbg1 = d_r2 + 2048; float* d_dbg2 = d_dbg1 + n_out;
    const int nfr = int((n_out + size_t(4 * zc) / 2 * 2 - size_t(4 * zc)) / size_t(zc)) + 1;  // rt_utils.rs:93-101
    if (nfr > 2048 || nfr < 2) return ctx->fail(RVC_ERR_BAD_SHAPE, "too many rms frames");
    CK(cudaMemcpyAsync(d_in, input, n_out * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_out, output, n_out * 4, cudaMemcpyHostToDevice, s));
    launch_rms(d_in, int(n_out), 4 * zc, zc, d_r1, nfr, s);
    launch_rms(d_out, int(n_out), 4 * zc, zc, d_r2, nfr, s);
    launch_envelop_mix(d_out, int(n_out), d_r1, d_r2, nfr, float(1.0 - mix_rate), (rms1 && rms2) ? d_dbg1 : nullptr,
                       (rms1 && rms2) ? d_dbg2 : nullptr, s);
    ctx->total_launches += 3;
    CK(cudaMemcpyAsync(output, d_out, n_out * 4, cudaMemcpyDeviceToHost, s));
    if (rms1 && rms2) {
        CK(cudaMemcpyAsync(rms1, d_dbg1, n_out * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(rms2, d_dbg2, n_out * 4, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    return RVC_OK;
}

static int sola_common(rvc_ctx* ctx, const float* x, size_t n, const float* sola, uint32_t buf, uint32_t search, float** d_x_out,
                       float** d_sola_out, int** d_off_out) {
    if (!x || !sola || buf == 0 || size_t(buf) + search > n || 2 * n + buf + search + 96 > size_t(StateLayout::AUDIO_CAP))   // x | sola | cor | offset | block (<= n)
        return ctx->fail(RVC_ERR_BAD_SHAPE, "bad SOLA shape");
    cudaStream_t s = ctx->streams[0];
    float* base = state_audio(ctx);
    float* d_x = base; float* d_sola = base + n; float* d_cor = d_sola + buf; int* d_off = reinterpret_cast<int*>(d_cor + search + 1);
    CK(cudaMemcpyAsync(d_x, x, n * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_sola, sola, size_t(buf) * 4, cudaMemcpyHostToDevice, s));
    launch_sola(d_x, d_sola, int(buf), int(search), d_cor, d_off, s);
    ctx->total_launches += 2;
    *d_x_out = d_x; *d_sola_out = d_sola; *d_off_out = d_off;
    return RVC_OK;
}

int rvc_sola_offset(rvc_ctx* ctx, const float* input_buffer, size_t n, const float* sola_buffer, uint32_t buffer_frame_size,
                    uint32_t search_frame_size, uint32_t* offset) {
    int rc = enter(ctx); if (rc) return rc;
    if (!offset) return ctx->fail(RVC_ERR_INVALID_ARG, "null offset");
    float *dx, *ds; int* doff;
    rc = sola_common(ctx, input_buffer, n, sola_buffer, buffer_frame_size, search_frame_size, &dx, &ds, &doff); if (rc) return rc;
    int h = 0;
    CK(cudaMemcpyAsync(&h, doff, 4, cudaMemcpyDeviceToHost, ctx->streams[0]));
    CK(cudaStreamSynchronize(ctx->streams[0]));
    *offset = uint32_t(h);
    return RVC_OK;
}

int rvc_sola_crossfade(rvc_ctx* ctx, const float* infer_out, size_t n, float* sola_buffer, uint32_t buffer_frame_size,
                       uint32_t search_frame_size, uint32_t sample_frame_size, float* block_out, uint32_t* offset) {
    int rc = enter(ctx); if (rc) return rc;
    if (!block_out || size_t(search_frame_size) + sample_frame_size + buffer_frame_size > n)
        return ctx->fail(RVC_ERR_BAD_SHAPE, "infer output shorter than search + block + fade (lib.rs:789)");
    float *dx, *ds; int* doff;
    rc = sola_common(ctx, infer_out, n, sola_buffer, buffer_frame_size, search_frame_size, &dx, &ds, &doff); if (rc) return rc;
    cudaStream_t s = ctx->streams[0];
    float* d_block = reinterpret_cast<float*>(doff) + 16;
    launch_sola_crossfade(dx, doff, ds, int(buffer_frame_size), int(sample_frame_size), d_block, s);
    ctx->total_launches += 2;
    int h = 0;
    CK(cudaMemcpyAsync(block_out, d_block, size_t(sample_frame_size) * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(sola_buffer, ds, size_t(buffer_frame_size) * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&h, doff, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (offset) *offset = uint32_t(h);
    return RVC_OK;
}

// ---- streaming loop around the call: obs-rvc/src/lib.rs:186-300 (state), :659-795 (process_one_frame) ----
static int upload_resampler(rvc_ctx* ctx, int fs_in, int fs_out, int chunk, ResamplerDev& r) {
    r.release();
    if (!build_resample_table(fs_in, fs_out, chunk, r.t)) return ctx->fail(RVC_ERR_BAD_SHAPE, "bad resampler arguments");
    if (size_t(2 * r.t.n_in + 31 * r.t.a + 1) * 4 > size_t(220 * 1024))
        return ctx->fail(RVC_ERR_BAD_SHAPE, "resampler chunk too long for the polyphase kernel (" + std::to_string(r.t.n_in) + " samples)");
    CK(cudaMalloc(&r.kappa, r.t.kappa.size() * sizeof(float)));
    CK(cudaMemcpy(r.kappa, r.t.kappa.data(), r.t.kappa.size() * sizeof(float), cudaMemcpyHostToDevice));
    for (int i = 0; i < 2; ++i) {
        CK(cudaMalloc(&r.overlap[i], size_t(r.t.n_out) * sizeof(float)));
        CK(cudaMemset(r.overlap[i], 0, size_t(r.t.n_out) * sizeof(float)));
    }
    r.cur = 0;
    std::vector<float>().swap(r.t.kappa);
    CK(cudaDeviceSynchronize());   // the memsets above run on the legacy stream: they must not trail the first launch on the context's own stream
    return RVC_OK;
}

static void run_resampler(rvc_ctx* ctx, ResamplerDev& r, const float* x, float* out, cudaStream_t s) {
    launch_resample(x, r.kappa, r.overlap[r.cur], r.overlap[r.cur ^ 1], out, r.t.a, r.t.b, r.t.period, r.t.n_in, r.t.n_out, s);
    r.cur ^= 1;
    ctx->total_launches++;
}

void rvc_stream_config_default(rvc_stream_config* c) {
    if (!c) return;
    std::memset(c, 0, sizeof(*c));
    // defaults of the OBS filter's settings (lib.rs:179,188-190,262-265)
    c->sample_rate = 48000; c->sample_length = 0.30; c->crossfade_length = 0.07; c->extra_inference_time = 2.00;
    c->rms_mix_rate = 0.0; c->pitch_shift = 12; c->skip_inference = 0;
}

static long long rust_round(double x) { return (long long)(x >= 0 ? std::floor(x + 0.5) : -std::floor(-x + 0.5)); }

int rvc_stream_open(rvc_ctx* ctx, const rvc_stream_config* cfg, uint32_t* sample_frame_size) {
    int rc = enter(ctx); if (rc) return rc;
    if (!cfg || cfg->sample_rate < 8000 || cfg->sample_rate % 100 != 0) return ctx->fail(RVC_ERR_INVALID_ARG, "bad stream config");
    if (!cfg->skip_inference && !ctx->syn.loaded) return ctx->fail(RVC_ERR_MODEL_NOT_LOADED, "ModelNotLoaded");
    StreamState& st = ctx->stream;
    ctx->sync_all();
    st.release();
    st.cfg = *cfg;
    const int sr = int(cfg->sample_rate), zc = sr / 100;
    st.zc = zc;
    st.sample_frame_time = int(rust_round(cfg->sample_length * sr / zc));
    st.sample_frame_size = st.sample_frame_time * zc;
    st.sample_frame_16k = st.sample_frame_time * 160;
    st.crossfade = int(rust_round(cfg->crossfade_length * sr / zc)) * zc;
    st.sola_buf = std::min(st.crossfade, 4 * zc);
    st.sola_search = zc;
    st.extra = int(rust_round(cfg->extra_inference_time * sr / zc)) * zc;
    st.in_size = st.extra + st.crossfade + st.sola_search + st.sample_frame_size;
    st.in16k_size = 160 * st.in_size / zc;
    st.ret_len = (st.sample_frame_size + st.sola_buf + st.sola_search) / zc;
    st.model_sr = cfg->skip_inference ? 16000 : ctx->syi.sr;
    st.ret_size = st.ret_len * (st.model_sr / 100);
    st.skip_head = st.extra / zc;
    if (st.sample_frame_time <= 0 || st.sola_buf <= 0 || st.in16k_size > int(StateLayout::PCM_CAP))
        return ctx->fail(RVC_ERR_BAD_SHAPE, "stream geometry out of range");
    rc = upload_resampler(ctx, sr, 16000, st.sample_frame_size + 2 * zc, st.down); if (rc) return rc;
    rc = upload_resampler(ctx, st.model_sr, sr, st.ret_size, st.up); if (rc) return rc;
    // the loop only works when the resamplers consume exactly the chunk they are given (rubato would return an error otherwise)
    if (st.down.t.n_in != st.sample_frame_size + 2 * zc || st.down.t.n_out != (st.sample_frame_time + 2) * 160 || st.up.t.n_in != st.ret_size)
        return ctx->fail(RVC_ERR_BAD_SHAPE, "resampler chunk sizes do not tile the frame");
    st.up_out = st.up.t.n_out;
    if (st.up_out < st.sola_search + st.sample_frame_size + st.sola_buf) return ctx->fail(RVC_ERR_BAD_SHAPE, "upsampled block shorter than search + frame + fade");
    for (int i = 0; i < 2; ++i) {
        CK(cudaMalloc(&st.inbuf[i], size_t(st.in_size) * 4)); CK(cudaMemset(st.inbuf[i], 0, size_t(st.in_size) * 4));
        CK(cudaMalloc(&st.in16k[i], size_t(st.in16k_size) * 4)); CK(cudaMemset(st.in16k[i], 0, size_t(st.in16k_size) * 4));
    }
    st.scratch_floats = size_t(st.sample_frame_size) * 2 + size_t(st.down.t.n_out) + size_t(st.ret_size) + size_t(st.up_out) + 2 * 2048 +
                        size_t(st.sola_search) + 64 + size_t(st.sola_buf) + 1024;
    CK(cudaMalloc(&st.scratch, st.scratch_floats * 4));
    CK(cudaMemset(st.scratch, 0, st.scratch_floats * 4));
    CK(cudaDeviceSynchronize());   // (memsets on the legacy stream, launches on the context's stream)
    st.cur = 0; st.frames = 0; st.open = true;
    if (sample_frame_size) *sample_frame_size = uint32_t(st.sample_frame_size);
    return RVC_OK;
}

int rvc_stream_close(rvc_ctx* ctx) {
    int rc = enter(ctx); if (rc) return rc;
    ctx->sync_all();
    ctx->stream.release();
    return RVC_OK;
}

int rvc_stream_set(rvc_ctx* ctx, int32_t pitch_shift, double rms_mix_rate) {
    if (!ctx || !ctx->stream.open) return RVC_ERR_INVALID_ARG;
    ctx->stream.cfg.pitch_shift = pitch_shift; ctx->stream.cfg.rms_mix_rate = rms_mix_rate;
    return RVC_OK;
}

int rvc_stream_info(rvc_ctx* ctx, char* out, size_t cap_bytes, size_t* out_bytes) {
    if (!ctx || !ctx->stream.open) return RVC_ERR_INVALID_ARG;
    const StreamState& st = ctx->stream;
    char buf[640];
    int n = std::snprintf(buf, sizeof(buf),
        "{\"sample_rate\": %u, \"sample_frame_size\": %d, \"sample_frame_16k\": %d, \"crossfade_frame_size\": %d, \"sola_buffer_frame_size\": %d, "
        "\"sola_search_frame_size\": %d, \"extra_frame_size\": %d, \"input_buffer_size\": %d, \"input_buffer_16k_size\": %d, "
        "\"model_return_length\": %d, \"model_return_size\": %d, \"model_sample_rate\": %d, \"skip_head\": %d, \"down\": [%d, %d], \"up\": [%d, %d], \"frames\": %llu}",
        st.cfg.sample_rate, st.sample_frame_size, st.sample_frame_16k, st.crossfade, st.sola_buf, st.sola_search, st.extra, st.in_size, st.in16k_size,
        st.ret_len, st.ret_size, st.model_sr, st.skip_head, st.down.t.n_in, st.down.t.n_out, st.up.t.n_in, st.up.t.n_out, (unsigned long long)st.frames);
    if (out_bytes) *out_bytes = size_t(n);
    if (out && cap_bytes > 0) { size_t m = size_t(n) < cap_bytes - 1 ? size_t(n) : cap_bytes - 1; std::memcpy(out, buf, m); out[m] = 0; }
    return RVC_OK;
}

// process_one_frame (lib.rs:659-795): one H2D of the new block, everything else device-resident, one D2H of the result
int rvc_process_frame(rvc_ctx* ctx, const float* input, float* output, uint32_t* sola_offset) {
    int rc = enter(ctx); if (rc) return rc;
    StreamState& st = ctx->stream;
    if (!st.open) return ctx->fail(RVC_ERR_INVALID_ARG, "no open stream (rvc_stream_open)");
    if (!input || !output) return ctx->fail(RVC_ERR_INVALID_ARG, "null input / output");
    cudaStream_t s = ctx->streams[0];
    float* d_frame = st.scratch;
    float* d_down = d_frame + st.sample_frame_size;
    float* d_model = d_down + st.down.t.n_out;
    float* d_up = d_model + st.ret_size;
    float* d_r1 = d_up + st.up_out; float* d_r2 = d_r1 + 2048;
    float* d_cor = d_r2 + 2048;
    int* d_off = reinterpret_cast<int*>(d_cor + st.sola_search + 32);
    float* d_sola = d_cor + st.sola_search + 64;
    float* d_block = d_sola + st.sola_buf + 512;
    const int p = st.cur, q = p ^ 1;
    CK(cudaMemcpyAsync(d_frame, input, size_t(st.sample_frame_size) * 4, cudaMemcpyHostToDevice, s));
    // lib.rs:661-669: both ring buffers move left by one frame; the new block lands at the end of the OBS-rate one
    launch_shift_append(st.inbuf[q], st.inbuf[p], st.in_size, st.sample_frame_size, d_frame, st.sample_frame_size, s);
    launch_shift_append(st.in16k[q], st.in16k[p], st.in16k_size, st.sample_frame_16k, nullptr, 0, s);
    ctx->total_launches += 2;
    // lib.rs:671-683: the last frame + 2 zc samples -> 16 kHz; the first 160 samples of the result are dropped
    run_resampler(ctx, st.down, st.inbuf[q] + (st.in_size - st.sample_frame_size - 2 * st.zc), d_down, s);
    const int copy_n = (st.sample_frame_time + 1) * 160;
    CK(cudaMemcpyAsync(st.in16k[q] + (st.in16k_size - copy_n), d_down + 160, size_t(copy_n) * 4, cudaMemcpyDeviceToDevice, s));
    st.cur = q;
    if (st.cfg.skip_inference) {
        CK(cudaMemcpyAsync(d_model, st.in16k[q] + (st.in16k_size - st.ret_size), size_t(st.ret_size) * 4, cudaMemcpyDeviceToDevice, s));
    } else {
        size_t got = 0;
        rc = enqueue_infer(ctx, st.in16k[q], size_t(st.in16k_size), true, uint32_t(st.sample_frame_16k), st.cfg.pitch_shift, uint32_t(st.skip_head),
                           uint32_t(st.ret_len), d_model, true, size_t(st.ret_size), &got);
        if (rc) return rc;
        if (int(got) != st.ret_size) return ctx->fail(RVC_ERR_BAD_SHAPE, "model output size mismatch (lib.rs:728)");
    }
    run_resampler(ctx, st.up, d_model, d_up, s);
    if (st.cfg.rms_mix_rate < 1.0) {   // lib.rs:758-765
        const int n_out = st.up_out, zc = st.zc;
        const int nfr = (n_out + (4 * zc) / 2 * 2 - 4 * zc) / zc + 1;
        if (nfr > 2048 || nfr < 2) return ctx->fail(RVC_ERR_BAD_SHAPE, "too many rms frames");
        launch_rms(st.inbuf[q] + st.extra, n_out, 4 * zc, zc, d_r1, nfr, s);
        launch_rms(d_up, n_out, 4 * zc, zc, d_r2, nfr, s);
        launch_envelop_mix(d_up, n_out, d_r1, d_r2, nfr, float(1.0 - st.cfg.rms_mix_rate), nullptr, nullptr, s);
        ctx->total_launches += 3;
    }
    // lib.rs:767-794: SOLA offset, sin^2 cross-fade with the previous tail, new tail, the emitted block
    launch_sola(d_up, d_sola, st.sola_buf, st.sola_search, d_cor, d_off, s);
    launch_sola_crossfade(d_up, d_off, d_sola, st.sola_buf, st.sample_frame_size, d_block, s);
    ctx->total_launches += 4;
    CK(cudaGetLastError());
    int h = 0;
    CK(cudaMemcpyAsync(output, d_block, size_t(st.sample_frame_size) * 4, cudaMemcpyDeviceToHost, s));
    if (sola_offset) CK(cudaMemcpyAsync(&h, d_off, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (sola_offset) *sola_offset = uint32_t(h);
    st.frames++;
    return RVC_OK;
}

// one chunk through a rubato-equivalent resampler (unit entry for the parity tests): overlap_inout carries the
// overlap-add state between calls ([n_out] floats, zeros at the start of a stream)
int rvc_resample_chunk(rvc_ctx* ctx, uint32_t fs_in, uint32_t fs_out, const float* in, size_t n_in, float* overlap_inout, float* out,
                       size_t cap, size_t* n_out) {
    int rc = enter(ctx); if (rc) return rc;
    if (!in || !out || !overlap_inout || n_in == 0) return ctx->fail(RVC_ERR_INVALID_ARG, "null buffers");
    const std::string key = std::to_string(fs_in) + ">" + std::to_string(fs_out) + ":" + std::to_string(n_in);
    auto it = ctx->resamplers.find(key);
    if (it == ctx->resamplers.end()) {
        ResamplerDev r;
        rc = upload_resampler(ctx, int(fs_in), int(fs_out), int(n_in), r); if (rc) { r.release(); return rc; }
        it = ctx->resamplers.emplace(key, r).first;
    }
    ResamplerDev& r = it->second;
    if (size_t(r.t.n_in) != n_in) return ctx->fail(RVC_ERR_BAD_SHAPE, "chunk must be a multiple of fs_in / gcd(fs_in, fs_out) (FftFixedInOut::input_frames_next)");
    if (size_t(r.t.n_out) > cap || size_t(r.t.n_in + r.t.n_out) + 64 > size_t(StateLayout::AUDIO_CAP)) return ctx->fail(RVC_ERR_INVALID_ARG, "output buffer too small");
    cudaStream_t s = ctx->streams[0];
    float* d_in = state_audio(ctx); float* d_out = d_in + r.t.n_in;
    CK(cudaMemcpyAsync(d_in, in, n_in * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(r.overlap[r.cur], overlap_inout, size_t(r.t.n_out) * 4, cudaMemcpyHostToDevice, s));
    run_resampler(ctx, r, d_in, d_out, s);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, d_out, size_t(r.t.n_out) * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(overlap_inout, r.overlap[r.cur], size_t(r.t.n_out) * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (n_out) *n_out = size_t(r.t.n_out);
    return RVC_OK;
}

int rvc_event_record(rvc_ctx* ctx, int slot) {
    int rc = enter(ctx); if (rc) return rc;
    if (slot < 0 || slot >= 8) return ctx->fail(RVC_ERR_INVALID_ARG, "bad timer slot");
    if (!ctx->timers[slot]) CK(cudaEventCreate(&ctx->timers[slot]));
    CK(cudaEventRecord(ctx->timers[slot], ctx->streams[0]));
    return RVC_OK;
}

int rvc_event_elapsed_ms(rvc_ctx* ctx, int a, int b, float* ms) {
    int rc = enter(ctx); if (rc) return rc;
    if (a < 0 || a >= 8 || b < 0 || b >= 8 || !ms || !ctx->timers[a] || !ctx->timers[b]) return ctx->fail(RVC_ERR_INVALID_ARG, "bad timer slot");
    CK(cudaEventSynchronize(ctx->timers[b]));
    CK(cudaEventElapsedTime(ms, ctx->timers[a], ctx->timers[b]));
    return RVC_OK;
}

int rvc_profile_ops(rvc_ctx* ctx, int iters, char* out, size_t cap_bytes, size_t* out_bytes) {
    int rc = enter(ctx); if (rc) return rc;
    if (!ctx->last || iters <= 0) return ctx->fail(RVC_ERR_INVALID_ARG, "nothing has run yet");
    PlanEntry& e = *ctx->last;
    ctx->sync_all();
    const DeviceBases B = ctx->bases(e);
    cudaStream_t s = ctx->streams[0];
    cudaEvent_t t0, t1;
    CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
    static const char* KN[] = {"gemm", "layernorm", "attn", "relattn", "conv0_stats", "conv0_apply", "stftmel", "avgpool", "gru",
                               "f0decode", "f0post", "embed", "zp", "sinegen", "avg3", "convpost", "knn_scan", "knn_select",
                               "knn_blend", "gather_rows", "fill", "wait"};
    std::string js = "[";
    bool first = true;
    double st_flops = 0, st_wbytes = 0, st_iobytes = 0;   // the persistent ContentVec stack: one launch, timed as a whole
    for (const Op& op : e.plan.ops) {
        if (op.kind == OP_WAIT || op.kind == OP_FILL) continue;
        if (op.kind == OP_F0POST) continue;  // stateful (rolls the pitch cache)
        if (op.fuse == 1) continue;          // covered by the fused residual-block kernel, timed under the block's last op
        if (op.stack && e.cvs.grid > 0) {
            if (op.kind == OP_GEMM) {
                const GemmOp& g = op.gemm;
                st_flops += 2.0 * g.M * double(g.N) * g.K; st_wbytes += 4.0 * double(g.N) * g.K;
                st_iobytes += 4.0 * (double(g.M) * g.K + double(g.M) * g.N * (g.R.null() ? 1 : 2));
            }
            if (&op != &e.plan.ops[size_t(e.plan.cvstack.first + e.plan.cvstack.count - 1)]) continue;
            launch_cvstack(e.cvs, s);
            CK(cudaStreamSynchronize(s));
            CK(cudaEventRecord(t0, s));
            for (int i = 0; i < iters; ++i) launch_cvstack(e.cvs, s);   // cooperative launches are not captured into a graph here
            CK(cudaEventRecord(t1, s));
            CK(cudaEventSynchronize(t1));
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, t0, t1));
            char buf[512];
            std::snprintf(buf, sizeof(buf), "%s{\"name\": \"cv.stack\", \"kind\": \"cvstack\", \"lane\": %d, \"us\": %.3f, \"flops\": %.0f, \"wbytes\": %.0f, \"iobytes\": %.0f, \"M\": %d, \"variant\": -1, \"splitk\": 1, \"chain\": -1}",
                          first ? "" : ", ", op.lane, double(ms) * 1e3 / iters, st_flops, st_wbytes, st_iobytes, e.plan.cvstack.T);
            js += buf; first = false;
            continue;
        }
        int n = 0;
        issue_one(ctx, op, B, s, &n);  // warm (also sets kernel attributes outside capture)
        CK(cudaStreamSynchronize(s));
        // `iters` launches captured into one graph: device time without host launch overhead
        cudaGraph_t g = nullptr; cudaGraphExec_t ge = nullptr;
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < iters; ++i) issue_one(ctx, op, B, s, &n);
        CK(cudaStreamEndCapture(s, &g));
        CK(cudaGraphInstantiate(&ge, g, 0));
        CK(cudaGraphLaunch(ge, s));  // warm replay
        CK(cudaEventRecord(t0, s));
        CK(cudaGraphLaunch(ge, s));
        CK(cudaEventRecord(t1, s));
        CK(cudaEventSynchronize(t1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, t0, t1));
        cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
        double flops = 0, wbytes = 0, iobytes = 0; long long grid = 0; int variant = -1, splitk = 1;
        if (op.kind == OP_GEMM) {
            variant = op.gemm.sched_variant; splitk = op.gemm.splitk;
            const GemmOp& g = op.gemm;
            flops = 2.0 * g.M * double(g.N) * g.K * g.batch;
            wbytes = 4.0 * double(g.N) * g.K * g.batch;
            double arows = double(g.M) * std::min<double>(double(g.K), double(g.lda > 0 ? g.lda : g.K)) + g.K;
            iobytes = 4.0 * (arows * g.batch + double(g.M) * g.N * g.batch * (1 + (g.R.null() ? 0 : 1) + (g.C2.null() ? 0 : 1)));
            grid = g.M;
        } else if (op.kind == OP_KNN_SCAN) {
            flops = 3.0 * double(op.kd.N) * op.kd.C * op.kd.Q; wbytes = 4.0 * double(op.kd.N) * op.kd.C;
        }
        char buf[512];
        std::snprintf(buf, sizeof(buf), "%s{\"name\": \"%s\", \"kind\": \"%s\", \"lane\": %d, \"us\": %.3f, \"flops\": %.0f, \"wbytes\": %.0f, \"iobytes\": %.0f, \"M\": %lld, \"variant\": %d, \"splitk\": %d, \"chain\": %d}",
                      first ? "" : ", ", op.name.c_str(), KN[op.kind], op.lane, double(ms) * 1e3 / iters, flops, wbytes, iobytes, grid, variant, splitk,
                      (op.chain >= 0 && size_t(op.chain) < e.chains.size()) ? op.chain : -1);
        js += buf; first = false;
    }
    js += "]";
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    if (out_bytes) *out_bytes = js.size();
    if (out && cap_bytes > 0) {
        if (js.size() + 1 > cap_bytes) return ctx->fail(RVC_ERR_INVALID_ARG, "profile buffer too small");
        std::memcpy(out, js.data(), js.size()); out[js.size()] = 0;
    }
    return RVC_OK;
}

// Timeline of one graph-replayed window: an external timed event is recorded after every op on the
// op's own lane inside the captured graph, so the end time of each op under real multi-lane
// concurrency can be read back (critical-path analysis; the event nodes add a little overhead).
int rvc_profile_timeline(rvc_ctx* ctx, char* out, size_t cap_bytes, size_t* out_bytes) {
    int rc = enter(ctx); if (rc) return rc;
    if (!ctx->last) return ctx->fail(RVC_ERR_INVALID_ARG, "nothing has run yet");
    PlanEntry& e = *ctx->last;
    ctx->sync_all();
    const DeviceBases B = ctx->bases(e);
    std::vector<cudaEvent_t> evs;
    std::vector<const Op*> which;
    std::string marks_s;
    { const char* m = getenv("RVC_TL_MARKS"); if (m && *m) marks_s = std::string(",") + m + ","; }
    const char* marks = marks_s.empty() ? nullptr : marks_s.c_str();
    cudaEvent_t t0; CK(cudaEventCreate(&t0));
    cudaStream_t s0 = ctx->streams[0];
    CK(cudaStreamBeginCapture(s0, cudaStreamCaptureModeThreadLocal));
    CK(cudaEventRecordWithFlags(t0, s0, cudaEventRecordExternal));
    size_t ev = 0; int n = 0;
    for (const Op& op : e.plan.ops) {
        if (op.kind == OP_WAIT) {
            if (ev >= ctx->events.size()) { cudaEvent_t x; CK(cudaEventCreateWithFlags(&x, cudaEventDisableTiming)); ctx->events.push_back(x); }
            CK(cudaEventRecord(ctx->events[ev], ctx->streams[op.wait.src_lane]));
            CK(cudaStreamWaitEvent(ctx->streams[op.wait.dst_lane], ctx->events[ev], 0));
            ++ev; continue;
        }
        if (op.stack && e.cvs.grid > 0) {
            if (&op != &e.plan.ops[size_t(e.plan.cvstack.first + e.plan.cvstack.count - 1)]) continue;
            n += launch_cvstack(e.cvs, ctx->streams[op.lane]);   // reported under the stack's LAST op
        } else if (op.chain >= 0 && size_t(op.chain) < e.chains.size()) {
            if (&op != &e.plan.ops[size_t(e.plan.chains[size_t(op.chain)].first + e.plan.chains[size_t(op.chain)].count - 1)]) continue;
            n += launch_chain(e.chains[size_t(op.chain)], ctx->streams[op.lane]);  // reported under the chain's LAST op
        } else {
            issue_one(ctx, op, B, ctx->streams[op.lane], &n);
        }
        // RVC_TL_MARKS=name,name,...: events only after these ops (a sparse timeline barely perturbs the window; an event
        // after every op serialises the lanes' launches and inflates it by a third)
        if (marks && std::strstr(marks, ("," + op.name + ",").c_str()) == nullptr) continue;
        cudaEvent_t x; CK(cudaEventCreate(&x));
        CK(cudaEventRecordWithFlags(x, ctx->streams[op.lane], cudaEventRecordExternal));
        evs.push_back(x); which.push_back(&op);
    }
    cudaGraph_t g = nullptr; cudaGraphExec_t ge = nullptr;
    CK(cudaStreamEndCapture(s0, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    for (int i = 0; i < 3; ++i) CK(cudaGraphLaunch(ge, s0));
    CK(cudaStreamSynchronize(s0));
    std::string js = "[";
    for (size_t i = 0; i < evs.size(); ++i) {
        float ms = 0.f; cudaEventElapsedTime(&ms, t0, evs[i]);
        char buf[256];
        std::snprintf(buf, sizeof(buf), "%s{\"name\": \"%s\", \"lane\": %d, \"end_us\": %.3f}", i ? ", " : "", which[i]->name.c_str(),
                      which[i]->lane, double(ms) * 1e3);
        js += buf; cudaEventDestroy(evs[i]);
    }
    js += "]";
    cudaEventDestroy(t0); cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
    if (out_bytes) *out_bytes = js.size();
    if (out && cap_bytes > 0) {
        if (js.size() + 1 > cap_bytes) return ctx->fail(RVC_ERR_INVALID_ARG, "profile buffer too small");
        std::memcpy(out, js.data(), js.size()); out[js.size()] = 0;
    }
    return RVC_OK;
}

// Phase-by-phase device time of every persistent chain of the last plan (globaltimer stamps written by
// CTA 0 at each phase start during the most recent run): JSON [{"chain","lane","grid","phases":[{"ops","us"}]}].
int rvc_profile_chains(rvc_ctx* ctx, char* out, size_t cap_bytes, size_t* out_bytes) {
    int rc = enter(ctx); if (rc) return rc;
    if (!ctx->last) return ctx->fail(RVC_ERR_INVALID_ARG, "nothing has run yet");
    PlanEntry& e = *ctx->last;
    ctx->sync_all();
    std::string js = "[";
    for (size_t c = 0; c < e.chains.size(); ++c) {
        const ChainInfo& ci = e.plan.chains[c];
        std::vector<unsigned long long> t(size_t(ci.n_phases + 1));
        CK(cudaMemcpy(t.data(), e.chains[c].d_dbg, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        char buf[256];
        std::snprintf(buf, sizeof(buf), "%s{\"chain\": %zu, \"lane\": %d, \"grid\": %d, \"t0_ns\": %llu, \"t1_ns\": %llu, \"phases\": [", c ? ", " : "", c, ci.lane, e.chains[c].grid,
                      t[0], t[size_t(ci.n_phases)]);
        js += buf;
        for (int ph = 0; ph < ci.n_phases; ++ph) {
            std::string names;
            for (int k = 0; k < ci.count; ++k)
                if (ci.phase[size_t(k)] == ph) {
                    const Op& op = e.plan.ops[size_t(ci.first + k)];
                    names += (names.empty() ? "" : ",") + op.name;
                    if (op.kind == OP_GEMM) names += "[v" + std::to_string(op.gemm.ch_variant) + "x" + std::to_string(op.gemm.ch_tiles_m * op.gemm.ch_tiles_n * op.gemm.batch) + "k" + std::to_string(op.gemm.ch_splitk) + "]";
                }
            std::snprintf(buf, sizeof(buf), "%s{\"ops\": \"%s\", \"us\": %.3f}", ph ? ", " : "", names.c_str(), double(t[size_t(ph + 1)] - t[size_t(ph)]) * 1e-3);
            js += buf;
        }
        js += "]}";
    }
    js += "]";
    if (out_bytes) *out_bytes = js.size();
    if (out && cap_bytes > 0) {
        if (js.size() + 1 > cap_bytes) return ctx->fail(RVC_ERR_INVALID_ARG, "profile buffer too small");
        std::memcpy(out, js.data(), js.size()); out[js.size()] = 0;
    }
    return RVC_OK;
}

// [phase][4] clock64 stamps (worker start, work done, arrived) of CTA 0 of the last persistent ContentVec stack launch
int rvc_debug_cvstack_stamps(rvc_ctx* ctx, long long* out, int n) {
    int rc = enter(ctx); if (rc) return rc;
    ctx->sync_all();
    if (!ctx->last || ctx->last->cvs.grid <= 0) return RVC_ERR_INVALID_ARG;
    cvstack_debug_read(out, n < 512 ? n : 512);
    if (n >= 512 + 1024) cvstack_debug_read2(out + 512, 1024);
    return ctx->last->cvs.n_phases;
}

// %globaltimer (ns) of the marker kernels of the last window: STFT start, F0 decode start, pitch cache start, retrieval
// gather start, conv_post end, RMVPE pool 0..4 start, GRU start, sine source start (12 values).  Written by the kernels themselves: no event nodes, the graph replays unperturbed.
int rvc_debug_lane_stamps(rvc_ctx* ctx, unsigned long long* out12) {
    int rc = enter(ctx); if (rc) return rc;
    ctx->sync_all();
    unsigned long long d[4];
    dsp_read_stamps(d);
    out12[0] = d[0]; out12[1] = d[1]; out12[2] = d[2]; out12[11] = d[3];
    misc_read_stamps(out12 + 3);
    return RVC_OK;
}

int rvc_debug_chain_stamps(rvc_ctx* ctx, int chain, long long* out2048) {
    int rc = enter(ctx); if (rc) return rc;
    if (!ctx->last || chain < 0 || size_t(chain) >= ctx->last->chains.size()) return RVC_ERR_INVALID_ARG;
    ctx->sync_all();
    launch_chain(ctx->last->chains[size_t(chain)], ctx->streams[0]);   // stand-alone replay on whatever the arena holds
    CK(cudaStreamSynchronize(ctx->streams[0]));
    chain_debug_read(out2048, 2048);
    chain_debug_read2(out2048 + 2048, 512);
    return RVC_OK;
}

int rvc_debug_knn_stamps(rvc_ctx* ctx, long long* out256) {
    int rc = enter(ctx); if (rc) return rc;
    ctx->sync_all();
    rvc::knn_umma_debug_read(out256);
    return RVC_OK;
}

int rvc_debug_umma_timing(rvc_ctx* ctx, const char* op_name, long long* out16) {
    int rc = enter(ctx); if (rc) return rc;
    if (!ctx->last) return RVC_ERR_INVALID_ARG;
    ctx->sync_all();
    const DeviceBases B = ctx->bases(*ctx->last);
    for (const Op& op : ctx->last->plan.ops) {
        std::string want(op_name); if (want.size() > 3 && want.substr(want.size() - 3) == "@v2") want = want.substr(0, want.size() - 3);
        if (op.name == want) { int n = 0; issue_one(ctx, op, B, ctx->streams[0], &n); issue_one(ctx, op, B, ctx->streams[0], &n); break; }
    }
    CK(cudaStreamSynchronize(ctx->streams[0]));
    if (std::string(op_name).find("@v2") != std::string::npos) rvc::v2_debug_read(out16); else { rvc::umma_debug_read(out16); rvc::umma_debug_read2(out16 + 16); }
    return RVC_OK;
}

}  // extern "C"
