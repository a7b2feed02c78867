// model.h - RVCW container reader, weight packing and plan construction (host-only C++).
//
// Replaces the reference's ORT session builders (rvc/src/models.rs:7-76): instead of handing an
// opaque .onnx graph to ONNX Runtime, the tensors are read from an `.rvcw` container, re-laid
// out once for the implicit-GEMM kernels (channels-last, tap-major K, BN / weight-norm / constant
// conditioning folded) and kept resident in HBM.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "ops.h"

namespace rvc {

struct HostTensor {
    std::vector<int64_t> shape;
    int32_t dtype = 0;  // 0 f32, 1 i32
    const void* data = nullptr;
    int64_t numel() const { int64_t n = 1; for (auto d : shape) n *= d; return n; }
    const float* f() const { return static_cast<const float*>(data); }
    const int32_t* i() const { return static_cast<const int32_t*>(data); }
};

struct RvcwFile {
    std::vector<uint8_t> blob;
    std::map<std::string, HostTensor> t;
    bool load(const std::string& path, std::string& err);
    const HostTensor* find(const std::string& name) const;
};

// One weight arena in host staging memory; uploaded verbatim to the device.
struct Packed {
    std::vector<float> host;
    std::map<std::string, int64_t> off;  // element offsets
    int64_t add(const std::string& name, int64_t elems);  // 256-byte aligned, zero-filled
    int64_t at(const std::string& name) const;
    float* p(int64_t o) { return host.data() + o; }
};

struct CvInfo { int32_t n_layers = 12, out_dim = 768; bool final_proj = false; };
struct F0Info { float in_scale = 1.f, in_shift = 0.f; int32_t mel_nnz = 0; };
struct SynInfo {
    int32_t sr = 40000, phone_dim = 768; float lin_w = 1.f, lin_b = 0.f;
    // generator config by output rate (upstream RVC configs: 32k / 40k / 48k): ConvTranspose1d strides and kernel sizes
    int32_t rates[4] = {10, 10, 2, 2}, up_kernels[4] = {16, 16, 4, 4};
};
// fills rates / up_kernels for a supported output rate; false otherwise
inline bool syn_config_for_rate(SynInfo& s) {
    static const int R[3][9] = {{32000, 10, 8, 2, 2, 20, 16, 4, 4}, {40000, 10, 10, 2, 2, 16, 16, 4, 4}, {48000, 12, 10, 2, 2, 24, 20, 4, 4}};
    for (const auto& r : R)
        if (r[0] == s.sr) { for (int i = 0; i < 4; ++i) { s.rates[i] = r[1 + i]; s.up_kernels[i] = r[5 + i]; } return true; }
    return false;
}

bool pack_contentvec(const RvcwFile& f, Packed& out, CvInfo& info, std::string& err);
bool pack_rmvpe(const RvcwFile& f, Packed& out, F0Info& info, std::string& err);
bool pack_synth(const RvcwFile& f, Packed& out, SynInfo& info, std::string& err);

// Geometry of one call (obs-rvc/src/lib.rs:200-227).
struct Geometry {
    int32_t n16k = 0, sf16k = 0, skip_head = 0, return_length = 0;
    bool operator<(const Geometry& o) const {
        if (n16k != o.n16k) return n16k < o.n16k;
        if (sf16k != o.sf16k) return sf16k < o.sf16k;
        if (skip_head != o.skip_head) return skip_head < o.skip_head;
        return return_length < o.return_length;
    }
};

struct PlanOptions {
    int32_t index_k = 8;
    int32_t upstream_cents_window = 0;
    bool with_index = false; int32_t index_rows = 0;
    bool multi_lane = true;  // independent branches on separate stream lanes
    bool allow_umma = true;  // dense GEMMs on the tcgen05 3xTF32 kernel (needs hi/lo weight copies)
    // persistent chain kernels (chain.h): CTA budget of a chain on lane 0 (nothing else running) and
    // on a side lane (shares the GPU with the tcgen05 GEMMs of lane 0); 0 disables chains
    int32_t chain_grid_main = 0, chain_grid_side = 0;
    int32_t chain_side_max_m = 0;  // side-lane GEMMs join a chain only up to this many rows (0 = no limit)
    // Batched plans (PLAN_INFER only): `nb` windows per launch.  Work arena and state block are replicated per window
    // (Plan::work_bytes / Plan::state_block apart), weights and index are shared.  sequential = the windows are
    // consecutive windows of ONE stream (offline conversion): one pitch cache, updated in window order; otherwise they
    // belong to nb independent streams with their own caches.
    int32_t nb = 1;
    bool sequential = false;
    int32_t index_cols = 0;        // width of the loaded retrieval index (must equal the ContentVec width)
    int64_t index_planes_off = 0;  // > 0: the index carries fp16 planes (tensor-core candidate pass, kernels_knn_umma.cu)
    float index_ymax2 = 0.f;       // max |y|^2 over the index rows (error bound of the candidate pass)
    bool cv_stack = false;         // ContentVec transformer layers as one persistent tcgen05 kernel (single-window plans, T <= 128)
    int fuse_cbr = 0;              // RMVPE residual blocks of U-Net levels 0 / 1 as one kernel each (kernels_cbr.cu): 1 all, 2 decoder only
    bool f0_umma = false;          // RMVPE's wide levels on the tcgen05 FP16-split kernel (batched plans; needs the f0 weight planes)
};

struct Plan {
    std::vector<Op> ops;
    std::vector<NamedBuf> bufs;
    std::vector<ChainInfo> chains;
    CvStackInfo cvstack;
    int64_t work_bytes = 0;
    int32_t n_lanes = 1;
    // well-known buffers
    Ref pcm, audio, params, cache;
    int32_t hubert_T = 0, hubert_C = 0, f0_T = 0, audio_len = 0, knn_q = 0;
    int32_t nb = 1;               // windows per launch
    int64_t state_block = 0;      // bytes between the state blocks of consecutive windows (batched plans; 0 = the context's own state arena)
    const NamedBuf* find(const std::string& name) const;
};

enum PlanKind : int32_t { PLAN_INFER = 0, PLAN_HUBERT = 1, PLAN_PITCH = 2, PLAN_MEL = 3,
                          PLAN_KNN = 4, PLAN_FEATURE = 5 };

// Builds the op list.  `cv`/`f0`/`syn` may be null when the plan kind does not need them.
// Persistent buffers (params, pitch cache, pcm in, audio out, knn queries) live in SP_STATE at
// the fixed offsets of `StateLayout`, identical for every plan of a context.
bool build_plan(PlanKind kind, const Geometry& g, const PlanOptions& opt, const Packed* cv,
                const CvInfo* cvi, const Packed* f0, const F0Info* f0i, const Packed* syn,
                const SynInfo* syi, Plan& plan, std::string& err);

struct StateLayout {
    static const int64_t PCM_CAP = 1 << 20;     // samples (65 s @16 kHz)
    static const int64_t AUDIO_CAP = 1 << 22;   // floats: audio out / generic result staging
    static const int64_t CACHE_LEN = 1024;      // rvc.rs:42
    static const int64_t off_params = 0;                                   // RunParams (64 B)
    static const int64_t off_cache = 256;                                  // f32[1024]
    static const int64_t off_pcm = off_cache + CACHE_LEN * 4;              // f32[PCM_CAP]
    static const int64_t off_audio = off_pcm + PCM_CAP * 4;                // f32[AUDIO_CAP]
    static const int64_t bytes = off_audio + AUDIO_CAP * 4;
};

}  // namespace rvc
