// rvc_b200.hpp - C++ host mirror of the reference's `rvc` crate public API over the C ABI.
//
// Same method set, argument meaning and error behaviour as `rvc::RvcInfer`
// (/root/reference/rvc/src/rvc.rs:18-220, exported by rvc/src/lib.rs:5): `Result<_, RvcInferError>`
// becomes a thrown `rvc::RvcInferError` carrying the `rvc_status` code.  Header-only; link with
// -lrvc_b200.  (The reference host language is Rust; no Rust toolchain exists in the build image,
// so the compiled host side is C++ - the equivalent Rust crate sources are in obs-rvc_b200/rust/.)
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "rvc_b200.h"

namespace rvc {

// rvc-common/src/enums.rs:3-28
enum class RvcModelVersion { V1 = RVC_MODEL_V1, V2 = RVC_MODEL_V2 };
enum class PitchAlgorithm { Rmvpe = RVC_PITCH_RMVPE };

// rvc-common/src/errors.rs:1-20
struct RvcInferError : std::runtime_error {
    int code;
    RvcInferError(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

class RvcInfer {
public:
    // RvcInfer::new(data_path) - rvc.rs:30-44
    explicit RvcInfer(const std::string& data_path, const rvc_config* cfg = nullptr) {
        int rc = rvc_create(data_path.c_str(), cfg, &ctx_);
        if (rc != RVC_OK) throw RvcInferError(rc, rvc_last_create_error());
    }
    ~RvcInfer() { rvc_destroy(ctx_); }
    RvcInfer(const RvcInfer&) = delete;
    RvcInfer& operator=(const RvcInfer&) = delete;

    void load_contentvec(RvcModelVersion v) { chk(rvc_load_contentvec(ctx_, int(v))); }        // rvc.rs:46
    void load_model(const std::string& path) { chk(rvc_load_model(ctx_, path.c_str())); }      // rvc.rs:56
    void load_f0(PitchAlgorithm a) { chk(rvc_load_f0(ctx_, int(a))); }                         // rvc.rs:62
    void unload_model() { chk(rvc_unload_model(ctx_)); }                                       // rvc.rs:77
    void load_index(const std::string& path, float rate) { chk(rvc_load_index(ctx_, path.c_str(), rate)); }

    // rvc.rs:81-97: returns (1, C, T) flattened; shape in c/t
    std::vector<float> hubert(const std::vector<float>& pcm, size_t& c, size_t& t) {
        std::vector<float> out(1024 * (pcm.size() / 320 + 2));
        chk(rvc_hubert(ctx_, pcm.data(), pcm.size(), out.data(), out.size(), &c, &t));
        out.resize(c * t);
        return out;
    }
    // rvc.rs:99-109: (1, 2T+1, C)
    std::vector<float> extract_feature(const std::vector<float>& pcm, size_t& frames, size_t& c) {
        std::vector<float> out(1024 * (2 * (pcm.size() / 320 + 2) + 1));
        chk(rvc_extract_feature(ctx_, pcm.data(), pcm.size(), out.data(), out.size(), &frames, &c));
        out.resize(frames * c);
        return out;
    }
    // rvc.rs:111-131
    std::vector<float> pitch(const std::vector<float>& pcm, int32_t pitch_shift, size_t sample_frame_16k_size) {
        std::vector<float> out(4096);
        size_t n = 0;
        chk(rvc_pitch(ctx_, pcm.data(), pcm.size(), pitch_shift, sample_frame_16k_size, out.data(), out.size(), &n));
        out.resize(n);
        return out;
    }
    // rvc.rs:133-220 (argument order of rvcadapter.rs:60-67)
    std::vector<float> infer(const float* pcm, size_t n, size_t sample_frame_16k_size, int32_t pitch_shift, uint32_t skip_head,
                             uint32_t return_length) {
        std::vector<float> out(size_t(return_length) * 480 + 16);
        size_t len = 0;
        chk(rvc_infer(ctx_, pcm, n, uint32_t(sample_frame_16k_size), pitch_shift, skip_head, return_length, out.data(), out.size(), &len));
        out.resize(len);
        return out;
    }
    rvc_ctx* handle() { return ctx_; }

private:
    void chk(int rc) { if (rc != RVC_OK) throw RvcInferError(rc, rvc_last_error(ctx_)); }
    rvc_ctx* ctx_ = nullptr;
};

}  // namespace rvc
