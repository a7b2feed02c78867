// rvc_rpc.cpp - wire-compatible replacement of the reference's `rvc-rpc` child process
// (/root/reference/rvc-rpc/src/main.rs:8-103), driving the B200 engine through the C++ mirror.
//
// CLI:      rvc-rpc <version: v1|v2> <f0: rmvpe> <model> <data>           (main.rs:12-22)
// Request:  u32 LE nbytes, f32 LE[nbytes/4], u32 sf16k, i32 shift, u32 skip_head, u32 return_length
// Response: u32 LE nbytes, f32 LE[]                                      (main.rs:64-100,
//           obs-rvc/src/rvcadapter.rs:69-118).  Blocking, one request in flight; any error exits
//           the process like the reference's unwrap()/panic!, which the adapter answers by respawning
//           the child (obs-rvc/src/lib.rs:716-724).  An optional 5th argument is an index file and a
//           6th the index rate (the reference stores both and never uses them, lib.rs:78,81,264).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rvc_b200.hpp"

static bool read_exact(void* dst, size_t n) { return std::fread(dst, 1, n, stdin) == n; }

int main(int argc, char** argv) {
    if (argc < 5) {
        std::fprintf(stderr, "Usage: rvc-rpc <version> <f0_algorithm> <model> <data> [index [index_rate]]\n");
        return 0;  // main.rs:14-17 returns without error
    }
    // enums.rs:66-83,126-141: unknown strings fall back to V2 / Rmvpe
    const rvc::RvcModelVersion ver = std::strcmp(argv[1], "v1") == 0 ? rvc::RvcModelVersion::V1 : rvc::RvcModelVersion::V2;
    try {
        rvc::RvcInfer eng(argv[4]);
        eng.load_contentvec(ver);
        eng.load_f0(rvc::PitchAlgorithm::Rmvpe);
        eng.load_model(argv[3]);
        if (argc > 5) eng.load_index(argv[5], argc > 6 ? float(std::atof(argv[6])) : 0.0f);
        static char ibuf[1 << 20], obuf[1 << 20];          // 1 MiB buffers as main.rs:59-60
        std::setvbuf(stdin, ibuf, _IOFBF, sizeof(ibuf));
        std::setvbuf(stdout, obuf, _IOFBF, sizeof(obuf));
        std::fprintf(stderr, "Ready to receive input\n");
        std::vector<float> pcm;
        for (;;) {
            uint32_t nbytes, sf16k, skip_head, return_length; int32_t shift;
            if (!read_exact(&nbytes, 4)) return 1;           // EOF = parent gone
            // exactly nbytes payload bytes are consumed, as main.rs:66-70 does (read_exact of the whole buffer, then
            // chunks_exact(4) drops a ragged tail): a length that is not a multiple of 4 must not desynchronise the stream
            std::vector<unsigned char> raw(nbytes);
            if (nbytes && !read_exact(raw.data(), nbytes)) return 1;
            pcm.resize(nbytes / 4);
            if (!pcm.empty()) std::memcpy(pcm.data(), raw.data(), size_t(nbytes / 4) * 4);
            if (!read_exact(&sf16k, 4) || !read_exact(&shift, 4) || !read_exact(&skip_head, 4) || !read_exact(&return_length, 4)) return 1;
            std::vector<float> out = eng.infer(pcm.data(), pcm.size(), sf16k, shift, skip_head, return_length);
            const uint32_t obytes = uint32_t(out.size() * 4);
            std::fwrite(&obytes, 4, 1, stdout);
            std::fwrite(out.data(), 4, out.size(), stdout);
            std::fflush(stdout);
        }
    } catch (const rvc::RvcInferError& e) {
        std::fprintf(stderr, "rvc-rpc: %s (status %d)\n", e.what(), e.code);
        return 101;  // Rust panic exit code
    } catch (const std::exception& e) {   // e.g. bad_alloc from an absurd length field of an untrusted peer
        std::fprintf(stderr, "rvc-rpc: %s\n", e.what());
        return 101;
    }
}
