// noise.h - stateless counter-based Gaussian noise for the synthesizer's random sources
// (enc_p posterior noise and SineGen additive noise).  The reference bakes these into the ONNX
// graph (rvc/src/rvc.rs:186-191,200-203 - the `rnd` input is commented out); here they are an
// explicit, reproducible function of (seed, window counter, kind, index).  The formula is
// mirrored by oracle/noise.py for parity tests.
#pragma once
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define RVC_HD __host__ __device__ __forceinline__
#else
#define RVC_HD inline
#endif

namespace rvc {

RVC_HD uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

RVC_HD uint64_t noise_key(uint64_t seed, uint64_t window, uint64_t kind) {
    return seed * 0x9E3779B97F4A7C15ull + window * 0xBF58476D1CE4E5B9ull + kind * 0x94D049BB133111EBull;
}

RVC_HD float noise_gauss(uint64_t key, uint64_t idx) {
    uint64_t x = splitmix64(key + idx);
    double u1 = (double(x >> 40) + 0.5) / 16777216.0;
    double u2 = (double((x >> 16) & 0xFFFFFFull) + 0.5) / 16777216.0;
    return float(sqrt(-2.0 * log(u1)) * cos(2.0 * 3.14159265358979323846 * u2));
}

}  // namespace rvc
