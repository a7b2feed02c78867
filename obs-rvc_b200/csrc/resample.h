// resample.h - rubato 0.15.0 `FftFixedInOut` (obs-rvc/src/lib.rs:236-242) as a direct-form polyphase filter table.
//
// rubato resamples one chunk as  rfft(zero-padded chunk) x rfft(windowed sinc), truncated / zero-extended to the output
// length, inverse rfft, overlap-add (src/synchro.rs; the algorithm is restated with its sources in oracle/resample.py).
// That is a fixed LINEAR map of the chunk:  y[m] = sum_n x[n] kappa[(a m - b n) mod L],  a = fs_in / gcd, b = fs_out / gcd,
// L = 2 a b u (u = chunks of the gcd grid), kappa = the length-L inverse real FFT of the truncated filter spectrum.  The
// table is built once per stream in double precision on the host (mixed-radix FFT below), stored by phase
// (kappa_ph[r][q] = kappa[q b + r]) so that one output reads a contiguous, descending run of it, and applied on the GPU
// by resample_kernel (kernels_rt.cu).  No golden exists for the resamplers (parity unpinned).
#pragma once
#include <cstdint>
#include <vector>

namespace rvc {

struct ResampleTable {
    int fs_in = 0, fs_out = 0, a = 0, b = 0, u = 0;
    int n_in = 0, n_out = 0;     // fft_size_in / fft_size_out of rubato = samples consumed / produced per chunk
    int period = 0;              // entries per phase = 2 a u
    std::vector<float> kappa;    // [b][period]
};

// chunk_size_in as passed to FftFixedInOut::new; returns false for degenerate arguments
bool build_resample_table(int fs_in, int fs_out, int chunk_size_in, ResampleTable& out);

}  // namespace rvc
