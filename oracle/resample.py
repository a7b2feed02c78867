"""TEST INFRASTRUCTURE (oracle).  numpy restatement of rubato 0.15.0 `FftFixedInOut` (synchronous FFT resampler), the
crate the reference's streaming loop uses for 48 k -> 16 k and model-rate -> 48 k (obs-rvc/src/lib.rs:236-242, call
sites :673-683 and :742-756).  rubato is a crates.io dependency (Cargo.lock:1223, version 0.15.0) that is NOT vendored in
/root/reference, so this follows its published algorithm (src/synchro.rs `FftResampler::new` / `resample_unit`,
src/sinc.rs `make_sincs`, src/windows.rs `blackman_harris` squared):

  fft_chunks  = ceil(chunk_size_in / (fs_in / gcd));  fft_size_in = fft_chunks * fs_in / gcd;  fft_size_out likewise
  cutoff      = 0.4^(16 / fft_size_in) * (fft_size_out / fft_size_in  if downsampling else 1)
  filter_t    = blackman_harris2-windowed sinc over fft_size_in points, sum-normalised, / (2 fft_size_in), zero-padded to 2 fft_size_in
  per chunk   : X = rfft([x, 0...0] (2 fft_size_in));  Y[:new_len] = X[:new_len] * rfft(filter_t)[:new_len], rest 0
                (new_len = fft_size_in + 1 when upsampling, fft_size_out when downsampling)
                y = unnormalised irfft(Y, 2 fft_size_out);  out = y[:fft_size_out] + overlap;  overlap = y[fft_size_out:]

PARITY UNPINNED: the reference holds no golden vector for the resamplers and rubato cannot be built here (no Rust
toolchain); the restatement is checked through properties only (tests/test_stream_cpu.py: a band-limited sine comes
out as the same sine at the new rate, delayed by fft_size_in / 2 input samples; linearity; chunk-to-chunk continuity)."""
import math

import numpy as np


def blackman_harris2(npoints: int) -> np.ndarray:
    """windows.rs `blackman_harris` (periodic form, divides by npoints), squared."""
    x = np.arange(npoints, dtype=np.float64)
    w = (0.35875 - 0.48829 * np.cos(2 * np.pi * x / npoints) + 0.14128 * np.cos(4 * np.pi * x / npoints)
         - 0.01168 * np.cos(6 * np.pi * x / npoints))
    return w * w


def make_sinc(npoints: int, f_cutoff: float) -> np.ndarray:
    """sinc.rs `make_sincs(npoints, 1, f_cutoff, BlackmanHarris2)[0]`."""
    x = np.arange(npoints, dtype=np.float64) - (npoints // 2)
    y = blackman_harris2(npoints) * np.sinc(x * f_cutoff)     # np.sinc(t) = sin(pi t) / (pi t)
    return y / y.sum()


class FftFixedInOut:
    def __init__(self, fs_in: int, fs_out: int, chunk_size_in: int):
        g = math.gcd(fs_in, fs_out)
        self.fft_chunks = -(-chunk_size_in // (fs_in // g))
        self.fft_size_in = self.fft_chunks * (fs_in // g)
        self.fft_size_out = self.fft_chunks * (fs_out // g)
        nin, nout = self.fft_size_in, self.fft_size_out
        cutoff = float(np.float32(0.4) ** np.float32(16.0 / nin))          # computed in f32 by rubato
        if nin > nout:
            cutoff = float(np.float32(cutoff) * np.float32(nout) / np.float32(nin))
        filter_t = np.zeros(2 * nin)
        filter_t[:nin] = make_sinc(nin, cutoff) / (2 * nin)
        self.filter_f = np.fft.rfft(filter_t)
        self.new_len = nin + 1 if nin < nout else nout
        self.overlap = np.zeros(nout)

    def input_frames_next(self) -> int:
        return self.fft_size_in

    def output_frames_max(self) -> int:
        return self.fft_size_out

    def process(self, wave_in: np.ndarray) -> np.ndarray:
        nin, nout = self.fft_size_in, self.fft_size_out
        assert wave_in.shape[0] == nin
        buf = np.zeros(2 * nin)
        buf[:nin] = wave_in
        spec = np.fft.rfft(buf)
        out_f = np.zeros(nout + 1, dtype=np.complex128)
        out_f[:self.new_len] = spec[:self.new_len] * self.filter_f[:self.new_len]
        # realfft's inverse ignores the imaginary parts of the DC / Nyquist bins and does not normalise
        y = np.fft.irfft(out_f, 2 * nout) * (2 * nout)
        out = y[:nout] + self.overlap
        self.overlap = y[nout:].copy()
        return out.astype(np.float32)
