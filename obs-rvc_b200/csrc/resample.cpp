// resample.cpp - host-side construction of the resampler tables (resample.h).
#include "resample.h"

#include <cmath>
#include <complex>
#include <numeric>

namespace rvc {

namespace {

typedef std::complex<double> cd;
const double PI = 3.14159265358979323846264338327950288;

// in-place mixed-radix decimation-in-time FFT (any length; O(n * sum of prime factors)); sign = -1 forward, +1 inverse
void fft_rec(std::vector<cd>& x, int sign) {
    const size_t n = x.size();
    if (n <= 1) return;
    size_t p = 2;
    while (p * p <= n && n % p != 0) ++p;
    if (n % p != 0) p = n;   // prime length: plain DFT below
    const size_t m = n / p;
    std::vector<std::vector<cd>> sub(p, std::vector<cd>(m));
    for (size_t i = 0; i < n; ++i) sub[i % p][i / p] = x[i];
    if (m > 1) for (size_t r = 0; r < p; ++r) fft_rec(sub[r], sign);
    std::vector<cd> wp(p);   // p-th roots of unity
    for (size_t r = 0; r < p; ++r) wp[r] = std::polar(1.0, sign * 2.0 * PI * double(r) / double(p));
    for (size_t k = 0; k < m; ++k) {
        std::vector<cd> t(p);
        for (size_t r = 0; r < p; ++r) t[r] = sub[r][k] * std::polar(1.0, sign * 2.0 * PI * double(r * k) / double(n));
        for (size_t q = 0; q < p; ++q) {
            cd s = 0.0;
            for (size_t r = 0; r < p; ++r) s += t[r] * wp[(r * q) % p];
            x[k + q * m] = s;
        }
    }
}

double sinc(double t) { return t == 0.0 ? 1.0 : std::sin(PI * t) / (PI * t); }

}  // namespace

bool build_resample_table(int fs_in, int fs_out, int chunk_size_in, ResampleTable& T) {
    if (fs_in <= 0 || fs_out <= 0 || chunk_size_in <= 0) return false;
    const int g = std::gcd(fs_in, fs_out);
    T.fs_in = fs_in; T.fs_out = fs_out; T.a = fs_in / g; T.b = fs_out / g;
    T.u = (chunk_size_in + T.a - 1) / T.a;
    T.n_in = T.u * T.a; T.n_out = T.u * T.b;
    const int nin = T.n_in, nout = T.n_out;
    // cutoff in f32 arithmetic, as rubato computes it (synchro.rs FftResampler::new)
    float cutoff = std::pow(0.4f, 16.0f / float(nin));
    if (nin > nout) cutoff = cutoff * float(nout) / float(nin);
    // blackman_harris^2-windowed sinc over nin points, sum-normalised (sinc.rs make_sincs, windows.rs), / (2 nin)
    std::vector<double> h(size_t(nin), 0.0);
    double sum = 0.0;
    for (int n = 0; n < nin; ++n) {
        const double x = double(n) / double(nin);
        double w = 0.35875 - 0.48829 * std::cos(2 * PI * x) + 0.14128 * std::cos(4 * PI * x) - 0.01168 * std::cos(6 * PI * x);
        w *= w;
        h[size_t(n)] = w * sinc(double(n - nin / 2) * double(cutoff));
        sum += h[size_t(n)];
    }
    std::vector<cd> f(size_t(2 * nin), cd(0.0, 0.0));
    for (int n = 0; n < nin; ++n) f[size_t(n)] = h[size_t(n)] / sum / double(2 * nin);
    fft_rec(f, -1);
    const int new_len = nin < nout ? nin + 1 : nout;
    // kappa = unnormalised inverse real FFT of length L = 2 a b u of the truncated spectrum (Hermitian extension)
    const long long L = 2LL * T.a * T.b * T.u;
    std::vector<cd> spec(size_t(L), cd(0.0, 0.0));
    spec[0] = cd(f[0].real(), 0.0);   // realfft ignores the imaginary part of the DC bin
    for (int k = 1; k < new_len; ++k) { spec[size_t(k)] = f[size_t(k)]; spec[size_t(L - k)] = std::conj(f[size_t(k)]); }
    fft_rec(spec, +1);
    T.period = int(L / T.b);
    T.kappa.assign(size_t(L), 0.f);
    for (long long j = 0; j < L; ++j) T.kappa[size_t((j % T.b) * T.period + j / T.b)] = float(spec[size_t(j)].real());
    return true;
}

}  // namespace rvc
