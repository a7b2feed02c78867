// cpu_exec.cpp - scalar CPU interpreter of the engine's op descriptors.
//
// TEST INFRASTRUCTURE ONLY (lives under oracle/; never linked into librvc_b200.so).  It replays
// the SAME plan (obs-rvc_b200/csrc/plan.cpp) over the SAME packed weights (model.cpp) on host
// memory, one straightforward loop nest per op, accumulating in double.  Uses:
//   * here (no GPU): plan + packing + op semantics are validated against the torch oracle;
//   * on the GPU box: per-op reference for every CUDA kernel (first divergent op is named).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../obs-rvc_b200/csrc/model.h"
#include "../../obs-rvc_b200/csrc/noise.h"

using namespace rvc;

namespace {

struct Bases { uint8_t* b[SP_COUNT] = {nullptr}; };

// no OpenMP runtime in this image: plain std::thread fan-out
template <typename F> void parallel_for(int n, F fn) {
    int nt = int(std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u));
    if (n < 2 * nt) { for (int i = 0; i < n; ++i) fn(i); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
        th.emplace_back([=]() { for (int i = t; i < n; i += nt) fn(i); });
    for (auto& x : th) x.join();
}

template <typename T> T* P(const Bases& B, const Ref& r) {
    return r.null() ? nullptr : reinterpret_cast<T*>(B.b[r.space] + r.off);
}

float act_f(int act, float v) {
    switch (act) {
        case ACT_GELU: return float(0.5 * double(v) * (1.0 + std::erf(double(v) / std::sqrt(2.0))));
        case ACT_RELU: return v > 0 ? v : 0.f;
        case ACT_LRELU01: return v > 0 ? v : 0.1f * v;
        case ACT_LRELU001: return v > 0 ? v : 0.01f * v;
        case ACT_SIGMOID: return float(1.0 / (1.0 + std::exp(-double(v))));
        case ACT_TANH: return float(std::tanh(double(v)));
        default: return v;
    }
}

void run_gemm(const GemmOp& g, const Bases& B) {
    for (int bt = 0; bt < g.batch; ++bt) {
        const float* A = P<float>(B, g.A) + bt * g.sA;
        const float* W = P<float>(B, g.W) + bt * g.sW;
        const float* bias = g.bias.null() ? nullptr : P<float>(B, g.bias) + bt * g.sBias;
        float* C = P<float>(B, g.C) + bt * g.sC;
        float* C2 = g.C2.null() ? nullptr : P<float>(B, g.C2) + bt * g.sC;
        const float* R = g.R.null() ? nullptr : P<float>(B, g.R) + bt * g.sR;
        parallel_for(g.M, [&](int m) {
            std::vector<float> row(g.N);
            for (int n = 0; n < g.N; ++n) {
                double acc = 0.0;
                const float* w = W + int64_t(n) * g.ldw;
                for (int k0 = 0, s = 0; k0 < g.K; k0 += g.seg_len, ++s) {
                    const float* a = A + int64_t(m) * g.lda + int64_t(s) * g.seg_stride;
                    int len = std::min(g.seg_len, g.K - k0);
                    for (int k = 0; k < len; ++k) acc += double(a[k]) * double(w[k0 + k]);
                }
                row[n] = float(double(g.alpha) * acc + (bias ? double(bias[n]) : 0.0));
            }
            const bool masked = g.mask_period > 0 && (m % g.mask_period) >= g.mask_valid;
            const int ncols = g.act == ACT_GATE ? g.N / 2 : g.N;
            for (int col = 0; col < ncols; ++col) {
                float v;
                if (g.act == ACT_GATE) v = float(std::tanh(double(row[2 * col])) * (1.0 / (1.0 + std::exp(-double(row[2 * col + 1])))));
                else v = act_f(g.act, row[col]);
                int64_t idx, idx2;
                if (g.out_mode == OUT_PLAIN) {
                    idx = int64_t(m) * g.ldc + col; idx2 = int64_t(m) * g.ldc2 + col;
                    if (R) v += R[int64_t(m) * g.ldr + col];
                } else if (g.out_mode == OUT_PIXSHUF2) {
                    int qt = m / g.om_a, qf = m % g.om_a;
                    if (qf >= g.om_a - 2) break;
                    int ph = col / g.om_b, co = col % g.om_b, rt = ph / 2, rf = ph % 2;
                    idx = (int64_t(2 * qt + rt) * g.om_c + (2 * qf + rf)) * g.ldc + co; idx2 = idx;
                } else {
                    int r = col / g.om_b, co = col % g.om_b;
                    int o = m * g.om_a + r - g.om_c;
                    if (o < 0 || o >= g.om_d) continue;
                    idx = int64_t(o) * g.ldc + co; idx2 = int64_t(o) * g.ldc2 + co;
                }
                if (masked) v = 0.f;
                C[idx] = v;
                if (C2) C2[idx2] = masked ? 0.f : act_f(g.act2, v);
            }
        });
    }
}

void run_layernorm(const LayerNormOp& o, const Bases& B) {
    const float* X = P<float>(B, o.X); float* Y = P<float>(B, o.Y);
    const float* g = P<float>(B, o.gamma); const float* b = P<float>(B, o.beta);
    for (int r = 0; r < o.rows; ++r) {
        const float* x = X + int64_t(r) * o.ldx;
        double mean = 0, var = 0;
        for (int c = 0; c < o.cols; ++c) mean += x[c];
        mean /= o.cols;
        for (int c = 0; c < o.cols; ++c) var += (x[c] - mean) * (x[c] - mean);
        var /= o.cols;
        double rstd = 1.0 / std::sqrt(var + double(o.eps));
        for (int c = 0; c < o.cols; ++c) Y[int64_t(r) * o.ldy + c] = float((x[c] - mean) * rstd * g[c] + b[c]);
    }
}

void run_attn(const AttnOp& o, const Bases& B) {
    const float* qkv = P<float>(B, o.qkv); float* out = P<float>(B, o.out);
    const int HD = o.heads * o.dim;
    std::vector<double> s(o.T);
    for (int h = 0; h < o.heads; ++h)
        for (int i = 0; i < o.T; ++i) {
            const float* q = qkv + int64_t(i) * o.ldqkv + h * o.dim;
            double mx = -1e300;
            for (int j = 0; j < o.T; ++j) {
                const float* k = qkv + int64_t(j) * o.ldqkv + HD + h * o.dim;
                double a = 0; for (int d = 0; d < o.dim; ++d) a += double(q[d]) * k[d];
                s[j] = a; mx = std::max(mx, a);
            }
            double sum = 0; for (int j = 0; j < o.T; ++j) { s[j] = std::exp(s[j] - mx); sum += s[j]; }
            for (int d = 0; d < o.dim; ++d) {
                double a = 0;
                for (int j = 0; j < o.T; ++j) a += s[j] * qkv[int64_t(j) * o.ldqkv + 2 * HD + h * o.dim + d];
                out[int64_t(i) * o.ldo + h * o.dim + d] = float(a / sum);
            }
        }
}

void run_relattn(const RelAttnOp& o, const Bases& B) {
    const float* qkv = P<float>(B, o.qkv); float* out = P<float>(B, o.out);
    const float* rk = P<float>(B, o.rel_k); const float* rv = P<float>(B, o.rel_v);
    const int HD = o.heads * o.dim, Wn = o.window;
    std::vector<double> s(o.T);
    for (int h = 0; h < o.heads; ++h)
        for (int i = 0; i < o.T; ++i) {
            const float* q = qkv + int64_t(i) * o.ldqkv + h * o.dim;
            double mx = -1e300;
            for (int j = 0; j < o.T; ++j) {
                const float* k = qkv + int64_t(j) * o.ldqkv + HD + h * o.dim;
                double a = 0; for (int d = 0; d < o.dim; ++d) a += double(q[d]) * k[d];
                int rel = j - i + Wn;
                if (rel >= 0 && rel <= 2 * Wn) for (int d = 0; d < o.dim; ++d) a += double(q[d]) * rk[rel * o.dim + d];
                s[j] = a; mx = std::max(mx, a);
            }
            double sum = 0; for (int j = 0; j < o.T; ++j) { s[j] = std::exp(s[j] - mx); sum += s[j]; }
            for (int d = 0; d < o.dim; ++d) {
                double a = 0;
                for (int j = 0; j < o.T; ++j) {
                    double p = s[j];
                    a += p * qkv[int64_t(j) * o.ldqkv + 2 * HD + h * o.dim + d];
                    int rel = j - i + Wn;
                    if (rel >= 0 && rel <= 2 * Wn) a += p * rv[rel * o.dim + d];
                }
                out[int64_t(i) * o.ldo + h * o.dim + d] = float(a / sum);
            }
        }
}

void run_conv0_stats(const Conv0StatsOp& o, const Bases& B) {
    const float* x = P<float>(B, o.pcm); const float* w = P<float>(B, o.w); float* st = P<float>(B, o.stats);
    parallel_for(o.C, [&](int c) {
        double s = 0, ss = 0;
        for (int t = 0; t < o.T; ++t) {
            double a = 0; for (int j = 0; j < o.k; ++j) a += double(x[t * o.stride + j]) * w[c * o.k + j];
            s += a; ss += a * a;
        }
        double mean = s / o.T, var = ss / o.T - mean * mean;
        st[2 * c] = float(mean); st[2 * c + 1] = float(1.0 / std::sqrt(var + double(o.eps)));
    });
}

void run_conv0_apply(const Conv0ApplyOp& o, const Bases& B) {
    const float* x = P<float>(B, o.pcm); const float* w = P<float>(B, o.w); const float* st = P<float>(B, o.stats);
    const float* g = P<float>(B, o.gamma); const float* b = P<float>(B, o.beta); float* Y = P<float>(B, o.Y);
    parallel_for(o.T, [&](int t) {
        for (int c = 0; c < o.C; ++c) {
            double a = 0; for (int j = 0; j < o.k; ++j) a += double(x[t * o.stride + j]) * w[c * o.k + j];
            float v = float((a - double(st[2 * c])) * double(st[2 * c + 1]) * g[c] + b[c]);
            Y[int64_t(t) * o.C + c] = act_f(ACT_GELU, v);
        }
    });
}

void run_stftmel(const StftMelOp& o, const Bases& B) {
    const float* x = P<float>(B, o.pcm); const float* win = P<float>(B, o.window);
    const int32_t* bs = P<int32_t>(B, o.band_start); const int32_t* bc = P<int32_t>(B, o.band_count);
    const int32_t* bo = P<int32_t>(B, o.band_off); const float* bw = P<float>(B, o.band_w);
    float* mel = P<float>(B, o.mel); float* out2 = P<float>(B, o.out2);
    std::vector<double> ct(1024), st(1024);
    for (int i = 0; i < 1024; ++i) { ct[i] = std::cos(2 * M_PI * i / 1024.0); st[i] = std::sin(2 * M_PI * i / 1024.0); }
    parallel_for(o.T, [&](int t) {
        std::vector<double> fr(1024), mag(513);
        for (int j = 0; j < 1024; ++j) {
            int p = t * 160 + j - 512;                       // index into the unpadded signal
            if (p < 0) p = -p;                               // reflect (rmvpe.rs:59-61)
            if (p >= o.L) p = 2 * (o.L - 1) - p;             // rmvpe.rs:64-66
            fr[j] = double(float(x[p] * win[j]));
        }
        for (int k = 0; k < 513; ++k) {
            double re = 0, im = 0;
            for (int j = 0; j < 1024; ++j) { int idx = (k * j) & 1023; re += fr[j] * ct[idx]; im -= fr[j] * st[idx]; }
            mag[k] = std::sqrt(re * re + im * im);
        }
        for (int m = 0; m < 128; ++m) {
            double a = 0;
            for (int j = 0; j < bc[m]; ++j) a += double(bw[bo[m] + j]) * mag[bs[m] + j];
            float v = std::log(std::max(float(a), o.clamp));
            mel[t * 128 + m] = v;
            if (out2) out2[int64_t(t) * o.out2_pitch + m] = v * o.scale + o.shift;
        }
    });
}

void run_avgpool(const AvgPoolOp& o, const Bases& B) {
    const float* in = P<float>(B, o.in); float* out = P<float>(B, o.out);
    const int To = o.T / 2, Fo = o.F / 2;
    for (int t = 0; t < To; ++t) for (int f = 0; f < Fo; ++f) for (int c = 0; c < o.C; ++c) {
        auto at = [&](int tt, int ff) { return in[(int64_t(tt + 1) * (o.F + 2) + ff + 1) * o.ldin + c]; };
        float v = (at(2 * t, 2 * f) + at(2 * t, 2 * f + 1) + at(2 * t + 1, 2 * f) + at(2 * t + 1, 2 * f + 1)) * 0.25f;
        out[(int64_t(t + 1) * (Fo + 2) + f + 1) * o.C + c] = v;
    }
}

void run_gru(const GruOp& o, const Bases& B) {
    const float* gi = P<float>(B, o.gi); const float* wt = P<float>(B, o.whh_t); const float* bh = P<float>(B, o.bhh);
    float* out = P<float>(B, o.out);
    const int H = o.H;
    for (int d = 0; d < 2; ++d) {
        std::vector<double> h(H, 0.0), gh(3 * H), hn(H);
        for (int s = 0; s < o.T; ++s) {
            int t = d == 0 ? s : o.T - 1 - s;
            for (int g = 0; g < 3 * H; ++g) {
                double a = bh[d * 3 * H + g];
                for (int k = 0; k < H; ++k) a += double(wt[(int64_t(d) * H + k) * 3 * H + g]) * h[k];
                gh[g] = a;
            }
            const float* x = gi + int64_t(t) * 6 * H + d * 3 * H;
            for (int j = 0; j < H; ++j) {
                double r = 1.0 / (1.0 + std::exp(-(x[j] + gh[j])));
                double z = 1.0 / (1.0 + std::exp(-(x[H + j] + gh[H + j])));
                double n = std::tanh(x[2 * H + j] + r * gh[2 * H + j]);
                hn[j] = (1.0 - z) * n + z * h[j];
            }
            h = hn;
            for (int j = 0; j < H; ++j) out[int64_t(t) * 2 * H + d * H + j] = float(h[j]);
        }
    }
}

void run_f0decode(const F0DecodeOp& o, const Bases& B) {
    const float* sal = P<float>(B, o.salience); float* f0 = P<float>(B, o.f0); int32_t* am = P<int32_t>(B, o.argmax);
    const RunParams* rp = P<RunParams>(B, o.params);
    for (int t = 0; t < o.T; ++t) {
        const float* s = sal + int64_t(t) * o.bins;
        int c = 0; float mx = s[0];
        for (int i = 1; i < o.bins; ++i) if (s[i] > mx) { mx = s[i]; c = i; }
        if (!(mx > 0.0f)) c = -4;  // padded-row argmax of an all-<=0 row is index 0 (never with sigmoid)
        am[t] = c;
        float ps = 0.f, ws = 0.f;
        for (int y = 0; y < 9; ++y) {
            int bi, ci;  // salience bin, cents bin (index into the 368-entry padded table)
            if (o.upstream_window) { bi = c - 4 + y; ci = c + y; } else { bi = c + 4 + y; ci = c + 4 + y; }
            if (bi < 0 || bi >= o.bins) continue;
            float cents = (float(ci) - 4.0f) * 20.0f + 1997.3794084376191f;
            ps += s[bi] * cents; ws += s[bi];
        }
        float cents = ws != 0.f ? ps / ws : 0.f;
        if (!(mx > o.threshold)) cents = 0.f;
        float f = 10.0f * exp2f(cents / 1200.0f);
        if (f == 10.0f) f = 0.f;
        f0[t] = f * rp->uppower;
    }
}

void run_f0post(const F0PostOp& o, const Bases& B) {
    const float* f0 = P<float>(B, o.f0); float* cache = P<float>(B, o.cache);
    int32_t* pitch = P<int32_t>(B, o.pitch); float* pitchf = P<float>(B, o.pitchf);
    const int n = o.cache_len;
    std::memmove(cache, cache + o.shift, sizeof(float) * (n - o.shift));
    int start = n + 4 - o.pitch_len;
    for (int i = 3; i < o.pitch_len - 1; ++i) cache[start + i - 3] = f0[i];
    int a = n - o.hubert_length + o.skip_head;
    for (int r = 0; r < o.return_length; ++r) {
        float f = cache[a + r];
        float m = std::log(f / 700.0f + 1.0f) * 1127.0f;
        if (!(m <= 0.f)) m = (m - o.mel_min) * 254.0f / (o.mel_max - o.mel_min) + 1.0f;
        m = std::min(std::max(m, 1.0f), 255.0f);
        pitch[r] = int32_t(std::floor(double(m) + 0.5));
        pitchf[r] = f;
    }
}

void run_embed(const EmbedOp& o, const Bases& B) {
    const float* ph = P<float>(B, o.phone); const int32_t* pi = P<int32_t>(B, o.pitch);
    const float* wp = P<float>(B, o.wp); const float* bp = P<float>(B, o.bp); const float* ep = P<float>(B, o.emb_pitch);
    float* out = P<float>(B, o.out);
    const float sc = std::sqrt(float(o.H));
    const float* pre = o.pre.null() ? nullptr : P<float>(B, o.pre);
    for (int r = 0; r < o.R; ++r) for (int h = 0; h < o.H; ++h) {
        double a = bp[h];
        if (pre) a = pre[int64_t(r) * o.H + h];
        else for (int k = 0; k < o.Cin; ++k) a += double(ph[int64_t(r) * o.Cin + k]) * wp[int64_t(h) * o.Cin + k];
        float v = (float(a) + ep[pi[r] * o.H + h]) * sc;
        out[int64_t(r) * o.ldo + h] = v > 0 ? v : 0.1f * v;
    }
}

void run_zp(const ZpOp& o, const Bases& B) {
    const float* st = P<float>(B, o.stats); float* out = P<float>(B, o.out); const RunParams* rp = P<RunParams>(B, o.params);
    uint64_t key = noise_key(rp->noise_seed, rp->window, NOISE_KIND_Z);
    for (int r = 0; r < o.R; ++r) for (int c = 0; c < o.H; ++c) {
        float m = st[int64_t(r) * 2 * o.H + c], lg = st[int64_t(r) * 2 * o.H + o.H + c];
        float nz = rp->noise_mode ? noise_gauss(key, uint64_t(r) * o.H + c) : 0.f;
        out[int64_t(r) * o.ldo + c] = m + std::exp(lg) * nz * 0.66666f;
    }
}

void run_sinegen(const SineGenOp& o, const Bases& B) {
    const float* f0 = P<float>(B, o.pitchf); float* out = P<float>(B, o.out); float* dbg = P<float>(B, o.sine_dbg);
    const RunParams* rp = P<RunParams>(B, o.params);
    const int T = o.R, L = o.R * o.upp;
    std::vector<float> rad(T), cum(T), tmp(L);
    double acc = 0;
    for (int t = 0; t < T; ++t) { rad[t] = std::fmod(f0[t] / o.sr, 1.0f); acc += rad[t]; cum[t] = float(acc) * float(o.upp); }
    const float scale = L > 1 ? float(T - 1) / float(L - 1) : 0.f;
    for (int i = 0; i < L; ++i) {
        float src = scale * float(i);
        int i0 = std::min(int(src), T - 1), i1 = i0 + (i0 < T - 1 ? 1 : 0);
        float l1 = std::min(std::max(src - float(i0), 0.f), 1.f), l0 = 1.f - l1;
        tmp[i] = std::fmod(l0 * cum[i0] + l1 * cum[i1], 1.0f);
    }
    uint64_t key = noise_key(rp->noise_seed, rp->window, NOISE_KIND_SINE);
    double ph = 0;
    for (int i = 0; i < L; ++i) {
        float sh = (i > 0 && (tmp[i] - tmp[i - 1]) < 0.f) ? -1.f : 0.f;
        ph += double(rad[i / o.upp] + sh);
        float s = std::sin(float(ph) * 2.0f * 3.14159265358979323846f) * 0.1f;
        float uv = f0[i / o.upp] > 0.f ? 1.f : 0.f;
        float amp = uv * 0.003f + (1.f - uv) * 0.1f / 3.f;
        float nz = rp->noise_mode ? noise_gauss(key, uint64_t(i)) : 0.f;
        float v = s * uv + amp * nz;
        if (dbg) dbg[i] = v;
        out[i] = std::tanh(v * o.lin_w + o.lin_b);
    }
}

void run_avg3(const Avg3Op& o, const Bases& B) {
    const float* a = P<float>(B, o.a); const float* b = P<float>(B, o.b); const float* c = P<float>(B, o.c);
    float* out = P<float>(B, o.out); float* raw = P<float>(B, o.raw);
    for (int t = 0; t < o.T; ++t) for (int ch = 0; ch < o.C; ++ch) {
        int64_t i = int64_t(t) * o.ld + ch;
        float s = (a[i] + b[i] + c[i]) / 3.0f;
        if (raw) raw[int64_t(t) * o.ldraw + ch] = s;
        out[int64_t(t) * o.ldo + ch] = s > 0 ? s : o.slope * s;
    }
}

void run_convpost(const ConvPostOp& o, const Bases& B) {
    const float* in = P<float>(B, o.in); const float* w = P<float>(B, o.w); float* out = P<float>(B, o.out);
    for (int t = 0; t < o.T; ++t) {
        double a = 0;
        for (int j = 0; j < o.k * o.C; ++j) a += double(in[int64_t(t) * o.C + j]) * w[j];
        out[t] = float(std::tanh(a));
    }
}

void run_knn_scan(const KnnScanOp& o, const Bases& B) {
    const float* idx = P<float>(B, o.index); const float* q = P<float>(B, o.queries);
    float* cd = P<float>(B, o.cand_d); int32_t* ci = P<int32_t>(B, o.cand_i);
    parallel_for(o.parts, [&](int p) {
        for (int qi = 0; qi < o.Q; ++qi) {
            std::vector<std::pair<float, int32_t>> best;
            for (int n = p; n < o.N; n += o.parts) {
                double a = 0;
                for (int c = 0; c < o.C; ++c) { double d = double(q[int64_t(qi) * o.ldq + c]) - idx[int64_t(n) * o.C + c]; a += d * d; }
                best.emplace_back(float(a), n);
            }
            std::sort(best.begin(), best.end());
            for (int j = 0; j < o.k; ++j) {
                int64_t e = (int64_t(qi) * o.parts + p) * o.k + j;
                if (j < int(best.size())) { cd[e] = best[j].first; ci[e] = best[j].second; } else { cd[e] = 3.4e38f; ci[e] = -1; }
            }
        }
    });
}

void run_knn_select(const KnnSelectOp& o, const Bases& B) {
    const float* cd = P<float>(B, o.cand_d); const int32_t* ci = P<int32_t>(B, o.cand_i);
    int32_t* idx = P<int32_t>(B, o.idx); float* d2 = P<float>(B, o.d2);
    const int M = o.parts * o.k;
    for (int q = 0; q < o.Q; ++q) {
        std::vector<std::pair<float, int32_t>> all;
        for (int m = 0; m < M; ++m) if (ci[int64_t(q) * M + m] >= 0) all.emplace_back(cd[int64_t(q) * M + m], ci[int64_t(q) * M + m]);
        std::sort(all.begin(), all.end());
        for (int i = 0; i < o.k; ++i) { idx[q * o.k + i] = all[i].second; d2[q * o.k + i] = all[i].first; }
    }
}

void run_knn_blend(const KnnBlendOp& o, const Bases& B) {
    const float* index = P<float>(B, o.index); const int32_t* idx = P<int32_t>(B, o.idx); const float* d2 = P<float>(B, o.d2);
    const float* x = P<float>(B, o.x); float* out = P<float>(B, o.out); const RunParams* rp = P<RunParams>(B, o.params);
    for (int q = 0; q < o.Q; ++q) {
        std::vector<double> w(o.k); double ws = 0;
        for (int i = 0; i < o.k; ++i) { double r = 1.0 / double(d2[q * o.k + i]); w[i] = r * r; ws += w[i]; }
        for (int c = 0; c < o.C; ++c) {
            double a = 0;
            for (int i = 0; i < o.k; ++i) a += w[i] / ws * index[int64_t(idx[q * o.k + i]) * o.C + c];
            out[int64_t(q) * o.C + c] = float(double(rp->index_rate) * a + (1.0 - double(rp->index_rate)) * x[int64_t(q) * o.ldx + c]);
        }
    }
}

void run_gather(const GatherRowsOp& o, const Bases& B) {
    const float* src = P<float>(B, o.src); float* out = P<float>(B, o.out);
    for (int r = 0; r < o.R; ++r) {
        int s = std::min((o.skip + r) / 2, o.T - 1) - o.row0;
        std::memcpy(out + int64_t(r) * o.C, src + int64_t(s) * o.lds, sizeof(float) * o.C);
    }
}

void run_op(const Op& op, const Bases& B) {
    switch (op.kind) {
        case OP_GEMM: run_gemm(op.gemm, B); break;
        case OP_LAYERNORM: run_layernorm(op.ln, B); break;
        case OP_ATTN: run_attn(op.attn, B); break;
        case OP_RELATTN: run_relattn(op.relattn, B); break;
        case OP_CONV0_STATS: run_conv0_stats(op.c0s, B); break;
        case OP_CONV0_APPLY: run_conv0_apply(op.c0a, B); break;
        case OP_STFTMEL: run_stftmel(op.stft, B); break;
        case OP_AVGPOOL: run_avgpool(op.pool, B); break;
        case OP_GRU: run_gru(op.gru, B); break;
        case OP_F0DECODE: run_f0decode(op.f0d, B); break;
        case OP_F0POST: run_f0post(op.f0p, B); break;
        case OP_EMBED: run_embed(op.embed, B); break;
        case OP_ZP: run_zp(op.zp, B); break;
        case OP_SINEGEN: run_sinegen(op.sine, B); break;
        case OP_AVG3: run_avg3(op.avg3, B); break;
        case OP_CONVPOST: run_convpost(op.cpost, B); break;
        case OP_KNN_SCAN: run_knn_scan(op.kd, B); break;
        case OP_KNN_SELECT: run_knn_select(op.ks, B); break;
        case OP_KNN_BLEND: run_knn_blend(op.kb, B); break;
        case OP_GATHER_ROWS: run_gather(op.gather, B); break;
        case OP_FILL: std::memset(B.b[op.fill.dst.space] + op.fill.dst.off, 0, op.fill.bytes); break;
        case OP_WAIT: break;
    }
}

struct Exec {
    std::string data_dir, err;
    Packed cv, f0, syn; CvInfo cvi; F0Info f0i; SynInfo syi;
    bool has_cv = false, has_f0 = false, has_syn = false;
    std::vector<float> index; int index_rows = 0, index_c = 0;
    std::vector<uint8_t> state, work;
    Plan plan;
    RunParams rp{};
    int index_k = 8;
    bool allow_umma = true;
    int chain_main = 0, chain_side = 0, chain_max_m = 0;   // persistent chains off unless pe_set_chain asks
    Exec() : state(StateLayout::bytes, 0) { rp.uppower = 1.f; rp.noise_mode = 1; }
    Bases bases() {
        Bases B;
        B.b[SP_CV] = reinterpret_cast<uint8_t*>(cv.host.data());
        B.b[SP_F0] = reinterpret_cast<uint8_t*>(f0.host.data());
        B.b[SP_SYN] = reinterpret_cast<uint8_t*>(syn.host.data());
        B.b[SP_IDX] = reinterpret_cast<uint8_t*>(index.data());
        B.b[SP_WORK] = work.data();
        B.b[SP_STATE] = state.data();
        return B;
    }
};

}  // namespace

extern "C" {

void* pe_create(const char* data_dir) { Exec* e = new Exec(); e->data_dir = data_dir; return e; }
void pe_destroy(void* h) { delete static_cast<Exec*>(h); }
const char* pe_error(void* h) { return static_cast<Exec*>(h)->err.c_str(); }

int pe_load(void* h, int which, const char* path) {
    Exec* e = static_cast<Exec*>(h);
    RvcwFile f;
    if (!f.load(path, e->err)) return 1;
    if (which == 0) { e->cv = Packed{}; e->has_cv = pack_contentvec(f, e->cv, e->cvi, e->err); return e->has_cv ? 0 : 1; }
    if (which == 1) { e->f0 = Packed{}; e->has_f0 = pack_rmvpe(f, e->f0, e->f0i, e->err); return e->has_f0 ? 0 : 1; }
    if (which == 2) { e->syn = Packed{}; e->has_syn = pack_synth(f, e->syn, e->syi, e->err); return e->has_syn ? 0 : 1; }
    return 1;
}

int pe_set_index(void* h, const float* rows, int n, int c, float rate) {
    Exec* e = static_cast<Exec*>(h);
    e->index.assign(rows, rows + int64_t(n) * c); e->index_rows = n; e->index_c = c; e->rp.index_rate = rate;
    return 0;
}

void pe_set_params(void* h, uint64_t seed, int noise_mode, int index_k) {
    Exec* e = static_cast<Exec*>(h);
    e->rp.noise_seed = seed; e->rp.noise_mode = noise_mode; e->index_k = index_k;
}

// kind: PlanKind; runs the whole plan on the CPU.  The pitch cache / window counter persist.
int pe_run(void* h, int kind, const float* pcm, int n, int sf16k, int pitch_shift, int skip_head, int return_length) {
    Exec* e = static_cast<Exec*>(h);
    Geometry g{n, sf16k, skip_head, return_length};
    PlanOptions opt; opt.index_k = e->index_k; opt.allow_umma = e->allow_umma; opt.with_index = e->index_rows > 0; opt.index_rows = e->index_rows; opt.index_cols = e->index_c;
    opt.chain_grid_main = e->chain_main; opt.chain_grid_side = e->chain_side; opt.chain_side_max_m = e->chain_max_m;
    if (!build_plan(PlanKind(kind), g, opt, e->has_cv ? &e->cv : nullptr, &e->cvi, e->has_f0 ? &e->f0 : nullptr, &e->f0i,
                    e->has_syn ? &e->syn : nullptr, &e->syi, e->plan, e->err)) return 1;
    e->work.assign(size_t(e->plan.work_bytes), 0);
    e->rp.uppower = std::pow(2.0f, float(pitch_shift / 12));
    std::memcpy(e->state.data() + StateLayout::off_params, &e->rp, sizeof(RunParams));
    std::memcpy(e->state.data() + StateLayout::off_pcm, pcm, sizeof(float) * n);
    Bases B = e->bases();
    // Ops outside chains run in plan order.  A chain (chain.h) runs phase by phase, and INSIDE a phase in REVERSE plan
    // order: the plan builder claims the ops of a phase are independent, so any order must give the same bits - a wrong
    // dependency analysis shows up as a different result (tests/test_host_cpu.py).
    const size_t n_ops = e->plan.ops.size();
    for (size_t i = 0; i < n_ops;) {
        const Op& op = e->plan.ops[i];
        if (op.chain < 0) { run_op(op, B); ++i; continue; }
        const ChainInfo& ci = e->plan.chains[size_t(op.chain)];
        for (int ph = 0; ph < ci.n_phases; ++ph)
            for (int k = ci.count - 1; k >= 0; --k)
                if (ci.phase[size_t(k)] == ph) run_op(e->plan.ops[size_t(ci.first + k)], B);
        i = size_t(ci.first + ci.count);
    }
    if (kind == PLAN_INFER) e->rp.window++;
    return 0;
}

// chains of the last plan: CTA budgets for lane 0 / side lanes and the largest M of a side-lane GEMM that may join
void pe_set_chain(void* h, int grid_main, int grid_side, int side_max_m) {
    Exec* e = static_cast<Exec*>(h);
    e->chain_main = grid_main; e->chain_side = grid_side; e->chain_max_m = side_max_m;
}
int pe_chain_stats(void* h, int* out /* n_chains, n_phases, n_chain_ops, max ops in one phase */) {
    Exec* e = static_cast<Exec*>(h);
    int phases = 0, ops = 0, widest = 0;
    for (const ChainInfo& ci : e->plan.chains) {
        phases += ci.n_phases; ops += ci.count;
        for (int ph = 0; ph < ci.n_phases; ++ph) {
            int w = 0;
            for (int k = 0; k < ci.count; ++k) w += ci.phase[size_t(k)] == ph;
            widest = std::max(widest, w);
        }
    }
    out[0] = int(e->plan.chains.size()); out[1] = phases; out[2] = ops; out[3] = widest;
    return 0;
}

int pe_num_ops(void* h) { return int(static_cast<Exec*>(h)->plan.ops.size()); }
int pe_num_bufs(void* h) { return int(static_cast<Exec*>(h)->plan.bufs.size()); }
const char* pe_buf_name(void* h, int i) { return static_cast<Exec*>(h)->plan.bufs[i].name.c_str(); }

// copies a named buffer (float or int32 raw bits); returns element count or -1
long pe_get(void* h, const char* name, void* out, long cap_elems) {
    Exec* e = static_cast<Exec*>(h);
    Bases B = e->bases();
    std::string nm(name);
    if (nm == "audio") { long n = e->plan.audio_len; if (n > cap_elems) return -1; std::memcpy(out, B.b[SP_STATE] + StateLayout::off_audio, n * 4); return n; }
    if (nm == "cache") { if (1024 > cap_elems) return -1; std::memcpy(out, B.b[SP_STATE] + StateLayout::off_cache, 4096); return 1024; }
    const NamedBuf* nb = e->plan.find(nm);
    if (!nb) return -1;
    if (nb->elems > cap_elems) return -1;
    std::memcpy(out, B.b[nb->ref.space] + nb->ref.off, nb->elems * 4);
    return long(nb->elems);
}

int pe_plan_dims(void* h, int* out /* hubert_T, hubert_C, f0_T, audio_len, knn_q, n_ops, n_lanes */) {
    Exec* e = static_cast<Exec*>(h);
    out[0] = e->plan.hubert_T; out[1] = e->plan.hubert_C; out[2] = e->plan.f0_T; out[3] = e->plan.audio_len;
    out[4] = e->plan.knn_q; out[5] = int(e->plan.ops.size()); out[6] = e->plan.n_lanes;
    return 0;
}

}  // extern "C"
