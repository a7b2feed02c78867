// model.cpp - RVCW reader and weight packing (host-only).  See model.h.
#include "model.h"

#include <cmath>
#include <cstdio>
#include <cstring>

namespace rvc {

// ------------------------------------------------------------------------------------------
// RVCW container (layout documented in oracle/weights.py and DESIGN.md)
// ------------------------------------------------------------------------------------------

bool RvcwFile::load(const std::string& path, std::string& err) {
    FILE* fp = std::fopen(path.c_str(), "rb");
    if (!fp) { err = "cannot open " + path; return false; }
    std::fseek(fp, 0, SEEK_END);
    long sz = std::ftell(fp);
    std::fseek(fp, 0, SEEK_SET);
    if (sz < 32) { std::fclose(fp); err = "truncated file " + path; return false; }
    blob.resize(static_cast<size_t>(sz));
    size_t got = std::fread(blob.data(), 1, blob.size(), fp);
    std::fclose(fp);
    if (got != blob.size()) { err = "short read " + path; return false; }
    if (std::memcmp(blob.data(), "RVCW0001", 8) != 0) { err = "bad magic in " + path; return false; }
    uint32_t n, tb; uint64_t doff, dbytes;
    std::memcpy(&n, &blob[8], 4); std::memcpy(&tb, &blob[12], 4);
    std::memcpy(&doff, &blob[16], 8); std::memcpy(&dbytes, &blob[24], 8);
    if (32ull + tb > blob.size() || doff + dbytes > blob.size()) { err = "corrupt header in " + path; return false; }
    size_t p = 32, end = 32 + tb;
    for (uint32_t i = 0; i < n; ++i) {
        if (p + 2 > end) { err = "corrupt table in " + path; return false; }
        uint16_t nl; std::memcpy(&nl, &blob[p], 2); p += 2;
        if (p + nl + 2 > end) { err = "corrupt table in " + path; return false; }
        std::string name(reinterpret_cast<const char*>(&blob[p]), nl); p += nl;
        uint8_t dt = blob[p], nd = blob[p + 1]; p += 2;
        if (p + 4ull * nd + 16 > end) { err = "corrupt table in " + path; return false; }
        HostTensor ht; ht.dtype = dt;
        for (int d = 0; d < nd; ++d) { uint32_t v; std::memcpy(&v, &blob[p], 4); p += 4; ht.shape.push_back(v); }
        uint64_t off, nb; std::memcpy(&off, &blob[p], 8); std::memcpy(&nb, &blob[p + 8], 8); p += 16;
        if (doff + off + nb > blob.size() || nb != static_cast<uint64_t>(ht.numel()) * 4) {
            err = "tensor out of range: " + name; return false;
        }
        ht.data = blob.data() + doff + off;
        t[name] = ht;
    }
    return true;
}

const HostTensor* RvcwFile::find(const std::string& name) const {
    auto it = t.find(name);
    return it == t.end() ? nullptr : &it->second;
}

int64_t Packed::add(const std::string& name, int64_t elems) {
    int64_t o = (static_cast<int64_t>(host.size()) + 63) & ~int64_t(63);
    host.resize(static_cast<size_t>(o + elems), 0.0f);
    off[name] = o;
    return o;
}

int64_t Packed::at(const std::string& name) const {
    auto it = off.find(name);
    return it == off.end() ? -1 : it->second;
}

namespace {

struct Loader {
    const RvcwFile& f; std::string& err; bool ok = true;
    const float* get(const std::string& name, std::initializer_list<int64_t> shape) {
        const HostTensor* t = f.find(name);
        if (!t) { fail("missing tensor " + name); return nullptr; }
        if (t->dtype != 0 || t->shape.size() != shape.size()) { fail("bad rank/dtype for " + name); return nullptr; }
        size_t i = 0;
        for (auto d : shape) { if (t->shape[i++] != d) { fail("bad shape for " + name); return nullptr; } }
        return t->f();
    }
    int32_t meta(const std::string& name, int32_t dflt) {
        const HostTensor* t = f.find(name);
        return (t && t->dtype == 1 && t->numel() >= 1) ? t->i()[0] : dflt;
    }
    void fail(const std::string& m) { if (ok) err = m; ok = false; }
};

void copy_vec(Packed& out, const std::string& name, const float* src, int64_t n) {
    int64_t o = out.add(name, n);
    if (src) std::memcpy(out.p(o), src, sizeof(float) * n);
}

// conv1d weight [cout, cin, k] -> [cout][k*cin] tap-major / channel-minor (channels-last A rows)
void pack_conv1d(Packed& out, const std::string& name, const float* w, int cout, int cin, int k) {
    int64_t o = out.add(name, int64_t(cout) * k * cin);
    if (!w) return;
    float* d = out.p(o);
    for (int n = 0; n < cout; ++n)
        for (int c = 0; c < cin; ++c)
            for (int j = 0; j < k; ++j)
                d[(int64_t(n) * k + j) * cin + c] = w[(int64_t(n) * cin + c) * k + j];
}

}  // namespace

// ------------------------------------------------------------------------------------------
// ContentVec / HuBERT-base
// ------------------------------------------------------------------------------------------

bool pack_contentvec(const RvcwFile& f, Packed& out, CvInfo& info, std::string& err) {
    Loader L{f, err};
    // layer count: the meta tensor when the file has one, otherwise the encoder layers actually present (a converted
    // ONNX / state_dict checkpoint carries no meta tensors; the engine checks the count against the requested version)
    int present = 0;
    while (f.find("encoder.layers." + std::to_string(present) + ".self_attn.q_proj.weight") != nullptr) ++present;
    info.n_layers = L.meta("meta.n_layers", present > 0 ? present : 12);
    info.final_proj = f.find("final_proj.weight") != nullptr;
    info.out_dim = info.final_proj ? 256 : 768;
    static const int K[7] = {10, 3, 3, 3, 3, 2, 2};
    copy_vec(out, "conv0.w", L.get("feature_extractor.conv_layers.0.0.weight", {512, 1, 10}), 5120);
    copy_vec(out, "gn.g", L.get("feature_extractor.conv_layers.0.2.weight", {512}), 512);
    copy_vec(out, "gn.b", L.get("feature_extractor.conv_layers.0.2.bias", {512}), 512);
    for (int i = 1; i < 7; ++i) {
        std::string n = "feature_extractor.conv_layers." + std::to_string(i) + ".0.weight";
        pack_conv1d(out, "conv" + std::to_string(i) + ".w", L.get(n, {512, 512, K[i]}), 512, 512, K[i]);
    }
    copy_vec(out, "ln.g", L.get("layer_norm.weight", {512}), 512);
    copy_vec(out, "ln.b", L.get("layer_norm.bias", {512}), 512);
    copy_vec(out, "proj.w", L.get("post_extract_proj.weight", {768, 512}), 768 * 512);
    copy_vec(out, "proj.b", L.get("post_extract_proj.bias", {768}), 768);
    {   // grouped pos-conv [768, 48, 128] -> [16 groups][48 out][128 taps * 48 cin]
        const float* w = L.get("encoder.pos_conv.0.weight", {768, 48, 128});
        int64_t o = out.add("pos.w", int64_t(768) * 128 * 48);
        if (w) {
            float* d = out.p(o);
            for (int n = 0; n < 768; ++n)
                for (int c = 0; c < 48; ++c)
                    for (int j = 0; j < 128; ++j)
                        d[(int64_t(n) * 128 + j) * 48 + c] = w[(int64_t(n) * 48 + c) * 128 + j];
        }
        copy_vec(out, "pos.b", L.get("encoder.pos_conv.0.bias", {768}), 768);
    }
    copy_vec(out, "eln.g", L.get("encoder.layer_norm.weight", {768}), 768);
    copy_vec(out, "eln.b", L.get("encoder.layer_norm.bias", {768}), 768);
    for (int i = 0; i < info.n_layers && L.ok; ++i) {
        std::string s = "encoder.layers." + std::to_string(i) + ".", d = "L" + std::to_string(i) + ".";
        int64_t ow = out.add(d + "qkv.w", int64_t(2304) * 768), ob = out.add(d + "qkv.b", 2304);
        const char* nm[3] = {"q_proj", "k_proj", "v_proj"};
        for (int j = 0; j < 3; ++j) {
            const float* w = L.get(s + "self_attn." + nm[j] + ".weight", {768, 768});
            const float* b = L.get(s + "self_attn." + nm[j] + ".bias", {768});
            if (!w || !b) break;
            float sc = j == 0 ? 0.125f : 1.0f;  // fairseq: q = q_proj(x) * head_dim^-0.5 (exact)
            for (int64_t e = 0; e < 768 * 768; ++e) out.p(ow)[int64_t(j) * 768 * 768 + e] = w[e] * sc;
            for (int e = 0; e < 768; ++e) out.p(ob)[j * 768 + e] = b[e] * sc;
        }
        copy_vec(out, d + "o.w", L.get(s + "self_attn.out_proj.weight", {768, 768}), 768 * 768);
        copy_vec(out, d + "o.b", L.get(s + "self_attn.out_proj.bias", {768}), 768);
        copy_vec(out, d + "ln1.g", L.get(s + "self_attn_layer_norm.weight", {768}), 768);
        copy_vec(out, d + "ln1.b", L.get(s + "self_attn_layer_norm.bias", {768}), 768);
        copy_vec(out, d + "fc1.w", L.get(s + "fc1.weight", {3072, 768}), 3072 * 768);
        copy_vec(out, d + "fc1.b", L.get(s + "fc1.bias", {3072}), 3072);
        copy_vec(out, d + "fc2.w", L.get(s + "fc2.weight", {768, 3072}), 768 * 3072);
        copy_vec(out, d + "fc2.b", L.get(s + "fc2.bias", {768}), 768);
        copy_vec(out, d + "ln2.g", L.get(s + "final_layer_norm.weight", {768}), 768);
        copy_vec(out, d + "ln2.b", L.get(s + "final_layer_norm.bias", {768}), 768);
    }
    if (info.final_proj) {
        copy_vec(out, "fp.w", L.get("final_proj.weight", {256, 768}), 256 * 768);
        copy_vec(out, "fp.b", L.get("final_proj.bias", {256}), 256);
    }
    return L.ok;
}

// ------------------------------------------------------------------------------------------
// RMVPE
// ------------------------------------------------------------------------------------------

namespace {

struct Bn { const float *g, *b, *m, *v; };

Bn get_bn(Loader& L, const std::string& p, int c) {
    return Bn{L.get(p + ".weight", {c}), L.get(p + ".bias", {c}), L.get(p + ".running_mean", {c}),
              L.get(p + ".running_var", {c})};
}

// conv2d 3x3 [cout,cin,3,3] (+ optional eval-mode BN fold) -> W[cout][(dt*3+df)*cin + c], bias[cout]
void pack_conv3x3(Loader& L, Packed& out, const std::string& dst, const std::string& wname,
                  const std::string& bnname, const std::string& biasname, int cout, int cin) {
    const float* w = L.get(wname, {cout, cin, 3, 3});
    int64_t ow = out.add(dst + ".w", int64_t(cout) * 9 * cin), ob = out.add(dst + ".b", cout);
    if (!w) return;
    std::vector<double> s(cout, 1.0), sh(cout, 0.0);
    if (!bnname.empty()) {
        Bn bn = get_bn(L, bnname, cout);
        if (!L.ok) return;
        for (int n = 0; n < cout; ++n) {
            s[n] = double(bn.g[n]) / std::sqrt(double(bn.v[n]) + 1e-5);
            sh[n] = double(bn.b[n]) - double(bn.m[n]) * s[n];
        }
    }
    if (!biasname.empty()) {
        const float* b = L.get(biasname, {cout});
        if (!b) return;
        for (int n = 0; n < cout; ++n) sh[n] += double(b[n]) * s[n];
    }
    float* d = out.p(ow);
    for (int n = 0; n < cout; ++n) {
        for (int c = 0; c < cin; ++c)
            for (int t = 0; t < 9; ++t)
                d[(int64_t(n) * 9 + t) * cin + c] = float(double(w[(int64_t(n) * cin + c) * 9 + t]) * s[n]);
        out.p(ob)[n] = float(sh[n]);
    }
}

void pack_convblockres(Loader& L, Packed& out, const std::string& src, const std::string& dst, int cin,
                       int cout) {
    pack_conv3x3(L, out, dst + "c1", src + "conv.0.weight", src + "conv.1", "", cout, cin);
    pack_conv3x3(L, out, dst + "c2", src + "conv.3.weight", src + "conv.4", "", cout, cout);
    if (cin != cout) {
        copy_vec(out, dst + "sc.w", L.get(src + "shortcut.weight", {cout, cin, 1, 1}), int64_t(cout) * cin);
        copy_vec(out, dst + "sc.b", L.get(src + "shortcut.bias", {cout}), cout);
    }
}

// vendor/mel-spec/mel_spec/src/mel.rs:149-237 restated (f64), htk scale + Slaney norm as called
// from rvc/src/f0/rmvpe.rs:146-148,220: mel(16000, 1024, 128, 30, 8000, htk=true, norm=true).
double hz_to_mel_htk(double f) { return 2595.0 * std::log10(1.0 + f / 700.0); }
double mel_to_hz_htk(double m) { return 700.0 * (std::pow(10.0, m / 2595.0) - 1.0); }

void mel_filterbank_htk(double sr, int n_fft, int n_mels, double fmin, double fmax,
                        std::vector<float>& weights /* [n_mels][n_fft/2+1] */) {
    int nb = n_fft / 2 + 1;
    std::vector<double> fft(nb), melf(n_mels + 2);
    double step = sr / double(n_fft);
    for (int i = 0; i < nb; ++i) fft[i] = step * double(i);
    double lo = hz_to_mel_htk(fmin), hi = hz_to_mel_htk(fmax);
    double mstep = (hi - lo) / double(n_mels + 1);  // ndarray linspace: start + i*step
    for (int i = 0; i < n_mels + 2; ++i) melf[i] = mel_to_hz_htk(i == n_mels + 1 ? hi : lo + mstep * double(i));
    weights.assign(size_t(n_mels) * nb, 0.0f);
    for (int i = 0; i < n_mels; ++i) {
        double fd0 = melf[i + 1] - melf[i], fd1 = melf[i + 2] - melf[i + 1];
        double enorm = 2.0 / (melf[i + 2] - melf[i]);
        for (int j = 0; j < nb; ++j) {
            double lower = -(melf[i] - fft[j]) / fd0;
            double upper = (melf[i + 2] - fft[j]) / fd1;
            lower = std::fmin(std::fmax(lower, 0.0), 1.0);
            upper = std::fmin(std::fmax(upper, 0.0), 1.0);
            weights[size_t(i) * nb + j] = float(std::fmin(lower, upper) * enorm);
        }
    }
}

}  // namespace

bool pack_rmvpe(const RvcwFile& f, Packed& out, F0Info& info, std::string& err) {
    Loader L{f, err};
    {   // DSP constants: periodic Hann (rmvpe.rs:33-37) and the sparse mel filterbank
        int64_t ow = out.add("window", 1024);
        for (int i = 0; i < 1024; ++i) {
            float c = float(std::cos(2.0 * 3.14159265358979323846 * double(i) / 1024.0));
            out.p(ow)[i] = 0.5f * (1.0f - c);
        }
        std::vector<float> mb;
        mel_filterbank_htk(16000.0, 1024, 128, 30.0, 8000.0, mb);
        std::vector<int32_t> start(128), count(128), boff(128);
        std::vector<float> w;
        for (int i = 0; i < 128; ++i) {
            int a = -1, b = -1;
            for (int j = 0; j < 513; ++j) if (mb[size_t(i) * 513 + j] != 0.0f) { if (a < 0) a = j; b = j; }
            start[i] = a < 0 ? 0 : a; count[i] = a < 0 ? 0 : b - a + 1; boff[i] = int32_t(w.size());
            for (int j = 0; j < count[i]; ++j) w.push_back(mb[size_t(i) * 513 + start[i] + j]);
        }
        info.mel_nnz = int32_t(w.size());
        int64_t o;
        o = out.add("mel.start", 128); std::memcpy(out.p(o), start.data(), 512);
        o = out.add("mel.count", 128); std::memcpy(out.p(o), count.data(), 512);
        o = out.add("mel.off", 128); std::memcpy(out.p(o), boff.data(), 512);
        o = out.add("mel.w", int64_t(w.size())); std::memcpy(out.p(o), w.data(), w.size() * 4);
        o = out.add("mel.dense", int64_t(128) * 513); std::memcpy(out.p(o), mb.data(), mb.size() * 4);
    }
    {
        Bn bn = get_bn(L, "unet.encoder.bn", 1);
        if (L.ok) {
            double s = double(bn.g[0]) / std::sqrt(double(bn.v[0]) + 1e-5);
            info.in_scale = float(s); info.in_shift = float(double(bn.b[0]) - double(bn.m[0]) * s);
        }
    }
    int cin = 1, cout = 16;
    for (int i = 0; i < 5 && L.ok; ++i) {
        for (int j = 0; j < 4; ++j)
            pack_convblockres(L, out, "unet.encoder.layers." + std::to_string(i) + ".conv." + std::to_string(j) + ".",
                              "enc" + std::to_string(i) + "." + std::to_string(j) + ".", j == 0 ? cin : cout, cout);
        cin = cout; cout *= 2;
    }
    for (int i = 0; i < 4 && L.ok; ++i)
        for (int j = 0; j < 4; ++j)
            pack_convblockres(L, out, "unet.intermediate.layers." + std::to_string(i) + ".conv." + std::to_string(j) + ".",
                              "mid" + std::to_string(i) + "." + std::to_string(j) + ".", (i == 0 && j == 0) ? 256 : 512, 512);
    cin = 512;
    for (int i = 0; i < 5 && L.ok; ++i) {
        cout = cin / 2;
        std::string s = "unet.decoder.layers." + std::to_string(i) + ".", d = "dec" + std::to_string(i) + ".";
        const float* w = L.get(s + "conv1.0.weight", {cin, cout, 3, 3});
        Bn bn = get_bn(L, s + "conv1.1", cout);
        int64_t ow = out.add(d + "up.w", int64_t(4) * cout * 4 * cin), ob = out.add(d + "up.b", 4 * cout);
        if (w && L.ok) {
            // ConvTranspose2d(k3,s2,p1,op1): out[2q+r] gets taps {r=0: (a=0,k=1)}, {r=1: (a=0,k=2),(a=1,k=0)}
            auto tap = [](int r, int a) { return r == 0 ? (a == 0 ? 1 : -1) : (a == 0 ? 2 : 0); };
            for (int rt = 0; rt < 2; ++rt) for (int rf = 0; rf < 2; ++rf) for (int co = 0; co < cout; ++co) {
                double s1 = double(bn.g[co]) / std::sqrt(double(bn.v[co]) + 1e-5);
                int64_t n = (int64_t(rt) * 2 + rf) * cout + co;
                out.p(ob)[n] = float(double(bn.b[co]) - double(bn.m[co]) * s1);
                for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) {
                    int kt = tap(rt, a), kf = tap(rf, b);
                    if (kt < 0 || kf < 0) continue;
                    for (int ci = 0; ci < cin; ++ci)
                        out.p(ow)[n * (4 * cin) + (int64_t(a) * 2 + b) * cin + ci] =
                            float(double(w[((int64_t(ci) * cout + co) * 3 + kt) * 3 + kf]) * s1);
                }
            }
        }
        for (int j = 0; j < 4; ++j)
            pack_convblockres(L, out, s + "conv2." + std::to_string(j) + ".", d + std::to_string(j) + ".",
                              j == 0 ? cout * 2 : cout, cout);
        cin = cout;
    }
    pack_conv3x3(L, out, "cnn", "cnn.weight", "", "cnn.bias", 3, 16);
    {   // BiGRU(384 -> 256): input GEMM weights re-indexed onto the padded cnn rows [(F+2)*3]
        const int F = 128, H = 256, KI = (F + 2) * 4;  // cnn map pixels are padded to 4 channels (plan.cpp)
        int64_t ow = out.add("gru.wih", int64_t(2) * 3 * H * KI), ob = out.add("gru.bih", 2 * 3 * H);
        int64_t oh = out.add("gru.whh_t", int64_t(2) * H * 3 * H), obh = out.add("gru.bhh", 2 * 3 * H);
        const char* sfx[2] = {"", "_reverse"};
        for (int d = 0; d < 2 && L.ok; ++d) {
            const float* wi = L.get(std::string("fc.0.gru.weight_ih_l0") + sfx[d], {3 * H, 3 * F});
            const float* wh = L.get(std::string("fc.0.gru.weight_hh_l0") + sfx[d], {3 * H, H});
            const float* bi = L.get(std::string("fc.0.gru.bias_ih_l0") + sfx[d], {3 * H});
            const float* bh = L.get(std::string("fc.0.gru.bias_hh_l0") + sfx[d], {3 * H});
            if (!wi || !wh || !bi || !bh) break;
            for (int g = 0; g < 3 * H; ++g) {
                for (int c = 0; c < 3; ++c) for (int fq = 0; fq < F; ++fq)
                    out.p(ow)[(int64_t(d) * 3 * H + g) * KI + (fq + 1) * 4 + c] = wi[int64_t(g) * 3 * F + c * F + fq];
                out.p(ob)[d * 3 * H + g] = bi[g];
                out.p(obh)[d * 3 * H + g] = bh[g];
                for (int k = 0; k < H; ++k) out.p(oh)[(int64_t(d) * H + k) * 3 * H + g] = wh[int64_t(g) * H + k];
            }
        }
    }
    copy_vec(out, "fc.w", L.get("fc.1.weight", {360, 512}), 360 * 512);
    copy_vec(out, "fc.b", L.get("fc.1.bias", {360}), 360);
    return L.ok;
}

// ------------------------------------------------------------------------------------------
// SynthesizerTrnMs768NSFsid (32k / 40k / 48k generator configs)
// ------------------------------------------------------------------------------------------

bool pack_synth(const RvcwFile& f, Packed& out, SynInfo& info, std::string& err) {
    Loader L{f, err};
    const int H = 192;
    info.sr = L.meta("meta.sr", 40000);
    info.phone_dim = L.meta("meta.phone_dim", 768);
    int sid = L.meta("meta.sid", 0);
    if (!syn_config_for_rate(info)) { err = "synthesizer output rate " + std::to_string(info.sr) + " is not one of the generator configs (32k / 40k / 48k)"; return false; }
    const int PD = info.phone_dim;
    copy_vec(out, "emb.wp", L.get("enc_p.emb_phone.weight", {H, PD}), int64_t(H) * PD);
    copy_vec(out, "emb.bp", L.get("enc_p.emb_phone.bias", {H}), H);
    copy_vec(out, "emb.pitch", L.get("enc_p.emb_pitch.weight", {256, H}), 256 * H);
    for (int i = 0; i < 6 && L.ok; ++i) {
        std::string s = "enc_p.encoder.attn_layers." + std::to_string(i) + ".", d = "E" + std::to_string(i) + ".";
        int64_t ow = out.add(d + "qkv.w", int64_t(3) * H * H), ob = out.add(d + "qkv.b", 3 * H);
        const char* nm[3] = {"conv_q", "conv_k", "conv_v"};
        for (int j = 0; j < 3; ++j) {
            const float* w = L.get(s + nm[j] + ".weight", {H, H, 1});
            const float* b = L.get(s + nm[j] + ".bias", {H});
            if (!w || !b) break;
            double sc = j == 0 ? 1.0 / std::sqrt(96.0) : 1.0;  // query / sqrt(k_channels)
            for (int e = 0; e < H * H; ++e) out.p(ow)[j * H * H + e] = float(double(w[e]) * sc);
            for (int e = 0; e < H; ++e) out.p(ob)[j * H + e] = float(double(b[e]) * sc);
        }
        copy_vec(out, d + "relk", L.get(s + "emb_rel_k", {1, 21, 96}), 21 * 96);
        copy_vec(out, d + "relv", L.get(s + "emb_rel_v", {1, 21, 96}), 21 * 96);
        copy_vec(out, d + "o.w", L.get(s + "conv_o.weight", {H, H, 1}), H * H);
        copy_vec(out, d + "o.b", L.get(s + "conv_o.bias", {H}), H);
        std::string n1 = "enc_p.encoder.norm_layers_1." + std::to_string(i), n2 = "enc_p.encoder.norm_layers_2." + std::to_string(i);
        copy_vec(out, d + "ln1.g", L.get(n1 + ".gamma", {H}), H);
        copy_vec(out, d + "ln1.b", L.get(n1 + ".beta", {H}), H);
        copy_vec(out, d + "ln2.g", L.get(n2 + ".gamma", {H}), H);
        copy_vec(out, d + "ln2.b", L.get(n2 + ".beta", {H}), H);
        std::string fs = "enc_p.encoder.ffn_layers." + std::to_string(i) + ".";
        pack_conv1d(out, d + "ffn1.w", L.get(fs + "conv_1.weight", {768, H, 3}), 768, H, 3);
        copy_vec(out, d + "ffn1.b", L.get(fs + "conv_1.bias", {768}), 768);
        pack_conv1d(out, d + "ffn2.w", L.get(fs + "conv_2.weight", {H, 768, 3}), H, 768, 3);
        copy_vec(out, d + "ffn2.b", L.get(fs + "conv_2.bias", {H}), H);
    }
    copy_vec(out, "proj.w", L.get("enc_p.proj.weight", {2 * H, H, 1}), 2 * H * H);
    copy_vec(out, "proj.b", L.get("enc_p.proj.bias", {2 * H}), 2 * H);

    const float* embg = L.get("emb_g.weight", {109, 256});
    if (!L.ok) return false;
    const float* g = embg + int64_t(sid) * 256;

    // Flow (reverse): flips are folded into the weights - flows 3 and 1 see a channel-reversed z.
    for (int fl = 0; fl < 4 && L.ok; ++fl) {
        bool flipped = (fl == 3 || fl == 1);
        std::string s = "flow.flows." + std::to_string(2 * fl) + ".", d = "F" + std::to_string(fl) + ".";
        const float* wpre = L.get(s + "pre.weight", {H, 96, 1});
        const float* bpre = L.get(s + "pre.bias", {H});
        int64_t ow = out.add(d + "pre.w", int64_t(2) * H * 96), ob = out.add(d + "pre.b", 2 * H);
        if (wpre && bpre)
            for (int n = 0; n < H; ++n) {
                for (int k = 0; k < 96; ++k) out.p(ow)[n * 96 + k] = wpre[n * 96 + (flipped ? 95 - k : k)];
                out.p(ob)[n] = bpre[n];
            }  // rows [H, 2H) stay zero: they clear the WN skip accumulator
        const float* wc = L.get(s + "enc.cond_layer.weight", {2 * H * 3, 256, 1});
        const float* bc = L.get(s + "enc.cond_layer.bias", {2 * H * 3});
        std::vector<double> gc(2 * H * 3, 0.0);
        if (wc && bc)
            for (int n = 0; n < 2 * H * 3; ++n) {
                double a = bc[n];
                for (int k = 0; k < 256; ++k) a += double(wc[n * 256 + k]) * double(g[k]);
                gc[n] = a;
            }
        for (int i = 0; i < 3 && L.ok; ++i) {
            const float* w = L.get(s + "enc.in_layers." + std::to_string(i) + ".weight", {2 * H, H, 5});
            const float* b = L.get(s + "enc.in_layers." + std::to_string(i) + ".bias", {2 * H});
            int64_t o2 = out.add(d + "in" + std::to_string(i) + ".w", int64_t(2) * H * 5 * H);
            int64_t o3 = out.add(d + "in" + std::to_string(i) + ".b", 2 * H);
            if (w && b)
                for (int n = 0; n < 2 * H; ++n) {
                    int src = (n & 1) ? H + n / 2 : n / 2;  // interleave (tanh_c, sigmoid_c) pairs
                    for (int c = 0; c < H; ++c) for (int j = 0; j < 5; ++j)
                        out.p(o2)[(int64_t(n) * 5 + j) * H + c] = w[(int64_t(src) * H + c) * 5 + j];
                    out.p(o3)[n] = float(double(b[src]) + gc[i * 2 * H + src]);
                }
            int rs = i < 2 ? 2 * H : H;
            copy_vec(out, d + "rs" + std::to_string(i) + ".w", L.get(s + "enc.res_skip_layers." + std::to_string(i) + ".weight", {rs, H, 1}), int64_t(rs) * H);
            copy_vec(out, d + "rs" + std::to_string(i) + ".b", L.get(s + "enc.res_skip_layers." + std::to_string(i) + ".bias", {rs}), rs);
        }
        const float* wpo = L.get(s + "post.weight", {96, H, 1});
        const float* bpo = L.get(s + "post.bias", {96});
        int64_t o4 = out.add(d + "post.w", int64_t(96) * H), o5 = out.add(d + "post.b", 96);
        if (wpo && bpo)
            for (int p = 0; p < 96; ++p) {
                int src = flipped ? 95 - p : p;
                for (int k = 0; k < H; ++k) out.p(o4)[p * H + k] = wpo[src * H + k];
                out.p(o5)[p] = -bpo[src];  // x1 - m with alpha = -1
            }
    }

    // GeneratorNSF
    {
        const float* lw = L.get("dec.m_source.l_linear.weight", {1, 1});
        const float* lb = L.get("dec.m_source.l_linear.bias", {1});
        if (lw && lb) { info.lin_w = lw[0]; info.lin_b = lb[0]; }
        pack_conv1d(out, "pre.w", L.get("dec.conv_pre.weight", {512, H, 7}), 512, H, 7);
        const float* b = L.get("dec.conv_pre.bias", {512});
        const float* wc = L.get("dec.cond.weight", {512, 256, 1});
        const float* bc = L.get("dec.cond.bias", {512});
        int64_t ob = out.add("pre.b", 512);
        if (b && wc && bc)
            for (int n = 0; n < 512; ++n) {
                double a = double(b[n]) + double(bc[n]);
                for (int k = 0; k < 256; ++k) a += double(wc[n * 256 + k]) * double(g[k]);
                out.p(ob)[n] = float(a);
            }
    }
    const int* RATES = info.rates; const int* UK = info.up_kernels;
    static const int RK[3] = {3, 7, 11};
    for (int i = 0; i < 4 && L.ok; ++i) {
        int cin = 512 >> i, cout = 512 >> (i + 1), k = UK[i], u = RATES[i];
        if (k > 2 * u) { err = "transposed conv kernel wider than two strides"; return false; }
        std::string d = "U" + std::to_string(i) + ".";
        const float* w = L.get("dec.ups." + std::to_string(i) + ".weight", {cin, cout, k});
        const float* b = L.get("dec.ups." + std::to_string(i) + ".bias", {cout});
        // out[q*u + r - p] = x[q-1].w[:, :, r+u] + x[q].w[:, :, r]   (DESIGN.md "ConvTranspose1d")
        int64_t ow = out.add(d + "up.w", int64_t(u) * cout * 2 * cin), ob = out.add(d + "up.b", u * cout);
        if (w && b)
            for (int r = 0; r < u; ++r) for (int co = 0; co < cout; ++co) {
                int64_t n = int64_t(r) * cout + co;
                out.p(ob)[n] = b[co];
                for (int ci = 0; ci < cin; ++ci) {
                    if (r + u < k) out.p(ow)[n * 2 * cin + ci] = w[(int64_t(ci) * cout + co) * k + r + u];
                    out.p(ow)[n * 2 * cin + cin + ci] = w[(int64_t(ci) * cout + co) * k + r];
                }
            }
        int sf = 1; for (int j = i + 1; j < 4; ++j) sf *= RATES[j];
        int nk = (i + 1 < 4) ? 2 * sf : 1;
        copy_vec(out, d + "noise.w", L.get("dec.noise_convs." + std::to_string(i) + ".weight", {cout, 1, nk}), int64_t(cout) * nk);
        copy_vec(out, d + "noise.b", L.get("dec.noise_convs." + std::to_string(i) + ".bias", {cout}), cout);
        for (int j = 0; j < 3; ++j) for (int dd = 0; dd < 3; ++dd) {
            std::string s = "dec.resblocks." + std::to_string(i * 3 + j) + ".";
            std::string dn = d + "rb" + std::to_string(j) + "." + std::to_string(dd) + ".";
            pack_conv1d(out, dn + "c1.w", L.get(s + "convs1." + std::to_string(dd) + ".weight", {cout, cout, RK[j]}), cout, cout, RK[j]);
            copy_vec(out, dn + "c1.b", L.get(s + "convs1." + std::to_string(dd) + ".bias", {cout}), cout);
            pack_conv1d(out, dn + "c2.w", L.get(s + "convs2." + std::to_string(dd) + ".weight", {cout, cout, RK[j]}), cout, cout, RK[j]);
            copy_vec(out, dn + "c2.b", L.get(s + "convs2." + std::to_string(dd) + ".bias", {cout}), cout);
        }
    }
    pack_conv1d(out, "post.w", L.get("dec.conv_post.weight", {1, 32, 7}), 1, 32, 7);
    return L.ok;
}

const NamedBuf* Plan::find(const std::string& name) const {
    for (const auto& b : bufs) if (b.name == name) return &b;
    return nullptr;
}

}  // namespace rvc
