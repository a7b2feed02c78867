// plan.cpp - builds the per-window op list (host-only).  See ops.h / model.h.
//
// Layout rules used throughout (DESIGN.md "Data layout in HBM"):
//   * activations are fp32 channels-last; 1-D sequences are [rows, C], 2-D maps are
//     [T+2, F+2, C] with a one-pixel zero halo (+ a small zero guard after the last row);
//   * halo rows/pixels are zeroed once when the work arena is created and every op preserves
//     them (masked epilogues), so "same" padding costs nothing at run time;
//   * every convolution is the implicit GEMM of ops.h over overlapping / segmented rows.
#include <algorithm>
#include <cmath>
#include <cstdio>

#include "gemm_sched.h"
#include "model.h"

namespace rvc {

namespace {

struct PB {
    Plan& plan;
    std::string& err;
    bool ok = true;
    int64_t work = 0;
    int lane = 0;
    int sc_lane = -1;   // >= 0: the 1x1 shortcut convs of RMVPE's residual blocks run on this lane, beside c1
    bool allow_umma = true;
    bool f0_umma = false;
    bool cv_stack = false;
    int fuse_cbr = 0;
    int cv_gate = -1;   // >= 0: lane 0 (ContentVec) waits for RMVPE encoder level cv_gate

    void fail(const std::string& m) { if (ok) err = m; ok = false; }

    std::vector<int64_t> alloc_starts;  // byte offsets of every work-arena allocation (ascending)

    Ref alloc(const std::string& name, int64_t elems, bool is_int = false) {
        work = (work + 255) & ~int64_t(255);
        Ref r{SP_WORK, work};
        alloc_starts.push_back(work);
        work += elems * 4;
        if (!name.empty()) plan.bufs.push_back(NamedBuf{name, r, elems, is_int ? 1 : 0});
        return r;
    }
    void alias(const std::string& name, Ref r, int64_t elems, bool is_int = false) {
        plan.bufs.push_back(NamedBuf{name, r, elems, is_int ? 1 : 0});
    }
    // halo-padded NHWC map: interior T x F, C channels; 4 guard pixels at the end
    Ref pad2d(const std::string& name, int T, int F, int C) {
        return alloc(name, (int64_t(T + 2) * (F + 2) + 4) * C);
    }
    Ref w(const Packed* p, Space sp, const std::string& name) {
        int64_t o = p ? p->at(name) : -1;
        if (o < 0) { fail("packed tensor missing: " + name); return Ref{}; }
        return Ref{sp, o * 4};
    }
    Op& add(OpKind k, const std::string& name) {
        plan.ops.emplace_back();
        Op& op = plan.ops.back();
        op.kind = k; op.lane = lane; op.name = name;
        return op;
    }
    void wait(int src, int dst) {
        if (src == dst) return;
        Op& op = add(OP_WAIT, "wait");
        op.wait.src_lane = src; op.wait.dst_lane = dst;
        if (src + 1 > plan.n_lanes) plan.n_lanes = src + 1;
        if (dst + 1 > plan.n_lanes) plan.n_lanes = dst + 1;
    }
    GemmOp& gemm(const std::string& name, Ref A, int64_t lda, int seg_len, int64_t seg_stride, Ref W,
                 int64_t ldw, Ref bias, Ref C, int64_t ldc, int M, int N, int K, int act) {
        Op& op = add(OP_GEMM, name);
        GemmOp& g = op.gemm;
        g.A = A; g.lda = lda; g.seg_len = seg_len; g.seg_stride = seg_stride;
        g.W = W; g.ldw = ldw; g.bias = bias; g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K; g.act = act;
        return g;
    }
    void layernorm(const std::string& name, Ref X, int64_t ldx, Ref Y, int64_t ldy, Ref g, Ref b, int rows,
                   int cols) {
        Op& op = add(OP_LAYERNORM, name);
        op.ln.X = X; op.ln.ldx = ldx; op.ln.Y = Y; op.ln.ldy = ldy; op.ln.gamma = g; op.ln.beta = b;
        op.ln.rows = rows; op.ln.cols = cols; op.ln.eps = 1e-5f;
    }
};

std::string S(int i) { return std::to_string(i); }

// ------------------------------------------------------------------------------------------
// ContentVec (rvc.rs:81-97 `hubert`): pcm[N] -> [T, C]
// ------------------------------------------------------------------------------------------
Ref build_contentvec(PB& b, const Packed* P, const CvInfo& info, Ref pcm, int N, int& T_out) {
    static const int K[7] = {10, 3, 3, 3, 3, 2, 2};
    auto W = [&](const std::string& n) { return b.w(P, SP_CV, n); };
    if (N < 400) { b.fail("input shorter than one HuBERT frame"); return Ref{}; }
    int T0 = (N - 10) / 5 + 1;
    Ref stats = b.alloc("cv.gn_stats", 1024);
    {
        Op& op = b.add(OP_CONV0_STATS, "cv.conv0_stats");
        op.c0s.pcm = pcm; op.c0s.w = W("conv0.w"); op.c0s.stats = stats; op.c0s.T = T0; op.c0s.C = 512;
        op.c0s.k = 10; op.c0s.stride = 5; op.c0s.eps = 1e-5f;
    }
    Ref prev = b.alloc("cv.conv0", int64_t(T0) * 512);
    {
        Op& op = b.add(OP_CONV0_APPLY, "cv.conv0");
        op.c0a.pcm = pcm; op.c0a.w = W("conv0.w"); op.c0a.stats = stats; op.c0a.gamma = W("gn.g");
        op.c0a.beta = W("gn.b"); op.c0a.Y = prev; op.c0a.T = T0; op.c0a.C = 512; op.c0a.k = 10; op.c0a.stride = 5;
    }
    int Tp = T0;
    for (int i = 1; i < 7; ++i) {
        int Ti = (Tp - K[i]) / 2 + 1;
        Ref y = b.alloc("cv.conv" + S(i), int64_t(Ti) * 512);
        // stride-2 conv over channels-last rows: row t of the im2col matrix is the contiguous
        // span x[2t .. 2t+k) -> plain GEMM with overlapping rows (lda = 2*512)
        b.gemm("cv.conv" + S(i), prev, 2 * 512, K[i] * 512, 0, W("conv" + S(i) + ".w"), K[i] * 512, Ref{}, y,
               512, Ti, 512, K[i] * 512, ACT_GELU);
        prev = y; Tp = Ti;
    }
    const int T = Tp;
    T_out = T;
    Ref ln0 = b.alloc("cv.ln0", int64_t(T) * 512);
    b.layernorm("cv.ln0", prev, 512, ln0, 512, W("ln.g"), W("ln.b"), T, 512);
    Ref xpad = b.alloc("", int64_t(T + 128) * 768);  // 64 zero rows either side (pos_conv pad)
    Ref x = xpad.plus(64 * 768);
    b.alias("cv.proj", x, int64_t(T) * 768);
    b.gemm("cv.proj", ln0, 512, 512, 0, W("proj.w"), 512, W("proj.b"), x, 768, T, 768, 512, ACT_NONE);
    Ref pc = b.alloc("cv.posconv", int64_t(T) * 768);
    {   // grouped Conv1d(768,768,k=128,pad=64,groups=16), last output dropped, GELU, + x
        GemmOp& g = b.gemm("cv.posconv", xpad, 768, 48, 768, W("pos.w"), 128 * 48, W("pos.b"), pc, 768, T, 48,
                           128 * 48, ACT_GELU);
        g.R = x; g.ldr = 768; g.batch = 16; g.sA = 48; g.sW = int64_t(48) * 128 * 48; g.sBias = 48; g.sC = 48;
        g.sR = 48; g.cta_budget = 128;   // 16 groups x split-K 8: 6144-deep contraction, 24 k-blocks per CTA
    }
    Ref cur = b.alloc("cv.enc_in", int64_t(T) * 768);
    const int stack_first = int(b.plan.ops.size());
    b.layernorm("cv.enc_in", pc, 768, cur, 768, W("eln.g"), W("eln.b"), T, 768);
    for (int i = 0; i < info.n_layers; ++i) {
        std::string d = "L" + S(i) + ".", n = "cv.L" + S(i) + ".";
        Ref qkv = b.alloc(n + "qkv", int64_t(T) * 2304);
        b.gemm(n + "qkv", cur, 768, 768, 0, W(d + "qkv.w"), 768, W(d + "qkv.b"), qkv, 2304, T, 2304, 768, ACT_NONE);
        Ref a = b.alloc(n + "attn", int64_t(T) * 768);
        {
            Op& op = b.add(OP_ATTN, n + "attn");
            op.attn.qkv = qkv; op.attn.ldqkv = 2304; op.attn.out = a; op.attn.ldo = 768; op.attn.T = T;
            op.attn.heads = 12; op.attn.dim = 64;
        }
        Ref t1 = b.alloc(n + "o", int64_t(T) * 768);
        { GemmOp& g = b.gemm(n + "o", a, 768, 768, 0, W(d + "o.w"), 768, W(d + "o.b"), t1, 768, T, 768, 768, ACT_NONE);
          g.R = cur; g.ldr = 768; }
        Ref x1 = b.alloc(n + "ln1", int64_t(T) * 768);
        b.layernorm(n + "ln1", t1, 768, x1, 768, W(d + "ln1.g"), W(d + "ln1.b"), T, 768);
        Ref h = b.alloc(n + "fc1", int64_t(T) * 3072);
        b.gemm(n + "fc1", x1, 768, 768, 0, W(d + "fc1.w"), 768, W(d + "fc1.b"), h, 3072, T, 3072, 768, ACT_GELU);
        Ref t2 = b.alloc(n + "fc2", int64_t(T) * 768);
        { GemmOp& g = b.gemm(n + "fc2", h, 3072, 3072, 0, W(d + "fc2.w"), 3072, W(d + "fc2.b"), t2, 768, T, 768, 3072, ACT_NONE);
          g.R = x1; g.ldr = 768; }
        Ref x2 = b.alloc("cv.layer" + S(i), int64_t(T) * 768);
        b.layernorm("cv.layer" + S(i), t2, 768, x2, 768, W(d + "ln2.g"), W(d + "ln2.b"), T, 768);
        cur = x2;
    }
    if (b.cv_stack && b.allow_umma && T <= 128 && info.n_layers * 7 + 1 <= 128) {
        // the layers run as one persistent tcgen05 kernel (kernels_cvstack.cu); the ops stay in the list for the CPU
        // plan interpreter and the per-op tools, the CUDA launcher replaces them by one launch
        CvStackInfo& cs = b.plan.cvstack;
        cs.first = stack_first; cs.count = int(b.plan.ops.size()) - stack_first; cs.T = T; cs.width = 768; cs.ffn = 3072;
        for (int k = 0; k < cs.count; ++k) b.plan.ops[size_t(stack_first + k)].stack = 1;
        cs.planes_x = b.alloc("cv.dbg_planes_x", int64_t(2) * 128 * 768 / 2);     // [hi | lo'][128][768] halves
        cs.planes_x1 = b.alloc("", int64_t(2) * 128 * 768 / 2);
        cs.planes_a = b.alloc("", int64_t(2) * 128 * 768 / 2);
        cs.planes_h = b.alloc("", int64_t(2) * 128 * 3072 / 2);
        cs.partial = b.alloc("cv.dbg_partial", int64_t(4) * 128 * 768);          // [z][128][768] fp32
    }
    if (info.final_proj) {
        Ref o = b.alloc("cv.out", int64_t(T) * 256);
        b.gemm("cv.out", cur, 768, 768, 0, W("fp.w"), 768, W("fp.b"), o, 256, T, 256, 768, ACT_NONE);
        cur = o;
    } else {
        b.alias("cv.out", cur, int64_t(T) * 768);
    }
    return cur;
}

// ------------------------------------------------------------------------------------------
// RMVPE (rmvpe.rs:250-261 `pitch`): last L samples -> salience [T,360] -> f0 [T]
// ------------------------------------------------------------------------------------------

struct Map2d { Ref base; int T, F, C; };  // halo-padded NHWC, C = full pixel stride

int64_t interior(const Map2d& m) { return (int64_t(m.F + 2) + 1) * m.C; }

void conv3x3(PB& b, const Packed* P, const std::string& name, const std::string& wname, const Map2d& in,
             int cout, Ref dst_interior, int64_t ld_dst, int act, Ref R, int64_t ldr) {
    Ref A = in.base, Wt = b.w(P, SP_F0, wname + ".w");
    int K = 9 * in.C;
    if (in.T == 1) {
        // a single time row: kernel rows dt=0 and dt=2 only ever see the zero halo, so only the
        // middle third of every filter is streamed (3x less weight traffic at the U-Net bottleneck)
        A = in.base.plus(int64_t(in.F + 2) * in.C); Wt = Wt.plus(3 * in.C); K = 3 * in.C;
    }
    GemmOp& g = b.gemm(name, A, in.C, 3 * in.C, int64_t(in.F + 2) * in.C, Wt, 9 * in.C, b.w(P, SP_F0, wname + ".b"),
                       dst_interior, ld_dst, in.T * (in.F + 2), cout, K, act);
    g.mask_period = in.F + 2; g.mask_valid = in.F; g.R = R; g.ldr = ldr;
}

// ConvBlockRes: relu(bn(conv(relu(bn(conv(x)))))) + (shortcut(x) | x); result written into the
// interior of `dst` (which may be a channel slice of a wider map: ld_dst = its pixel stride).
void conv_block_res(PB& b, const Packed* P, const std::string& name, const std::string& wp, const Map2d& in,
                    int cout, Ref dst_interior, int64_t ld_dst) {
    Map2d t1{b.pad2d(name + "t1", in.T, in.F, cout), in.T, in.F, cout};
    Ref R; int64_t ldr;
    const int M = in.T * (in.F + 2);
    // levels 0 / 1 of the U-Net: the whole block is one kernel on the GPU (kernels_cbr.cu).  The three GEMM ops stay in
    // the plan - they ARE the block's definition (CPU interpreter, cost accounting) - and are marked as covered.
    const bool fuse = (b.fuse_cbr == 1 || (b.fuse_cbr == 2 && name.compare(0, 6, "rm.dec") == 0)) && cbr_supported(cout, in.C, in.F);
    const size_t first = b.plan.ops.size();
    if (in.C != cout) {
        Ref sc = b.alloc(name + "sc", int64_t(M) * cout);
        const int home = b.lane;
        const bool side = !fuse && b.sc_lane >= 0 && b.sc_lane != home;
        if (side) { b.wait(home, b.sc_lane); b.lane = b.sc_lane; }   // shortcut and c1 both only read `in`
        b.gemm(name + "sc", in.base.plus(interior(in)), in.C, in.C, 0, b.w(P, SP_F0, wp + "sc.w"), in.C,
               b.w(P, SP_F0, wp + "sc.b"), sc, cout, M, cout, in.C, ACT_NONE);
        if (side) b.lane = home;
        conv3x3(b, P, name + "c1", wp + "c1", in, cout, t1.base.plus(interior(t1)), cout, ACT_RELU, Ref{}, 0);
        if (side) b.wait(b.sc_lane, home);
        R = sc; ldr = cout;
    } else {
        conv3x3(b, P, name + "c1", wp + "c1", in, cout, t1.base.plus(interior(t1)), cout, ACT_RELU, Ref{}, 0);
        R = in.base.plus(interior(in)); ldr = in.C;
    }
    conv3x3(b, P, name + "c2", wp + "c2", t1, cout, dst_interior, ld_dst, ACT_RELU, R, ldr);
    if (fuse && b.ok) {
        for (size_t i = first; i < b.plan.ops.size(); ++i) b.plan.ops[i].fuse = 1;
        Op& last = b.plan.ops.back();
        last.fuse = 2;
        CbrOp& c = last.cbr;
        c.in = in.base; c.T = in.T; c.F = in.F; c.Cin = in.C; c.C = cout; c.dst = dst_interior; c.ld_dst = ld_dst;
        c.w1 = b.w(P, SP_F0, wp + "c1.w"); c.b1 = b.w(P, SP_F0, wp + "c1.b");
        c.w2 = b.w(P, SP_F0, wp + "c2.w"); c.b2 = b.w(P, SP_F0, wp + "c2.b");
        if (in.C != cout) { c.wsc = b.w(P, SP_F0, wp + "sc.w"); c.bsc = b.w(P, SP_F0, wp + "sc.b"); }
    }
}

struct F0Out { Ref salience, f0, argmax, mel; int T; };

F0Out build_rmvpe(PB& b, const Packed* P, const F0Info& info, Ref pcm_window, int L, Ref params,
                  bool mel_only, int upstream_window) {
    F0Out o{};
    const int T = 1 + L / 160, F = 128;
    o.T = T;
    auto W = [&](const std::string& n) { return b.w(P, SP_F0, n); };
    o.mel = b.alloc("mel", int64_t(T) * 128);
    Map2d in0{Ref{}, T, F, 1};
    if (!mel_only) {
        if (T % 32 != 0) { b.fail("mel frame count must be a multiple of 32 (rmvpe.rs:227)"); return o; }
        in0.base = b.pad2d("rm.in", T, F, 1);
    }
    {
        Op& op = b.add(OP_STFTMEL, "mel");
        StftMelOp& s = op.stft;
        s.pcm = pcm_window; s.L = L; s.T = T; s.window = W("window"); s.band_start = W("mel.start");
        s.band_count = W("mel.count"); s.band_off = W("mel.off"); s.band_w = W("mel.w"); s.mel = o.mel;
        if (!mel_only) { s.out2 = in0.base.plus(interior(in0)); s.out2_pitch = F + 2; }
        s.scale = info.in_scale; s.shift = info.in_shift; s.clamp = 1e-5f;
    }
    if (mel_only) return o;

    // encoder: 5 levels x 4 ConvBlockRes, 2x2 average pool; the un-pooled output of level i is
    // written straight into the upper channel half of the decoder's concat map
    Map2d cat[5];
    Map2d cur = in0;
    int C = 16, Tl = T, Fl = F;
    for (int i = 0; i < 5; ++i) {
        cat[i] = Map2d{b.pad2d("rm.cat" + S(i), Tl, Fl, 2 * C), Tl, Fl, 2 * C};
        for (int j = 0; j < 4; ++j) {
            std::string nm = "rm.enc" + S(i) + "." + S(j) + ".", wp = "enc" + S(i) + "." + S(j) + ".";
            if (j < 3) {
                Map2d y{b.pad2d(nm + "y", Tl, Fl, C), Tl, Fl, C};
                conv_block_res(b, P, nm, wp, cur, C, y.base.plus(interior(y)), C);
                cur = y;
            } else {
                conv_block_res(b, P, nm, wp, cur, C, cat[i].base.plus(interior(cat[i]) + C), 2 * C);
            }
        }
        Map2d pooled{b.pad2d("rm.pool" + S(i), Tl / 2, Fl / 2, C), Tl / 2, Fl / 2, C};
        {
            Op& op = b.add(OP_AVGPOOL, "rm.pool" + S(i));
            op.pool.in = cat[i].base.plus(C); op.pool.ldin = 2 * C; op.pool.out = pooled.base;
            op.pool.T = Tl; op.pool.F = Fl; op.pool.C = C;
        }
        cur = pooled; Tl /= 2; Fl /= 2; C *= 2;
        // ContentVec's conv stem (wide grids, ~0.2 ms) starts once this level is done: beside it the full-resolution
        // convs of level 0 take twice as long, and the ContentVec lane has the slack (PB::cv_gate)
        if (i == b.cv_gate && b.lane != 0) b.wait(b.lane, 0);
    }
    // intermediate: 16 blocks at (T/32) x 4, 256 -> 512
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            std::string nm = "rm.mid" + S(i) + "." + S(j) + ".", wp = "mid" + S(i) + "." + S(j) + ".";
            Map2d y{b.pad2d(nm + "y", Tl, Fl, 512), Tl, Fl, 512};
            conv_block_res(b, P, nm, wp, cur, 512, y.base.plus(interior(y)), 512);
            cur = y;
        }
    b.alias("rm.inter", cur.base, (int64_t(Tl + 2) * (Fl + 2)) * 512);
    // decoder: ConvTranspose2d(k3,s2) as one GEMM over the 2x2 input neighbourhood with the
    // four output phases stacked along N, scattered pixel-shuffle style into the concat map
    int cin = 512;
    for (int i = 0; i < 5; ++i) {
        int cout = cin / 2;
        const Map2d& ct = cat[4 - i];
        std::string d = "dec" + S(i) + ".";
        {
            GemmOp& g = b.gemm("rm.dec" + S(i) + ".up", cur.base.plus(interior(cur)), cin, 2 * cin,
                               int64_t(cur.F + 2) * cin, W(d + "up.w"), 4 * cin, W(d + "up.b"),
                               ct.base.plus(interior(ct)), 2 * cout, cur.T * (cur.F + 2), 4 * cout, 4 * cin,
                               ACT_RELU);
            g.out_mode = OUT_PIXSHUF2; g.om_a = cur.F + 2; g.om_b = cout; g.om_c = ct.F + 2;
        }
        cur = ct;
        for (int j = 0; j < 4; ++j) {
            std::string nm = "rm.dec" + S(i) + "." + S(j) + ".";
            Map2d y{b.pad2d(nm + "y", ct.T, ct.F, cout), ct.T, ct.F, cout};
            conv_block_res(b, P, nm, d + S(j) + ".", cur, cout, y.base.plus(interior(y)), cout);
            cur = y;
        }
        cin = cout;
    }
    b.alias("rm.dec4", cur.base, (int64_t(T + 2) * (F + 2)) * 16);
    // 3 output channels stored with a pixel stride of 4 (4th stays zero): rows of (F+2)*4 floats are 16-byte
    // aligned, so the GRU input projection runs on the vectorised split-K kernel instead of the scalar one
    Map2d cn{b.pad2d("rm.cnn", T, F, 4), T, F, 4};
    conv3x3(b, P, "rm.cnn", "cnn", cur, 3, cn.base.plus(interior(cn)), 4, ACT_NONE, Ref{}, 0);
    // BiGRU input projections for both directions: rows of the padded cnn map are the GEMM rows
    const int KI = (F + 2) * 4;
    Ref gi = b.alloc("rm.gi", int64_t(T) * 1536);
    b.gemm("rm.gi", cn.base.plus(KI), KI, KI, 0, W("gru.wih"), KI, W("gru.bih"), gi, 1536, T, 1536, KI, ACT_NONE);
    Ref h = b.alloc("rm.gru", int64_t(T) * 512);
    {
        Op& op = b.add(OP_GRU, "rm.gru");
        op.gru.gi = gi; op.gru.whh_t = W("gru.whh_t"); op.gru.bhh = W("gru.bhh"); op.gru.out = h; op.gru.T = T;
        op.gru.H = 256;
    }
    o.salience = b.alloc("rm.salience", int64_t(T) * 360);
    b.gemm("rm.salience", h, 512, 512, 0, W("fc.w"), 512, W("fc.b"), o.salience, 360, T, 360, 512, ACT_SIGMOID);
    o.f0 = b.alloc("f0", T);
    o.argmax = b.alloc("f0_argmax", T, true);
    {
        Op& op = b.add(OP_F0DECODE, "f0");
        op.f0d.salience = o.salience; op.f0d.f0 = o.f0; op.f0d.argmax = o.argmax; op.f0d.params = params;
        op.f0d.T = T; op.f0d.bins = 360; op.f0d.threshold = 0.03f; op.f0d.upstream_window = upstream_window;
    }
    return o;
}

// ------------------------------------------------------------------------------------------
// Synthesizer (rvc.rs:193-214): phone [R,C], pitch i32[R], pitchf [R] -> audio [R*400]
// ------------------------------------------------------------------------------------------
Ref build_synth(PB& b, const Packed* P, const SynInfo& info, Ref phone, Ref pitch, Ref pitchf, Ref params,
                Ref audio, int R, bool multi_lane, Ref emb_pre = Ref{}) {
    const int H = 192;
    auto W = [&](const std::string& n) { return b.w(P, SP_SYN, n); };
    // ---- NSF sine source ------------------------------------------------------------------
    // It needs only pitchf: forked onto lane 1 HERE, before enc_p / flow are emitted - a fork emitted after them would
    // record lane 0's position behind the flow chain and the kernel would run on the critical path, beside conv_pre
    // (measured with a lane stamp: sine source at 2044 us, right after the flow chain; RVC_SINE_EARLY=0 restores that).
    const int upp = info.sr / 100, L = R * upp, HH = 32;   // samples per 10 ms feature frame = product of the upsample rates
    Ref harpad = b.alloc("", L + 2 * HH);
    Ref har = harpad.plus(HH);
    b.alias("sy.har", har, L);
    Ref sine_dbg = b.alloc("sy.sine", L);
    auto emit_sine = [&]() {
        if (multi_lane) { b.wait(0, 1); b.lane = 1; }
        Op& op = b.add(OP_SINEGEN, "sy.sine");
        op.sine.pitchf = pitchf; op.sine.out = har; op.sine.sine_dbg = sine_dbg; op.sine.params = params;
        op.sine.R = R; op.sine.upp = upp; op.sine.sr = float(info.sr); op.sine.lin_w = info.lin_w;
        op.sine.lin_b = info.lin_b;
        b.lane = 0;
    };
    static const bool sine_early = sched_env("RVC_SINE_EARLY", 1) != 0;
    if (sine_early) emit_sine();
    // ---- enc_p ---------------------------------------------------------------------------
    Ref xpad = b.alloc("", int64_t(R + 2) * H);  // 1 zero row either side (FFN k=3)
    Ref x = xpad.plus(H);
    b.alias("sy.emb", x, int64_t(R) * H);
    {
        Op& op = b.add(OP_EMBED, "sy.emb");
        op.embed.phone = phone; op.embed.pitch = pitch; op.embed.wp = W("emb.wp"); op.embed.bp = W("emb.bp");
        op.embed.emb_pitch = W("emb.pitch"); op.embed.out = x; op.embed.ldo = H; op.embed.R = R;
        op.embed.Cin = info.phone_dim; op.embed.H = H; op.embed.pre = emb_pre;
    }
    for (int i = 0; i < 6; ++i) {
        std::string d = "E" + S(i) + ".", n = "sy.E" + S(i) + ".";
        Ref qkv = b.alloc(n + "qkv", int64_t(R) * 3 * H);
        b.gemm(n + "qkv", x, H, H, 0, W(d + "qkv.w"), H, W(d + "qkv.b"), qkv, 3 * H, R, 3 * H, H, ACT_NONE);
        Ref a = b.alloc(n + "attn", int64_t(R) * H);
        {
            Op& op = b.add(OP_RELATTN, n + "attn");
            op.relattn.qkv = qkv; op.relattn.ldqkv = 3 * H; op.relattn.out = a; op.relattn.ldo = H;
            op.relattn.rel_k = W(d + "relk"); op.relattn.rel_v = W(d + "relv"); op.relattn.T = R;
            op.relattn.heads = 2; op.relattn.dim = 96; op.relattn.window = 10;
        }
        Ref t1 = b.alloc(n + "o", int64_t(R) * H);
        { GemmOp& g = b.gemm(n + "o", a, H, H, 0, W(d + "o.w"), H, W(d + "o.b"), t1, H, R, H, H, ACT_NONE);
          g.R = x; g.ldr = H; }
        Ref x1pad = b.alloc("", int64_t(R + 2) * H);
        Ref x1 = x1pad.plus(H);
        b.alias(n + "ln1", x1, int64_t(R) * H);
        b.layernorm(n + "ln1", t1, H, x1, H, W(d + "ln1.g"), W(d + "ln1.b"), R, H);
        Ref hpad = b.alloc("", int64_t(R + 2) * 768);
        Ref h = hpad.plus(768);
        b.alias(n + "ffn1", h, int64_t(R) * 768);
        b.gemm(n + "ffn1", x1pad, H, 3 * H, 0, W(d + "ffn1.w"), 3 * H, W(d + "ffn1.b"), h, 768, R, 768, 3 * H, ACT_RELU);
        Ref t2 = b.alloc(n + "ffn2", int64_t(R) * H);
        { GemmOp& g = b.gemm(n + "ffn2", hpad, 768, 3 * 768, 0, W(d + "ffn2.w"), 3 * 768, W(d + "ffn2.b"), t2, H, R, H, 3 * 768, ACT_NONE);
          g.R = x1; g.ldr = H; }
        Ref x2pad = b.alloc("", int64_t(R + 2) * H);
        Ref x2 = x2pad.plus(H);
        b.alias("sy.enc" + S(i), x2, int64_t(R) * H);
        b.layernorm("sy.enc" + S(i), t2, H, x2, H, W(d + "ln2.g"), W(d + "ln2.b"), R, H);
        xpad = x2pad; x = x2;
    }
    Ref stats = b.alloc("sy.stats", int64_t(R) * 2 * H);
    b.gemm("sy.stats", x, H, H, 0, W("proj.w"), H, W("proj.b"), stats, 2 * H, R, 2 * H, H, ACT_NONE);
    // ---- z_p + flow (reverse) --------------------------------------------------------------
    Ref zpad = b.alloc("", int64_t(R + 6) * H);  // 3 zero rows either side (conv_pre k=7)
    Ref z = zpad.plus(3 * H);
    b.alias("sy.z", z, int64_t(R) * H);
    {
        Op& op = b.add(OP_ZP, "sy.z_p");
        op.zp.stats = stats; op.zp.out = z; op.zp.ldo = H; op.zp.params = params; op.zp.R = R; op.zp.H = H;
    }
    for (int fl = 3; fl >= 0; --fl) {
        const bool flipped = (fl == 3 || fl == 1);
        std::string d = "F" + S(fl) + ".", n = "sy.F" + S(fl) + ".";
        Ref xopad = b.alloc("", int64_t(R + 4) * 2 * H);  // [x | skip-sum], 2 zero rows either side
        Ref xo = xopad.plus(2 * 2 * H);
        b.alias(n + "xo", xo, int64_t(R) * 2 * H);
        b.gemm(n + "pre", flipped ? z.plus(96) : z, H, 96, 0, W(d + "pre.w"), 96, W(d + "pre.b"), xo, 2 * H, R,
               2 * H, 96, ACT_NONE);
        for (int i = 0; i < 3; ++i) {
            Ref acts = b.alloc(n + "acts" + S(i), int64_t(R) * H);
            b.gemm(n + "in" + S(i), xopad, 2 * H, H, 2 * H, W(d + "in" + S(i) + ".w"), 5 * H,
                   W(d + "in" + S(i) + ".b"), acts, H, R, 2 * H, 5 * H, ACT_GATE);
            int rs = i < 2 ? 2 * H : H;
            Ref dst = i < 2 ? xo : xo.plus(H);
            GemmOp& g = b.gemm(n + "rs" + S(i), acts, H, H, 0, W(d + "rs" + S(i) + ".w"), H,
                               W(d + "rs" + S(i) + ".b"), dst, 2 * H, R, rs, H, ACT_NONE);
            g.R = dst; g.ldr = 2 * H;
        }
        Ref x1 = flipped ? z : z.plus(96);
        GemmOp& g = b.gemm("sy.flow" + S(fl), xo.plus(H), 2 * H, H, 0, W(d + "post.w"), H, W(d + "post.b"), x1, H, R,
                           96, H, ACT_NONE);
        g.alpha = -1.0f; g.R = x1; g.ldr = H;
    }
    // ---- GeneratorNSF ----------------------------------------------------------------------
    if (!sine_early) emit_sine();
    const int* RATES = info.rates; const int* UK = info.up_kernels;
    static const int RK[3] = {3, 7, 11}, RD[3] = {1, 3, 5};
    if (RATES[0] * RATES[1] * RATES[2] * RATES[3] != upp) { b.fail("generator upsample rates do not multiply to sr / 100"); return audio; }
    // the three ResBlocks of a stage run on three lanes at once: each conv is scheduled for a third of the GPU
    // so the lanes really overlap instead of queueing behind each other's full-width grids
    static const int rb_budget_env = sched_env("RVC_RB_WANT", 0);
    const int rb_budget = multi_lane ? rb_budget_env : 0;
    // conv_pre (+ cond(g) folded into the bias); lrelu'd copy feeds ups[0] (1 halo row)
    Ref pre_raw = b.alloc("sy.conv_pre", int64_t(R) * 512);
    Ref upin_pad = b.alloc("", int64_t(R + 2) * 512);
    {
        GemmOp& g = b.gemm("sy.conv_pre", zpad, H, 7 * H, 0, W("pre.w"), 7 * H, W("pre.b"), pre_raw, 512, R, 512,
                           7 * H, ACT_NONE);
        g.C2 = upin_pad.plus(512); g.ldc2 = 512; g.act2 = ACT_LRELU01;
    }
    if (multi_lane) b.wait(1, 0);  // har is consumed by the noise convs
    int Tin = R, cin = 512;
    for (int i = 0; i < 4; ++i) {
        const int u = RATES[i], k = UK[i], p = (k - u) / 2, cout = cin / 2, Tout = Tin * u;
        std::string d = "U" + S(i) + ".", n = "sy.U" + S(i) + ".";
        Ref xu_pad = b.alloc("", int64_t(Tout + 2 * HH) * cout);
        Ref xu = xu_pad.plus(int64_t(HH) * cout);
        b.alias("sy.up" + S(i), xu, int64_t(Tout) * cout);
        Ref xa_pad = b.alloc("", int64_t(Tout + 2 * HH) * cout);
        Ref xa = xa_pad.plus(int64_t(HH) * cout);
        {   // ConvTranspose1d: rows (q-1, q) of the input -> u output phases per row q
            GemmOp& g = b.gemm(n + "up", upin_pad, cin, 2 * cin, 0, W(d + "up.w"), 2 * cin, W(d + "up.b"), xu, cout,
                               Tin + 1, u * cout, 2 * cin, ACT_NONE);
            g.out_mode = OUT_CONVT1D; g.om_a = u; g.om_b = cout; g.om_c = p; g.om_d = Tout;
        }
        {   // + noise_convs[i](har): Conv1d(1,cout,2sf,stride sf,pad sf/2) == GEMM over har rows
            int sf = 1; for (int j = i + 1; j < 4; ++j) sf *= RATES[j];
            int nk = (i + 1 < 4) ? 2 * sf : 1, npad = (i + 1 < 4) ? sf / 2 : 0;
            GemmOp& g = b.gemm(n + "noise", har.plus(-npad), (i + 1 < 4) ? sf : 1, nk, 0, W(d + "noise.w"), nk,
                               W(d + "noise.b"), xu, cout, Tout, cout, nk, ACT_NONE);
            g.R = xu; g.ldr = cout; g.C2 = xa; g.ldc2 = cout; g.act2 = ACT_LRELU01;
        }
        Ref ys[3];
        // RVC_SY_RB_LANES (experiment): 3 = the three ResBlocks of a stage on three lanes (default), 2 = rk 3 + rk 7 share
        // lane 1 beside rk 11 on lane 0 (equal work: 3 + 7 ~ 11 taps), 1 = one lane
        static const int rb_lanes = sched_env("RVC_SY_RB_LANES", 3);
        const int rbl = multi_lane ? rb_lanes : 1;
        // RVC_SY_RB_ORDER=1 (experiment): the longest ResBlock (rk 11) stays on lane 0 - no fork event in front of it - and is
        // emitted first; the short ones take the forked lanes
        static const int rb_order = sched_env("RVC_SY_RB_ORDER", 0);
        auto rb_lane = [&](int j) { return rbl >= 3 ? (rb_order ? 2 - j : j) : (rbl == 2 ? (j == 2 ? 0 : 1) : 0); };
        if (rbl >= 2) b.wait(0, 1);
        if (rbl >= 3) b.wait(0, 2);  // the three ResBlocks run concurrently
        for (int jj = 0; jj < 3; ++jj) {
            const int j = (rb_order && rbl >= 3) ? 2 - jj : jj;
            b.lane = rb_lane(j);
            const int rk = RK[j];
            std::string rn = n + "rb" + S(j) + ".";
            Ref ta_pad = b.alloc("", int64_t(Tout + 2 * HH) * cout);
            Ref ya_pad = b.alloc("", int64_t(Tout + 2 * HH) * cout);
            Ref y = b.alloc(rn + "y", int64_t(Tout) * cout);
            ys[j] = y;
            Ref in_act_pad = xa_pad, res = xu;
            for (int dd = 0; dd < 3; ++dd) {
                const int dil = RD[dd], pad1 = (rk * dil - dil) / 2, pad2 = (rk - 1) / 2;
                std::string wn = d + "rb" + S(j) + "." + S(dd) + ".";
                GemmOp& g1 = b.gemm(rn + S(dd) + ".c1", in_act_pad.plus(int64_t(HH - pad1) * cout), cout, cout, int64_t(dil) * cout,
                       W(wn + "c1.w"), rk * cout, W(wn + "c1.b"), ta_pad.plus(int64_t(HH) * cout), cout, Tout, cout,
                       rk * cout, ACT_LRELU01);
                g1.cta_budget = rb_budget;
                GemmOp& g = b.gemm(rn + S(dd) + ".c2", ta_pad.plus(int64_t(HH - pad2) * cout), cout, rk * cout, 0,
                                   W(wn + "c2.w"), rk * cout, W(wn + "c2.b"), y, cout, Tout, cout, rk * cout,
                                   ACT_NONE);
                g.R = res; g.ldr = cout; g.cta_budget = rb_budget;
                if (dd < 2) { g.C2 = ya_pad.plus(int64_t(HH) * cout); g.ldc2 = cout; g.act2 = ACT_LRELU01; }
                in_act_pad = ya_pad; res = y;
            }
            b.lane = 0;
        }
        if (rbl >= 2) b.wait(1, 0);
        if (rbl >= 3) b.wait(2, 0);
        Ref raw = b.alloc("sy.stage" + S(i), int64_t(Tout) * cout);
        Op& op = b.add(OP_AVG3, "sy.stage" + S(i));
        op.avg3.a = ys[0]; op.avg3.b = ys[1]; op.avg3.c = ys[2]; op.avg3.ld = cout; op.avg3.raw = raw;
        op.avg3.ldraw = cout; op.avg3.T = Tout; op.avg3.C = cout;
        if (i < 3) {
            upin_pad = b.alloc("", int64_t(Tout + 2) * cout);
            op.avg3.out = upin_pad.plus(cout); op.avg3.ldo = cout; op.avg3.slope = 0.1f;
        } else {
            Ref post_pad = b.alloc("", int64_t(Tout + 6) * cout);
            op.avg3.out = post_pad.plus(3 * cout); op.avg3.ldo = cout; op.avg3.slope = 0.01f;  // F.leaky_relu default
            Op& cp = b.add(OP_CONVPOST, "sy.audio");
            cp.cpost.in = post_pad; cp.cpost.w = W("post.w"); cp.cpost.out = audio; cp.cpost.T = Tout;
            cp.cpost.C = cout; cp.cpost.k = 7;
        }
        Tin = Tout; cin = cout;
    }
    return audio;
}

// second stage of a tensor-core retrieval: exact re-rank of the candidates (kernels_knn_umma.cu)
void knn_rerank_setup(KnnSelectOp& ks, bool um, const KnnScanOp& kd, const PlanOptions& opt) {
    if (!um) return;
    ks.rerank = 1; ks.index = kd.index; ks.queries = kd.queries; ks.ldq = kd.ldq; ks.N = kd.N; ks.C = kd.C; ks.ymax2 = opt.index_ymax2;
    ks.planes_off = opt.index_planes_off;
    ks.fallback_off = opt.index_planes_off + knn_umma_counters_off(kd.N, kd.C);
}

// chooses the kernel variant / split-K factor of every GEMM and allocates its scratch
void schedule_gemms(PB& b, int nb = 1) {
    for (Op& op : b.plan.ops) {
        if (op.kind != OP_GEMM) continue;
        GemmOp& g = op.gemm;
        GemmSched s = gemm_schedule(g, b.allow_umma, nb, b.f0_umma);
        g.sched_variant = s.variant; g.splitk = s.splitk;
        if (s.variant > 0 && s.variant < 5 && s.splitk > 1) {  // v2: partial tiles + one arrival counter per tile
            g.scratch = b.alloc("", int64_t(s.splitk) * g.batch * g.M * g.N);
            g.counters = b.alloc("", s.tiles, true);
        }
    }
    // tcgen05 kernel: the split-K partial tiles of a cluster are exchanged through L2.  Ops of one lane run one after
    // the other, so a lane needs one scratch area, sized for its largest op ([tiles][splitk][128][BN] floats).
    int64_t need[8] = {0};
    for (Op& op : b.plan.ops) {
        if (op.kind != OP_GEMM || op.gemm.sched_variant < 5 || op.gemm.splitk <= 1 || op.lane >= 8) continue;
        GemmSched s = gemm_schedule(op.gemm, b.allow_umma, nb, b.f0_umma);
        need[op.lane] = std::max(need[op.lane], int64_t(s.tiles) * s.splitk * 128 * s.bn);
    }
    Ref lane_scratch[8];
    for (int l = 0; l < 8; ++l) if (need[l] > 0) lane_scratch[l] = b.alloc("", need[l]);
    for (Op& op : b.plan.ops)
        if (op.kind == OP_GEMM && op.gemm.sched_variant >= 5 && op.gemm.splitk > 1 && op.lane < 8) op.gemm.scratch = lane_scratch[op.lane];
}

// ------------------------------------------------------------------------------------------
// Persistent chains (chain.h): runs of small same-lane ops that one cooperative kernel executes
// with grid barriers instead of kernel boundaries.
// ------------------------------------------------------------------------------------------
bool gemm_aligned(const GemmOp& g) {
    return g.A.off % 16 == 0 && g.W.off % 16 == 0 && g.lda % 4 == 0 && g.seg_len % 4 == 0 && g.seg_stride % 4 == 0 &&
           g.K % 4 == 0 && g.ldw % 4 == 0 && g.sA % 4 == 0 && g.sW % 4 == 0;
}

bool chain_eligible(const Op& op, int side_max_m) {
    if (op.stack) return false;   // runs inside the persistent ContentVec stack kernel
    if (op.fuse) return false;    // runs inside the fused residual-block kernel
    switch (op.kind) {
        case OP_GEMM: {
            const GemmOp& g = op.gemm;
            if (g.sched_variant >= 5) return false;              // tcgen05 kernel
            // side lanes share the GPU with lane 0: only their weight-streaming ops (tiny M, one tile per CTA of a
            // small grid) are worth a resident chain; the wide ones want all SMs for a few microseconds each
            if (op.lane != 0 && side_max_m > 0 && g.M > side_max_m) return false;
            if (gemm_aligned(g) && g.K >= 64) return true;       // tile path
            return g.K < 64 && g.batch == 1;                     // scalar path: tiny contractions only
        }
        case OP_AVGPOOL: return op.lane == 0 || side_max_m <= 0;
        case OP_LAYERNORM: return op.ln.cols <= 1024;
        case OP_RELATTN:  // q, p, K, V and both relative tables staged in the chain's shared memory (90 KB)
            return (op.relattn.dim + op.relattn.T + 4 + (2 * op.relattn.T + 2 * (2 * op.relattn.window + 1)) * op.relattn.dim) * 4 <= 88 * 1024;
        default: return false;
    }
}

// buffer identity at allocation granularity (work arena) / whole space (everything else)
int64_t buf_id(const PB& b, const Ref& r) {
    if (r.null()) return INT64_MIN;
    if (r.space != SP_WORK) return -int64_t(r.space) - 1;
    auto it = std::upper_bound(b.alloc_starts.begin(), b.alloc_starts.end(), r.off);
    return int64_t(it - b.alloc_starts.begin()) - 1;
}

void op_buffers(const PB& b, const Op& op, std::vector<int64_t>& rd, std::vector<int64_t>& wr) {
    rd.clear(); wr.clear();
    auto R = [&](const Ref& r) { if (!r.null()) rd.push_back(buf_id(b, r)); };
    auto Wt = [&](const Ref& r) { if (!r.null()) wr.push_back(buf_id(b, r)); };
    switch (op.kind) {
        case OP_GEMM: R(op.gemm.A); R(op.gemm.R); Wt(op.gemm.C); Wt(op.gemm.C2); break;
        case OP_AVGPOOL: R(op.pool.in); Wt(op.pool.out); break;
        case OP_LAYERNORM: R(op.ln.X); Wt(op.ln.Y); break;
        case OP_RELATTN: R(op.relattn.qkv); Wt(op.relattn.out); break;
        default: break;
    }
}

// Tile shape + split-K of a GEMM inside a chain of G CTAs: minimise a small cost model (cycles) over the
// eight tile variants of chain.h - waves x (K x max(FMA rate, L2->SM load rate) + per-k-tile and per-tile
// overheads), with an optional split-K round trip when even the smallest useful tile leaves CTAs idle.
void chain_schedule_gemm(PB& b, GemmOp& g, int G) {
    if (!(gemm_aligned(g) && g.K >= 64)) { g.ch_variant = -1; g.ch_splitk = 1; return; }
    // id (chain.h ChainTile), BM, BN, LK (k-quads across lanes), BK
    static const struct { int id, bm, bn, lk, bk; } V[8] = {{0, 32, 32, 1, 64},  {1, 16, 64, 1, 64},  {2, 8, 128, 1, 32}, {3, 32, 16, 2, 64},
                                                            {4, 32, 8, 4, 128},  {5, 8, 32, 1, 128},  {6, 8, 16, 2, 128}, {7, 16, 16, 2, 128}};
    double best = 1e30;
    for (const auto& v : V) {
        const int tm = (g.M + v.bm - 1) / v.bm, tn = (g.N + v.bn - 1) / v.bn, tiles = tm * tn * g.batch;
        const int nkt = (g.K + v.bk - 1) / v.bk;
        const int ct = v.bn / (32 / v.lk);
        // cycles per unit of K for one CTA: FMA issue (~60% of 128 lanes), shared-memory wavefronts (one 16-byte
        // broadcast per row + 4 per W column, 8 * LK quads in flight), L2 -> SM bytes at ~64 B/clk; + per-k-tile sync
        // (RVC_CHAIN_BW = assumed L2 -> SM bytes per clock of the cp.async ring, RVC_CHAIN_SK = cycles charged for a split-K round trip)
        static const double kBw = double(sched_env("RVC_CHAIN_BW", 64)), kSk = double(sched_env("RVC_CHAIN_SK", 4000));
        const double per_k = std::max(std::max(v.bm * v.bn / 77.0, 1.5 * (v.bm + 4.0 * ct) / (4.0 * v.lk)), (v.bm + v.bn) * 4 / kBw) + 150.0 / v.bk;
        for (int sk = 1; sk <= 8; sk *= 2) {
            if (sk > 1 && (tiles * sk > G || nkt / sk < 2)) break;
            const int waves = (tiles * sk + G - 1) / G;
            const double cost = waves * (double(nkt / sk + (nkt % sk ? 1 : 0)) * v.bk * per_k + 2500.0) + (sk > 1 ? kSk : 0.0);
            if (cost < best - 1e-9) {
                best = cost; g.ch_variant = v.id; g.ch_tiles_m = tm; g.ch_tiles_n = tn; g.ch_splitk = sk;
            }
        }
    }
    if (g.ch_splitk > 1) {
        g.ch_scratch = b.alloc("", int64_t(g.ch_splitk) * g.batch * g.M * g.N);
        g.ch_counters = b.alloc("", g.ch_tiles_m * g.ch_tiles_n * g.batch, true);
    }
}

void form_chains(PB& b, const PlanOptions& opt) {
    std::vector<Op>& ops = b.plan.ops;
    const int n = int(ops.size());
    std::vector<int64_t> rd, wr;
    int i = 0;
    while (i < n) {
        if (!chain_eligible(ops[i], opt.chain_side_max_m)) { ++i; continue; }
        int j = i;
        // side lanes: a run of skinny GEMMs of one width (RMVPE's bottleneck) ends where that shape ends, so that the whole
        // chain can run on the weight-streaming kernel (kernels_wstream.cu) instead of the generic tile engine
        const bool ws_on = sched_env("RVC_WSTREAM", 1) != 0;
        auto ws_sig = [&](const Op& o) {
            if (!ws_on || o.lane == 0 || o.kind != OP_GEMM) return -1;
            const GemmOp& g = o.gemm;
            return (g.M <= 8 && g.out_mode == OUT_PLAIN && g.seg_len >= g.K && g.N % 8 == 0 && g.K % 128 == 0 && g.K <= 1536 && g.batch == 1) ? g.N : -1;
        };
        const int sig0 = ws_sig(ops[i]);
        while (j < n && j - i < 256 && ops[j].lane == ops[i].lane && chain_eligible(ops[j], opt.chain_side_max_m) && ws_sig(ops[j]) == sig0) ++j;  // 256 = table size of the kernel
        const int G = ops[i].lane == 0 ? opt.chain_grid_main : opt.chain_grid_side;
        if (j - i >= 4 && G > 0) {
            ChainInfo c;
            c.first = i; c.count = j - i; c.lane = ops[i].lane; c.grid = G;
            std::vector<int64_t> cur_rd, cur_wr;
            int phase = 0;
            for (int k = i; k < j; ++k) {
                op_buffers(b, ops[k], rd, wr);
                bool conflict = false;
                for (int64_t x : rd) conflict |= std::find(cur_wr.begin(), cur_wr.end(), x) != cur_wr.end();
                for (int64_t x : wr)
                    conflict |= std::find(cur_wr.begin(), cur_wr.end(), x) != cur_wr.end() ||
                                std::find(cur_rd.begin(), cur_rd.end(), x) != cur_rd.end();
                if (conflict) { ++phase; cur_rd.clear(); cur_wr.clear(); }
                cur_rd.insert(cur_rd.end(), rd.begin(), rd.end());
                cur_wr.insert(cur_wr.end(), wr.begin(), wr.end());
                c.phase.push_back(phase);
                ops[k].chain = int(b.plan.chains.size());
                if (ops[k].kind == OP_GEMM) chain_schedule_gemm(b, ops[k].gemm, G);
            }
            c.n_phases = phase + 1;
            b.plan.chains.push_back(c);
        }
        i = j;
    }
}

}  // namespace

bool build_plan(PlanKind kind, const Geometry& g, const PlanOptions& opt, const Packed* cv, const CvInfo* cvi,
                const Packed* f0, const F0Info* f0i, const Packed* syn, const SynInfo* syi, Plan& plan,
                std::string& err) {
    plan = Plan{};
    PB b{plan, err};
    b.allow_umma = opt.allow_umma;
    b.f0_umma = opt.f0_umma;
    b.cv_stack = opt.cv_stack && opt.nb <= 1;
    b.fuse_cbr = opt.fuse_cbr;
    plan.params = Ref{SP_STATE, StateLayout::off_params};
    plan.cache = Ref{SP_STATE, StateLayout::off_cache};
    plan.pcm = Ref{SP_STATE, StateLayout::off_pcm};
    plan.audio = Ref{SP_STATE, StateLayout::off_audio};
    plan.n_lanes = 1;
    plan.nb = opt.nb > 1 ? opt.nb : 1;
    const int N = g.n16k;
    if (plan.nb > 1) {
        // compact per-window state block: params | pitch cache | pcm window | audio out
        if (kind != PLAN_INFER || !syi) { err = "batched plans exist for infer only"; return false; }
        if (plan.nb > 64) { err = "at most 64 windows per launch"; return false; }
        auto up = [](int64_t v) { return (v + 255) & ~int64_t(255); };
        const int64_t off_pcm = StateLayout::off_cache + StateLayout::CACHE_LEN * 4;
        const int64_t off_audio = off_pcm + up(int64_t(N > 0 ? N : 0) * 4);
        plan.pcm = Ref{SP_STATE, off_pcm};
        plan.audio = Ref{SP_STATE, off_audio};
        plan.state_block = off_audio + up(int64_t(g.return_length > 0 ? g.return_length : 0) * (syi->sr / 100) * 4);
    }
    if (kind != PLAN_KNN && (N <= 0 || N > StateLayout::PCM_CAP)) { err = "bad input length"; return false; }

    if (kind == PLAN_KNN) {
        // queries staged in the audio buffer region by the caller: [Q, C]; results -> work
        const int Q = g.n16k, C = g.sf16k, k = g.return_length, Nrows = opt.index_rows;
        if (Q <= 0 || C <= 0 || C % 4 != 0 || C > 1024 || k <= 0 || k > 16 || Nrows < k) { err = "bad kNN shape"; return false; }
        const bool um = opt.index_planes_off > 0 && knn_umma_ok(C, k);
        const int parts = um ? knn_umma_parts(Nrows) : knn_parts(Q, C, k, Nrows), kc = um ? KNN_UMMA_KC : k;
        Ref cd = b.alloc("knn_cand_d", int64_t(Q) * parts * kc), ci = b.alloc("knn_cand_i", int64_t(Q) * parts * kc, true);
        Ref idx = b.alloc("knn_idx", int64_t(Q) * k, true), d2 = b.alloc("knn_d2", int64_t(Q) * k);
        Op& a = b.add(OP_KNN_SCAN, "knn_scan");
        a.kd.index = Ref{SP_IDX, 0}; a.kd.queries = plan.audio; a.kd.ldq = C; a.kd.cand_d = cd; a.kd.cand_i = ci;
        a.kd.N = Nrows; a.kd.C = C; a.kd.Q = Q; a.kd.k = k; a.kd.parts = parts; a.kd.umma = um ? 1 : 0; a.kd.planes_off = opt.index_planes_off;
        Op& s = b.add(OP_KNN_SELECT, "knn_select");
        s.ks.cand_d = cd; s.ks.cand_i = ci; s.ks.idx = idx; s.ks.d2 = d2; s.ks.Q = Q; s.ks.k = k; s.ks.parts = parts;
        knn_rerank_setup(s.ks, um, a.kd, opt);
        plan.knn_q = Q;
        schedule_gemms(b);
    form_chains(b, opt);
        plan.work_bytes = (b.work + 256 + 255) & ~int64_t(255);
        return b.ok;
    }
    if (kind == PLAN_MEL) {
        if (!f0) { err = "f0 model not loaded"; return false; }
        if (N < 1024) { err = "input shorter than one FFT frame"; return false; }
        F0Out o = build_rmvpe(b, f0, *f0i, plan.pcm, N, plan.params, true, 0);
        plan.f0_T = o.T;
        schedule_gemms(b);
    form_chains(b, opt);
        plan.work_bytes = (b.work + 256 + 255) & ~int64_t(255);
        return b.ok;
    }
    if (kind == PLAN_HUBERT || kind == PLAN_FEATURE) {
        if (!cv) { err = "contentvec not loaded"; return false; }
        int T = 0;
        Ref x = build_contentvec(b, cv, *cvi, plan.pcm, N, T);
        plan.hubert_T = T; plan.hubert_C = cvi->out_dim;
        if (kind == PLAN_FEATURE && b.ok) {
            Op& op = b.add(OP_GATHER_ROWS, "feature");
            op.gather.src = x; op.gather.lds = cvi->out_dim; op.gather.out = plan.audio; op.gather.T = T;
            op.gather.C = cvi->out_dim; op.gather.skip = 0; op.gather.R = 2 * T + 1;
            if (int64_t(2 * T + 1) * cvi->out_dim > StateLayout::AUDIO_CAP) b.fail("feature too large");
        }
        schedule_gemms(b);
    form_chains(b, opt);
        plan.work_bytes = (b.work + 256 + 255) & ~int64_t(255);
        return b.ok;
    }
    // PLAN_PITCH / PLAN_INFER need the f0 window
    if (!f0) { err = "f0 model not loaded"; return false; }
    const int Lf0 = 5120 * ((g.sf16k + 800 - 1) / 5120 + 1) - 160;  // rmvpe.rs:256
    if (Lf0 > N) { err = "input shorter than the f0 window (rmvpe.rs:257)"; return false; }
    if (kind == PLAN_PITCH) {
        // RVC_PITCH_ML=1 (experiment): the pitch plan with the lane structure the F0 branch has inside an infer plan
        // (lane 1, shortcut convs on lane 3, side-lane chains) - its stand-alone time is what ContentVec's concurrency costs
        static const bool as_side = sched_env("RVC_PITCH_ML", 0) != 0;
        if (as_side && opt.multi_lane) { b.wait(0, 1); b.lane = 1; b.sc_lane = sched_env("RVC_SC_LANE", 1) != 0 ? 3 : -1; }
        F0Out o = build_rmvpe(b, f0, *f0i, plan.pcm.plus(N - Lf0), Lf0, plan.params, false, opt.upstream_cents_window);
        if (as_side && opt.multi_lane) { b.lane = 0; b.sc_lane = -1; b.wait(1, 0); }
        plan.f0_T = o.T;
        schedule_gemms(b);
    form_chains(b, opt);
        plan.work_bytes = (b.work + 256 + 255) & ~int64_t(255);
        return b.ok;
    }
    // ---- PLAN_INFER (rvc.rs:133-220) ---------------------------------------------------------
    if (!cv) { err = "contentvec not loaded"; return false; }
    if (!syn) { err = "model not loaded"; return false; }
    if (cvi->out_dim != syi->phone_dim) { err = "contentvec width does not match the voice model"; return false; }
    const int R = g.return_length, skip = g.skip_head;
    if (R <= 0) { err = "return_length must be positive"; return false; }
    const bool ml = opt.multi_lane;
    // lane 1 (+ lane 3 for the shortcut convs): F0 chain, concurrently with ContentVec on lane 0.  The fork is
    // recorded first, ContentVec is emitted before the F0 ops: graph nodes are created in op order and the lane
    // that is emitted first was observed to start first (the other way round ContentVec began ~0.5 ms late).
    if (ml) b.wait(0, 1);
    int T = 0;
    Ref x{};
    F0Out fo{};
    static const bool f0_first = sched_env("RVC_F0_FIRST", 0) != 0;
    auto emit_f0 = [&]() {
        static const bool sc_side = sched_env("RVC_SC_LANE", 1) != 0;   // 0: shortcut convs stay on the F0 lane (same chain phase as c1)
        static const int cv_gate = sched_env("RVC_CV_GATE", -1);
        if (ml) { b.lane = 1; b.sc_lane = sc_side ? 3 : -1; b.cv_gate = (f0_first && opt.nb <= 1) ? cv_gate : -1; }
        fo = build_rmvpe(b, f0, *f0i, plan.pcm.plus(N - Lf0), Lf0, plan.params, false, opt.upstream_cents_window);
        b.lane = 0; b.sc_lane = -1; b.cv_gate = -1;
    };
    // The lane whose ops are emitted (= whose graph nodes are created) first gets the SMs first when both lanes have
    // work ready.  With the weight-streaming bottleneck kernel the F0 branch ends ~170 us before ContentVec + retrieval:
    // ContentVec goes first again (RVC_F0_FIRST=1: the order of the rounds in which the F0 chain was the longer branch).
    if (f0_first) emit_f0();
    const size_t cv_first_op = b.plan.ops.size();
    x = build_contentvec(b, cv, *cvi, plan.pcm, N, T);
    if (!b.ok) return false;
    {
        // CTA budget of ContentVec's conv stem / pos-conv GEMMs beside the F0 lane (RVC_CV_WANT; 0 = scheduler default 96).
        // 80 while the F0 lane was the critical branch; since the weight-streaming bottleneck kernel ContentVec + retrieval
        // is, and 120 measures best (2.529 vs 2.539 ms; with ContentVec emitted first 2.504)
        static const int cv_want = sched_env("RVC_CV_WANT", 120);
        if (cv_want > 0 && ml && opt.nb <= 1)
            for (size_t i = cv_first_op; i < b.plan.ops.size(); ++i)
                if (b.plan.ops[i].kind == OP_GEMM && b.plan.ops[i].gemm.cta_budget == 0) b.plan.ops[i].gemm.cta_budget = cv_want;
    }
    if (!f0_first) emit_f0();
    plan.f0_T = fo.T;
    const int C = cvi->out_dim;
    plan.hubert_T = T; plan.hubert_C = C;
    const int ext = 2 * T + 1;
    const int hubert_length = std::min(N / 160, ext);              // rvc.rs:153
    if (skip + R > ext) { err = "skip_head + return_length exceeds the feature length (rvc.rs:155)"; return false; }
    if (hubert_length > 1024 || 1024 - hubert_length + skip + R > 1024) { err = "pitch cache range (rvc.rs:176)"; return false; }
    if (fo.T < 4 || 1024 + 4 - fo.T < 0) { err = "pitch length"; return false; }
    // retrieval on the 20 ms frames that survive the slice (rvc.rs:159 TODO; upstream semantics)
    const int first = std::min(skip / 2, T - 1), last = std::min((skip + R - 1) / 2, T - 1);
    const int Q = last - first + 1;
    Ref src = x; int row0 = 0;
    if (opt.with_index) {
        const int k = opt.index_k, Nrows = opt.index_rows;
        if (k <= 0 || k > 16 || Nrows < k || C % 4 != 0 || C > 1024) { err = "bad index / k"; return false; }
        if (opt.index_cols != C) {
            err = "retrieval index is " + std::to_string(opt.index_cols) + " wide, the ContentVec features are " + std::to_string(C);
            return false;
        }
        const bool um = opt.index_planes_off > 0 && knn_umma_ok(C, k);
        int parts = um ? knn_umma_parts(Nrows) : knn_parts(Q, C, k, Nrows);
        const int kc = um ? KNN_UMMA_KC : k;
        {
            // the scan runs beside the F0 lane (the longer branch): RVC_KNN_WPARTS caps its CTAs (0 = one or two per SM)
            static const int wparts = sched_env("RVC_KNN_WPARTS", 0);
            if (wparts > 0 && ml && !um && opt.nb <= 1 && wparts < parts) parts = wparts;
        }
        Ref cd = b.alloc("knn_cand_d", int64_t(Q) * parts * kc), ci = b.alloc("knn_cand_i", int64_t(Q) * parts * kc, true);
        Ref idx = b.alloc("knn_idx", int64_t(Q) * k, true), d2 = b.alloc("knn_d2", int64_t(Q) * k);
        Ref xb = b.alloc("knn_blend", int64_t(Q) * C);
        Op& a = b.add(OP_KNN_SCAN, "knn_scan");
        a.kd.index = Ref{SP_IDX, 0}; a.kd.queries = x.plus(int64_t(first) * C); a.kd.ldq = C; a.kd.cand_d = cd;
        a.kd.cand_i = ci; a.kd.N = Nrows; a.kd.C = C; a.kd.Q = Q; a.kd.k = k; a.kd.parts = parts; a.kd.umma = um ? 1 : 0; a.kd.planes_off = opt.index_planes_off;
        Op& s = b.add(OP_KNN_SELECT, "knn_select");
        s.ks.cand_d = cd; s.ks.cand_i = ci; s.ks.idx = idx; s.ks.d2 = d2; s.ks.Q = Q; s.ks.k = k; s.ks.parts = parts;
        knn_rerank_setup(s.ks, um, a.kd, opt);
        Op& m = b.add(OP_KNN_BLEND, "knn_blend");
        m.kb.index = Ref{SP_IDX, 0}; m.kb.idx = idx; m.kb.d2 = d2; m.kb.x = x.plus(int64_t(first) * C); m.kb.ldx = C;
        m.kb.out = xb; m.kb.params = plan.params; m.kb.C = C; m.kb.Q = Q; m.kb.k = k;
        src = xb; row0 = first;
        plan.knn_q = Q;
    }
    Ref phone = b.alloc("phone", int64_t(R) * C);
    {
        Op& op = b.add(OP_GATHER_ROWS, "phone");
        op.gather.src = src; op.gather.lds = C; op.gather.out = phone; op.gather.T = T; op.gather.C = C;
        op.gather.skip = skip; op.gather.R = R; op.gather.row0 = row0;
    }
    // enc_p's phone projection needs no pitch: it runs here, while the F0 lane is still busy, and the embedding op after the
    // join only adds the pitch embedding (RVC_EMB_PRE=0: one op after the join, as the reference orders it)
    Ref emb_pre{};
    static const bool emb_early = sched_env("RVC_EMB_PRE", 1) != 0;
    if (emb_early && ml) {
        emb_pre = b.alloc("sy.emb_pre", int64_t(R) * 192);
        b.gemm("sy.emb_pre", phone, C, C, 0, b.w(syn, SP_SYN, "emb.wp"), C, b.w(syn, SP_SYN, "emb.bp"), emb_pre, 192, R, 192, C, ACT_NONE);
    }
    if (ml) b.wait(1, 0);
    Ref pitch = b.alloc("pitch", R, true), pitchf = b.alloc("pitchf", R);
    {
        Op& op = b.add(OP_F0POST, "pitch");
        F0PostOp& p = op.f0p;
        p.f0 = fo.f0; p.cache = plan.cache; p.pitch = pitch; p.pitchf = pitchf; p.pitch_len = fo.T;
        p.shift = g.sf16k / 160; p.hubert_length = hubert_length; p.skip_head = skip; p.return_length = R;
        p.cache_len = 1024; p.sequential = opt.sequential ? 1 : 0;
        p.mel_min = std::log(50.0f / 700.0f + 1.0f) * 1127.0f;   // rvc.rs:31-34 (f32)
        p.mel_max = std::log(500.0f / 700.0f + 1.0f) * 1127.0f;
    }
    const int audio_len = R * (syi->sr / 100);
    if (audio_len > StateLayout::AUDIO_CAP) { err = "output too long"; return false; }
    build_synth(b, syn, *syi, phone, pitch, pitchf, plan.params, plan.audio, R, ml, emb_pre);
    plan.audio_len = audio_len;
    schedule_gemms(b, plan.nb);
    form_chains(b, opt);
    plan.work_bytes = (b.work + 256 + 255) & ~int64_t(255);
    return b.ok;
}

}  // namespace rvc
