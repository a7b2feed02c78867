"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol include/rvc_b200.h
declares; the plan builder + weight packing, replayed by the test-only CPU interpreter, match the
torch oracle; host mirror error mapping; oracle network bodies cross-checked against independent
implementations (torch.nn.GRU, torchaudio HuBERT)."""
import ctypes
import os
import re
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="session")
def built():
    import __graft_entry__ as ge
    ge.build()
    return os.path.join(ROOT, "obs-rvc_b200", "librvc_b200.so")


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "rvc_b200.h")).read()
    names = set(re.findall(r"\b(rvc_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 25
    L = ctypes.CDLL(built)
    for n in sorted(names):
        assert hasattr(L, n), f"{n} declared in include/rvc_b200.h but not exported"
    L.rvc_version.restype = ctypes.c_char_p
    assert b"sm_100a" in L.rvc_version()


def test_create_fails_loudly_without_gpu(built):
    """No CPU fallback: without a CUDA device the engine refuses to construct."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import rvc_b200
    with pytest.raises(rvc_b200.CudaError):
        rvc_b200.RvcInfer(tempfile.gettempdir())


def test_config_struct_layout(built):
    import rvc_b200
    L = rvc_b200.lib()
    cfg = rvc_b200.Config()
    L.rvc_config_default(ctypes.byref(cfg))
    assert (cfg.device, cfg.noise_mode, cfg.index_k, cfg.use_cuda_graph) == (0, 1, 8, 1)
    assert ctypes.sizeof(rvc_b200.Config) == 64   # matches sizeof(rvc_config)


@pytest.fixture(scope="session")
def data():
    from oracle import weights
    root = os.path.join(tempfile.gettempdir(), "rvc_b200_data_seed7")
    return weights.make_data_dir(root, seed=7, index_rows=40000)


def test_plan_on_cpu_matches_oracle(built, data):
    """Plan + packing + op semantics (shared with the CUDA engine) vs the torch oracle, two
    consecutive BASELINE windows with retrieval: waveform RMS error < 1e-5, integers exact."""
    import planexec
    from oracle import pipeline
    from oracle.weights import read_rvcw
    idx = read_rvcw(data["index"])["big_npy"]
    pe = planexec.PlanExec(data["data"])
    pe.load(0, data["contentvec"]); pe.load(1, data["f0"]); pe.load(2, data["model"])
    pe.set_index(idx, 0.5); pe.set_params(seed=0, noise_mode=1, index_k=8)
    ora = pipeline.RvcInfer(data["data"], noise_seed=0)
    ora.load_contentvec(2); ora.load_f0(1); ora.load_model(data["model"]); ora.set_index(idx, 0.5)
    g = pipeline.BASELINE_GEOM
    pcm = pipeline.synthetic_pcm(g["n16k"] + g["sf16k"] * 2)
    for w in range(2):
        x = pcm[w * g["sf16k"]: w * g["sf16k"] + g["n16k"]]
        pe.run(planexec.PLAN_INFER, x, g["sf16k"], 12, g["skip_head"], g["return_length"])
        want = ora.infer(x, g["sf16k"], 12, g["skip_head"], g["return_length"])
        got = pe.get("audio")
        assert np.sqrt(np.mean((got - want) ** 2)) < 1e-5
        np.testing.assert_array_equal(pe.get("f0_argmax", np.int32), ora.last["argmax"])
        np.testing.assert_array_equal(pe.get("pitch", np.int32), ora.last["pitch"])
        np.testing.assert_array_equal(pe.get("knn_idx", np.int32).reshape(-1, 8), ora.last["knn_idx"])
    d = pe.dims()
    assert (d["hubert_T"], d["hubert_C"], d["f0_T"], d["audio_len"], d["knn_q"]) == (111, 768, 32, 8400, 11)
    pe.close()


def test_chain_phases_are_order_independent(built, data):
    """Persistent chains (chain.h): plan.cpp groups runs of small same-lane ops into phases whose ops it claims are
    independent.  The CPU interpreter runs each phase in REVERSE plan order: the window must come out bit-identical
    to the plain in-order run, for the default grids and for a plan that chains the whole RMVPE U-Net."""
    import planexec
    from oracle import pipeline
    g = pipeline.BASELINE_GEOM
    x = pipeline.synthetic_pcm(g["n16k"], seed=3)
    outs = {}
    for name, cfg in (("off", None), ("default", (148, 32, 8)), ("wide", (96, 48, 0))):
        pe = planexec.PlanExec(data["data"])
        pe.load(0, data["contentvec"]); pe.load(1, data["f0"]); pe.load(2, data["model"])
        pe.set_params(seed=0, noise_mode=1, index_k=8)
        if cfg:
            pe.set_chain(*cfg)
        pe.run(planexec.PLAN_INFER, x, g["sf16k"], 12, g["skip_head"], g["return_length"])
        outs[name] = (pe.get("audio"), pe.get("f0_argmax", np.int32), pe.get("rm.salience"), pe.chain_stats())
        pe.close()
    assert outs["off"][3]["n_chains"] == 0
    st = outs["default"][3]
    assert st["n_chains"] == 3 and st["n_chain_ops"] >= 100 and st["n_phases"] <= st["n_chain_ops"], st
    assert outs["wide"][3]["n_chain_ops"] > st["n_chain_ops"] + 80, outs["wide"][3]      # the whole U-Net joined
    for name in ("default", "wide"):
        np.testing.assert_array_equal(outs[name][0], outs["off"][0])
        np.testing.assert_array_equal(outs[name][1], outs["off"][1])
        np.testing.assert_array_equal(outs[name][2], outs["off"][2])
    # the single-lane pitch plan keeps the shortcut convs on the F0 lane: c1 and its shortcut share a phase there
    sal = {}
    for name, cfg in (("off", None), ("chained", (148, 32, 0))):
        pe = planexec.PlanExec(data["data"])
        pe.load(1, data["f0"])
        if cfg:
            pe.set_chain(*cfg)
        pe.run(planexec.PLAN_PITCH, x, g["sf16k"], 12, 0, 0)
        sal[name] = (pe.get("rm.salience"), pe.get("f0"), pe.chain_stats())
        pe.close()
    assert sal["chained"][2]["widest_phase"] >= 2 and sal["chained"][2]["n_chain_ops"] > 120, sal["chained"][2]
    np.testing.assert_array_equal(sal["chained"][0], sal["off"][0])
    np.testing.assert_array_equal(sal["chained"][1], sal["off"][1])


def test_plan_shape_contract(built, data):
    """SURVEY 8b shape contract: violations are plan errors, not crashes."""
    import planexec
    pe = planexec.PlanExec(data["data"])
    pe.load(0, data["contentvec"]); pe.load(1, data["f0"]); pe.load(2, data["model"])
    x = np.zeros(35840, np.float32)
    with pytest.raises(RuntimeError):
        pe.run(planexec.PLAN_INFER, x, 2560, 0, 220, 21)      # rvc.rs:155
    with pytest.raises(RuntimeError):
        pe.run(planexec.PLAN_PITCH, x[:3000], 2560, 0, 0, 0)  # rmvpe.rs:257
    with pytest.raises(RuntimeError):
        pe.run(planexec.PLAN_HUBERT, x[:100], 0, 0, 0, 0)
    pe.close()


def test_gru_matches_torch():
    """oracle.nets.gru_direction vs torch.nn.GRU (pins the restated gate order)."""
    import torch
    from oracle import nets
    torch.manual_seed(0)
    gru = torch.nn.GRU(24, 16, batch_first=True, bidirectional=True)
    x = torch.randn(9, 24)
    with torch.no_grad():
        want = gru(x[None])[0][0]
        f = nets.gru_direction(x, gru.weight_ih_l0, gru.weight_hh_l0, gru.bias_ih_l0, gru.bias_hh_l0, False)
        b = nets.gru_direction(x, gru.weight_ih_l0_reverse, gru.weight_hh_l0_reverse, gru.bias_ih_l0_reverse,
                               gru.bias_hh_l0_reverse, True)
    assert torch.allclose(torch.cat([f, b], 1), want, atol=1e-6)


def test_hubert_matches_torchaudio_structure():
    """Independent structure check: oracle HuBERT == torchaudio.models.hubert_base with
    the same weights copied over (frame geometry 35840 -> 111 and numerics)."""
    torchaudio = pytest.importorskip("torchaudio")
    import torch
    from oracle import nets, weights
    w = weights.synth_contentvec(3)
    m = torchaudio.models.hubert_base().eval()
    sd = m.state_dict()
    tw = nets.to_torch(w)
    mp = {}
    for i in range(7):
        mp[f"feature_extractor.conv_layers.{i}.conv.weight"] = tw[f"feature_extractor.conv_layers.{i}.0.weight"]
    mp["feature_extractor.conv_layers.0.layer_norm.weight"] = tw["feature_extractor.conv_layers.0.2.weight"]
    mp["feature_extractor.conv_layers.0.layer_norm.bias"] = tw["feature_extractor.conv_layers.0.2.bias"]
    mp["encoder.feature_projection.layer_norm.weight"] = tw["layer_norm.weight"]
    mp["encoder.feature_projection.layer_norm.bias"] = tw["layer_norm.bias"]
    mp["encoder.feature_projection.projection.weight"] = tw["post_extract_proj.weight"]
    mp["encoder.feature_projection.projection.bias"] = tw["post_extract_proj.bias"]
    wv = tw["encoder.pos_conv.0.weight"]
    mp["encoder.transformer.pos_conv_embed.conv.bias"] = tw["encoder.pos_conv.0.bias"]
    mp["encoder.transformer.layer_norm.weight"] = tw["encoder.layer_norm.weight"]
    mp["encoder.transformer.layer_norm.bias"] = tw["encoder.layer_norm.bias"]
    for i in range(12):
        s, d = f"encoder.layers.{i}.", f"encoder.transformer.layers.{i}."
        for a, b in (("self_attn.q_proj", "attention.q_proj"), ("self_attn.k_proj", "attention.k_proj"),
                     ("self_attn.v_proj", "attention.v_proj"), ("self_attn.out_proj", "attention.out_proj"),
                     ("self_attn_layer_norm", "layer_norm"), ("fc1", "feed_forward.intermediate_dense"),
                     ("fc2", "feed_forward.output_dense"), ("final_layer_norm", "final_layer_norm")):
            mp[d + b + ".weight"] = tw[s + a + ".weight"]
            mp[d + b + ".bias"] = tw[s + a + ".bias"]
    # torchaudio keeps pos_conv under weight norm (dim=2): w = g * v / |v|; g = |v| makes w == v
    g = wv.norm(dim=(0, 1), keepdim=True)
    for k in sd:
        if "pos_conv_embed.conv" not in k or k.endswith("bias"):
            continue
        if k.endswith("original0") or k.endswith("weight_g"):
            mp[k] = g
        elif k.endswith("original1") or k.endswith("weight_v"):
            mp[k] = wv
    missing = [k for k in sd if k not in mp]
    assert not missing, missing
    m.load_state_dict(mp)
    x = torch.from_numpy(np.random.default_rng(0).standard_normal(35840).astype(np.float32) * 0.1)
    with torch.no_grad():
        want = m.extract_features(x[None])[0][-1][0]
    got = nets.hubert_forward(tw, x)
    assert got.shape == want.shape == (111, 768)
    assert float((got - want).abs().max()) < 2e-4


def test_rust_sys_mirror_is_complete():
    """Seam 1/2 boundary artefacts: the raw Rust FFI crate is generated from include/rvc_b200.h (tools/gen_rust_sys.py)
    and mirrors EVERY declared entry point; the Seam-2 adapter keeps the reference adapter's public surface
    (obs-rvc/src/rvcadapter.rs:33-67,122-126)."""
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert subprocess.run([sys.executable, os.path.join(root, "tools", "gen_rust_sys.py"), "--check"]).returncode == 0, \
        "rvc-cuda-sys/src/lib.rs is stale: run python tools/gen_rust_sys.py"
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "rvc_b200.h")).read(), flags=re.S)
    declared = set(re.findall(r"\b(rvc_[a-z0-9_]+)\s*\(", hdr))
    rust = open(os.path.join(root, "obs-rvc_b200", "rust", "rvc-cuda-sys", "src", "lib.rs")).read()
    mirrored = set(re.findall(r"pub fn (rvc_[a-z0-9_]+)\(", rust))
    assert declared and declared == mirrored, sorted(declared ^ mirrored)
    adapter = open(os.path.join(root, "obs-rvc_b200", "rust", "obs-rvc-adapter", "src", "rvcadapter.rs")).read()
    for item in ("pub struct RvcInfer", "pub enum RvcAdapterError", "RvcInferError(RvcInferError)", "IoError(std::io::Error)",
                 "pub fn new(_binary_path: PathBuf, model_version: RvcModelVersion, pitch_algorithm: PitchAlgorithm, model_path: PathBuf",
                 "pub fn infer(&mut self, input: ndarray::ArrayView1<f32>, sample_frame_16k_size: usize, pitch_shift: i32, skip_head: u32",
                 "impl Drop for RvcInfer"):
        assert item in adapter, item


def test_generator_body_equals_transformers_hifigan():
    """Independent structural check of the synthesizer's vocoder body (the oracle's network bodies are
    unpinned): oracle/nets.py `generator_nsf` with the NSF additions switched off (noise convs and speaker conditioning
    zeroed) must be the plain HiFi-GAN v1 generator - checked against `transformers.SpeechT5HifiGan`, an implementation
    written by other people, with the same weights: conv_pre, 4 x (leaky_relu 0.1, ConvTranspose1d, three ResBlocks of
    three (dilated conv, conv) pairs averaged), leaky_relu 0.01, conv_post, tanh."""
    import torch
    transformers = pytest.importorskip("transformers")
    from oracle import nets, weights
    torch.manual_seed(0)
    w = nets.to_torch(weights.synth_voice(3456, 40000, 768))
    for i in range(4):
        w[f"dec.noise_convs.{i}.weight"] = torch.zeros_like(w[f"dec.noise_convs.{i}.weight"])
        w[f"dec.noise_convs.{i}.bias"] = torch.zeros_like(w[f"dec.noise_convs.{i}.bias"])
    w["dec.cond.weight"] = torch.zeros_like(w["dec.cond.weight"])
    w["dec.cond.bias"] = torch.zeros_like(w["dec.cond.bias"])
    cfg = transformers.SpeechT5HifiGanConfig(model_in_dim=192, sampling_rate=40000, upsample_initial_channel=512,
                                             upsample_rates=[10, 10, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4],
                                             resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5]] * 3,
                                             leaky_relu_slope=0.1, normalize_before=False)
    m = transformers.SpeechT5HifiGan(cfg).eval()
    sd = {"conv_pre.weight": w["dec.conv_pre.weight"], "conv_pre.bias": w["dec.conv_pre.bias"],
          "conv_post.weight": w["dec.conv_post.weight"], "conv_post.bias": torch.zeros(1),
          "mean": torch.zeros(192), "scale": torch.ones(192)}
    for i in range(4):
        sd[f"upsampler.{i}.weight"] = w[f"dec.ups.{i}.weight"]
        sd[f"upsampler.{i}.bias"] = w[f"dec.ups.{i}.bias"]
    for r in range(12):
        for d in range(3):
            for c in ("convs1", "convs2"):
                for t in ("weight", "bias"):
                    sd[f"resblocks.{r}.{c}.{d}.{t}"] = w[f"dec.resblocks.{r}.{c}.{d}.{t}"]
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not [k for k in missing if not k.startswith(("mean", "scale"))] and not unexpected, (missing, unexpected)
    T = 21
    z = torch.randn(1, 192, T)
    with torch.no_grad():
        want = m(z.transpose(1, 2)[0])                                    # (T, 192) -> waveform
        got = nets.generator_nsf(w, z, torch.full((T,), 220.0), torch.zeros(1, 256, 1), 40000, torch.zeros(T * 400))
    assert got.shape == want.shape == (T * 400,)
    assert float((got - want).abs().max()) < 2e-5 and float(want.abs().max()) > 0.01


def _wn_param(sd, prefix, w):
    """plain conv weight -> HF's weight_norm parametrisation (g = ||v|| per output channel, v = w)."""
    sd[prefix + ".parametrizations.weight.original1"] = w
    sd[prefix + ".parametrizations.weight.original0"] = w.flatten(1).norm(dim=1).reshape(-1, 1, 1)


def test_text_encoder_and_flow_equal_transformers_vits():
    """Independent structural check of enc_p (6 relative-position attention layers, window 10, FFN kernel 3) and of the
    reverse flow (4 mean-only residual coupling layers with a 3-layer WaveNet each, Flip between them): the oracle's
    functional restatement against `transformers`' VITS modules (VitsEncoder, VitsResidualCouplingBlock) carrying the
    same weights.  The RVC-specific parts around them (phone / pitch embedding, the projection to m / logs) are plain
    linear algebra and stay outside the comparison."""
    import torch
    pytest.importorskip("transformers")
    from transformers import VitsConfig
    from transformers.models.vits import modeling_vits as mv
    from oracle import nets, weights
    torch.manual_seed(1)
    w = nets.to_torch(weights.synth_voice(3456, 40000, 768))
    cfg = VitsConfig(hidden_size=192, num_hidden_layers=6, num_attention_heads=2, window_size=10, ffn_dim=768, ffn_kernel_size=3,
                     hidden_act="relu", layerdrop=0.0, hidden_dropout=0.0, attention_dropout=0.0, activation_dropout=0.0,
                     use_bias=True, flow_size=192, prior_encoder_num_flows=4, prior_encoder_num_wavenet_layers=3,
                     wavenet_kernel_size=5, wavenet_dilation_rate=1, speaker_embedding_size=256, num_speakers=2, layer_norm_eps=1e-5)
    # ---- encoder
    enc = mv.VitsEncoder(cfg).eval()
    sd = {}
    for i in range(6):
        a, h = f"enc_p.encoder.attn_layers.{i}.", f"layers.{i}.attention."
        for o, n in (("conv_q", "q_proj"), ("conv_k", "k_proj"), ("conv_v", "v_proj"), ("conv_o", "out_proj")):
            sd[h + n + ".weight"] = w[a + o + ".weight"][:, :, 0]
            sd[h + n + ".bias"] = w[a + o + ".bias"]
        sd[h + "emb_rel_k"] = w[a + "emb_rel_k"]; sd[h + "emb_rel_v"] = w[a + "emb_rel_v"]
        sd[f"layers.{i}.layer_norm.weight"] = w[f"enc_p.encoder.norm_layers_1.{i}.gamma"]
        sd[f"layers.{i}.layer_norm.bias"] = w[f"enc_p.encoder.norm_layers_1.{i}.beta"]
        sd[f"layers.{i}.final_layer_norm.weight"] = w[f"enc_p.encoder.norm_layers_2.{i}.gamma"]
        sd[f"layers.{i}.final_layer_norm.bias"] = w[f"enc_p.encoder.norm_layers_2.{i}.beta"]
        for c in ("conv_1", "conv_2"):
            sd[f"layers.{i}.feed_forward.{c}.weight"] = w[f"enc_p.encoder.ffn_layers.{i}.{c}.weight"]
            sd[f"layers.{i}.feed_forward.{c}.bias"] = w[f"enc_p.encoder.ffn_layers.{i}.{c}.bias"]
    missing, unexpected = enc.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    T = 21
    phone = torch.randn(T, 768) * 0.5
    pitch = torch.randint(1, 255, (T,))
    rec = {}
    with torch.no_grad():
        m_p, logs_p = nets.text_encoder(w, phone, pitch, rec)
        x0 = torch.from_numpy(rec["sy.emb"])[None] if not torch.is_tensor(rec["sy.emb"]) else rec["sy.emb"][None]   # (1, T, 192): the encoder's input
        hf = enc(x0.float(), torch.ones(1, T, 1), return_dict=True).last_hidden_state
        last = rec["sy.enc5"]
        last = torch.from_numpy(last) if not torch.is_tensor(last) else last
    assert float((hf[0] - last.float()).abs().max()) < 2e-5 and float(last.abs().max()) > 0.1
    # ---- flow (reverse)
    flow = mv.VitsResidualCouplingBlock(cfg).eval()
    sd = {}
    for f in range(4):
        p, h = f"flow.flows.{2 * f}.", f"flows.{f}."
        sd[h + "conv_pre.weight"] = w[p + "pre.weight"]; sd[h + "conv_pre.bias"] = w[p + "pre.bias"]
        sd[h + "conv_post.weight"] = w[p + "post.weight"]; sd[h + "conv_post.bias"] = w[p + "post.bias"]
        for i in range(3):
            _wn_param(sd, h + f"wavenet.in_layers.{i}", w[p + f"enc.in_layers.{i}.weight"])
            sd[h + f"wavenet.in_layers.{i}.bias"] = w[p + f"enc.in_layers.{i}.bias"]
            _wn_param(sd, h + f"wavenet.res_skip_layers.{i}", w[p + f"enc.res_skip_layers.{i}.weight"])
            sd[h + f"wavenet.res_skip_layers.{i}.bias"] = w[p + f"enc.res_skip_layers.{i}.bias"]
        _wn_param(sd, h + "wavenet.cond_layer", w[p + "enc.cond_layer.weight"])
        sd[h + "wavenet.cond_layer.bias"] = w[p + "enc.cond_layer.bias"]
    missing, unexpected = flow.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    z = torch.randn(1, 192, T)
    g = torch.randn(1, 256, 1)
    with torch.no_grad():
        want = flow(z, torch.ones(1, 1, T), global_conditioning=g, reverse=True)
        got = nets.flow_reverse(w, z, g)
    assert float((got - want).abs().max()) < 2e-5 and float(want.abs().max()) > 0.1


def test_rmvpe_body_equals_module_statement():
    """Independent structural check of the RMVPE network: the oracle's functional restatement (oracle/nets.py
    rmvpe_forward) against a module-based statement laid out like the published model code (tests/rmvpe_modules.py),
    loaded with the same weights under strict=True - the state_dict keys must be exactly the checkpoint's keys - in
    eval mode (BatchNorm running statistics, dropout off)."""
    import sys
    import torch
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import rmvpe_modules
    from oracle import nets, weights
    w = weights.synth_rmvpe(9)
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in w.items() if not k.startswith(("meta.", "window", "mel."))}
    model = rmvpe_modules.E2E().eval()
    own = model.state_dict()
    for k in own:
        if k.endswith("num_batches_tracked"):
            sd[k] = own[k]
    extra = sorted(set(sd) - set(own))
    assert not extra or all(not k.startswith(("unet.", "cnn.", "fc.")) for k in extra), extra[:5]
    model.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=True)
    torch.manual_seed(3)
    mel = torch.randn(128, 32) * 2.0 - 4.0
    with torch.no_grad():
        want = model(mel[None])[0]
        got = nets.rmvpe_forward(nets.to_torch(w), mel)
    assert got.shape == want.shape == (32, 360)
    assert float((got - want).abs().max()) < 1e-5 and float(want.max()) > 0.01
    assert torch.equal(got.argmax(dim=1), want.argmax(dim=1))
