"""GPU parity tests: the CUDA engine, called through the C ABI, against the CPU oracle on the same
seeded inputs (bit-exact for F0 argmax / coarse pitch / kNN indices; waveform within 1e-3 RMS -
BASELINE.json north_star).  Marked `gpu`: run on the B200 box (`pytest -m gpu`)."""
import os
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
WAVE_RMS_TOL = 1e-3      # north_star: output waveform within 1e-3 RMS
FEATS_ABS_TOL = 2e-3     # rvc/src/tests/hubert.rs:18 `epsilon = 2e-3`


@pytest.fixture(scope="module")
def env():
    import rvc_b200
    from oracle import pipeline, weights
    from oracle.weights import read_rvcw
    root = os.path.join(tempfile.gettempdir(), "rvc_b200_data_seed7")
    paths = weights.make_data_dir(root, seed=7, index_rows=40000)
    index = read_rvcw(paths["index"])["big_npy"]
    eng = rvc_b200.RvcInfer(root, noise_seed=0)
    eng.load_contentvec(2); eng.load_f0(1); eng.load_model(paths["model"])
    ora = pipeline.RvcInfer(root, noise_seed=0)
    ora.load_contentvec(2); ora.load_f0(1); ora.load_model(paths["model"])
    yield dict(eng=eng, ora=ora, paths=paths, index=index, pipeline=pipeline, rvc_b200=rvc_b200)
    eng.close()


def _rms(a):
    return float(np.sqrt(np.mean(np.asarray(a, np.float64) ** 2)))


KNN_TIE_REL = 2e-5   # fp32 accumulation noise of a 768-term squared distance (SURVEY section 7 "kNN bit-exact")


def assert_topk_exact_up_to_ties(got_idx, want_idx, want_d2, queries, index):
    """Top-k indices must be bit-exact wherever the float64 oracle separates neighbours by more than
    KNN_TIE_REL (relative, on d^2); inside such a near-tie the engine may return either member, but it
    must still be a true top-k candidate at that rank."""
    got_idx = np.asarray(got_idx).reshape(want_idx.shape)
    for q in range(want_idx.shape[0]):
        for j in np.nonzero(got_idx[q] != want_idx[q])[0]:
            d_got = float(((queries[q].astype(np.float64) - index[got_idx[q, j]].astype(np.float64)) ** 2).sum())
            assert abs(d_got - want_d2[q, j]) <= KNN_TIE_REL * want_d2[q, j], (
                f"query {q} rank {j}: got row {got_idx[q, j]} (d2={d_got}) vs oracle row {want_idx[q, j]} (d2={want_d2[q, j]})")
        assert len(set(got_idx[q].tolist())) == got_idx.shape[1]


def test_native_library_loaded(env):
    """The product path is the CUDA .so, not a fallback."""
    assert os.path.exists(env["rvc_b200"].LIB_PATH)
    assert env["eng"].cuda_stream() != 0


@pytest.mark.parametrize("src", ["synthetic", "input_wav2", "jfk"])
def test_mel_extract(env, src):
    """MelSpectrogram::mel_extract (rmvpe.rs:159-205) on synthetic and real audio."""
    if src == "synthetic":
        x = env["pipeline"].synthetic_pcm(4960, seed=3)
    elif src == "input_wav2":
        x = np.load(os.path.join(GOLDEN, "input_wav2.npy"))[-10080:]
    else:
        x = np.load(os.path.join(GOLDEN, "jfk_2p5s.npy"))[:4960]
    got = env["eng"].mel_extract(x)
    want = env["ora"].mel_extract(x)
    assert got.shape == want.shape == (128, 1 + len(x) // 160)
    # ln() amplifies error only where the band energy sits near the 1e-5 clamp: compare in the
    # linear domain everywhere and in the log domain away from the clamp
    assert np.abs(np.exp(got) - np.exp(want)).max() < 2e-5 * max(1.0, float(np.exp(want).max()))
    assert np.abs(got - want)[want > -9].max() < 2e-3


def test_hubert_geometry_and_values(env):
    """rvc/src/tests/hubert.rs:10-19: 38240 samples -> (1, 239, 768); values vs the oracle at the
    reference's own tolerance (2e-3 abs)."""
    x = np.load(os.path.join(GOLDEN, "input_wav2.npy"))
    h = env["eng"].hubert(x)
    assert h.shape == (1, 768, 119)
    want = env["ora"].hubert(x)
    assert np.abs(h - want).max() < FEATS_ABS_TOL
    f = env["eng"].extract_feature(x)
    assert f.shape == (1, 239, 768)
    assert np.abs(f - env["ora"].extract_feature(x)).max() < FEATS_ABS_TOL
    np.testing.assert_array_equal(f[0, -1], f[0, -3])      # rvc.rs:101-108: last frame three times


@pytest.mark.parametrize("shift,sf", [(13, 4800), (0, 2560), (-12, 2560)])
def test_pitch(env, shift, sf):
    """rvc/src/tests/pitch.rs:18-30 call shape (input_wav2.npy, 13, 4800); argmax bit-exact."""
    x = np.load(os.path.join(GOLDEN, "input_wav2.npy"))
    got = env["eng"].pitch(x, shift, sf)
    want = env["ora"].pitch(x, shift, sf)
    assert got.shape == want.shape
    np.testing.assert_array_equal(env["eng"].get_last("f0_argmax", np.int32), env["ora"].last["argmax"])
    assert np.abs(got - want).max() <= 1e-3 * max(1.0, float(np.abs(want).max()))


@pytest.mark.parametrize("geom_name,index_rate", [("BASELINE_GEOM", 0.0), ("BASELINE_GEOM", 0.5), ("DEFAULT_GEOM", 0.5)])
def test_infer_stream(env, geom_name, index_rate):
    """RvcInfer::infer (rvc.rs:133-220) over consecutive windows of one stream: pitch cache
    continuity, retrieval, synthesizer."""
    eng, ora = env["eng"], env["ora"]
    g = getattr(env["pipeline"], geom_name)
    eng.set_index(env["index"], index_rate)
    ora.set_index(env["index"], index_rate)
    eng.reset_state()
    ora.cache.buf[:] = 0
    ora.window = 0
    nwin = 4
    pcm = env["pipeline"].synthetic_pcm(g["n16k"] + g["sf16k"] * nwin, seed=1)
    for w in range(nwin):
        x = pcm[w * g["sf16k"]: w * g["sf16k"] + g["n16k"]]
        got = eng.infer(x, g["sf16k"], 12, g["skip_head"], g["return_length"]).copy()
        want = ora.infer(x, g["sf16k"], 12, g["skip_head"], g["return_length"])
        assert got.shape == want.shape == (g["return_length"] * 400,)
        np.testing.assert_array_equal(eng.get_last("f0_argmax", np.int32), ora.last["argmax"])
        np.testing.assert_array_equal(eng.get_last("pitch", np.int32), ora.last["pitch"])
        if index_rate > 0:
            assert_topk_exact_up_to_ties(eng.get_last("knn_idx", np.int32), ora.last["knn_idx"], ora.last["knn_d2"],
                                         ora.last["knn_q"], env["index"])
        assert np.abs(eng.get_last("phone").reshape(ora.last["phone"].shape) - ora.last["phone"]).max() < FEATS_ABS_TOL
        err = _rms(got - want)
        assert err < WAVE_RMS_TOL, f"window {w}: waveform RMS error {err}"
        assert _rms(want) > 0.05          # the comparison is not vacuous
    eng.set_index(None, 0.0)
    ora.set_index(None, 0.0)


def test_graph_replay_equals_eager(env):
    """Call 1 runs eagerly, later calls replay the captured CUDA graph: same inputs, same bits."""
    rb = env["rvc_b200"]
    g = env["pipeline"].BASELINE_GEOM
    x = env["pipeline"].synthetic_pcm(g["n16k"], seed=5)
    outs = []
    for use_graph in (False, True):
        e = rb.RvcInfer(env["paths"]["data"], noise_seed=3, use_cuda_graph=use_graph)
        e.load_contentvec(2); e.load_f0(1); e.load_model(env["paths"]["model"])
        seq = []
        for _ in range(3):
            e.reset_state()
            seq.append(e.infer(x, g["sf16k"], 0, g["skip_head"], g["return_length"]).copy())
        assert e.plan_info()["graph"] == int(use_graph)
        np.testing.assert_array_equal(seq[0], seq[1])
        np.testing.assert_array_equal(seq[1], seq[2])
        outs.append(seq[2])
        e.close()
    np.testing.assert_array_equal(outs[0], outs[1])


def test_persistent_chains_match_separate_kernels(env, monkeypatch):
    """The persistent chain kernel (chain.h: enc_p, flow, RMVPE bottleneck in one cooperative launch each) and the
    same ops as separate kernels give the same window: F0 argmax / coarse pitch identical, audio equal up to the
    order of fp32 partial sums.  RVC_CHAIN=2 forces chains although the fixture's context shares the device."""
    rb = env["rvc_b200"]
    g = env["pipeline"].BASELINE_GEOM
    x = env["pipeline"].synthetic_pcm(g["n16k"] + g["sf16k"], seed=11)
    res = {}
    for mode in ("2", "0"):
        monkeypatch.setenv("RVC_CHAIN", mode)
        e = rb.RvcInfer(env["paths"]["data"], noise_seed=5)
        e.load_contentvec(2); e.load_f0(1); e.load_model(env["paths"]["model"])
        outs = [e.infer(x[w * g["sf16k"]: w * g["sf16k"] + g["n16k"]], g["sf16k"], 12, g["skip_head"], g["return_length"]).copy()
                for w in range(2)]
        chains = e.profile_chains()
        res[mode] = dict(audio=outs, argmax=e.get_last("f0_argmax", np.int32).copy(), pitch=e.get_last("pitch", np.int32).copy(),
                         n_chains=len(chains), phases=sum(len(c["phases"]) for c in chains), launches=e.plan_info().get("launches", 0))
        e.close()
    monkeypatch.delenv("RVC_CHAIN")
    assert res["2"]["n_chains"] >= 3 and res["2"]["phases"] >= 100, res["2"]
    assert res["0"]["n_chains"] == 0
    np.testing.assert_array_equal(res["2"]["argmax"], res["0"]["argmax"])
    np.testing.assert_array_equal(res["2"]["pitch"], res["0"]["pitch"])
    for a, b in zip(res["2"]["audio"], res["0"]["audio"]):
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-4)
        assert _rms(a) > 0.01


def test_two_devices_one_process(env):
    """Contexts on different GPUs of one process are independent (kernel attributes are set per device, weights and
    state live on the context's device): the same window on cuda:0 and cuda:1 gives the same integers and the same
    audio up to fp32 summation order (the second context is not alone in the process, so it runs without chains)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rb = env["rvc_b200"]
    g = env["pipeline"].BASELINE_GEOM
    x = env["pipeline"].synthetic_pcm(g["n16k"], seed=13)
    res = []
    for dev in (0, 1):
        e = rb.RvcInfer(env["paths"]["data"], device=dev, noise_seed=2)
        e.load_contentvec(2); e.load_f0(1); e.load_model(env["paths"]["model"])
        a = e.infer(x, g["sf16k"], 12, g["skip_head"], g["return_length"]).copy()
        res.append((a, e.get_last("f0_argmax", np.int32).copy(), e.get_last("pitch", np.int32).copy()))
        e.close()
    np.testing.assert_array_equal(res[0][1], res[1][1])
    np.testing.assert_array_equal(res[0][2], res[1][2])
    np.testing.assert_allclose(res[0][0], res[1][0], rtol=0, atol=1e-4)
    assert _rms(res[0][0]) > 0.01


def test_noise_modes_and_pitch_shift_quirk(env):
    """noise_mode 0 = deterministic zeros; pitch shift is integer octaves (rvc.rs:121)."""
    rb = env["rvc_b200"]
    g = env["pipeline"].BASELINE_GEOM
    x = env["pipeline"].synthetic_pcm(g["n16k"], seed=6)
    e = rb.RvcInfer(env["paths"]["data"], noise_mode=0)
    e.load_contentvec(2); e.load_f0(1); e.load_model(env["paths"]["model"])
    f7 = e.pitch(x, 7, g["sf16k"]); f0 = e.pitch(x, 0, g["sf16k"]); f12 = e.pitch(x, 12, g["sf16k"])
    np.testing.assert_array_equal(f7, f0)
    np.testing.assert_allclose(f12, 2.0 * f0, rtol=0, atol=0)
    ora = env["pipeline"].RvcInfer(env["paths"]["data"])
    ora.load_contentvec(2); ora.load_f0(1); ora.load_model(env["paths"]["model"])
    ora.noise_enabled = False
    got = e.infer(x, g["sf16k"], 12, g["skip_head"], g["return_length"]).copy()
    want = ora.infer(x, g["sf16k"], 12, g["skip_head"], g["return_length"])
    assert _rms(got - want) < WAVE_RMS_TOL
    e.close()


def test_error_behaviour(env):
    """rvc-common/src/errors.rs + SURVEY 8b 'Shape contract': errors instead of panics."""
    rb = env["rvc_b200"]
    g = env["pipeline"].BASELINE_GEOM
    x = env["pipeline"].synthetic_pcm(g["n16k"], seed=7)
    e = rb.RvcInfer(env["paths"]["data"])
    with pytest.raises(rb.ModelNotLoaded):
        e.infer(x, g["sf16k"], 0, g["skip_head"], g["return_length"])
    with pytest.raises(rb.ContentvecNotLoaded):
        e.hubert(x)
    with pytest.raises(rb.F0NotLoaded):
        e.pitch(x, 0, g["sf16k"])
    with pytest.raises(rb.IoError):
        e.load_model("/nonexistent/voice.rvcw")
    e.load_contentvec(2); e.load_f0(1); e.load_model(env["paths"]["model"])
    with pytest.raises(rb.BadShape):          # skip_head + return_length > 2T+1 (rvc.rs:155 panics)
        e.infer(x, g["sf16k"], 0, 220, 21)
    with pytest.raises(rb.BadShape):          # input shorter than the f0 window (rmvpe.rs:257 panics)
        e.pitch(x[:3000], 0, g["sf16k"])
    e.unload_model()
    with pytest.raises(rb.ModelNotLoaded):
        e.infer(x, g["sf16k"], 0, g["skip_head"], g["return_length"])
    e.close()
    bad = rb.RvcInfer("/nonexistent-data-dir")
    with pytest.raises(rb.IoError):
        bad.load_contentvec(2)
    bad.close()


@pytest.mark.parametrize("n,c,q,k", [(40000, 768, 11, 8), (100000, 256, 128, 4), (5000, 256, 1, 8), (3000, 512, 40, 16),
                                     (40000, 256, 11, 8), (1000, 64, 3, 4), (20, 128, 5, 8)])
def test_knn_search_exact(env, n, c, q, k):
    """Brute-force L2 top-k: indices AND fp32 distances literally equal to the fp32 oracle that sums in the kernels'
    defined order (oracle/knn.py search_f32_ordered) - for the fp32 scan (C = 768, 512) and for the tcgen05 candidate
    pass + exact re-rank (C = 64 .. 256, k <= 8); also consistent with the float64 oracle up to fp32 near-ties."""
    from oracle import knn
    from oracle.weights import synth_index
    rb = env["rvc_b200"]
    rows = synth_index(11, n, c)
    rng = np.random.default_rng(5)
    queries = (rows[rng.integers(0, n, q)] + rng.standard_normal((q, c)).astype(np.float32) * 0.2).astype(np.float32)
    e = rb.RvcInfer(env["paths"]["data"], index_k=min(k, 8))
    e.set_index(rows, 0.0)
    d2, idx = e.knn_search(queries, k)
    od, oi = knn.search_f32_ordered(rows, queries, k)
    np.testing.assert_array_equal(idx, oi)
    np.testing.assert_array_equal(d2, od)
    wd, wi = knn.search(rows, queries, k)
    assert_topk_exact_up_to_ties(idx, wi, wd, queries, rows)
    if n >= 1000:
        assert e.knn_fallbacks() == 0     # the guard of the candidate pass held for every query (tiny indices may legitimately fall back)
    e.close()


def test_knn_search_1m_rows(env):
    """BASELINE configs[4] at its full size: 1 048 576 x 256 index, 128 queries, top-4 on one GPU (tcgen05 candidate
    pass + exact re-rank): literal equality with the ordered fp32 oracle on every query, no guard fallback."""
    from oracle import knn
    rb = env["rvc_b200"]
    n, c, q, k = 1 << 20, 256, 128, 4
    rng = np.random.default_rng(2)
    rows = rng.standard_normal((n, c), dtype=np.float32) * np.float32(0.34)
    queries = rng.standard_normal((q, c), dtype=np.float32) * np.float32(0.34)
    queries[:16] = rows[rng.integers(0, n, 16)] + rng.standard_normal((16, c), dtype=np.float32) * np.float32(0.05)   # some with a close hit
    e = rb.RvcInfer(env["paths"]["data"], index_k=k)
    e.set_index(rows, 0.0)
    d2, idx = e.knn_search(queries, k)
    od, oi = knn.search_f32_ordered(rows, queries, k, shortlist=32)
    np.testing.assert_array_equal(idx, oi)
    np.testing.assert_array_equal(d2, od)
    assert e.knn_fallbacks() == 0
    e.close()


def test_knn_guard_fallback_is_exact(env):
    """Near-duplicate rows defeat the candidate pass's guard (more rows inside the error bound than candidate slots):
    the re-rank kernel must notice and fall back to the exact scan - result still literally equal to the oracle."""
    from oracle import knn
    rb = env["rvc_b200"]
    n, c, k = 4096, 128, 8
    rng = np.random.default_rng(9)
    base = rng.standard_normal(c).astype(np.float32)
    rows = (base[None, :] + rng.standard_normal((n, c)).astype(np.float32) * np.float32(1e-5)).astype(np.float32)   # 4096 rows within fp16 noise of each other
    queries = (base[None, :] + rng.standard_normal((3, c)).astype(np.float32) * np.float32(0.5)).astype(np.float32)
    e = rb.RvcInfer(env["paths"]["data"], index_k=k)
    e.set_index(rows, 0.0)
    d2, idx = e.knn_search(queries, k)
    d_all = np.stack([knn.l2_f32_ordered(queries[i], rows) for i in range(3)])
    for i in range(3):
        order = np.lexsort((np.arange(n), d_all[i]))[:k]
        np.testing.assert_array_equal(idx[i], order)
        np.testing.assert_array_equal(d2[i], d_all[i][order])
    assert e.knn_fallbacks() > 0
    e.close()


def _fresh_engine(env, seed, index_rate=None):
    rb = env["rvc_b200"]
    e = rb.RvcInfer(env["paths"]["data"], noise_seed=seed)
    e.load_contentvec(2); e.load_f0(1); e.load_model(env["paths"]["model"])
    if index_rate is not None:
        e.load_index(env["paths"]["index"], index_rate)
    return e


BATCH_AUDIO_TOL = 1e-4   # batched vs single plans differ only in fp32 summation order (split-K factors, chains)


@pytest.mark.parametrize("n_streams", [2, 8])
def test_independent_streams_batch(env, n_streams):
    """SURVEY 8e / BASELINE configs[3]: independent live streams of one GPU driven through rvc_infer_batch (ONE batched
    plan: every kernel processes all streams) give what each stream gives alone - integers exact, audio to 1e-4 RMS -
    over consecutive windows (per-stream pitch cache and call counter stay with their context)."""
    rb = env["rvc_b200"]
    g = env["pipeline"].BASELINE_GEOM
    nwin = 3
    pcms = [env["pipeline"].synthetic_pcm(g["n16k"] + g["sf16k"] * nwin, seed=20 + i) for i in range(n_streams)]
    solo, solo_pitch, solo_arg = [], [], []
    for i in range(n_streams):
        e = _fresh_engine(env, i, 0.5)
        outs, ps, am = [], [], []
        for w in range(nwin):
            x = pcms[i][w * g["sf16k"]: w * g["sf16k"] + g["n16k"]]
            outs.append(e.infer(x, g["sf16k"], 12, g["skip_head"], g["return_length"]).copy())
            ps.append(e.get_last("pitch", np.int32)); am.append(e.get_last("f0_argmax", np.int32))
        solo.append(outs); solo_pitch.append(ps); solo_arg.append(am)
        e.close()
    es = [_fresh_engine(env, i, 0.5) for i in range(n_streams)]
    for w in range(nwin):
        xs = [pcms[i][w * g["sf16k"]: w * g["sf16k"] + g["n16k"]] for i in range(n_streams)]
        outs = rb.infer_batch(es, xs, g["sf16k"], 12, g["skip_head"], g["return_length"])
        assert es[0].plan_info()["windows"] == n_streams          # really one batched plan
        for i in range(n_streams):
            assert outs[i].shape == solo[i][w].shape
            np.testing.assert_array_equal(es[0].get_last_window(i, "pitch", np.int32), solo_pitch[i][w])
            np.testing.assert_array_equal(es[0].get_last_window(i, "f0_argmax", np.int32), solo_arg[i][w])
            assert _rms(outs[i] - solo[i][w]) < BATCH_AUDIO_TOL, (w, i)
            assert _rms(solo[i][w]) > 0.05
    for e in es:
        e.close()


@pytest.mark.parametrize("n_windows,max_batch", [(6, 4), (32, 32)])
def test_offline_windows_equal_single_calls(env, n_windows, max_batch):
    """BASELINE configs[2] (offline, 32 windows per launch): rvc_infer_windows == the same windows through successive
    rvc_infer calls (the reference's streaming loop, obs-rvc/src/lib.rs:659-707): coarse pitch and F0 argmax exact
    (the pitch cache is updated window by window inside the plan), kNN indices equal, audio to 1e-4 RMS."""
    g = env["pipeline"].BASELINE_GEOM
    pcm = env["pipeline"].synthetic_pcm(g["n16k"] + g["sf16k"] * n_windows, seed=5)
    e1 = _fresh_engine(env, 3, 0.5)
    singles, pitch, arg, idx = [], [], [], []
    for w in range(n_windows):
        x = pcm[w * g["sf16k"]: w * g["sf16k"] + g["n16k"]]
        singles.append(e1.infer(x, g["sf16k"], 12, g["skip_head"], g["return_length"]).copy())
        pitch.append(e1.get_last("pitch", np.int32)); arg.append(e1.get_last("f0_argmax", np.int32)); idx.append(e1.get_last("knn_idx", np.int32))
    e1.close()
    e2 = _fresh_engine(env, 3, 0.5)
    got = e2.infer_windows(pcm, g["n16k"], g["sf16k"], n_windows, 12, g["skip_head"], g["return_length"], max_batch)
    assert got.shape == (n_windows, g["return_length"] * 400)
    last_group = n_windows - ((n_windows - 1) // max_batch) * max_batch
    base = n_windows - last_group
    same_idx = []
    for b in range(last_group if last_group > 1 else 0):
        np.testing.assert_array_equal(e2.get_last_window(b, "pitch", np.int32), pitch[base + b])
        np.testing.assert_array_equal(e2.get_last_window(b, "f0_argmax", np.int32), arg[base + b])
        same_idx.append((e2.get_last_window(b, "knn_idx", np.int32) == idx[base + b]).mean())
    if same_idx:
        assert np.mean(same_idx) > 0.99
    for w in range(n_windows):
        assert _rms(got[w] - singles[w]) < BATCH_AUDIO_TOL, w
    # the stream continues seamlessly after the batched call (cache + call counter carried over)
    x = pcm[n_windows * g["sf16k"]: n_windows * g["sf16k"] + g["n16k"]]
    e3 = _fresh_engine(env, 3, 0.5)
    for w in range(n_windows):
        e3.infer(pcm[w * g["sf16k"]: w * g["sf16k"] + g["n16k"]], g["sf16k"], 12, g["skip_head"], g["return_length"])
    want = e3.infer(x, g["sf16k"], 12, g["skip_head"], g["return_length"]).copy()
    nxt = e2.infer(x, g["sf16k"], 12, g["skip_head"], g["return_length"]).copy()
    np.testing.assert_array_equal(e2.get_last("pitch", np.int32), e3.get_last("pitch", np.int32))
    assert _rms(nxt - want) < BATCH_AUDIO_TOL
    e2.close(); e3.close()


def test_rvc_rpc_wire_protocol(env):
    """Seam 3: the rvc-rpc replacement speaks the reference's pipe protocol
    (rvc-rpc/src/main.rs:64-100, obs-rvc/src/rvcadapter.rs:69-118) and returns the same audio as
    the in-process call."""
    import struct
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "obs-rvc_b200", "rvc-rpc")
    assert os.path.exists(exe)
    g = env["pipeline"].BASELINE_GEOM
    pcm = env["pipeline"].synthetic_pcm(g["n16k"] + g["sf16k"], seed=9)
    rb = env["rvc_b200"]
    e = rb.RvcInfer(env["paths"]["data"], noise_seed=0)
    e.load_contentvec(2); e.load_f0(1); e.load_model(env["paths"]["model"])
    want = [e.infer(pcm[w * g["sf16k"]: w * g["sf16k"] + g["n16k"]], g["sf16k"], 12, g["skip_head"], g["return_length"]).copy()
            for w in range(2)]
    e.close()
    p = subprocess.Popen([exe, "v2", "rmvpe", env["paths"]["model"], env["paths"]["data"]], stdin=subprocess.PIPE,
                         stdout=subprocess.PIPE)
    try:
        for w in range(2):
            x = pcm[w * g["sf16k"]: w * g["sf16k"] + g["n16k"]].astype("<f4").tobytes()
            p.stdin.write(struct.pack("<I", len(x)) + x + struct.pack("<IiII", g["sf16k"], 12, g["skip_head"], g["return_length"]))
            p.stdin.flush()
            (nb,) = struct.unpack("<I", p.stdout.read(4))
            got = np.frombuffer(p.stdout.read(nb), dtype="<f4")
            # the child process is alone on the GPU and runs the persistent-chain plan; this process holds other
            # contexts and runs the same ops as separate kernels (engine.cu live_contexts): fp32 sums in another order
            np.testing.assert_allclose(got, want[w], rtol=0, atol=1e-4)
    finally:
        p.kill()


# ---------------------------------------------------------------- "next" row #1: streaming glue, PINNED goldens


def test_sola_offset_golden(env):
    """obs-rvc/src/tests/sola.rs:10-16: get_sola_offset(infer_wav, sola_buffer, 1920, 480) == 321."""
    x = np.load(os.path.join(GOLDEN, "sola_infer_wav.npy"))
    b = np.load(os.path.join(GOLDEN, "sola_buffer.npy"))
    assert env["eng"].sola_offset(x, b, 1920, 480) == 321


def test_envelop_mixing_golden(env):
    """obs-rvc/src/tests/envelop_mixing.rs:8-36: rms1 / rms2 / mixed output vs the reference's .npy at 1e-6."""
    g = lambda n: np.load(os.path.join(GOLDEN, n))
    mixed, r1, r2 = env["eng"].envelop_mixing(g("envelop_input_wav.npy"), g("envelop_infer_wav.npy"), 48000, 0.8, want_rms=True)
    assert np.abs(r1 - g("envelop_rms1.npy")).max() < 1e-6
    assert np.abs(r2 - g("envelop_rms2.npy")).max() < 1e-6
    assert np.abs(mixed - g("envelop_infer_wav2.npy")).max() < 1e-6


def test_sola_crossfade_vs_oracle(env):
    """obs-rvc/src/lib.rs:768-794 on the reference's SOLA fixture: same offset, block and new tail."""
    from oracle import dsp
    x = np.load(os.path.join(GOLDEN, "sola_infer_wav.npy"))
    b = np.load(os.path.join(GOLDEN, "sola_buffer.npy"))
    frame = 14400  # 0.30 s block at 48 kHz (lib.rs:200-203)
    blk, tail, off = env["eng"].sola_crossfade(x, b, 1920, 480, frame)
    wblk, wtail, woff = dsp.sola_crossfade(x, b, 1920, 480, frame)
    assert off == woff == 321
    assert np.abs(blk - wblk).max() < 1e-6 and np.abs(tail - wtail).max() < 1e-6
