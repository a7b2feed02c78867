"""GPU parity tests added in round 2 (VERDICT r1 "parity hardening"): every F0 argmax bin incl. the bins where
the reference would panic, the exact-threshold case, both `upstream_*` conventions, the v1 model family
(256-wide ContentVec, 9 layers + final_proj, rvc-common/src/enums.rs:9-23), the persistent ContentVec stack
kernel, single-cluster chains, and two chained contexts sharing one device.  All through the C ABI."""
import os
import tempfile

import numpy as np
import pytest

from test_gpu_parity import assert_topk_exact_up_to_ties

pytestmark = pytest.mark.gpu

WAVE_RMS_TOL = 1e-3
FEATS_ABS_TOL = 2e-3


def _rms(a):
    return float(np.sqrt(np.mean(np.asarray(a, np.float64) ** 2)))


@pytest.fixture(scope="module")
def env():
    import rvc_b200
    from oracle import dsp, pipeline, weights
    from oracle.weights import read_rvcw
    root = os.path.join(tempfile.gettempdir(), "rvc_b200_data_seed7")
    paths = weights.make_data_dir(root, seed=7, index_rows=40000)
    return dict(rvc_b200=rvc_b200, pipeline=pipeline, weights=weights, read_rvcw=read_rvcw, dsp=dsp, paths=paths)


def _engine(env, **kw):
    e = env["rvc_b200"].RvcInfer(env["paths"]["data"], **kw)
    e.load_contentvec(2); e.load_f0(1); e.load_model(env["paths"]["model"])
    return e


def _crafted_salience(seed=0):
    """One row per argmax bin 0..359 (a bump over a noise floor, neighbours non-zero so that the literal window
    c+4..c+12 of rmvpe.rs:119-125 has mass), then the threshold cases and an all-zero row."""
    rng = np.random.default_rng(seed)
    rows = []
    for b in range(360):
        r = rng.uniform(0.001, 0.02, 360).astype(np.float32)
        lo, hi = max(0, b - 6), min(360, b + 14)
        r[lo:hi] += rng.uniform(0.05, 0.3, hi - lo).astype(np.float32)
        r[b] = np.float32(0.9)
        rows.append(r)
    thr = np.float32(0.03)
    for peak in (thr, np.nextafter(thr, np.float32(1)), np.nextafter(thr, np.float32(0))):
        r = np.full(360, 0.01, np.float32); r[100] = peak; r[104:113] = 0.02
        rows.append(r)
    rows.append(np.zeros(360, np.float32))
    tie = np.full(360, 0.2, np.float32); tie[[17, 200]] = 0.7            # two equal maxima: the first one wins
    rows.append(tie)
    return np.stack(rows)


@pytest.mark.parametrize("upstream_window", [False, True])
def test_decode_every_bin_threshold_and_ties(env, upstream_window):
    """Rmvpe::decode / to_local_average_cents (rmvpe.rs:118-133, 243-248) on crafted salience: argmax bins 0..359 -
    including c >= 348, where the reference indexes out of bounds and this build treats the missing taps as absent -
    the peak exactly at / one ulp around the 0.03 threshold, an all-zero row, a tie.  Both cents-window conventions."""
    sal = _crafted_salience()
    e = _engine(env, upstream_cents_window=upstream_window)
    f0, am = e.decode_salience(sal)
    e.close()
    want_f0, want_c = env["dsp"].decode(sal, 0.03, upstream_window)
    np.testing.assert_array_equal(am, want_c.astype(np.int32))
    assert len(set(am[:360].tolist())) == 360                             # the sweep visits every bin
    assert am[360 + 3] == -4 and am[360 + 4] == 17                        # zero row, first maximum of a tie
    np.testing.assert_allclose(f0, want_f0, rtol=3e-6, atol=0)
    assert f0[360] == 0.0 and f0[362] == 0.0 and (f0[361] > 0.0) == (want_f0[361] > 0.0)   # `>` threshold, not `>=`
    if not upstream_window:
        assert np.all(f0[356:360] == 0.0)                                 # no tap left in range: unvoiced


def test_upstream_flags_full_window(env):
    """`upstream_pitch_shift` (2^(shift/12) instead of the reference's integer octave, rvc.rs:121) and
    `upstream_cents_window` together, two consecutive windows against the oracle configured the same way."""
    pl = env["pipeline"]
    g = pl.BASELINE_GEOM
    e = _engine(env, noise_seed=4, upstream_pitch_shift=True, upstream_cents_window=True)
    ora = pl.RvcInfer(env["paths"]["data"], noise_seed=4, upstream_pitch_shift=True, upstream_cents_window=True)
    ora.load_contentvec(2); ora.load_f0(1); ora.load_model(env["paths"]["model"])
    pcm = pl.synthetic_pcm(g["n16k"] + 2 * g["sf16k"], seed=21)
    for w in range(2):
        x = pcm[w * g["sf16k"]: w * g["sf16k"] + g["n16k"]]
        got = e.infer(x, g["sf16k"], 7, g["skip_head"], g["return_length"]).copy()
        want = ora.infer(x, g["sf16k"], 7, g["skip_head"], g["return_length"])
        np.testing.assert_array_equal(e.get_last("f0_argmax", np.int32), ora.last["argmax"])
        np.testing.assert_array_equal(e.get_last("pitch", np.int32), ora.last["pitch"])
        assert _rms(got - want) < WAVE_RMS_TOL and _rms(want) > 0.05
    e.close()


def test_v1_model_family(env):
    """ModelVersion::V1 (rvc-common/src/enums.rs:9-23): vec-256-layer-9 = 9 transformer layers + final_proj to 256,
    a 256-wide voice model and a 256-wide retrieval index (the tensor-core candidate pass + exact re-rank)."""
    rb, pl = env["rvc_b200"], env["pipeline"]
    root = os.path.join(tempfile.gettempdir(), "rvc_b200_data_v1_seed9")
    paths = env["weights"].make_data_dir(root, seed=9, v1=True, index_rows=20000)
    index = env["read_rvcw"](paths["index"])["big_npy"]
    assert index.shape == (20000, 256)
    e = rb.RvcInfer(root, noise_seed=1)
    e.load_contentvec(1); e.load_f0(1); e.load_model(paths["model"])
    ora = pl.RvcInfer(root, noise_seed=1)
    ora.load_contentvec(1); ora.load_f0(1); ora.load_model(paths["model"])
    g = pl.BASELINE_GEOM
    pcm = pl.synthetic_pcm(g["n16k"] + 2 * g["sf16k"], seed=31)
    h = e.hubert(pcm[:g["n16k"]])
    want_h = ora.hubert(pcm[:g["n16k"]])
    assert h.shape == want_h.shape and h.shape[1] == 256
    assert np.abs(h - want_h).max() < FEATS_ABS_TOL
    e.set_index(index, 0.5); ora.set_index(index, 0.5)
    for w in range(2):
        x = pcm[w * g["sf16k"]: w * g["sf16k"] + g["n16k"]]
        got = e.infer(x, g["sf16k"], 12, g["skip_head"], g["return_length"]).copy()
        want = ora.infer(x, g["sf16k"], 12, g["skip_head"], g["return_length"])
        np.testing.assert_array_equal(e.get_last("f0_argmax", np.int32), ora.last["argmax"])
        np.testing.assert_array_equal(e.get_last("pitch", np.int32), ora.last["pitch"])
        assert_topk_exact_up_to_ties(e.get_last("knn_idx", np.int32), ora.last["knn_idx"], ora.last["knn_d2"], ora.last["knn_q"], index)
        assert np.abs(e.get_last("phone").reshape(ora.last["phone"].shape) - ora.last["phone"]).max() < FEATS_ABS_TOL
        assert _rms(got - want) < WAVE_RMS_TOL and _rms(want) > 0.05
    assert e.knn_fallbacks() == 0
    # a 768-wide index on the 256-wide model must be refused, not read out of bounds
    with pytest.raises(rb.RvcInferError):
        e.set_index(np.zeros((64, 768), np.float32), 0.5)
    e.close()


def _two_windows(e, env, seed):
    pl = env["pipeline"]
    g = pl.BASELINE_GEOM
    x = pl.synthetic_pcm(g["n16k"] + g["sf16k"], seed=seed)
    outs = [e.infer(x[w * g["sf16k"]: w * g["sf16k"] + g["n16k"]], g["sf16k"], 12, g["skip_head"], g["return_length"]).copy()
            for w in range(2)]
    return outs, e.get_last("f0_argmax", np.int32).copy(), e.get_last("pitch", np.int32).copy()


def test_cvstack_kernel_matches_separate_kernels(env, monkeypatch):
    """The persistent ContentVec stack kernel (kernels_cvstack.cu: 12 layers in one cooperative tcgen05 launch; the default
    for infer plans, RVC_CVSTACK=1 extends it to the other plan kinds) against the oracle at the reference's feature tolerance and against the same window on the
    separate kernels: integers identical, audio equal up to fp32 summation order."""
    pl = env["pipeline"]
    g = pl.BASELINE_GEOM
    monkeypatch.setenv("RVC_CHAIN", "2")
    monkeypatch.setenv("RVC_CVSTACK", "1")
    e1 = _engine(env, noise_seed=6)
    x = pl.synthetic_pcm(g["n16k"], seed=41)
    h = e1.hubert(x)
    assert e1.plan_info().get("cvstack", 0) == 1
    ora = pl.RvcInfer(env["paths"]["data"], noise_seed=6)
    ora.load_contentvec(2)
    assert np.abs(h - ora.hubert(x)).max() < FEATS_ABS_TOL
    a1, am1, p1 = _two_windows(e1, env, 43)
    e1.close()
    monkeypatch.setenv("RVC_CVSTACK", "0")
    e0 = _engine(env, noise_seed=6)
    a0, am0, p0 = _two_windows(e0, env, 43)
    assert e0.plan_info().get("cvstack", 0) == 0
    e0.close()
    np.testing.assert_array_equal(am1, am0)
    np.testing.assert_array_equal(p1, p0)
    for a, b in zip(a1, a0):
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-4)
        assert _rms(a) > 0.01


def test_single_cluster_chains_match_separate_kernels(env, monkeypatch):
    """Chains of <= 16 CTAs run as ONE thread-block cluster with the hardware cluster barrier between phases
    (kernels_chain.cu): same integers / audio as the separate kernels."""
    monkeypatch.setenv("RVC_CHAIN", "2")
    monkeypatch.setenv("RVC_CHAIN_MAIN", "16")
    monkeypatch.setenv("RVC_CHAIN_SIDE", "16")
    monkeypatch.setenv("RVC_WSTREAM", "0")   # the bottleneck on the generic chain kernel too (its own kernel always takes 64 CTAs)
    e1 = _engine(env, noise_seed=8)
    a1, am1, p1 = _two_windows(e1, env, 47)
    chains = e1.profile_chains()
    assert len(chains) >= 3 and all(c["grid"] == 16 for c in chains)
    e1.close()
    monkeypatch.setenv("RVC_CHAIN", "0")
    e0 = _engine(env, noise_seed=8)
    a0, am0, p0 = _two_windows(e0, env, 47)
    e0.close()
    np.testing.assert_array_equal(am1, am0)
    np.testing.assert_array_equal(p1, p0)
    for a, b in zip(a1, a0):
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-4)


def test_two_chained_contexts_share_one_device(env, monkeypatch):
    """Two contexts on ONE device, both with their persistent chains forced on (cooperative launches from two
    contexts), windows interleaved: each stream's integers and audio equal the stream run alone (per-stream state,
    RvcInfer struct rvc.rs:18-27) - the single-GPU stand-in for the two-device test."""
    monkeypatch.setenv("RVC_CHAIN", "2")
    pl = env["pipeline"]
    g = pl.BASELINE_GEOM
    pcm = [pl.synthetic_pcm(g["n16k"] + 2 * g["sf16k"], seed=51 + s) for s in range(2)]

    def win(s, w):
        return pcm[s][w * g["sf16k"]: w * g["sf16k"] + g["n16k"]]

    solo = []
    for s in range(2):
        e = _engine(env, noise_seed=10 + s)
        solo.append([(e.infer(win(s, w), g["sf16k"], 12, g["skip_head"], g["return_length"]).copy(),
                      e.get_last("pitch", np.int32).copy()) for w in range(3)])
        e.close()
    es = [_engine(env, noise_seed=10 + s) for s in range(2)]
    for w in range(3):
        for s in (1, 0) if w % 2 else (0, 1):
            a = es[s].infer(win(s, w), g["sf16k"], 12, g["skip_head"], g["return_length"]).copy()
            np.testing.assert_array_equal(es[s].get_last("pitch", np.int32), solo[s][w][1])
            np.testing.assert_allclose(a, solo[s][w][0], rtol=0, atol=1e-4)
    for e in es:
        e.close()


# ---- the streaming loop around the call ("next" row 1): resamplers + process_one_frame on the device ----------------

@pytest.mark.parametrize("fs_in,fs_out,chunk", [(48000, 16000, 15360), (40000, 48000, 14000), (44100, 16000, 441 * 32),
                                                  (32000, 48000, 3200), (48000, 48000, 4800)])
def test_resampler_matches_rubato_restatement(env, fs_in, fs_out, chunk):
    """rubato 0.15.0 FftFixedInOut (obs-rvc/src/lib.rs:236-242) on the device as a polyphase filter against the numpy
    FFT restatement (oracle/resample.py; parity unpinned - no golden exists): four consecutive chunks incl. the
    overlap-add state, 1e-5 of full scale."""
    from oracle.resample import FftFixedInOut
    rng = np.random.default_rng(5)
    e = env["rvc_b200"].RvcInfer(env["paths"]["data"])
    ora = FftFixedInOut(fs_in, fs_out, chunk)
    assert ora.fft_size_in == chunk
    t = np.arange(4 * chunk) / fs_in
    x = (0.4 * np.sin(2 * np.pi * 440 * t) + 0.2 * np.sin(2 * np.pi * 3100 * t) + 0.05 * rng.standard_normal(4 * chunk)).astype(np.float32)
    ov = np.zeros(ora.fft_size_out, np.float32)
    for i in range(4):
        got = e.resample_chunk(fs_in, fs_out, x[i * chunk:(i + 1) * chunk], ov)
        want = ora.process(x[i * chunk:(i + 1) * chunk])
        assert got.shape == want.shape == (ora.fft_size_out,)
        assert np.abs(got - want).max() < 1e-5, (i, float(np.abs(got - want).max()))
        assert np.abs(ov - ora.overlap.astype(np.float32)).max() < 1e-5
    assert _rms(want) > 0.1
    e.close()


@pytest.mark.parametrize("skip_inference,mix", [(True, 0.0), (False, 0.0), (False, 1.0)])
def test_process_frame_matches_streaming_loop(env, skip_inference, mix):
    """process_one_frame (obs-rvc/src/lib.rs:659-795) device-resident behind rvc_process_frame - ring buffers, down-sample,
    infer, up-sample, envelope mixing, SOLA, cross-fade - against the same loop restated on the oracle (oracle/stream.py)
    for consecutive frames: same SOLA offsets, block within the waveform tolerance."""
    from oracle.stream import Stream, StreamGeometry
    pl = env["pipeline"]
    e = _engine(env, noise_seed=12)
    kw = dict(sample_rate=48000, sample_length=0.16, crossfade_length=0.04, extra_inference_time=2.0)
    frame = e.stream_open(pitch_shift=12, rms_mix_rate=mix, skip_inference=skip_inference, **kw)
    g = StreamGeometry(model_sample_rate=40000, skip_inference=skip_inference, **kw)
    info = e.stream_info()
    assert frame == g.sample_frame_size and info["input_buffer_16k_size"] == g.input_buffer_16k_size
    assert info["model_return_length"] == g.model_return_length and info["skip_head"] == g.skip_head
    ora_eng = None
    if not skip_inference:
        ora_eng = pl.RvcInfer(env["paths"]["data"], noise_seed=12)
        ora_eng.load_contentvec(2); ora_eng.load_f0(1); ora_eng.load_model(env["paths"]["model"])
    ora = Stream(ora_eng, g, pitch_shift=12, rms_mix_rate=mix)
    nfr = 4 if skip_inference else 3
    x48 = pl.synthetic_pcm(frame * nfr, seed=61)          # any band-limited test signal; used at the OBS rate here
    for i in range(nfr):
        blk = x48[i * frame:(i + 1) * frame]
        got = e.process_frame(blk)
        want = ora.process_one_frame(blk)
        assert got.shape == want.shape == (frame,)
        if _rms(want) > 1e-3:
            assert e.last_sola_offset == ora.last["sola_offset"], (i, e.last_sola_offset, ora.last["sola_offset"])
        assert _rms(got - want) < WAVE_RMS_TOL, (i, _rms(got - want), _rms(want))
    assert _rms(want) > 0.01
    e.stream_close()
    e.close()


def test_c_smoke_program_links_and_runs(env):
    """A plain-C caller of the boundary (obs-rvc_b200/examples/smoke.c: rvc_create ... rvc_infer, rvc_process_frame),
    built against include/rvc_b200.h and librvc_b200.so only."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "obs-rvc_b200", "smoke")
    assert os.path.exists(exe), "make -C obs-rvc_b200 smoke"
    r = subprocess.run([exe, env["paths"]["data"], env["paths"]["model"]], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "SMOKE_C ok" in r.stdout and "rvc_infer: 8400 samples" in r.stdout


@pytest.mark.parametrize("sr", [32000, 48000])
def test_other_generator_rates(env, sr):
    """The 32 kHz and 48 kHz generator configs of upstream RVC (upsample rates 10-8-2-2 / 12-10-2-2; the OBS filter
    offers 16-48 kHz destination rates, obs-rvc/src/lib.rs:345-352): two windows against the oracle, plus one block
    through the streaming loop (model rate -> 48 kHz resampler with that rate)."""
    rb, pl = env["rvc_b200"], env["pipeline"]
    root = os.path.join(tempfile.gettempdir(), f"rvc_b200_data_sr{sr}")
    paths = env["weights"].make_data_dir(root, seed=11, sr=sr)
    e = rb.RvcInfer(root, noise_seed=2)
    e.load_contentvec(2); e.load_f0(1); e.load_model(paths["model"])
    ora = pl.RvcInfer(root, noise_seed=2)
    ora.load_contentvec(2); ora.load_f0(1); ora.load_model(paths["model"])
    g = pl.BASELINE_GEOM
    pcm = pl.synthetic_pcm(g["n16k"] + 2 * g["sf16k"], seed=71)
    for w in range(2):
        x = pcm[w * g["sf16k"]: w * g["sf16k"] + g["n16k"]]
        got = e.infer(x, g["sf16k"], 12, g["skip_head"], g["return_length"]).copy()
        want = ora.infer(x, g["sf16k"], 12, g["skip_head"], g["return_length"])
        assert got.shape == want.shape == (g["return_length"] * sr // 100,)
        np.testing.assert_array_equal(e.get_last("pitch", np.int32), ora.last["pitch"])
        assert _rms(got - want) < WAVE_RMS_TOL and _rms(want) > 0.05
    frame = e.stream_open(sample_rate=48000, sample_length=0.16, crossfade_length=0.04, extra_inference_time=2.0)
    info = e.stream_info()
    assert info["model_sample_rate"] == sr and info["up"][0] == info["model_return_length"] * sr // 100
    out = e.process_frame(pl.synthetic_pcm(frame, seed=72))
    assert out.shape == (frame,) and np.isfinite(out).all()
    e.close()


def test_slab_chains_match_separate_kernels(env, monkeypatch):
    """Slab chains (chain.h, opt-in with RVC_SLAB=1): enc_p, flow and RMVPE's bottleneck as single 16-CTA clusters whose
    GEMMs are split by output columns, weights re-packed into per-CTA streams pulled by a producer warp ahead of the
    cluster barriers.  Same integers / audio as the separate kernels."""
    monkeypatch.setenv("RVC_CHAIN", "2")
    monkeypatch.setenv("RVC_SLAB", "1")
    e1 = _engine(env, noise_seed=9)
    a1, am1, p1 = _two_windows(e1, env, 49)
    chains = e1.profile_chains()
    assert len(chains) >= 3 and all(c["grid"] == 16 for c in chains)
    e1.close()
    monkeypatch.setenv("RVC_SLAB", "0")
    monkeypatch.setenv("RVC_CHAIN", "0")
    e0 = _engine(env, noise_seed=9)
    a0, am0, p0 = _two_windows(e0, env, 49)
    e0.close()
    np.testing.assert_array_equal(am1, am0)
    np.testing.assert_array_equal(p1, p0)
    for a, b in zip(a1, a0):
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-4)


@pytest.mark.gpu
def test_fused_residual_blocks_match_separate_gemms(env, monkeypatch):
    """kernels_cbr.cu: RMVPE's ConvBlockRes at U-Net levels 0 / 1 as one kernel per block (opt-in: RVC_CBR=2 the decoder's
    blocks, RVC_CBR=1 the encoder's too; default RVC_CBR=0: three implicit GEMMs per block).  Same pitch bins / argmax,
    audio within 1e-4, and fewer launches."""
    out = {}
    for mode in ("0", "1", "2"):
        monkeypatch.setenv("RVC_CBR", mode)
        e = _engine(env, noise_seed=9)
        a, am, p = _two_windows(e, env, 49)
        out[mode] = (a, am, p, e.kernel_launches())
        e.close()
    assert out["1"][3] < out["2"][3] < out["0"][3]
    for mode in ("1", "2"):
        np.testing.assert_array_equal(out[mode][1], out["0"][1])
        np.testing.assert_array_equal(out[mode][2], out["0"][2])
        for a, b in zip(out[mode][0], out["0"][0]):
            np.testing.assert_allclose(a, b, rtol=0, atol=1e-4)


@pytest.mark.gpu
def test_weight_streaming_bottleneck_matches_generic_chain(env, monkeypatch):
    """kernels_wstream.cu: RMVPE's U-Net bottleneck (31 skinny GEMMs, 4 valid rows x 512 x 1536) on the weight-streaming
    kernel (default) against the same ops on the generic chain kernel (RVC_WSTREAM=0) and on separate launches
    (RVC_CHAIN=0): identical argmax / pitch bins, audio within 1e-4."""
    out = {}
    for tag, envs in (("ws", {}), ("chain", {"RVC_WSTREAM": "0"}), ("plain", {"RVC_CHAIN": "0"})):
        for k in ("RVC_WSTREAM", "RVC_CHAIN"):
            monkeypatch.delenv(k, raising=False)
        for k, v in envs.items():
            monkeypatch.setenv(k, v)
        e = _engine(env, noise_seed=9)
        a, am, p = _two_windows(e, env, 53)
        grids = sorted(c["grid"] for c in e.profile_chains())
        out[tag] = (a, am, p, grids)
        e.close()
    assert 64 in out["ws"][3] and 64 not in out["chain"][3] and out["plain"][3] == []
    for tag in ("chain", "plain"):
        np.testing.assert_array_equal(out["ws"][1], out[tag][1])
        np.testing.assert_array_equal(out["ws"][2], out[tag][2])
        for a, b in zip(out["ws"][0], out[tag][0]):
            np.testing.assert_allclose(a, b, rtol=0, atol=1e-4)
