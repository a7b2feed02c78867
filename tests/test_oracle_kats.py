"""Pins oracle/dsp.py to the reference's own known-answer tests.

Each test cites the reference test it restates (paths relative to /root/reference).
"""
import json
import os

import numpy as np
import pytest

from oracle import dsp

F32 = np.float32


def test_stft_table():
    """rvc/src/f0/rmvpe.rs:269-291 `test_stft` (expected table "from torch"; the reference's
    own TODO admits exact equality fails - tolerance 4e-5 abs)."""
    signal = np.linspace(0.0, 1.0, 500, dtype=F32)
    window = dsp.get_hann_window_periodic(16)
    expected = np.array([
        [3.7801e-02, 2.5651e+00, 5.1303e+00, 7.6954e+00],
        [5.7373e-03, 1.2829e+00, 2.5653e+00, 3.8478e+00],
        [1.4787e-02, 6.7956e-03, 6.7958e-03, 6.7957e-03],
        [3.2463e-03, 1.6874e-03, 1.6874e-03, 1.6875e-03],
        [2.3478e-03, 6.6042e-04, 6.6042e-04, 6.6054e-04],
        [1.4494e-03, 3.1195e-04, 3.1202e-04, 3.1184e-04],
        [1.2455e-03, 1.5500e-04, 1.5485e-04, 1.5491e-04],
        [1.0416e-03, 6.5722e-05, 6.5798e-05, 6.5790e-05],
        [1.0417e-03, 0.0000e+00, 0.0000e+00, 2.3842e-07]], dtype=F32)
    out = dsp.stft(signal, 16, 160, window, True)
    assert out.shape == expected.shape
    assert np.abs(out - expected).max() < 4e-5


def test_stft_matches_torch():
    """Cross-check: equals torch.stft(center=True, reflect, periodic hann) magnitude."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(0)
    x = rng.standard_normal(4960).astype(F32) * 0.1
    out = dsp.stft(x, 1024, 160, dsp.get_hann_window_periodic(1024), True)
    ref = torch.stft(torch.from_numpy(x), 1024, 160, 1024,
                     torch.hann_window(1024, periodic=True), center=True, pad_mode="reflect",
                     return_complex=True).abs().numpy()
    assert out.shape == ref.shape == (513, 32)
    assert np.abs(out - ref).max() < 2e-5 * ref.max()


def test_pad_reflect():
    """rmvpe.rs:293-308."""
    np.testing.assert_array_equal(dsp.pad_reflect(np.array([1.0, 2.0, 3.0]), 2),
                                  [3.0, 2.0, 1.0, 2.0, 3.0, 2.0, 1.0])
    np.testing.assert_array_equal(dsp.pad_reflect(np.array([4.0, 5.0]), 1), [5.0, 4.0, 5.0, 4.0])


def test_pad_constant():
    """rmvpe.rs:310-326."""
    np.testing.assert_array_equal(dsp.pad_constant(np.array([1.0, 2.0, 3.0]), 2, 0.0),
                                  [0.0, 0.0, 1.0, 2.0, 3.0, 0.0, 0.0])
    np.testing.assert_array_equal(dsp.pad_constant(np.array([4.0, 5.0]), 1, 2.0),
                                  [2.0, 4.0, 5.0, 2.0])


def test_hz_mel_helpers():
    """vendor/mel-spec/mel_spec/src/mel.rs:266-315."""
    assert abs(dsp.hz_to_mel(60.0, False) - 0.9) < 1e-3
    assert dsp.mel_to_hz(3.0, False) == 200.0
    got = dsp.mels_to_hz(np.array([1.0, 2.0, 3.0, 4.0, 5.0]), False)
    assert np.abs(got - [66.667, 133.333, 200., 266.667, 333.333]).max() < 1e-3
    want = np.array([
        0., 85.317, 170.635, 255.952, 341.269, 426.586, 511.904, 597.221, 682.538, 767.855,
        853.173, 938.49, 1024.856, 1119.114, 1222.042, 1334.436, 1457.167, 1591.187, 1737.532,
        1897.337, 2071.84, 2262.393, 2470.47, 2697.686, 2945.799, 3216.731, 3512.582, 3835.643,
        4188.417, 4573.636, 4994.285, 5453.621, 5955.205, 6502.92, 7101.009, 7754.107,
        8467.272, 9246.028, 10096.408, 11025.])
    assert np.abs(dsp.mel_frequencies(40, 0.0, 11025.0, False) - want).max() < 5e-3
    want = [0., 1378.125, 2756.25, 4134.375, 5512.5, 6890.625, 8268.75, 9646.875, 11025.]
    assert np.abs(dsp.fft_frequencies(22050.0, 16) - want).max() < 1e-3


def test_mel_filters_npz(golden_dir):
    """mel.rs:317-330 `test_mel`: mel(16000,400,80,None,None,false,true) vs mel_filters.npz
    at 1e-7."""
    want = np.load(os.path.join(golden_dir, "mel_filters.npz"))["mel_80"].astype(np.float64)
    got = dsp.mel_filterbank(16000.0, 400, 80, None, None, False, True)
    assert got.shape == (80, 201)
    assert np.abs(got - want).max() <= 1.0e-7


def test_hot_path_mel_basis_shape():
    """rmvpe.rs:220: the hot-path basis mel(16000,1024,128,30,8000,htk,norm) - no reference
    golden exists (pitch.rs:12-16 only prints); structural facts from SURVEY section 8a."""
    mb = dsp.MelSpectrogram().mel_basis
    assert mb.shape == (128, 513) and mb.dtype == F32
    assert int((mb != 0).sum()) == 1010
    assert abs(float(mb.max()) - 0.0568) < 1e-3


def test_frame_geometry(golden_dir):
    """rvc/src/tests/hubert.rs:10-19: input_wav.npy[38240] <-> feats.npy (1,239,768):
    T=(N-400)/320+1=119, 2T+1=239 (rvc.rs:101-108)."""
    meta = json.load(open(os.path.join(golden_dir, "MANIFEST.json")))["feats_meta"]
    n = np.load(os.path.join(golden_dir, "input_wav.npy")).shape[0]
    T = (n - 400) // 320 + 1
    assert n == 38240 and T == 119 and meta["shape"] == [1, 2 * T + 1, 768]
    raw = np.arange(T * 3, dtype=F32).reshape(T, 3)
    ext = dsp.extend_feature_2x(raw)
    assert ext.shape == (239, 3)
    np.testing.assert_array_equal(ext[0], raw[0])
    np.testing.assert_array_equal(ext[1], raw[0])
    np.testing.assert_array_equal(ext[2], raw[1])
    np.testing.assert_array_equal(ext[-1], raw[-1])
    np.testing.assert_array_equal(ext[-3], raw[-1])
    np.testing.assert_array_equal(ext[-4], raw[-2])


def test_f0_window_geometry():
    """rmvpe.rs:256 + SURVEY section 8 geometry table."""
    assert dsp.f0_extractor_frame(2560) == 4960
    assert dsp.f0_extractor_frame(4800) == 10080
    assert 1 + 4960 // 160 == 32 and 1 + 10080 // 160 == 64


def test_local_average_cents_literal():
    """rmvpe.rs:118-133 literal window (Appendix B3): taps c+4..c+12 with cents of c..c+8."""
    sal = np.full((3, 360), 1e-3, F32)
    sal[0, 150] = 0.9
    sal[0, 154:163] = np.linspace(0.1, 0.5, 9, dtype=F32)
    sal[1, 100] = 0.02            # below threshold -> 0
    sal[2, 355] = 0.8             # c >= 348: reference panics; here taps past 359 are absent
    cents, c = dsp.to_local_average_cents(sal, dsp.cents_mapping(), 0.03)
    assert list(c) == [150, 100, 355]
    w = sal[0, 154:163].astype(np.float64)
    cm = 1997.3794084376191 + 20.0 * np.arange(150, 159)
    assert abs(cents[0] - (w * cm).sum() / w.sum()) < 1e-2
    assert cents[1] == 0.0
    w = sal[2, 359:360].astype(np.float64)
    assert abs(cents[2] - (1997.3794084376191 + 20.0 * 355)) < 1e-2
    f0, _ = dsp.decode(sal, 0.03)
    assert f0[1] == 0.0 and f0[0] > 0


def test_get_f0_post():
    """f0/mod.rs:7-12: unvoiced -> 1, clamp to [1,255], half-away rounding."""
    f0 = np.array([0.0, 50.0, 500.0, 1000.0, 10.0, 220.0], F32)
    coarse, f = dsp.get_f0_post(f0)
    assert coarse[0] == 1 and coarse[1] == 1 and coarse[2] == 255 and coarse[3] == 255
    assert coarse[4] == 1
    m = 1127.0 * np.log(1 + 220.0 / 700.0)
    want = (m - float(dsp.F0_MEL_MIN)) * 254.0 / float(dsp.F0_MEL_MAX - dsp.F0_MEL_MIN) + 1
    assert coarse[5] == int(np.floor(want + 0.5))
    assert f is not None and f.dtype == F32


def test_pitch_shift_integer_octaves():
    """rvc.rs:121 (Appendix B1)."""
    assert dsp.pitch_shift_factor(12) == 2.0
    assert dsp.pitch_shift_factor(7) == 1.0
    assert dsp.pitch_shift_factor(-11) == 1.0
    assert dsp.pitch_shift_factor(-12) == 0.5
    assert dsp.pitch_shift_factor(25) == 4.0


def test_pitch_cache_alignment():
    """rvc.rs:167-179 with the BASELINE geometry: write [996..1024),
    read [1001..1022)."""
    cache = dsp.PitchCache()
    cache.buf[:] = np.arange(1024, dtype=F32)
    pitchf = 1000.0 + np.arange(32, dtype=F32)
    got = cache.update_and_slice(pitchf, 2560, 223, 200, 21)
    assert cache.buf[0] == 16 and cache.buf[995] == 1011
    np.testing.assert_array_equal(cache.buf[996:], pitchf[3:31])
    np.testing.assert_array_equal(got, cache.buf[1001:1022])


# ------------------------------------------------------------------ "next" row #1 goldens


def test_rms_kat():
    """obs-rvc/src/rt_utils.rs:138-147."""
    y = np.arange(1, 11, dtype=F32)
    want = np.array([1.118034, 2.738613, 4.6368093, 6.595453, 8.573215, 6.726812], F32)
    assert np.abs(dsp.rms(y, 4, 2) - want).max() < 1e-6


def test_linear_interpolate_kat():
    """obs-rvc/src/rt_utils.rs:149-158."""
    x = np.array([0.2353, 0.9068, 0.7870, 0.5878, 0.0097, 0.7160, 0.5812, 0.8901, 0.8822,
                  0.8547], F32)
    want3 = np.array([0.2353, 0.36285, 0.8547], F32)
    want15 = np.array([0.2353, 0.66697854, 0.8725714, 0.79555714, 0.6731714, 0.4639215,
                       0.09228568, 0.36285, 0.6967429, 0.6100857, 0.7135856, 0.8895357,
                       0.8844571, 0.8723786, 0.8547], F32)
    assert np.abs(dsp.linear_interpolate_align_corners(x, 3) - want3).max() < 1e-6
    assert np.abs(dsp.linear_interpolate_align_corners(x, 15) - want15).max() < 1e-6


def test_sola_golden(golden_dir):
    """obs-rvc/src/tests/sola.rs:10-16 -> 321."""
    x = np.load(os.path.join(golden_dir, "sola_infer_wav.npy"))
    b = np.load(os.path.join(golden_dir, "sola_buffer.npy"))
    assert dsp.get_sola_offset(x, b, 1920, 480) == 321


def test_envelop_mixing_golden(golden_dir):
    """obs-rvc/src/tests/envelop_mixing.rs:8-36 (zc=480, mix 0.8, tol 1e-6)."""
    g = lambda n: np.load(os.path.join(golden_dir, n))
    mixed, r1, r2 = dsp.envelop_mixing(g("envelop_input_wav.npy"), g("envelop_infer_wav.npy"),
                                       48000, 0.8)
    assert np.abs(r1 - g("envelop_rms1.npy")).max() < 1e-6
    assert np.abs(r2 - g("envelop_rms2.npy")).max() < 1e-6
    assert np.abs(mixed - g("envelop_infer_wav2.npy")).max() < 1e-6
