"""Literal numpy restatement of the reference's Rust DSP and glue (TEST INFRASTRUCTURE).

Every function cites the reference lines it follows (paths relative to /root/reference).
Arithmetic is float32 wherever the reference computes in f32 and float64 where it computes in
f64; the reference's quirks are reproduced, not fixed.

Pinned by tests/test_oracle_kats.py against the reference's own known-answer tests.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32

# ----------------------------------------------------------------------------------------------
# vendor/mel-spec/mel_spec/src/mel.rs:149-237  (librosa-compatible mel filterbank, all f64)
# ----------------------------------------------------------------------------------------------


def hz_to_mel(frequency: float, htk: bool) -> float:
    """mel.rs:185-201."""
    if htk:
        return 2595.0 * np.log10(1.0 + frequency / 700.0)
    f_min, f_sp = 0.0, 200.0 / 3.0
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if frequency >= min_log_hz:
        return min_log_mel + np.log(frequency / min_log_hz) / logstep
    return (frequency - f_min) / f_sp


def mel_to_hz(mel: float, htk: bool) -> float:
    """mel.rs:203-219."""
    if htk:
        return 700.0 * (10.0 ** (mel / 2595.0) - 1.0)
    f_min, f_sp = 0.0, 200.0 / 3.0
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if mel >= min_log_mel:
        return min_log_hz * np.exp(logstep * (mel - min_log_mel))
    return f_min + f_sp * mel


def mels_to_hz(mels: np.ndarray, htk: bool) -> np.ndarray:
    """mel.rs:221-223."""
    return np.array([mel_to_hz(float(m), htk) for m in mels], dtype=np.float64)


def mel_frequencies(n_mels: int, fmin: float, fmax: float, htk: bool) -> np.ndarray:
    """mel.rs:225-231 (ndarray linspace is start + i*step with step=(end-start)/(n-1))."""
    min_mel = hz_to_mel(fmin, htk)
    max_mel = hz_to_mel(fmax, htk)
    mels = np.linspace(min_mel, max_mel, n_mels, dtype=np.float64)
    return mels_to_hz(mels, htk)


def fft_frequencies(sr: float, n_fft: int) -> np.ndarray:
    """mel.rs:233-237."""
    step = sr / n_fft
    return step * np.arange(n_fft // 2 + 1, dtype=np.float64)


def mel_filterbank(sr: float, n_fft: int, n_mels: int, f_min=None, f_max=None,
                   htk: bool = False, norm: bool = True) -> np.ndarray:
    """mel.rs:149-183 `mel()`; returns (n_mels, n_fft/2+1) float64."""
    fftfreqs = fft_frequencies(sr, n_fft)
    f_min = 0.0 if f_min is None else float(f_min)
    f_max = sr / 2.0 if f_max is None else float(f_max)
    mel_f = mel_frequencies(n_mels + 2, f_min, f_max, htk)
    fdiff = mel_f[1:n_mels + 2] - mel_f[:n_mels + 1]
    ramps = mel_f[:n_mels + 2, None] - fftfreqs[None, :]
    weights = np.zeros((n_mels, n_fft // 2 + 1), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        # mel.rs:164-172: clamp each side to [0,1] (max then min), then elementwise min
        lo = np.minimum(np.maximum(lower, 0.0), 1.0)
        up = np.minimum(np.maximum(upper, 0.0), 1.0)
        weights[i] = np.minimum(lo, up)
    if norm:
        enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
        weights *= enorm[:, None]
    return weights


# ----------------------------------------------------------------------------------------------
# rvc/src/f0/rmvpe.rs  (windows, padding, STFT, mel, decode)
# ----------------------------------------------------------------------------------------------


def get_hann_window(window_length: int) -> np.ndarray:
    """rmvpe.rs:27-31 (symmetric; unused by the hot path, kept for the KAT suite)."""
    i = np.arange(window_length, dtype=np.float64)
    c = np.cos(2.0 * np.pi * i / (float(window_length) - 1.0)).astype(F32)
    return (F32(0.5) * (F32(1.0) - c)).astype(F32)


def get_hann_window_periodic(window_length: int) -> np.ndarray:
    """rmvpe.rs:33-37: cos evaluated in f64, rounded to f32, then 0.5*(1-x) in f32."""
    i = np.arange(window_length, dtype=np.float64)
    c = np.cos(2.0 * np.pi * i / (float(window_length + 1) - 1.0)).astype(F32)
    return (F32(0.5) * (F32(1.0) - c)).astype(F32)


def pad_constant(x: np.ndarray, pad: int, value) -> np.ndarray:
    """rmvpe.rs:39-45."""
    out = np.full(x.shape[0] + 2 * pad, value, dtype=x.dtype)
    out[pad:pad + x.shape[0]] = x
    return out


def pad_reflect(x: np.ndarray, pad: int) -> np.ndarray:
    """rmvpe.rs:47-68: numpy 'reflect' (edge sample not repeated)."""
    n = x.shape[0]
    out = np.empty(n + 2 * pad, dtype=x.dtype)
    out[pad:pad + n] = x
    for i in range(pad):
        out[pad - i - 1] = x[i + 1]
        out[n + pad + i] = x[n - i - 2]
    return out


def stft(signal: np.ndarray, fft_size: int, hop_length: int, window: np.ndarray,
         center: bool = True) -> np.ndarray:
    """rmvpe.rs:80-116: magnitude STFT, returns (fft_size/2+1, T) float32.

    T = 1 + L/hop uses the UNPADDED length (rmvpe.rs:84-86).  The FFT is complex-f32 in the
    reference (rustfft 6.2.0); here numpy's pocketfft in single precision.
    """
    signal = np.asarray(signal, dtype=F32)
    L = signal.shape[0]
    N = fft_size // 2 + 1
    T = 1 + L // hop_length
    if center:
        signal = pad_reflect(signal, fft_size // 2)
    win_length = window.shape[0]
    if win_length < fft_size:  # rmvpe.rs:95-101 (never taken on the hot path)
        left = (fft_size - win_length) // 2
        window = window[left:left + win_length]
    idx = (np.arange(T) * hop_length)[:, None] + np.arange(fft_size)[None, :]
    frames = (signal[idx] * window.astype(F32)[None, :]).astype(F32)
    spec = np.fft.fft(frames.astype(np.complex64), axis=1)
    if spec.dtype != np.complex64:  # older numpy computes in double
        spec = spec.astype(np.complex64)
    re = spec.real.astype(F32)
    im = spec.imag.astype(F32)
    mag = np.sqrt((re * re + im * im).astype(F32)).astype(F32)  # norm_sqr().sqrt()
    return np.ascontiguousarray(mag.T[:N, :])


class MelSpectrogram:
    """rmvpe.rs:18-25,135-205."""

    def __init__(self, fft_size=1024, sample_rate=16000, n_mels=128, win_length=1024,
                 hop_length=160, f_min=30.0, f_max=8000.0, clamp=1e-5):
        # rmvpe.rs:146-148: filterbank in f64 (htk=true, norm=true) cast to f32
        self.mel_basis = mel_filterbank(float(sample_rate), fft_size, n_mels, f_min, f_max,
                                        True, True).astype(F32)
        self.fft_size = fft_size
        self.win_length = win_length
        self.hop_length = hop_length
        self.clamp = F32(clamp)
        self._window = {}

    def mel_extract(self, x: np.ndarray) -> np.ndarray:
        """rmvpe.rs:159-205 with keyshift=None, speed=None, center=true (the only call,
        rmvpe.rs:258).  Returns (n_mels, T) float32 = ln(max(mel_basis @ |STFT|, clamp))."""
        if 0 not in self._window:
            self._window[0] = get_hann_window_periodic(self.win_length)
        mag = stft(x, self.fft_size, self.hop_length, self._window[0], True)
        mel = (self.mel_basis @ mag).astype(F32)
        return np.log(np.maximum(mel, self.clamp)).astype(F32)


N_CLASS = 360


def cents_mapping() -> np.ndarray:
    """rmvpe.rs:212-216: 368 entries, f32 arithmetic."""
    i = np.arange(N_CLASS + 8, dtype=F32)
    return ((i - F32(4.0)) * F32(20.0) + F32(1997.3794084376191)).astype(F32)


def to_local_average_cents(salience: np.ndarray, cents: np.ndarray, threshold: float,
                           upstream_window: bool = False):
    """rmvpe.rs:118-133, LITERAL: `starts` is the argmax in the padded
    row (= c+4) but indexes the UNPADDED salience, so taps come from bins c+4..c+12 and are
    paired with the cents of bins c..c+8.  The reference panics for c >= 348; here taps past
    bin 359 are treated as absent and a frame with no tap left (c >= 356) is unvoiced
    (documented divergence from a crash).

    Returns (cents[T] f32, argmax c[T] int) - c is the index in the unpadded row.
    """
    salience = np.asarray(salience, dtype=F32)
    T, n = salience.shape
    padded = np.zeros((T, n + 8), dtype=F32)
    padded[:, 4:n + 4] = salience
    starts = np.argmax(padded, axis=1)  # first maximum wins (ndarray-stats argmax)
    out = np.zeros(T, dtype=F32)
    for t in range(T):
        s = int(starts[t])
        ps = F32(0.0)
        ws = F32(0.0)
        for y in range(9):
            if upstream_window:
                w = padded[t, s - 4 + y] if 0 <= s - 4 + y < n + 8 else F32(0.0)
                c = cents[s - 4 + y] if 0 <= s - 4 + y < n + 8 else F32(0.0)
            else:
                if s + y >= n:
                    continue
                w = salience[t, s + y]
                c = cents[s + y]
            ps = F32(ps + F32(w * c))
            ws = F32(ws + w)
        # c >= 356: every tap is past bin 359 (reference: OOB panic) -> no estimate -> 0 cents,
        # which decode() maps to f0 == 10.0 -> 0 (unvoiced)
        d = F32(ps / ws) if ws != 0 else F32(0.0)
        mx = salience[t].max()
        out[t] = d if mx > F32(threshold) else F32(0.0)
    return out, (starts - 4).astype(np.int64)


def decode(hidden: np.ndarray, threshold: float, upstream_window: bool = False):
    """rmvpe.rs:243-248: f0 = 10 * 2^(cents/1200); exactly 10.0 -> 0."""
    cents, c = to_local_average_cents(hidden, cents_mapping(), threshold, upstream_window)
    f0 = (F32(10.0) * np.exp2((cents / F32(1200.0)).astype(F32)).astype(F32)).astype(F32)
    f0[f0 == F32(10.0)] = F32(0.0)
    return f0, c


def f0_extractor_frame(sample_frame_16k_size: int) -> int:
    """rmvpe.rs:256."""
    return 5120 * ((sample_frame_16k_size + 800 - 1) // 5120 + 1) - 160


# ----------------------------------------------------------------------------------------------
# rvc/src/f0/mod.rs, rvc/src/rvc.rs glue
# ----------------------------------------------------------------------------------------------

F0_MEL_MIN = F32(np.log(F32(F32(50.0) / F32(700.0) + F32(1.0))) * F32(1127.0))  # rvc.rs:31-33
F0_MEL_MAX = F32(np.log(F32(F32(500.0) / F32(700.0) + F32(1.0))) * F32(1127.0))  # rvc.rs:34


def get_f0_post(f0: np.ndarray, f0_mel_min=F0_MEL_MIN, f0_mel_max=F0_MEL_MAX):
    """f0/mod.rs:7-12.  Rust `round` is half-away-from-zero; values are in [1,255]."""
    f0 = np.asarray(f0, dtype=F32)
    m = (np.log((f0 / F32(700.0) + F32(1.0)).astype(F32)).astype(F32) * F32(1127.0)).astype(F32)
    scaled = ((m - f0_mel_min) * F32(254.0) / F32(f0_mel_max - f0_mel_min) + F32(1.0)).astype(F32)
    m = np.where(m <= 0, m, scaled).astype(F32)
    m = np.clip(m, F32(1.0), F32(255.0))
    coarse = np.floor(m.astype(np.float64) + 0.5).astype(np.int32)
    return coarse, f0


def pitch_shift_factor(pitch_shift: int, upstream: bool = False) -> np.float32:
    """rvc.rs:121: `2.0f32.powi(pitch_shift / 12)` - i32 division truncating toward zero
    (SURVEY Appendix B1)."""
    if upstream:
        return F32(2.0 ** (pitch_shift / 12.0))
    q = int(pitch_shift / 12)  # trunc toward zero like Rust
    return F32(2.0 ** q)


def extend_feature_2x(raw: np.ndarray) -> np.ndarray:
    """rvc.rs:99-109: raw (T,C) -> (2T+1,C), out[k] = raw[min(k/2, T-1)]."""
    T = raw.shape[0]
    idx = np.minimum(np.arange(2 * T + 1) // 2, T - 1)
    return np.ascontiguousarray(raw[idx])


class PitchCache:
    """rvc.rs:26,42,167-179 + ndarray_ext.rs:9-32: 1024-entry sliding f0 cache (per stream)."""

    def __init__(self, n: int = 1024):
        self.buf = np.zeros(n, dtype=F32)

    def update_and_slice(self, pitchf: np.ndarray, sample_frame_16k_size: int,
                         hubert_length: int, skip_head: int, return_length: int) -> np.ndarray:
        n = self.buf.shape[0]
        pitch_len = pitchf.shape[0]
        shift = sample_frame_16k_size // 160
        self.buf[:n - shift] = self.buf[shift:].copy()            # copy_within(shift.., 0)
        start = n + 4 - pitch_len                                 # rvc.rs:172
        self.buf[start:] = pitchf[3:pitch_len - 1]                # rvc.rs:174
        a = n - hubert_length + skip_head                         # rvc.rs:176
        return self.buf[a:a + return_length].copy()


# ----------------------------------------------------------------------------------------------
# "next" row #1 - obs-rvc/src/rt_utils.rs (streaming glue around the call)
# ----------------------------------------------------------------------------------------------


def rms(y: np.ndarray, frame_length: int, hop_length: int) -> np.ndarray:
    """rt_utils.rs:93-102."""
    pad = frame_length // 2
    yp = np.concatenate([np.zeros(pad, F32), np.asarray(y, F32), np.zeros(pad, F32)])
    yp = (yp * yp).astype(F32)
    n = (yp.shape[0] - frame_length) // hop_length + 1
    out = np.empty(n, dtype=F32)
    for i in range(n):
        seg = yp[i * hop_length:i * hop_length + frame_length]
        out[i] = np.sqrt(F32(seg.sum(dtype=F32) / F32(frame_length)))
    return out


def linear_interpolate_align_corners(x: np.ndarray, size: int) -> np.ndarray:
    """rt_utils.rs:104-117."""
    x = np.asarray(x, F32)
    step = F32(F32(x.shape[0] - 1) / F32(size - 1))
    idx = (np.arange(size, dtype=F32) * step).astype(F32)
    fl = np.clip(np.floor(idx).astype(np.int64), 0, x.shape[0] - 1)
    ce = np.clip(np.ceil(idx).astype(np.int64), 0, x.shape[0] - 1)
    frac = (idx - fl.astype(F32)).astype(F32)
    return (x[fl] * (F32(1.0) - frac) + x[ce] * frac).astype(F32)


def envelop_mixing(inp: np.ndarray, out: np.ndarray, sample_rate: int, mix_rate: float):
    """rt_utils.rs:119-132; returns (mixed, rms1, rms2)."""
    zc = sample_rate // 100
    n = out.shape[0]
    r1 = rms(inp[:n], 4 * zc, zc)
    r2 = rms(out, 4 * zc, zc)
    r1 = linear_interpolate_align_corners(r1, n + 1)
    r2 = np.maximum(linear_interpolate_align_corners(r2, n + 1), F32(1e-3))
    p = F32(1.0 - mix_rate)
    mixed = (np.asarray(out, F32) * np.power((r1[:n] / r2[:n]).astype(F32), p)).astype(F32)
    return mixed, r1[:n], r2[:n]


def get_sola_offset(input_buffer: np.ndarray, sola_buffer: np.ndarray, buffer_frame_size: int,
                    search_frame_size: int) -> int:
    """rt_utils.rs:60-90: normalised cross-CORRELATION (ndarray-conv `conv_fft` does not flip
    the kernel - verified against the reference golden 321), LAST maximum
    wins ties (rt_utils.rs:82-88)."""
    n = buffer_frame_size + search_frame_size
    x = np.asarray(input_buffer[:n], np.float64)
    k = np.asarray(sola_buffer, np.float64)
    nom = np.correlate(x, k, mode="valid")
    den = np.sqrt(np.correlate(x * x, np.ones(buffer_frame_size), mode="valid") + 1e-8)
    cor = (nom / den).astype(F32)
    best, val = 0, cor[0]
    for i in range(cor.shape[0]):
        if not (val > cor[i]):
            best, val = i, cor[i]
    return best


def sola_crossfade(infer_out: np.ndarray, sola_buffer: np.ndarray, buffer_frame_size: int,
                   search_frame_size: int, sample_frame_size: int):
    """obs-rvc/src/lib.rs:231-233,768-794: offset search, sin^2 cross-fade with the previous tail,
    new sola_buffer, and the block of `sample_frame_size` samples that is emitted."""
    off = get_sola_offset(infer_out, sola_buffer, buffer_frame_size, search_frame_size)
    out = np.asarray(infer_out, F32)[off:].copy()
    lin = np.linspace(0.0, 1.0, buffer_frame_size, dtype=F32)
    fade_in = (np.sin(lin * F32(0.5) * F32(np.pi)).astype(F32) ** 2).astype(F32)
    fade_out = (F32(1.0) - fade_in).astype(F32)
    out[:buffer_frame_size] = out[:buffer_frame_size] * fade_in + np.asarray(sola_buffer, F32) * fade_out
    new_sola = out[sample_frame_size:sample_frame_size + buffer_frame_size].copy()
    return out[:sample_frame_size].copy(), new_sola, off
