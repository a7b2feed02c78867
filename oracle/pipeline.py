"""Oracle `RvcInfer`: the reference's public inference API restated on the CPU
(TEST INFRASTRUCTURE).  Method names, argument meaning and error behaviour follow
`rvc::RvcInfer` (/root/reference/rvc/src/rvc.rs:18-220).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import dsp, knn, nets, noise
from .weights import read_rvcw


class RvcInferError(Exception):
    """rvc-common/src/errors.rs:1-20."""


class ModelNotLoaded(RvcInferError):
    pass


class ContentvecNotLoaded(RvcInferError):
    pass


class F0NotLoaded(RvcInferError):
    pass


class BadShape(RvcInferError):
    """Where the reference panics on a slice out of range (SURVEY 8b 'Shape contract')."""


class RvcInfer:
    def __init__(self, data_path: str, noise_seed: int = 0, index_k: int = 8,
                 upstream_pitch_shift: bool = False, upstream_cents_window: bool = False):
        """rvc.rs:30-44."""
        self.data_path = data_path
        self.session = None
        self.contentvec = None
        self.rmvpe = None
        self.f0_mel_min = dsp.F0_MEL_MIN
        self.f0_mel_max = dsp.F0_MEL_MAX
        self.cache = dsp.PitchCache(1024)
        self.mel = dsp.MelSpectrogram()
        self.index = None
        self.index_rate = 0.0
        self.index_k = index_k
        self.noise_seed = noise_seed
        self.noise_enabled = True
        self.window = 0
        self.upstream_pitch_shift = upstream_pitch_shift
        self.upstream_cents_window = upstream_cents_window
        self.rec = None
        self.last = {}

    # -- loading (rvc.rs:46-79, models.rs:48-76; `.rvcw` instead of `.onnx`) ------------------
    def load_contentvec(self, version: int = 2):
        c, l = (256, 9) if version == 1 else (768, 12)        # enums.rs:9-23
        path = os.path.join(self.data_path, "contentvec", f"vec-{c}-layer-{l}.rvcw")
        self.contentvec = nets.to_torch(read_rvcw(path))

    def load_f0(self, algorithm: int = 1):
        self.rmvpe = nets.to_torch(read_rvcw(os.path.join(self.data_path, "f0", "rmvpe.rvcw")))

    def load_model(self, model_path: str):
        self.session = nets.to_torch(read_rvcw(model_path))

    def unload_model(self):
        self.session = None

    def set_index(self, rows: np.ndarray, index_rate: float):
        self.index = None if rows is None else np.ascontiguousarray(rows, np.float32)
        self.index_rate = float(index_rate)

    # -- inference ----------------------------------------------------------------------------
    def hubert(self, pcm: np.ndarray) -> np.ndarray:
        """rvc.rs:81-97: returns (1, C, T)."""
        if self.contentvec is None:
            raise ContentvecNotLoaded()
        out = nets.hubert_forward(self.contentvec, torch.from_numpy(np.asarray(pcm, np.float32)),
                                  self.rec)
        return np.ascontiguousarray(out.numpy().T[None])

    def extract_feature(self, pcm: np.ndarray) -> np.ndarray:
        """rvc.rs:99-109: returns (1, 2T+1, C)."""
        raw = self.hubert(pcm)[0].T
        return dsp.extend_feature_2x(raw)[None]

    def mel_extract(self, pcm: np.ndarray) -> np.ndarray:
        return self.mel.mel_extract(np.asarray(pcm, np.float32))

    def salience(self, pcm: np.ndarray, sample_frame_16k_size: int) -> np.ndarray:
        """rmvpe.rs:250-259 up to the network output (T,360)."""
        if self.rmvpe is None:
            raise F0NotLoaded()
        n = dsp.f0_extractor_frame(sample_frame_16k_size)
        if n > len(pcm):
            raise BadShape("input shorter than the f0 window")      # rmvpe.rs:257 panics
        mel = self.mel_extract(np.asarray(pcm, np.float32)[len(pcm) - n:])
        if self.rec is not None:
            self.rec["mel"] = mel.T.copy()
        # rmvpe.rs:227-233: T is a multiple of 32 by construction
        assert mel.shape[1] % 32 == 0
        return nets.rmvpe_forward(self.rmvpe, torch.from_numpy(mel), self.rec).numpy()

    def pitch(self, pcm: np.ndarray, pitch_shift: int, sample_frame_16k_size: int) -> np.ndarray:
        """rvc.rs:111-131 + rmvpe.rs:243-261."""
        sal = self.salience(pcm, sample_frame_16k_size)
        f0, c = dsp.decode(sal, 0.03, self.upstream_cents_window)
        self.last["argmax"] = c
        up = dsp.pitch_shift_factor(pitch_shift, self.upstream_pitch_shift)
        return (f0 * up).astype(np.float32)

    def infer(self, pcm: np.ndarray, sample_frame_16k_size: int, pitch_shift, skip_head: int,
              return_length: int) -> np.ndarray:
        """rvc.rs:133-220."""
        if self.session is None:
            raise ModelNotLoaded()
        pcm = np.asarray(pcm, np.float32)
        raw = self.hubert(pcm)[0].T                                    # (T,C)
        T = raw.shape[0]
        ext_len = 2 * T + 1
        hubert_length = min(len(pcm) // 160, ext_len)                  # rvc.rs:153
        if skip_head + return_length > ext_len or hubert_length > 1024:
            raise BadShape("skip_head + return_length exceeds the feature length")
        # kNN (rvc.rs:159 TODO; upstream semantics) on the 20 ms frames that survive the slice
        first = min(skip_head // 2, T - 1)
        last = min((skip_head + return_length - 1) // 2, T - 1)
        if self.index is not None and self.index_rate > 0.0 and return_length > 0:
            queries = raw[first:last + 1].copy()
            blended, d2, ix = knn.blend(self.index, queries, self.index_k, self.index_rate)
            raw = raw.copy()
            raw[first:last + 1] = blended
            self.last["knn_idx"] = ix
            self.last["knn_d2"] = d2
            self.last["knn_q"] = queries
        feats = dsp.extend_feature_2x(raw)
        phone = feats[skip_head:skip_head + return_length]             # rvc.rs:155
        shift = 0 if pitch_shift is None else int(pitch_shift)         # rvc.rs:163
        pitchf = self.pitch(pcm, shift, sample_frame_16k_size)
        sliced = self.cache.update_and_slice(pitchf, sample_frame_16k_size, hubert_length,
                                             skip_head, return_length)
        coarse, pf = dsp.get_f0_post(sliced, self.f0_mel_min, self.f0_mel_max)
        sr = int(self.session["meta.sr"][0])
        spf = sr // 100
        if self.noise_enabled:
            nz = noise.gauss(self.noise_seed, self.window, noise.KIND_Z, return_length * 192)
            ns = noise.gauss(self.noise_seed, self.window, noise.KIND_SINE, return_length * spf)
        else:
            nz = np.zeros(return_length * 192, np.float32)
            ns = np.zeros(return_length * spf, np.float32)
        self.window += 1
        self.last.update(phone=phone, pitch=coarse, pitchf=pf, f0=pitchf)
        audio = nets.synth_forward(self.session, torch.from_numpy(phone),
                                   torch.from_numpy(coarse), torch.from_numpy(pf),
                                   torch.from_numpy(nz.reshape(return_length, 192)),
                                   torch.from_numpy(ns), self.rec)
        return audio.numpy()


def synthetic_pcm(n: int, seed: int = 0) -> np.ndarray:
    """SURVEY section 8d primary synthetic input: harmonic 'voiced' glide 90->320 Hz, 8
    harmonics 1/h, 4 Hz AM, N(0,0.005^2) noise, 20 % unvoiced segments, peak 0.5."""
    rng = np.random.Generator(np.random.PCG64([seed, 16000]))
    t = np.arange(n) / 16000.0
    period = 1.7
    ph = (t % period) / period
    f0 = 90.0 + (320.0 - 90.0) * np.abs(2.0 * ph - 1.0)
    phase = 2.0 * np.pi * np.cumsum(f0) / 16000.0
    x = sum(np.sin(h * phase) / h for h in range(1, 9))
    x *= 0.6 + 0.4 * np.sin(2.0 * np.pi * 4.0 * t)
    seg = (t % 1.0) > 0.8                                   # 20 % unvoiced
    x = np.where(seg, 0.0, x)
    x = x + rng.standard_normal(n) * 0.005 * np.where(seg, 8.0, 1.0)
    x = 0.5 * x / np.abs(x).max()
    return x.astype(np.float32)


# BASELINE geometry (block 0.16 s): one call = the last 35 840 samples
BASELINE_GEOM = dict(n16k=35840, sf16k=2560, skip_head=200, return_length=21)
DEFAULT_GEOM = dict(n16k=38080, sf16k=4800, skip_head=200, return_length=35)
