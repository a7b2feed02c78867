"""TEST INFRASTRUCTURE (oracle).  The reference's streaming loop around one inference call, restated:
`RvcInferenceState` sizes (obs-rvc/src/lib.rs:200-245) and `process_one_frame` (lib.rs:659-795): ring buffers at the
OBS rate and at 16 kHz, rubato down / up sampling (oracle/resample.py - unpinned), RvcInfer::infer, envelope mixing,
SOLA offset search, sin^2 cross-fade (oracle/dsp.py - pinned by the reference's goldens)."""
import numpy as np

from . import dsp
from .resample import FftFixedInOut

F32 = np.float32


def rust_round(x: float) -> int:
    """f64::round: half away from zero."""
    return int(np.floor(abs(x) + 0.5) * (1 if x >= 0 else -1))


class StreamGeometry:
    """lib.rs:200-226."""

    def __init__(self, sample_rate=48000, sample_length=0.30, crossfade_length=0.07, extra_inference_time=2.0,
                 model_sample_rate=40000, skip_inference=False):
        self.sample_rate = sample_rate
        zc = self.zc = sample_rate // 100
        self.sample_frame_time = rust_round(sample_length * sample_rate / zc)
        self.sample_frame_size = self.sample_frame_time * zc
        self.sample_frame_16k = self.sample_frame_time * 160
        self.crossfade_frame_size = rust_round(crossfade_length * sample_rate / zc) * zc
        self.sola_buffer_frame_size = min(self.crossfade_frame_size, 4 * zc)
        self.sola_search_frame_size = zc
        self.extra_frame_size = rust_round(extra_inference_time * sample_rate / zc) * zc
        self.input_buffer_size = (self.extra_frame_size + self.crossfade_frame_size + self.sola_search_frame_size
                                  + self.sample_frame_size)
        self.input_buffer_16k_size = 160 * self.input_buffer_size // zc
        self.model_return_length = (self.sample_frame_size + self.sola_buffer_frame_size + self.sola_search_frame_size) // zc
        self.model_sample_rate = 16000 if skip_inference else model_sample_rate
        self.model_return_size = self.model_return_length * (self.model_sample_rate // 100)
        self.skip_head = self.extra_frame_size // zc
        self.skip_inference = skip_inference


class Stream:
    def __init__(self, engine, geom: StreamGeometry, pitch_shift=12, rms_mix_rate=0.0):
        self.g, self.engine, self.pitch_shift, self.rms_mix_rate = geom, engine, pitch_shift, rms_mix_rate
        g = geom
        self.input_buffer = np.zeros(g.input_buffer_size, F32)
        self.input_buffer_16k = np.zeros(g.input_buffer_16k_size, F32)
        self.sola_buffer = np.zeros(g.sola_buffer_frame_size, F32)
        self.down = FftFixedInOut(g.sample_rate, 16000, g.sample_frame_size + 2 * g.zc)
        self.up = FftFixedInOut(g.model_sample_rate, g.sample_rate, g.model_return_size)
        self.last = {}

    def process_one_frame(self, input_sample: np.ndarray) -> np.ndarray:
        g = self.g
        x = np.asarray(input_sample, F32)
        assert x.shape[0] == g.sample_frame_size
        self.input_buffer[:-g.sample_frame_size] = self.input_buffer[g.sample_frame_size:].copy()
        self.input_buffer[-g.sample_frame_size:] = x
        self.input_buffer_16k[:-g.sample_frame_16k] = self.input_buffer_16k[g.sample_frame_16k:].copy()
        res = self.down.process(self.input_buffer[-(g.sample_frame_size + 2 * g.zc):])
        copy_n = (g.sample_frame_size // g.zc + 1) * 160
        self.input_buffer_16k[-copy_n:] = res[160:]
        if g.skip_inference:
            out = self.input_buffer_16k[-g.model_return_size:].copy()
        else:
            out = self.engine.infer(self.input_buffer_16k, g.sample_frame_16k, self.pitch_shift, g.skip_head, g.model_return_length)
        self.last["model_out"] = np.asarray(out, F32).copy()
        out = self.up.process(np.asarray(out, F32))
        self.last["upsampled"] = out.copy()
        if self.rms_mix_rate < 1.0:
            out, _, _ = dsp.envelop_mixing(self.input_buffer[g.extra_frame_size:], out, g.sample_rate, self.rms_mix_rate)
        self.last["mixed"] = out.copy()
        block, self.sola_buffer, off = dsp.sola_crossfade(out, self.sola_buffer, g.sola_buffer_frame_size,
                                                          g.sola_search_frame_size, g.sample_frame_size)
        self.last["sola_offset"] = off
        return block
