"""Host-side mirror of the reference's `rvc` crate public API over the C ABI.

`RvcInfer` keeps the method set, argument meaning and error behaviour of `rvc::RvcInfer`
(/root/reference/rvc/src/rvc.rs:18-220; exported by rvc/src/lib.rs:5) and of the adapter the
OBS plugin calls (obs-rvc/src/rvcadapter.rs:33-67), but every call goes straight into
`librvc_b200.so` (include/rvc_b200.h) - hand-written sm_100a CUDA, no ONNX Runtime, no PyTorch,
no CPU fallback.  If the shared library or a CUDA device is missing this module raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_float, c_int32, c_size_t, c_uint32, c_uint64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "librvc_b200.so")

# rvc-common/src/enums.rs:3-28,32-39,97-103
MODEL_V1, MODEL_V2 = 1, 2
PITCH_RMVPE = 1


class RvcInferError(Exception):
    """rvc-common/src/errors.rs:1-20."""
    code = -1


class ModelNotLoaded(RvcInferError):
    code = 1


class ContentvecNotLoaded(RvcInferError):
    code = 2


class F0NotLoaded(RvcInferError):
    code = 3


class CudaError(RvcInferError):       # replaces RvcInferError::Ort
    code = 4


class BadShape(RvcInferError):        # RvcInferError::NdarrayShapeError / reference panics
    code = 5


class IoError(RvcInferError):
    code = 6


class InvalidArg(RvcInferError):
    code = 7


_ERRORS = {c.code: c for c in (ModelNotLoaded, ContentvecNotLoaded, F0NotLoaded, CudaError,
                               BadShape, IoError, InvalidArg)}


class StreamConfig(ctypes.Structure):
    """`rvc_stream_config` (include/rvc_b200.h): the OBS filter settings that size the streaming loop (lib.rs:186-226)."""
    _fields_ = [("sample_rate", ctypes.c_uint32), ("pitch_shift", c_int32), ("sample_length", ctypes.c_double),
                ("crossfade_length", ctypes.c_double), ("extra_inference_time", ctypes.c_double),
                ("rms_mix_rate", ctypes.c_double), ("skip_inference", c_int32), ("reserved", c_int32 * 7)]


class Config(ctypes.Structure):
    """`rvc_config` (include/rvc_b200.h)."""
    _fields_ = [("device", c_int32), ("noise_mode", c_int32), ("noise_seed", c_uint64),
                ("index_k", c_int32), ("upstream_pitch_shift", c_int32),
                ("upstream_cents_window", c_int32), ("use_cuda_graph", c_int32),
                ("debug_keep", c_int32), ("reserved", c_int32 * 7)]


_lib = None


def lib():
    """Loads librvc_b200.so; fails loudly when it has not been built (no fallback path)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing - run `python -c 'import __graft_entry__ as g; "
                              "g.build()'` (the engine has no CPU/PyTorch fallback)")
        L = ctypes.CDLL(LIB_PATH)
        L.rvc_version.restype = c_char_p
        L.rvc_last_error.restype = c_char_p
        L.rvc_last_error.argtypes = [c_void_p]
        L.rvc_last_create_error.restype = c_char_p
        L.rvc_cuda_stream.restype = c_void_p
        L.rvc_cuda_stream.argtypes = [c_void_p]
        L.rvc_destroy.argtypes = [c_void_p]
        L.rvc_destroy.restype = None
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class RvcInfer:
    """Drop-in for `rvc::RvcInfer`."""

    def __init__(self, data_path: str, device: int = 0, noise_seed: int = 0, noise_mode: int = 1,
                 index_k: int = 8, upstream_pitch_shift: bool = False,
                 upstream_cents_window: bool = False, use_cuda_graph: bool = True):
        """RvcInfer::new(data_path) - rvc.rs:30-44."""
        L = lib()
        cfg = Config()
        L.rvc_config_default(byref(cfg))
        cfg.device, cfg.noise_seed, cfg.noise_mode, cfg.index_k = device, noise_seed, noise_mode, index_k
        cfg.upstream_pitch_shift = int(upstream_pitch_shift)
        cfg.upstream_cents_window = int(upstream_cents_window)
        cfg.use_cuda_graph = int(use_cuda_graph)
        self._h = c_void_p()
        rc = L.rvc_create(str(data_path).encode(), byref(cfg), byref(self._h))
        if rc != 0:
            raise _ERRORS.get(rc, RvcInferError)(L.rvc_last_create_error().decode())
        self._L = L

    # ------------------------------------------------------------------ plumbing
    def _chk(self, rc):
        if rc != 0:
            raise _ERRORS.get(rc, RvcInferError)(self._L.rvc_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.rvc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    # ------------------------------------------------------------------ loading
    def load_contentvec(self, model_version: int = MODEL_V2):
        """rvc.rs:46-54."""
        self._chk(self._L.rvc_load_contentvec(self._h, c_int32(model_version)))

    def load_f0(self, pitch_algorithm: int = PITCH_RMVPE):
        """rvc.rs:62-75."""
        self._chk(self._L.rvc_load_f0(self._h, c_int32(pitch_algorithm)))

    def load_model(self, model_path: str):
        """rvc.rs:56-60."""
        self._chk(self._L.rvc_load_model(self._h, str(model_path).encode()))

    def unload_model(self):
        """rvc.rs:77-79."""
        self._chk(self._L.rvc_unload_model(self._h))

    def load_index(self, index_path: str, index_rate: float):
        self._chk(self._L.rvc_load_index(self._h, str(index_path).encode(), c_float(index_rate)))

    def set_index(self, rows, index_rate: float):
        if rows is None:
            self._chk(self._L.rvc_set_index(self._h, None, c_size_t(0), c_size_t(0), c_float(index_rate)))
            return
        rows = _f32(rows)
        self._chk(self._L.rvc_set_index(self._h, rows.ctypes.data_as(c_void_p), c_size_t(rows.shape[0]),
                                        c_size_t(rows.shape[1]), c_float(index_rate)))

    def set_index_rate(self, index_rate: float):
        self._chk(self._L.rvc_set_index_rate(self._h, c_float(index_rate)))

    # ------------------------------------------------------------------ inference
    def hubert(self, pcm) -> np.ndarray:
        """rvc.rs:81-97 -> (1, C, T)."""
        pcm = _f32(pcm)
        cap = 1024 * (pcm.shape[0] // 320 + 2)
        out = np.empty(cap, np.float32)
        c, t = c_size_t(), c_size_t()
        self._chk(self._L.rvc_hubert(self._h, pcm.ctypes.data_as(c_void_p), c_size_t(pcm.shape[0]),
                                     out.ctypes.data_as(c_void_p), c_size_t(cap), byref(c), byref(t)))
        return out[:c.value * t.value].reshape(1, c.value, t.value).copy()

    def extract_feature(self, pcm) -> np.ndarray:
        """rvc.rs:99-109 -> (1, 2T+1, C)."""
        pcm = _f32(pcm)
        cap = 1024 * (2 * (pcm.shape[0] // 320 + 2) + 1)
        out = np.empty(cap, np.float32)
        fr, c = c_size_t(), c_size_t()
        self._chk(self._L.rvc_extract_feature(self._h, pcm.ctypes.data_as(c_void_p),
                                              c_size_t(pcm.shape[0]), out.ctypes.data_as(c_void_p),
                                              c_size_t(cap), byref(fr), byref(c)))
        return out[:fr.value * c.value].reshape(1, fr.value, c.value).copy()

    def pitch(self, pcm, pitch_shift: int, sample_frame_16k_size: int) -> np.ndarray:
        """rvc.rs:111-131 -> f0[T] Hz."""
        pcm = _f32(pcm)
        out = np.empty(4096, np.float32)
        n = c_size_t()
        self._chk(self._L.rvc_pitch(self._h, pcm.ctypes.data_as(c_void_p), c_size_t(pcm.shape[0]),
                                    c_int32(pitch_shift), c_size_t(sample_frame_16k_size),
                                    out.ctypes.data_as(c_void_p), c_size_t(out.shape[0]), byref(n)))
        return out[:n.value].copy()

    def infer(self, pcm, sample_frame_16k_size: int, pitch_shift, skip_head: int,
              return_length: int, out: np.ndarray | None = None) -> np.ndarray:
        """rvc.rs:133-220 (`pitch_shift=None` -> 0, rvc.rs:163)."""
        pcm = _f32(pcm)
        if out is None:
            out = np.empty(return_length * 480 + 16, np.float32)
        n = c_size_t()
        self._chk(self._L.rvc_infer(self._h, pcm.ctypes.data_as(c_void_p), c_size_t(pcm.shape[0]),
                                    c_uint32(sample_frame_16k_size),
                                    c_int32(0 if pitch_shift is None else int(pitch_shift)),
                                    c_uint32(skip_head), c_uint32(return_length),
                                    out.ctypes.data_as(c_void_p), c_size_t(out.shape[0]), byref(n)))
        return out[:n.value]

    def infer_ptr(self, pcm_ptr: int, n: int, sample_frame_16k_size: int, pitch_shift: int,
                  skip_head: int, return_length: int, out_ptr: int, cap: int, device: bool) -> int:
        """Raw-pointer form (pinned host or device memory) used by bench.py."""
        ln = c_size_t()
        fn = self._L.rvc_infer_dev if device else self._L.rvc_infer
        self._chk(fn(self._h, c_void_p(pcm_ptr), c_size_t(n), c_uint32(sample_frame_16k_size),
                     c_int32(pitch_shift), c_uint32(skip_head), c_uint32(return_length),
                     c_void_p(out_ptr), c_size_t(cap), byref(ln)))
        return ln.value

    def infer_windows(self, pcm, n: int, sample_frame_16k_size: int, n_windows: int, pitch_shift, skip_head: int,
                      return_length: int, max_batch: int = 0) -> np.ndarray:
        """Offline conversion of this stream (BASELINE configs[2]): window w = pcm[w*sf16k, w*sf16k + n), `max_batch`
        (<= 32) windows per launch; same results as n_windows successive infer() calls.  -> (n_windows, audio_len)."""
        pcm = _f32(pcm)
        out = np.empty(n_windows * (return_length * 480 + 16), np.float32)
        al = c_size_t()
        self._chk(self._L.rvc_infer_windows(self._h, pcm.ctypes.data_as(c_void_p), c_size_t(pcm.shape[0]), c_size_t(n),
                                            c_uint32(sample_frame_16k_size), c_size_t(n_windows),
                                            c_int32(0 if pitch_shift is None else int(pitch_shift)), c_uint32(skip_head),
                                            c_uint32(return_length), out.ctypes.data_as(c_void_p), c_size_t(out.shape[0]),
                                            byref(al), c_int32(max_batch)))
        return out[:n_windows * al.value].reshape(n_windows, al.value)

    def infer_windows_ptr(self, pcm_ptr: int, n_pcm: int, n: int, sample_frame_16k_size: int, n_windows: int, pitch_shift: int,
                          skip_head: int, return_length: int, out_ptr: int, cap: int, device: bool, max_batch: int = 0) -> int:
        al = c_size_t()
        fn = self._L.rvc_infer_windows_dev if device else self._L.rvc_infer_windows
        self._chk(fn(self._h, c_void_p(pcm_ptr), c_size_t(n_pcm), c_size_t(n), c_uint32(sample_frame_16k_size), c_size_t(n_windows),
                     c_int32(pitch_shift), c_uint32(skip_head), c_uint32(return_length), c_void_p(out_ptr), c_size_t(cap),
                     byref(al), c_int32(max_batch)))
        return al.value

    def get_last_window(self, window: int, name: str, dtype=np.float32) -> np.ndarray:
        nb = c_size_t()
        self._chk(self._L.rvc_get_last_window(self._h, c_int32(window), name.encode(), None, c_size_t(0), byref(nb)))
        out = np.empty(nb.value // 4, dtype)
        self._chk(self._L.rvc_get_last_window(self._h, c_int32(window), name.encode(), out.ctypes.data_as(c_void_p),
                                              c_size_t(nb.value), byref(nb)))
        return out

    # ------------------------------------------------------------------ extras
    def mel_extract(self, pcm) -> np.ndarray:
        """rmvpe.rs:159-205 -> (128, T)."""
        pcm = _f32(pcm)
        cap = 128 * (pcm.shape[0] // 160 + 2)
        out = np.empty(cap, np.float32)
        t = c_size_t()
        self._chk(self._L.rvc_mel_extract(self._h, pcm.ctypes.data_as(c_void_p), c_size_t(pcm.shape[0]),
                                          out.ctypes.data_as(c_void_p), c_size_t(cap), byref(t)))
        return out[:128 * t.value].reshape(128, t.value).copy()

    # ------------------------------------------------------------------ streaming loop (obs-rvc/src/lib.rs:186-300, 659-795)
    def stream_open(self, sample_rate=48000, sample_length=0.30, crossfade_length=0.07, extra_inference_time=2.0,
                    pitch_shift=12, rms_mix_rate=0.0, skip_inference=False) -> int:
        """Builds RvcInferenceState on the device; returns sample_frame_size (samples per process_frame call)."""
        cfg = StreamConfig()
        self._L.rvc_stream_config_default(byref(cfg))
        cfg.sample_rate, cfg.pitch_shift, cfg.sample_length = sample_rate, pitch_shift, sample_length
        cfg.crossfade_length, cfg.extra_inference_time = crossfade_length, extra_inference_time
        cfg.rms_mix_rate, cfg.skip_inference = rms_mix_rate, int(skip_inference)
        n = ctypes.c_uint32()
        self._chk(self._L.rvc_stream_open(self._h, byref(cfg), byref(n)))
        self._frame = int(n.value)
        return self._frame

    def stream_info(self) -> dict:
        import json
        buf = ctypes.create_string_buffer(1024)
        n = c_size_t()
        self._chk(self._L.rvc_stream_info(self._h, buf, c_size_t(1024), byref(n)))
        return json.loads(buf.value.decode())

    def stream_close(self):
        self._chk(self._L.rvc_stream_close(self._h))

    def process_frame(self, block):
        """process_one_frame (lib.rs:659-795): sample_frame_size samples in, sample_frame_size samples out."""
        x = _f32(block)
        assert x.shape[0] == self._frame
        out = np.empty(self._frame, np.float32)
        off = ctypes.c_uint32()
        self._chk(self._L.rvc_process_frame(self._h, x.ctypes.data_as(c_void_p), out.ctypes.data_as(c_void_p), byref(off)))
        self.last_sola_offset = int(off.value)
        return out

    def resample_chunk(self, fs_in: int, fs_out: int, chunk, overlap):
        """rubato FftFixedInOut::process on one chunk; `overlap` ([n_out] f32) is updated in place."""
        x = _f32(chunk)
        cap = int(x.shape[0] * fs_out // fs_in + 16)
        out = np.empty(cap, np.float32)
        n = c_size_t()
        assert overlap.dtype == np.float32 and overlap.flags["C_CONTIGUOUS"]
        self._chk(self._L.rvc_resample_chunk(self._h, ctypes.c_uint32(fs_in), ctypes.c_uint32(fs_out), x.ctypes.data_as(c_void_p),
                                             c_size_t(x.shape[0]), overlap.ctypes.data_as(c_void_p), out.ctypes.data_as(c_void_p),
                                             c_size_t(cap), byref(n)))
        return out[:n.value].copy()

    def decode_salience(self, salience):
        """rmvpe.rs:118-133, 243-248 on given salience rows (T, 360) -> (f0[T] f32, argmax[T] i32)."""
        sal = _f32(salience)
        assert sal.ndim == 2 and sal.shape[1] == 360
        f0 = np.empty(sal.shape[0], np.float32)
        am = np.empty(sal.shape[0], np.int32)
        self._chk(self._L.rvc_decode_salience(self._h, sal.ctypes.data_as(c_void_p), c_size_t(sal.shape[0]),
                                              f0.ctypes.data_as(c_void_p), am.ctypes.data_as(c_void_p)))
        return f0, am

    def knn_search(self, queries, k: int):
        q = _f32(queries)
        d2 = np.empty((q.shape[0], k), np.float32)
        idx = np.empty((q.shape[0], k), np.int32)
        self._chk(self._L.rvc_knn_search(self._h, q.ctypes.data_as(c_void_p), c_size_t(q.shape[0]),
                                         c_size_t(q.shape[1]), c_int32(k), d2.ctypes.data_as(c_void_p),
                                         idx.ctypes.data_as(c_void_p)))
        return d2, idx

    def knn_fallbacks(self) -> int:
        """Queries whose tensor-core candidate set failed its guard and were recomputed by the exact scan."""
        n = c_uint64()
        self._chk(self._L.rvc_knn_fallbacks(self._h, byref(n)))
        return n.value

    # ------------------------------------------------------------------ streaming glue (next row #1)
    def envelop_mixing(self, inp, out, sample_rate: int, mix_rate: float, want_rms: bool = False):
        """rt_utils::envelop_mixing (obs-rvc/src/rt_utils.rs:119-132); returns the mixed output
        (and the interpolated rms1/rms2 when `want_rms`)."""
        inp, out = _f32(inp), _f32(out).copy()
        n = out.shape[0]
        r1 = np.empty(n, np.float32) if want_rms else None
        r2 = np.empty(n, np.float32) if want_rms else None
        self._chk(self._L.rvc_envelop_mixing(self._h, inp.ctypes.data_as(c_void_p), c_size_t(inp.shape[0]),
                                             out.ctypes.data_as(c_void_p), c_size_t(n), c_uint32(sample_rate),
                                             ctypes.c_double(mix_rate),
                                             r1.ctypes.data_as(c_void_p) if want_rms else None,
                                             r2.ctypes.data_as(c_void_p) if want_rms else None))
        return (out, r1, r2) if want_rms else out

    def sola_offset(self, input_buffer, sola_buffer, buffer_frame_size: int, search_frame_size: int) -> int:
        """rt_utils::get_sola_offset (obs-rvc/src/rt_utils.rs:60-90)."""
        x, sb = _f32(input_buffer), _f32(sola_buffer)
        off = c_uint32()
        self._chk(self._L.rvc_sola_offset(self._h, x.ctypes.data_as(c_void_p), c_size_t(x.shape[0]),
                                          sb.ctypes.data_as(c_void_p), c_uint32(buffer_frame_size),
                                          c_uint32(search_frame_size), byref(off)))
        return off.value

    def sola_crossfade(self, infer_out, sola_buffer, buffer_frame_size: int, search_frame_size: int,
                       sample_frame_size: int):
        """SOLA tail of process_one_frame (obs-rvc/src/lib.rs:768-794): returns (block, new sola_buffer, offset)."""
        x, sb = _f32(infer_out), _f32(sola_buffer).copy()
        block = np.empty(sample_frame_size, np.float32)
        off = c_uint32()
        self._chk(self._L.rvc_sola_crossfade(self._h, x.ctypes.data_as(c_void_p), c_size_t(x.shape[0]),
                                             sb.ctypes.data_as(c_void_p), c_uint32(buffer_frame_size),
                                             c_uint32(search_frame_size), c_uint32(sample_frame_size),
                                             block.ctypes.data_as(c_void_p), byref(off)))
        return block, sb, off.value

    def get_last(self, name: str, dtype=np.float32) -> np.ndarray:
        nb = c_size_t()
        self._chk(self._L.rvc_get_last(self._h, name.encode(), None, c_size_t(0), byref(nb)))
        out = np.empty(nb.value // 4, dtype)
        self._chk(self._L.rvc_get_last(self._h, name.encode(), out.ctypes.data_as(c_void_p),
                                       c_size_t(nb.value), byref(nb)))
        return out

    def buffer_names(self):
        nb = c_size_t()
        self._chk(self._L.rvc_debug_list(self._h, None, c_size_t(0), byref(nb)))
        buf = ctypes.create_string_buffer(nb.value + 1)
        self._chk(self._L.rvc_debug_list(self._h, buf, c_size_t(nb.value + 1), byref(nb)))
        return [s for s in buf.value.decode().split("\n") if s]

    def plan_info(self) -> dict:
        import json
        buf = ctypes.create_string_buffer(1024)
        nb = c_size_t()
        self._chk(self._L.rvc_plan_info(self._h, buf, c_size_t(1024), byref(nb)))
        return json.loads(buf.value.decode())

    def event_record(self, slot: int):
        self._chk(self._L.rvc_event_record(self._h, c_int32(slot)))

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = c_float()
        self._chk(self._L.rvc_event_elapsed_ms(self._h, c_int32(a), c_int32(b), byref(ms)))
        return ms.value

    def profile_ops(self, iters: int = 20) -> list:
        import json
        cap = 1 << 20
        buf = ctypes.create_string_buffer(cap)
        nb = c_size_t()
        self._chk(self._L.rvc_profile_ops(self._h, c_int32(iters), buf, c_size_t(cap), byref(nb)))
        return json.loads(buf.value.decode())

    def profile_timeline(self) -> list:
        import json
        cap = 1 << 20
        buf = ctypes.create_string_buffer(cap)
        nb = c_size_t()
        self._chk(self._L.rvc_profile_timeline(self._h, buf, c_size_t(cap), byref(nb)))
        return json.loads(buf.value.decode())

    def profile_chains(self) -> list:
        import json
        cap = 1 << 20
        buf = ctypes.create_string_buffer(cap)
        nb = c_size_t()
        self._chk(self._L.rvc_profile_chains(self._h, buf, c_size_t(cap), byref(nb)))
        return json.loads(buf.value.decode())

    def reset_state(self):
        self._chk(self._L.rvc_reset_state(self._h))

    def sync(self):
        self._chk(self._L.rvc_sync(self._h))

    def cuda_stream(self) -> int:
        return int(self._L.rvc_cuda_stream(self._h) or 0)

    def kernel_launches(self) -> int:
        n = c_uint64()
        self._chk(self._L.rvc_kernel_launches(self._h, byref(n)))
        return n.value


def infer_batch(engines, pcms, sample_frame_16k_size: int, pitch_shift: int, skip_head: int, return_length: int):
    """rvc_infer_batch: one window of each of `engines` (independent live streams, BASELINE configs[3]) in one call.
    Streams of one device that share their models run as ONE batched plan.  -> list of audio arrays."""
    L = lib()
    n_ctx = len(engines)
    pcms = [_f32(p) for p in pcms]
    n = pcms[0].shape[0]
    assert all(p.shape[0] == n for p in pcms)
    cap = return_length * 480 + 16
    outs = [np.empty(cap, np.float32) for _ in engines]
    H = (c_void_p * n_ctx)(*[e.handle for e in engines])
    P = (c_void_p * n_ctx)(*[p.ctypes.data_as(c_void_p) for p in pcms])
    O = (c_void_p * n_ctx)(*[o.ctypes.data_as(c_void_p) for o in outs])
    ln = c_size_t()
    rc = L.rvc_infer_batch(H, c_size_t(n_ctx), P, c_size_t(n), c_uint32(sample_frame_16k_size), c_int32(pitch_shift),
                           c_uint32(skip_head), c_uint32(return_length), O, c_size_t(cap), byref(ln))
    if rc != 0:
        raise _ERRORS.get(rc, RvcInferError)(L.rvc_last_error(engines[0].handle).decode())
    return [o[:ln.value] for o in outs]


def infer_batch_ptr(engines, pcm_ptrs, n: int, sample_frame_16k_size: int, pitch_shift: int, skip_head: int, return_length: int,
                    out_ptrs, cap: int, device: bool) -> int:
    """Raw-pointer form of rvc_infer_batch / rvc_infer_batch_dev (bench.py)."""
    L = lib()
    n_ctx = len(engines)
    H = (c_void_p * n_ctx)(*[e.handle for e in engines])
    P = (c_void_p * n_ctx)(*[c_void_p(p) for p in pcm_ptrs])
    O = (c_void_p * n_ctx)(*[c_void_p(o) for o in out_ptrs])
    ln = c_size_t()
    fn = L.rvc_infer_batch_dev if device else L.rvc_infer_batch
    rc = fn(H, c_size_t(n_ctx), P, c_size_t(n), c_uint32(sample_frame_16k_size), c_int32(pitch_shift), c_uint32(skip_head),
            c_uint32(return_length), O, c_size_t(cap), byref(ln))
    if rc != 0:
        raise _ERRORS.get(rc, RvcInferError)(L.rvc_last_error(engines[0].handle).decode())
    return ln.value
