"""CPU tests of the streaming-loop restatements (oracle/resample.py, oracle/stream.py): the rubato FftFixedInOut
restatement has no golden in the reference (parity unpinned), so it is held to the properties a synchronous FFT
resampler must have; the loop geometry is checked against the reference's formulas (obs-rvc/src/lib.rs:200-226)."""
import numpy as np
import pytest

from oracle.resample import FftFixedInOut
from oracle.stream import Stream, StreamGeometry


@pytest.mark.parametrize("fs_in,fs_out,chunk", [(48000, 16000, 15360), (40000, 48000, 14000), (32000, 48000, 3200)])
def test_resampler_delays_a_bandlimited_sine_by_half_a_chunk(fs_in, fs_out, chunk):
    r = FftFixedInOut(fs_in, fs_out, chunk)
    assert r.fft_size_in == chunk and r.fft_size_out * fs_in == chunk * fs_out
    n = 6
    t = np.arange(n * chunk) / fs_in
    x = np.sin(2 * np.pi * 997.0 * t).astype(np.float32)
    y = np.concatenate([r.process(x[i * chunk:(i + 1) * chunk]) for i in range(n)])
    to = (np.arange(y.shape[0]) * fs_in / fs_out - chunk / 2) / fs_in          # group delay = fft_size_in / 2 input samples
    ref = np.sin(2 * np.pi * 997.0 * to)
    assert np.abs(y[2 * r.fft_size_out:] - ref[2 * r.fft_size_out:]).max() < 2e-6


def test_resampler_is_linear_and_rejects_above_the_new_nyquist():
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal(15360).astype(np.float32), rng.standard_normal(15360).astype(np.float32)
    ya = FftFixedInOut(48000, 16000, 15360).process(a)
    yb = FftFixedInOut(48000, 16000, 15360).process(b)
    yab = FftFixedInOut(48000, 16000, 15360).process((2 * a - b).astype(np.float32))
    assert np.abs(yab - (2 * ya - yb)).max() < 1e-5
    r = FftFixedInOut(48000, 16000, 15360)
    t = np.arange(4 * 15360) / 48000.0
    hi = np.sin(2 * np.pi * 11000.0 * t).astype(np.float32)                    # above 8 kHz: must vanish
    y = np.concatenate([r.process(hi[i * 15360:(i + 1) * 15360]) for i in range(4)])
    assert np.abs(y[2 * 5120:]).max() < 1e-4


def test_stream_geometry_follows_the_reference_formulas():
    g = StreamGeometry()                                                        # OBS defaults: 48 kHz, 0.30 s, 0.07 s, 2.0 s
    assert (g.zc, g.sample_frame_time, g.sample_frame_size, g.sample_frame_16k) == (480, 30, 14400, 4800)
    assert (g.crossfade_frame_size, g.sola_buffer_frame_size, g.sola_search_frame_size, g.extra_frame_size) == (3360, 1920, 480, 96000)
    assert (g.input_buffer_size, g.input_buffer_16k_size) == (114240, 38080)
    assert (g.model_return_length, g.model_return_size, g.skip_head) == (35, 14000, 200)
    # the same numbers the per-call geometry of the engine tests uses (oracle/pipeline.py DEFAULT_GEOM)
    from oracle import pipeline
    d = pipeline.DEFAULT_GEOM
    assert (d["n16k"], d["sf16k"], d["skip_head"], d["return_length"]) == (g.input_buffer_16k_size, g.sample_frame_16k, g.skip_head, g.model_return_length)


def test_stream_loop_passthrough_level_and_block_structure():
    """skip_inference (lib.rs:198,706-708): resample down -> 16 kHz tail -> resample up -> SOLA.  The reference feeds its
    STATEFUL up-sampler overlapping chunks (every frame the last sample_frame + sola_buffer + sola_search samples,
    lib.rs:742-756), so each emitted block is [second half of the previous chunk | first half of this chunk] with a
    time step back of sola_buffer + sola_search samples in the middle - restated literally, not repaired.  Checked
    here: the level of a steady sine survives and exactly one such seam exists per block."""
    g = StreamGeometry(sample_length=0.16, crossfade_length=0.04, extra_inference_time=0.5, skip_inference=True)
    s = Stream(None, g, rms_mix_rate=1.0)
    n = g.sample_frame_size
    t = np.arange(12 * n) / 48000.0
    x = (0.5 * np.sin(2 * np.pi * 220.0 * t)).astype(np.float32)
    out = np.concatenate([s.process_one_frame(x[i * n:(i + 1) * n]) for i in range(12)])
    tail = out[8 * n:]
    assert 0.3 < np.sqrt(np.mean(tail ** 2)) < 0.4                              # 0.5 / sqrt(2)
    slope = 0.5 * 2 * np.pi * 220.0 / 48000.0
    jumps = np.nonzero(np.abs(np.diff(tail)) > 3 * slope)[0]
    seams = 1 + int(np.count_nonzero(np.diff(jumps) > 100))                     # clusters of steep samples
    assert 1 <= seams <= 4                                                      # one seam per emitted block, nothing else
