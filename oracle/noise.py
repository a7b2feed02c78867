"""Counter-based Gaussian noise shared by the oracle and the engine (TEST INFRASTRUCTURE).

The reference bakes the synthesizer's random sources into the ONNX graph (the `rnd` input is
commented out, rvc.rs:186-191,200-203), so its output is not reproducible sample-for-sample.
For parity both sides draw the SAME noise from this stateless generator:
    key  = seed*K1 + window*K2 + kind*K3            (mod 2^64)
    x    = splitmix64(key + idx)
    u1   = ((x >> 40) + 0.5) / 2^24,   u2 = (((x >> 16) & 0xFFFFFF) + 0.5) / 2^24
    z    = sqrt(-2 ln u1) * cos(2 pi u2)            (float64, rounded to float32)
The engine evaluates the same formula in `csrc/kernels_misc.cu` (`noise_gauss`).
"""
import numpy as np

K1 = np.uint64(0x9E3779B97F4A7C15)
K2 = np.uint64(0xBF58476D1CE4E5B9)
K3 = np.uint64(0x94D049BB133111EB)

KIND_Z = 1        # enc_p posterior noise, index = t*192 + c (channels-last)
KIND_SINE = 2     # SineGen additive noise, index = output sample


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)).astype(np.uint64)
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)).astype(np.uint64)
        return (x ^ (x >> np.uint64(31))).astype(np.uint64)


def gauss(seed: int, window: int, kind: int, n: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        key = (np.uint64(seed) * K1 + np.uint64(window) * K2 + np.uint64(kind) * K3)
        x = splitmix64((key + np.arange(n, dtype=np.uint64)).astype(np.uint64))
    u1 = ((x >> np.uint64(40)).astype(np.float64) + 0.5) / 16777216.0
    u2 = (((x >> np.uint64(16)) & np.uint64(0xFFFFFF)).astype(np.float64) + 0.5) / 16777216.0
    return (np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)).astype(np.float32)
