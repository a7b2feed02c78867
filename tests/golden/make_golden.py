"""Builds tests/golden/ from the reference tree (run in the dev container only).

The reference's own golden vectors / fixtures for the hot path are
copied verbatim (data files, not sources) so that the CPU test-suite and the GPU box - which
has no /root/reference - can pin the oracle against them.  Provenance is recorded in
MANIFEST.json.  Usage:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import shutil

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

COPIES = {
    # reference path -> fixture name
    "vendor/mel-spec/mel_spec/testdata/mel_filters.npz": "mel_filters.npz",   # mel.rs:318-330
    "rvc/src/tests/input_wav.npy": "input_wav.npy",                           # hubert.rs:14
    "rvc/src/tests/input_wav2.npy": "input_wav2.npy",                         # pitch.rs:25
    "obs-rvc/src/tests/infer_wav.npy": "sola_infer_wav.npy",                  # sola.rs:12
    "obs-rvc/src/tests/sola_buffer.npy": "sola_buffer.npy",                   # sola.rs:13
    "obs-rvc/src/tests/envelop_input_wav.npy": "envelop_input_wav.npy",       # envelop_mixing.rs:10
    "obs-rvc/src/tests/envelop_infer_wav.npy": "envelop_infer_wav.npy",
    "obs-rvc/src/tests/envelop_rms1.npy": "envelop_rms1.npy",
    "obs-rvc/src/tests/envelop_rms2.npy": "envelop_rms2.npy",
    "obs-rvc/src/tests/envelop_infer_wav2.npy": "envelop_infer_wav2.npy",
}


def main():
    manifest = {}
    for src, dst in COPIES.items():
        s = os.path.join(REF, src)
        d = os.path.join(HERE, dst)
        shutil.copyfile(s, d)
        os.chmod(d, 0o644)
        manifest[dst] = {"from": src, "sha256": hashlib.sha256(open(d, "rb").read()).hexdigest()}
    # feats.npy (734 KB) pins only geometry without pretrained weights: keep shape + rms.
    feats = np.load(os.path.join(REF, "rvc/src/tests/feats.npy"))
    manifest["feats_meta"] = {
        "from": "rvc/src/tests/feats.npy", "shape": list(feats.shape), "dtype": str(feats.dtype),
        "rms": float(np.sqrt((feats.astype(np.float64) ** 2).mean())),
        "input": "input_wav.npy", "note": "hubert.rs:10-19; 239 = 2*119+1, 119 = (38240-400)/320+1",
    }
    # 2.5 s of real speech from the mel-spec test WAV (f32le mono 16 kHz; 'data' chunk payload
    # starts at byte 114: RIFF(12) + fmt(8+40) + fact(8+4) + LIST(8+26) + data header(8))
    raw = np.fromfile(os.path.join(REF, "vendor/mel-spec/testdata/jfk_f32le.wav"), dtype=np.uint8)
    off = 114 + 4 * 8000   # skip the first 0.5 s of near-silence
    pcm = raw[off:off + 4 * 40000].copy().view(np.float32).copy()
    np.save(os.path.join(HERE, "jfk_2p5s.npy"), pcm)
    manifest["jfk_2p5s.npy"] = {"from": "vendor/mel-spec/testdata/jfk_f32le.wav",
                                "note": "samples [8000, 48000) of the data chunk"}
    json.dump(manifest, open(os.path.join(HERE, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    print("wrote", len(manifest), "entries")


if __name__ == "__main__":
    main()
