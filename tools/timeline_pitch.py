"""Sparse timeline (RVC_TL_MARKS) of the F0 lane ALONE (rvc_pitch plan): the same segments tools/timeline.py reports
under ContentVec's concurrency, for comparison."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "obs-rvc_b200")]
import rvc_b200
from oracle import pipeline, weights
p = weights.make_data_dir("/tmp/rvc_b200_data_seed7", seed=7, index_rows=40000)
eng = rvc_b200.RvcInfer(p["data"]); eng.load_contentvec(2); eng.load_f0(1); eng.load_model(p["model"])
g = pipeline.BASELINE_GEOM
x = pipeline.synthetic_pcm(g["n16k"])
for _ in range(4): eng.pitch(x, 12, g["sf16k"])
tl = eng.profile_timeline()
prev = 0.0
for o in tl:
    print(f"{o['name']:22s} lane {o['lane']} end {o['end_us']:9.1f}  +{o['end_us'] - prev:8.1f}")
    prev = o["end_us"]
