"""Stand-alone device time of the two front-end lanes: rvc_hubert (ContentVec only) and rvc_pitch (RMVPE only)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "obs-rvc_b200")]
import numpy as np
import rvc_b200
from oracle import pipeline, weights
p = weights.make_data_dir("/tmp/rvc_b200_data_seed7", seed=7, index_rows=40000)
eng = rvc_b200.RvcInfer(p["data"]); eng.load_contentvec(2); eng.load_f0(1); eng.load_model(p["model"])
g = pipeline.BASELINE_GEOM
x = pipeline.synthetic_pcm(g["n16k"])
for name, fn in (("hubert", lambda: eng.hubert(x)), ("pitch", lambda: eng.pitch(x, 12, g["sf16k"]))):
    for _ in range(5): fn()
    ts = []
    for _ in range(30):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    print(f"LANE {name}: median {np.median(ts)*1e6:.0f} us (host-observed, incl. copies + sync)")
