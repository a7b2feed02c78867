"""Runs one eager window (optionally RVC_SYNC_EACH=1) and prints the error, if any."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "obs-rvc_b200")]
import rvc_b200
from oracle import pipeline, weights
p = weights.make_data_dir("/tmp/rvc_b200_data_seed7", seed=7)
eng = rvc_b200.RvcInfer(p["data"], use_cuda_graph=False)
eng.load_contentvec(2); eng.load_f0(1); eng.load_model(p["model"])
g = pipeline.BASELINE_GEOM
x = pipeline.synthetic_pcm(g["n16k"])
try:
    out = eng.infer(x, g["sf16k"], 12, g["skip_head"], g["return_length"])
    print("ok", float(abs(out).max()))
except Exception as e:
    print("FAILED:", e)
