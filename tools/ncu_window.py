"""Three warm windows + one more of configs[1] through the graph-replayed engine; prints the kernel launches of
one window so that an ncu launch list of this command can be cut to exactly the last window."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "obs-rvc_b200")]
import rvc_b200
from oracle import pipeline, weights
p = weights.make_data_dir("/tmp/rvc_b200_data_seed7", seed=7, index_rows=40000)
eng = rvc_b200.RvcInfer(p["data"]); eng.load_contentvec(2); eng.load_f0(1); eng.load_model(p["model"])
eng.load_index(p["index"], 0.5)
g = pipeline.BASELINE_GEOM
x = pipeline.synthetic_pcm(g["n16k"] + 8 * g["sf16k"])
n = 0
for i in range(4):
    l0 = eng.kernel_launches()
    eng.infer(x[i * g["sf16k"]: i * g["sf16k"] + g["n16k"]], g["sf16k"], 12, g["skip_head"], g["return_length"])
    n = eng.kernel_launches() - l0
print(f"LAUNCHES_PER_WINDOW={n}")
