"""One (or N) windows of configs[1] through the C ABI, eager or graph: the command profilers / sanitizers wrap."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "obs-rvc_b200")]
import numpy as np
import rvc_b200
from oracle import pipeline, weights
p = weights.make_data_dir("/tmp/rvc_b200_data_seed7", seed=7, index_rows=40000)
eng = rvc_b200.RvcInfer(p["data"], use_cuda_graph=os.environ.get("GRAPH", "0") == "1")
eng.load_contentvec(2); eng.load_f0(1); eng.load_model(p["model"]); eng.load_index(p["index"], 0.5)
g = pipeline.BASELINE_GEOM
x = pipeline.synthetic_pcm(g["n16k"] + 8 * g["sf16k"])
for i in range(int(os.environ.get("WINDOWS", 1))):
    a = eng.infer(x[i * g["sf16k"]: i * g["sf16k"] + g["n16k"]], g["sf16k"], 12, g["skip_head"], g["return_length"])
print("ONE_WINDOW ok rms", float(np.sqrt(np.mean(a ** 2))), "launches", eng.kernel_launches(), eng.plan_info())
eng.close()
