import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "obs-rvc_b200")]
import rvc_b200
from oracle import pipeline, weights
p = weights.make_data_dir("/tmp/rvc_b200_data_seed7", seed=7)
eng = rvc_b200.RvcInfer(p["data"]); eng.load_contentvec(2); eng.load_f0(1); eng.load_model(p["model"])
g = pipeline.BASELINE_GEOM
x = pipeline.synthetic_pcm(g["n16k"])
eng.infer(x, g["sf16k"], 12, g["skip_head"], g["return_length"])
L = rvc_b200.lib()
for name in sys.argv[1:] or ["cv.conv6", "cv.conv1", "cv.L0.fc2", "cv.L0.qkv", "sy.U2.rb1.0.c1"]:
    out = (ctypes.c_longlong * 96)()
    rc = L.rvc_debug_umma_timing(eng.handle, name.encode(), out)
    t = np.array(list(out), dtype=np.int64)
    print(name, rc, "cycles since start:", [int(v - t[0]) for v in t[:16]])
    if not name.endswith("@v2"):
        ev = t[16:96].reshape(5, 16) - t[0]
        for lab, row in zip(("prod empty-ok", "prod tma-issued", "mma full-ok", "mma conv-ok", "mma issued"), ev):
            print(f"    {lab:16s}", [int(v) for v in row[:13]])
