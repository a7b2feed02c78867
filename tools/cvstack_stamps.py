"""clock64 stamps of CTA 0 inside the persistent ContentVec stack kernel: per phase worker start / work done / arrived."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "obs-rvc_b200")]
import rvc_b200
from oracle import pipeline, weights
p = weights.make_data_dir("/tmp/rvc_b200_data_seed7", seed=7, index_rows=40000)
eng = rvc_b200.RvcInfer(p["data"]); eng.load_contentvec(2)
g = pipeline.BASELINE_GEOM
x = pipeline.synthetic_pcm(g["n16k"])
for _ in range(3): eng.hubert(x)
L = rvc_b200.lib()
out = (ctypes.c_longlong * 512)()
n = L.rvc_debug_cvstack_stamps(eng.handle, out, ctypes.c_int(512))
t = np.array(list(out), dtype=np.int64).reshape(128, 4)
print("phases", n)
names = ["qkv", "attn", "o", "ln1", "fc1", "fc2", "ln2"]
for ph in range(min(n, 23)):
    nm = "enc_in" if ph == 0 else names[(ph - 1) % 7]
    nxt = t[ph + 1][0] if ph + 1 < n else t[ph][2]
    print(f"ph{ph:3d} {nm:7s} start@{t[ph][0]-t[0][0]:8d} gridwait {t[ph][3]-t[ph][0]:7d} work {t[ph][1]-t[ph][3]:7d} arrive +{t[ph][2]-t[ph][1]:6d} next-start +{nxt-t[ph][2]:6d}  total {nxt-t[ph][0]:7d}")
print("whole stack cycles:", t[n - 1][2] - t[0][0])
