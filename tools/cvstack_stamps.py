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
out = (ctypes.c_longlong * 1536)()
n = L.rvc_debug_cvstack_stamps(eng.handle, out, ctypes.c_int(1536))
t = np.array(list(out)[:512], dtype=np.int64).reshape(128, 4)
t2 = np.array(list(out)[512:], dtype=np.int64).reshape(128, 8)
print("phases", n)
names = ["qkv", "attn", "o", "ln1", "fc1", "fc2", "ln2"]
for ph in range(min(n, 23)):
    nm = "enc_in" if ph == 0 else names[(ph - 1) % 7]
    nxt = t[ph + 1][0] if ph + 1 < n else t[ph][2]
    print(f"ph{ph:3d} {nm:7s} start@{t[ph][0]-t[0][0]:8d} gridwait {t[ph][3]-t[ph][0]:7d} work {t[ph][1]-t[ph][3]:7d} arrive +{t[ph][2]-t[ph][1]:6d} next-start +{nxt-t[ph][2]:6d}  total {nxt-t[ph][0]:7d}")
for ph in range(8, 15):
    nm = names[(ph - 1) % 7]
    if nm in ("qkv", "o", "fc1", "fc2"):
        b = t[ph][0]
        print(f"ph{ph:3d} {nm:4s} worker gridwait over@{t[ph][3]-b:6d} | producer: wait over@{t2[ph][0]-b:6d} last issue@{t2[ph][1]-b:6d} | mma: first full@{t2[ph][2]-b:6d} last full@{t2[ph][3]-b:6d} committed@{t2[ph][4]-b:6d} | worker acc full@{t2[ph][5]-b:6d} done@{t[ph][1]-b:6d} arrived@{t[ph][2]-b:6d}")
for ph in (9, 16):
    b = t[ph][3]
    print(f"ph{ph:3d} attn (after grid wait): item start@{t2[ph][0]-b:6d} loads landed+K^T stored@{t2[ph][1]-b:6d} synced@{t2[ph][2]-b:6d} scores done@{t2[ph][3]-b:6d} synced@{t2[ph][4]-b:6d} PV done@{t2[ph][5]-b:6d}")
print("whole stack cycles:", t[n - 1][2] - t[0][0])
