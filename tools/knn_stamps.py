"""Per-tile pipeline stamps of the tcgen05 kNN scan (CTA 0): build with NVFLAGS+=-DRVC_KU_STAMPS."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "obs-rvc_b200")]
import numpy as np
import rvc_b200
from oracle import weights
p = weights.make_data_dir("/tmp/rvc_b200_data_seed7", seed=7, index_rows=40000)
n, c, q, k = 1 << 20, 256, 128, 4
eng = rvc_b200.RvcInfer(p["data"], index_k=k)
rng = np.random.default_rng(2)
rows = rng.standard_normal((n, c), dtype=np.float32) * np.float32(0.34)
eng.set_index(rows, 0.5)
x = rng.standard_normal((q, c), dtype=np.float32) * np.float32(0.34)
for _ in range(3): eng.knn_search(x, k)
buf = (ctypes.c_longlong * 256)()
rvc_b200.lib().rvc_debug_knn_stamps(eng.handle, buf)
st = np.array(buf[:], np.int64).reshape(8, 32)
t0 = st[st > 0].min()
names = ["Pe", "Pi", "Mae", "Mf", "Mi", "Eaf", "Eld", "Edone"]
for i in range(4, 16):
    print(f"tile {180+i:3d}: " + "  ".join(f"{names[e]}={st[e, i] - t0:7d}" for e in range(8)))
