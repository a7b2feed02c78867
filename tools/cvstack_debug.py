"""Numerical debugging of the persistent ContentVec stack: planes and first GEMM against the CPU plan interpreter."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "obs-rvc_b200"), os.path.join(ROOT, "tests")]
import planexec, rvc_b200
from oracle import pipeline, weights
p = weights.make_data_dir("/tmp/rvc_b200_data_seed7", seed=7, index_rows=40000)
eng = rvc_b200.RvcInfer(p["data"]); eng.load_contentvec(2)
pe = planexec.PlanExec(p["data"]); pe.load(0, p["contentvec"])
g = pipeline.BASELINE_GEOM
x = pipeline.synthetic_pcm(g["n16k"])
eng.hubert(x)
pe.run(planexec.PLAN_HUBERT, x, 0, 0, 0, 0)
T = 111
enc = pe.get("cv.enc_in").reshape(T, 768)
genc = eng.get_last("cv.enc_in").reshape(T, 768)
print("enc_in max err", np.abs(enc - genc).max())
pl = eng.get_last("cv.dbg_planes_x").view(np.float16).reshape(2, 128, 768).astype(np.float32)
rec = pl[0] + pl[1] / 2048.0
print("planes_x (LAST writer = ln2 of layer 11) rows>=T zero:", np.abs(rec[T:]).max())
last = eng.get_last("cv.layer11").reshape(T, 768)
print("planes_x vs gpu cv.layer11 max err", np.abs(rec[:T] - last).max())
q = pe.get("cv.L0.qkv").reshape(T, 2304); gq = eng.get_last("cv.L0.qkv").reshape(T, 2304)
err = np.abs(q - gq)
print("qkv err by 48-col tile (max):", np.round(err.reshape(T, 48, 48).max(axis=(0, 2)), 3))
print("qkv err by row (max) first 16:", np.round(err.max(axis=1)[:16], 3), " rows 96..111:", np.round(err.max(axis=1)[96:], 3))
print("qkv err by col within tile (max over tiles, rows):", np.round(err.reshape(T, 48, 48).max(axis=(0, 1)), 3))
print("sample ref", q[0, :6], "\nsample gpu", gq[0, :6])
print("ratio gpu/ref row0 first 6:", gq[0, :6] / q[0, :6])
