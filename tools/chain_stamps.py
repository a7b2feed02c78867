"""clock64 stamps of CTA 0 inside the chain kernel (stand-alone replay of one chain): per phase
[0 item start, 1 tile entry, 2 prologue issued, 3 k-tile 0 landed, 4 k-tile 1, 5 k-loop done, 6 epilogue done, 7 arrived]."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "obs-rvc_b200")]
import rvc_b200
from oracle import pipeline, weights
p = weights.make_data_dir("/tmp/rvc_b200_data_seed7", seed=7, index_rows=40000)
eng = rvc_b200.RvcInfer(p["data"]); eng.load_contentvec(2); eng.load_f0(1); eng.load_model(p["model"])
eng.load_index(p["index"], 0.5)
g = pipeline.BASELINE_GEOM
x = pipeline.synthetic_pcm(g["n16k"] + 8 * g["sf16k"])
for i in range(3):
    eng.infer(x[i * g["sf16k"]: i * g["sf16k"] + g["n16k"]], g["sf16k"], 12, g["skip_head"], g["return_length"])
chains = eng.profile_chains()
L = rvc_b200.lib()
for c in chains:
    out = (ctypes.c_longlong * 2560)()
    rc = L.rvc_debug_chain_stamps(eng.handle, ctypes.c_int(c["chain"]), out)
    t = np.array(list(out)[:2048], dtype=np.int64).reshape(256, 8); t2 = np.array(list(out)[2048:], dtype=np.int64).reshape(256, 2)
    print(f"chain {c['chain']} grid {c['grid']} rc={rc}")
    for ph, P in enumerate(c["phases"][:256]):
        if ph % int(os.environ.get("EVERY", 4)) and "pool" not in P["ops"]:
            continue
        r = t[ph]
        nxt = t[ph + 1][0] if ph + 1 < len(c["phases"]) else r[7]
        print(f"  ph{ph:3d} {P['ops'][:44]:44s} tile@{r[1]-r[0]:5d} issued@{r[2]-r[0]:6d} kt0@{r[3]-r[0]:6d} kt1@{r[4]-r[0]:6d} kdone@{r[5]-r[0]:6d} epi@{r[6]-r[0]:6d} arrive@{r[7]-r[0]:6d} spin@{t2[ph][0]-r[0]:6d} released@{t2[ph][1]-r[0]:6d} next@{nxt-r[0]:6d}")
