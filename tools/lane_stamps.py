"""Where the two front-end lanes and the tail stand inside a graph-replayed window, from %globaltimer stamps written by the
marker kernels themselves (rvc_debug_lane_stamps) - no event nodes in the graph, so the lanes are not perturbed
(tools/timeline.py is).  MODE=pitch: the F0 branch alone (rvc_pitch; RVC_PITCH_ML=1 gives it the lane structure it has
inside an infer plan)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "obs-rvc_b200")]
import numpy as np
import rvc_b200
from oracle import pipeline, weights
p = weights.make_data_dir("/tmp/rvc_b200_data_seed7", seed=7, index_rows=40000)
eng = rvc_b200.RvcInfer(p["data"]); eng.load_contentvec(2); eng.load_f0(1); eng.load_model(p["model"])
eng.load_index(p["index"], 0.5)
g = pipeline.BASELINE_GEOM
x = pipeline.synthetic_pcm(g["n16k"] + 64 * g["sf16k"])
L = rvc_b200.lib()
pitch_only = os.environ.get("MODE", "infer") == "pitch"
rows = []
for i in range(int(os.environ.get("N", "40"))):
    w = x[(i % 60) * g["sf16k"]: (i % 60) * g["sf16k"] + g["n16k"]]
    if pitch_only: eng.pitch(w, 12, g["sf16k"])
    else: eng.infer(w, g["sf16k"], 12, g["skip_head"], g["return_length"])
    out = (ctypes.c_ulonglong * 12)()
    L.rvc_debug_lane_stamps(eng.handle, out)
    t = np.array(list(out), dtype=np.float64)
    rows.append((t[1:] - t[0]) / 1e3)
    if i == int(os.environ.get("N", "40")) - 1:
        for c in eng.profile_chains():
            print(f"CHAIN lane {c['lane']} grid {c['grid']} phases {len(c['phases'])}: start {(c['t0_ns'] - t[0]) / 1e3:.0f} end {(c['t1_ns'] - t[0]) / 1e3:.0f} us  ({c['phases'][0]['ops'][:24]} ...)")
r = np.median(np.array(rows[8:]), axis=0)
f0 = f"pool0..4 {r[4]:.0f} {r[5]:.0f} {r[6]:.0f} {r[7]:.0f} {r[8]:.0f}, gru {r[9]:.0f}, f0 decode {r[0]:.0f}"
if pitch_only: print("STAMPS us after STFT start:", f0)
else: print(f"STAMPS us after STFT start: {f0}, pitch cache {r[1]:.0f}, retrieval gather {r[2]:.0f}, sine source {r[10]:.0f}, conv_post end {r[3]:.0f}")
