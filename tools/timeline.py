"""Per-op end times inside one graph-replayed window (rvc_profile_timeline) -> gpurun_out/timeline.json
plus a per-lane / per-section summary on stdout."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "obs-rvc_b200")]
import rvc_b200
from oracle import pipeline, weights
p = weights.make_data_dir("/tmp/rvc_b200_data_seed7", seed=7, index_rows=40000)
eng = rvc_b200.RvcInfer(p["data"]); eng.load_contentvec(2); eng.load_f0(1); eng.load_model(p["model"])
eng.load_index(p["index"], 0.5)
g = pipeline.BASELINE_GEOM
x = pipeline.synthetic_pcm(g["n16k"] + 8 * g["sf16k"])
for i in range(4):
    eng.infer(x[i * g["sf16k"]: i * g["sf16k"] + g["n16k"]], g["sf16k"], 12, g["skip_head"], g["return_length"])
tl = eng.profile_timeline()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(tl, open(os.path.join(ROOT, "gpurun_out", "timeline.json"), "w"))
marks = ["mel", "rm.pool4", "rm.inter", "rm.mid3.3.c2", "rm.cnn", "rm.gru", "f0", "cv.conv0", "cv.conv6", "cv.posconv", "cv.layer0", "cv.layer5", "cv.layer11",
         "knn_scan", "knn_select", "phone", "pitch", "sy.emb", "sy.enc5", "sy.z_p", "sy.flow0", "sy.conv_pre", "sy.stage0", "sy.stage1",
         "sy.stage2", "sy.stage3", "sy.audio"]
byname = {o["name"]: o for o in tl}
for m in marks:
    if m in byname:
        print(f"{m:16s} lane {byname[m]['lane']}  end {byname[m]['end_us']:9.1f} us")
print("last:", max(o["end_us"] for o in tl))
info = {}
prev_end = {}
print("--- chains / segments (end_us, delta since previous event on the lane) ---")
for o in tl:
    d = o["end_us"] - prev_end.get(o["lane"], 0.0)
    prev_end[o["lane"]] = o["end_us"]
    if d > 40:
        print(f"{o['name']:22s} lane {o['lane']} end {o['end_us']:9.1f}  +{d:8.1f}")
