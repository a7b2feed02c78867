"""Greedy search over the ops that get programmatic dependent launch (RVC_PDL_OPS): forward selection from a start set, then a
backward pass; one engine per candidate set, device ms / window over STEPS graph replays.  Log -> gpurun_out/pdl_search.log"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "obs-rvc_b200")]
import torch
import rvc_b200
from oracle import pipeline, weights
p = weights.make_data_dir("/tmp/rvc_b200_data_seed7", seed=7, index_rows=40000)
g = pipeline.BASELINE_GEOM
n16k, sf, skip, R = g["n16k"], g["sf16k"], g["skip_head"], g["return_length"]
steps, warm = int(os.environ.get("STEPS", 150)), 8
pcm = torch.from_numpy(pipeline.synthetic_pcm(n16k + sf * (steps + warm + 1))).cuda()
out = torch.empty(R * 400, dtype=torch.float32, device="cuda")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
log = open(os.path.join(ROOT, "gpurun_out", "pdl_search.log"), "w")
def say(*a):
    s = " ".join(str(x) for x in a); print(s, flush=True); log.write(s + "\n"); log.flush()
def measure(pats):
    os.environ["RVC_PDL_OPS"] = ",".join(sorted(pats)) if pats else "none"
    eng = rvc_b200.RvcInfer(p["data"]); eng.load_contentvec(2); eng.load_f0(1); eng.load_model(p["model"]); eng.load_index(p["index"], 0.5)
    def step(i): eng.infer_ptr(pcm.data_ptr() + 4 * i * sf, n16k, sf, 12, skip, R, out.data_ptr(), R * 400, True)
    for i in range(warm): step(i)
    eng.sync(); eng.event_record(0)
    for i in range(steps): step(warm + i)
    eng.event_record(1); eng.sync()
    ms = eng.event_elapsed_ms(0, 1) / steps
    eng.close()
    return ms * 1e3
atoms = ["rm.enc0", "rm.enc1", "rm.enc2", "rm.enc3", "rm.enc4", "rm.pool", "rm.cnn", "rm.gi", "rm.gru", "f0", "mel"]
for i in range(5): atoms += [f"rm.dec{i}*up", f"rm.dec{i}*sc", f"rm.dec{i}*c1", f"rm.dec{i}*c2"]
atoms += ["sy.stats", "sy.proj", "sy.enc", "sy.flow", "sy.E", "sy.F", "sy.up", "rm.mid"]
atoms += ["cv.conv", "cv.pos", "cv.", "knn_scan", "knn_select", "knn_blend", "phone", "pitch", "sy.emb", "sy.conv_pre", "sy.sine", "sy.har", "sy.z", "sy.stage", "sy.audio"]
for i in range(4): atoms += [f"sy.U{i}.up", f"sy.U{i}.noise", f"sy.U{i}.rb*c1", f"sy.U{i}.rb*c2"]
if os.environ.get("SETS"):
    for rep in range(2):
        for st in os.environ["SETS"].split(";"):
            say(f"{measure(set(x for x in st.split(',') if x)):.1f} us  {st}")
    sys.exit(0)
cur = set(x for x in os.environ.get("START", "rm.enc0,rm.enc1,rm.enc2,rm.enc3,rm.enc4,rm.pool").split(",") if x)
t_end = time.time() + float(os.environ.get("BUDGET_S", "1200"))
best = measure(cur); say(f"start {best:.1f} us  {sorted(cur)}")
base0 = measure(set()); say(f"no PDL {base0:.1f} us")
def backward():
    global best
    for a in sorted(cur):
        if time.time() > t_end + 240: break
        us = measure(cur - {a}); say(f"  - {a:18s} {us:.1f}  ({us - best:+.1f})")
        if us < best - 2.0: cur.discard(a); best = us; say(f"    dropped -> {best:.1f}")
if os.environ.get("BACKWARD_FIRST"): backward()
improved = True
while improved and time.time() < t_end:
    improved = False
    trial = []
    for a in atoms:
        if a in cur or time.time() > t_end: continue
        us = measure(cur | {a}); trial.append((us, a)); say(f"  + {a:18s} {us:.1f}  ({us - best:+.1f})")
    trial.sort()
    # take every atom that helps on its own by > 3 us, re-check the union, fall back to the single best
    good = [a for us, a in trial if us < best - 3.0]
    if good:
        us = measure(cur | set(good))
        if us < trial[0][0] - 1.0: cur |= set(good); best = us
        else: cur.add(trial[0][1]); best = trial[0][0]
        improved = True
        say(f"round -> {best:.1f} us  {sorted(cur)}")
backward()
say(f"final {best:.1f} us  RVC_PDL_OPS={','.join(sorted(cur))}")
