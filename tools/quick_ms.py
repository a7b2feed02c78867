"""Device ms/window of configs[1] for the current environment (schedule sweeps): prints one line."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "obs-rvc_b200")]
import numpy as np, torch
import rvc_b200
from oracle import pipeline, weights
p = weights.make_data_dir("/tmp/rvc_b200_data_seed7", seed=7, index_rows=40000)
eng = rvc_b200.RvcInfer(p["data"]); eng.load_contentvec(2); eng.load_f0(1); eng.load_model(p["model"])
eng.load_index(p["index"], 0.5)
g = pipeline.BASELINE_GEOM
n16k, sf, skip, R = g["n16k"], g["sf16k"], g["skip_head"], g["return_length"]
steps, warm = int(os.environ.get("STEPS", 200)), 10
pcm = torch.from_numpy(pipeline.synthetic_pcm(n16k + sf * (steps + warm + 1))).cuda()
out = torch.empty(R * 400, dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
def step(i): eng.infer_ptr(pcm.data_ptr() + 4 * i * sf, n16k, sf, 12, skip, R, out.data_ptr(), R * 400, True)
for i in range(warm): step(i)
eng.sync(); eng.event_record(0)
for i in range(steps): step(warm + i)
eng.event_record(1); eng.sync()
ms = eng.event_elapsed_ms(0, 1) / steps
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("RVC_"))
print(f"QUICK ms_per_window={ms:.4f} fps={1e3/ms:.1f} audio_rms={float(out.square().mean().sqrt()):.5f} [{tag}]")
if os.environ.get("DUMP_OPS"):
    import json
    json.dump({"step_us": ms * 1e3, "ops": eng.profile_ops(10)}, open(os.path.join(ROOT, "gpurun_out", os.environ["DUMP_OPS"]), "w"))
if os.environ.get("CHAINS"):
    for c in eng.profile_chains():
        tot = sum(p["us"] for p in c["phases"])
        print(f"chain {c['chain']} lane {c['lane']} grid {c['grid']} phases {len(c['phases'])} total {tot:.1f} us")
        if os.environ["CHAINS"] == "2":
            for p in c["phases"]:
                print(f"   {p['us']:7.2f}  {p['ops']}")
