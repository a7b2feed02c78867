"""Batched plans on the GPU: (a) rvc_infer_windows (consecutive windows of one stream, NB per launch) against the same
windows through single rvc_infer calls, (b) device time per group for NB in {1, 2, 4, 8, 16, 32}."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "obs-rvc_b200")]
import numpy as np, torch
import rvc_b200
from oracle import pipeline, weights
p = weights.make_data_dir("/tmp/rvc_b200_data_seed7", seed=7, index_rows=40000)
g = pipeline.BASELINE_GEOM
n16k, sf, skip, R = g["n16k"], g["sf16k"], g["skip_head"], g["return_length"]
def fresh(seed=3):
    e = rvc_b200.RvcInfer(p["data"], noise_seed=seed); e.load_contentvec(2); e.load_f0(1); e.load_model(p["model"]); e.load_index(p["index"], 0.5)
    return e
NW = int(os.environ.get("NW", 8))
pcm = pipeline.synthetic_pcm(n16k + sf * (NW + 600), seed=5)
if not os.environ.get("SKIP_CHECK"):
    e1 = fresh()
    singles = [e1.infer(pcm[w * sf: w * sf + n16k], sf, 12, skip, R).copy() for w in range(NW)]
    pitch1 = e1.get_last("pitch", np.int32)
    e1.close()
    for mb in [int(x) for x in os.environ.get("MB", "2,8").split(",")]:
        e2 = fresh()
        got = e2.infer_windows(pcm, n16k, sf, NW, 12, skip, R, mb)
        errs = [float(np.sqrt(np.mean((got[w] - singles[w]) ** 2))) for w in range(NW)]
        print(f"CHECK max_batch={mb}: max rms err {max(errs):.3e} (audio rms {float(np.sqrt(np.mean(singles[0]**2))):.3f}) info={e2.plan_info()}")
        e2.close()
# timing: device-resident
dev = torch.device("cuda")
pcm_dev = torch.from_numpy(pcm).to(dev)
for nb in [int(x) for x in os.environ.get("NBS", "1,2,4,8,16,32").split(",")]:
    e = fresh()
    out = torch.empty(nb * R * 400, dtype=torch.float32, device=dev)
    reps = max(3, 64 // nb)
    def step(i): e.infer_windows_ptr(pcm_dev.data_ptr() + 4 * i * sf * nb, pcm_dev.numel() - i * sf * nb, n16k, sf, nb, 12, skip, R, out.data_ptr(), out.numel(), True, nb)
    for i in range(3): step(i)
    e.sync(); e.event_record(0)
    for i in range(reps): step(3 + i)
    e.event_record(1); e.sync()
    ms = e.event_elapsed_ms(0, 1) / reps
    print(f"BATCH nb={nb}: {ms:.3f} ms per group, {nb / ms * 1e3:.1f} windows/s, {e.plan_info()}")
    if os.environ.get("DUMP_OPS") and nb == int(os.environ["DUMP_OPS"]):
        import json
        json.dump({"step_us": ms * 1e3, "nb": nb, "ops": e.profile_ops(5)}, open(os.path.join(ROOT, "gpurun_out", f"ops_nb{nb}.json"), "w"))
    e.close()
