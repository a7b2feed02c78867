"""Per-buffer comparison of the CUDA engine against the CPU plan interpreter and the torch oracle
(GPU box diagnostic; writes gpurun_out/compare.txt).  Usage: python tools/gpu_compare.py [windows]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "obs-rvc_b200"), os.path.join(ROOT, "tests")]

import planexec  # noqa: E402
import rvc_b200  # noqa: E402
from oracle import pipeline, weights  # noqa: E402
from oracle.weights import read_rvcw  # noqa: E402


def main():
    nwin = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "compare.txt"), "w")

    def P(*a):
        s = " ".join(str(x) for x in a)
        print(s)
        log.write(s + "\n")
        log.flush()

    t0 = time.time()
    p = weights.make_data_dir("/tmp/rvc_data", seed=7, index_rows=40000)
    idx = read_rvcw(p["index"])["big_npy"]
    P("weights ready", round(time.time() - t0, 1), "s")
    eng = rvc_b200.RvcInfer(p["data"], noise_seed=0, use_cuda_graph=True)
    eng.load_contentvec(2); eng.load_f0(1); eng.load_model(p["model"]); eng.set_index(idx, 0.5)
    pe = planexec.PlanExec(p["data"])
    pe.load(0, p["contentvec"]); pe.load(1, p["f0"]); pe.load(2, p["model"]); pe.set_index(idx, 0.5)
    pe.set_params(seed=0, noise_mode=1, index_k=8)
    ora = pipeline.RvcInfer(p["data"], noise_seed=0)
    ora.load_contentvec(2); ora.load_f0(1); ora.load_model(p["model"]); ora.set_index(idx, 0.5)
    g = pipeline.BASELINE_GEOM
    pcm = pipeline.synthetic_pcm(g["n16k"] + g["sf16k"] * nwin)
    for w in range(nwin):
        x = pcm[w * g["sf16k"]: w * g["sf16k"] + g["n16k"]]
        t1 = time.time()
        got = eng.infer(x, g["sf16k"], 12, g["skip_head"], g["return_length"]).copy()
        t2 = time.time()
        pe.run(planexec.PLAN_INFER, x, g["sf16k"], 12, g["skip_head"], g["return_length"])
        t3 = time.time()
        want = ora.infer(x, g["sf16k"], 12, g["skip_head"], g["return_length"])
        t4 = time.time()
        P(f"== window {w}: gpu {t2-t1:.4f}s  cpu-plan {t3-t2:.2f}s  oracle {t4-t3:.2f}s  plan {eng.plan_info()}")
        cpu_audio = pe.get("audio")
        P(f"audio rms: gpu-vs-oracle {np.sqrt(np.mean((got-want)**2)):.3e}  gpu-vs-cpuplan "
          f"{np.sqrt(np.mean((got-cpu_audio)**2)):.3e}  cpuplan-vs-oracle {np.sqrt(np.mean((cpu_audio-want)**2)):.3e}"
          f"  |audio| rms {np.sqrt(np.mean(want**2)):.3f}")
        bad = 0
        for name in pe.names():
            if name in ("knn_cand_d", "knn_cand_i"):
                continue
            isint = name in ("f0_argmax", "pitch", "knn_idx")
            a = eng.get_last(name, np.int32 if isint else np.float32)
            b = pe.get(name, np.int32 if isint else np.float32)
            if isint:
                ok = np.array_equal(a, b)
                P(f"{name:28s} int equal={ok}")
                bad += (not ok)
                continue
            fin = np.isfinite(a).all()
            d = np.abs(a.astype(np.float64) - b.astype(np.float64))
            scale = max(float(np.sqrt(np.mean(b.astype(np.float64) ** 2))), 1e-6)
            rel = float(d.max()) / scale
            flag = "" if (fin and rel < 2e-3) else "   <<<<<< MISMATCH"
            bad += bool(flag)
            P(f"{name:28s} n={a.size:9d} max|d|={d.max():.3e} rms_ref={scale:.3e} rel={rel:.2e} finite={fin}{flag}")
        P(f"window {w}: {bad} mismatching buffers")
        P("argmax eq oracle:", np.array_equal(eng.get_last("f0_argmax", np.int32), ora.last["argmax"]),
          " pitch eq:", np.array_equal(eng.get_last("pitch", np.int32), ora.last["pitch"]),
          " knn eq:", np.array_equal(eng.get_last("knn_idx", np.int32).reshape(-1, 8), ora.last["knn_idx"]))
    P("launches", eng.kernel_launches())
    # timing: 20 windows through the host API
    x = pcm[:g["n16k"]]
    ts = []
    for i in range(30):
        t1 = time.perf_counter()
        eng.infer(x, g["sf16k"], 12, g["skip_head"], g["return_length"])
        ts.append(time.perf_counter() - t1)
    P("host-API ms per window (last 20):", np.round(np.array(ts[10:]) * 1e3, 3).tolist())


if __name__ == "__main__":
    main()
