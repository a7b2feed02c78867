#!/usr/bin/env python
"""bench.py - headline benchmark of the per-audio-window hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one 160 ms @16 kHz window (2560 new samples; the call sees the last 35 840 samples)
through `rvc_infer` of configs[1]: single stream, batch 1, ContentVec-768 + RMVPE + 40k x 768
index (k=8, index_rate 0.5) + NSF-HiFiGAN 40k, seeded synthetic weights, synthetic PCM.
  value : windows/s with PCM and audio resident in HBM (rvc_infer_dev), CUDA events on the
          engine stream, MAX over ranks;
  e2e   : windows/s through the host-buffer C ABI call (rvc_infer): pinned H2D of the window and
          D2H of the audio inside the timed region, one window in flight (batch-1 latency bound);
  p50/p99_ms : per-window host-observed latency of that call.
N > 1 (torchrun): one process per GPU, each with its own independent stream (the path shards by
stream, SURVEY 8e - no data-path collective), NCCL only for the barrier and the max-time reduce.
`--impl reference` times the CPU restatement of the reference path (oracle/) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "obs-rvc_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "audio frames/sec (160 ms @16 kHz, batch 1)"
UNIT = "frames/s"
INDEX_ROWS = 40000
WORKLOAD = ("configs[1]: single stream batch=1, ContentVec-768 + RMVPE + 40k x 768 index (k=8, "
            "rate 0.5) + NSF-HiFiGAN 40k; 35840-sample window advanced by 2560")


def shard_streams(n_streams: int, rank: int, world: int):
    """Stream s lives on rank s mod world for its whole life (its pitch cache is device state)."""
    return [s for s in range(n_streams) if s % world == rank]


def aggregate_throughput(steps: int, world: int, max_seconds: float) -> float:
    """Whole-job windows/s: every rank processed `steps` windows; time = MAX over ranks."""
    return world * steps / max_seconds


def data_dir():
    from oracle import weights
    root = os.path.join(tempfile.gettempdir(), "rvc_b200_data_seed7")
    return weights.make_data_dir(root, seed=7, index_rows=INDEX_ROWS)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=float(d["hbm_gbs"]), bf16=float(d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, bf16=1590.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_reference(steps, warmup, paths, bounded_seconds=None):
    """Times the CPU restatement of the reference path (oracle/) - torch-CPU fp32 + numpy DSP,
    all host threads.  Returns (frames/s, ms list, cores)."""
    import torch

    from oracle import pipeline
    from oracle.weights import read_rvcw
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = pipeline.BASELINE_GEOM
    ora = pipeline.RvcInfer(paths["data"], noise_seed=0)
    ora.load_contentvec(2); ora.load_f0(1); ora.load_model(paths["model"])
    ora.set_index(read_rvcw(paths["index"])["big_npy"], 0.5)
    pcm = pipeline.synthetic_pcm(g["n16k"] + g["sf16k"] * (steps + warmup + 1))
    ts = []
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        x = pcm[i * g["sf16k"]: i * g["sf16k"] + g["n16k"]]
        t0 = time.perf_counter()
        ora.infer(x, g["sf16k"], 12, g["skip_head"], g["return_length"])
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
        if bounded_seconds and time.perf_counter() - t_start > bounded_seconds and len(ts) >= 3:
            break
    return len(ts) / sum(ts), [t * 1e3 for t in ts], cores, torch.get_num_threads()


def run_reference(args, rank, world):
    if rank != 0:
        return
    paths = data_dir()
    steps = min(args.steps, 60)
    fps, ms, cores, threads = cpu_reference(steps, min(args.warmup, 3), paths)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": len(ms),
        "warmup": min(args.warmup, 3), "ms_per_step": float(np.mean(ms)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU restatement of the reference path (torch-CPU + numpy), "
                   "not ONNX Runtime: the reference cannot be built here (SURVEY 8c)"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{len(ms)} consecutive windows of the same workload"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "p50_ms": float(np.percentile(ms, 50)), "p99_ms": float(np.percentile(ms, 99)), "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-ops", action="store_true", help="also dump per-op device times to gpurun_out/")
    ap.add_argument("--streams", type=int, default=8,
                    help="supplementary: independent live streams per GPU sharing one set of weights (configs[3]); 0 = skip")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import torch  # pinned host buffers + device buffers only (plumbing)

    import rvc_b200
    from oracle import pipeline  # synthetic input generator + geometry constants only
    from oracle.weights import read_rvcw
    if rank == 0:
        paths = data_dir()
    if dist is not None:
        dist.barrier()
    paths = data_dir()
    g = pipeline.BASELINE_GEOM
    n16k, sf, skip, R = g["n16k"], g["sf16k"], g["skip_head"], g["return_length"]
    total = args.warmup + args.steps
    pcm_all = pipeline.synthetic_pcm(n16k + sf * (total + 1), seed=rank)

    eng = rvc_b200.RvcInfer(paths["data"], device=local_rank, noise_seed=rank)
    eng.load_contentvec(2); eng.load_f0(1); eng.load_model(paths["model"])
    eng.load_index(paths["index"], 0.5)
    out_len = R * 400

    # ---- value: inputs resident in HBM --------------------------------------------------------
    dev = torch.device("cuda", local_rank)
    pcm_dev = torch.from_numpy(pcm_all).to(dev)
    out_dev = torch.empty(out_len, dtype=torch.float32, device=dev)
    torch.cuda.synchronize(dev)

    def step_dev(i):
        eng.infer_ptr(pcm_dev.data_ptr() + 4 * i * sf, n16k, sf, 12, skip, R, out_dev.data_ptr(), out_len, True)

    for i in range(args.warmup):
        step_dev(i)
    eng.sync()
    sampler = ClockSampler(local_rank)
    sampler.start()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    l0 = eng.kernel_launches()
    eng.event_record(0)
    for i in range(args.steps):
        step_dev(args.warmup + i)
    eng.event_record(1)
    eng.sync()
    torch.cuda.synchronize(dev)
    dev_ms = eng.event_elapsed_ms(0, 1)
    launches = eng.kernel_launches() - l0
    if dist is not None:
        t = torch.tensor([dev_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms_max = float(t.item())
    else:
        dev_ms_max = dev_ms
    value = aggregate_throughput(args.steps, world, dev_ms_max * 1e-3)

    # ---- e2e: host buffers through the C ABI call, one window in flight -------------------------
    eng.reset_state()
    pin_in = torch.from_numpy(pcm_all).pin_memory()
    pin_out = torch.empty(out_len, dtype=torch.float32).pin_memory()
    lat = []
    for i in range(args.warmup):
        eng.infer_ptr(pin_in.data_ptr() + 4 * i * sf, n16k, sf, 12, skip, R, pin_out.data_ptr(), out_len, False)
    if dist is not None:
        dist.barrier()
    t_e2e0 = time.perf_counter()
    for i in range(args.steps):
        t0 = time.perf_counter()
        eng.infer_ptr(pin_in.data_ptr() + 4 * (args.warmup + i) * sf, n16k, sf, 12, skip, R, pin_out.data_ptr(), out_len, False)
        lat.append((time.perf_counter() - t0) * 1e3)
    e2e_s = time.perf_counter() - t_e2e0
    if dist is not None:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    sampler.stop_flag = True
    sampler.join(timeout=2)
    e2e = aggregate_throughput(args.steps, world, e2e_s)

    # ---- supplementary: S independent streams on this GPU (configs[3] shape: 8 per GPU), weights shared --
    multi = None
    if args.streams > 1:
        engs = [eng]
        for i in range(1, args.streams):
            e2 = rvc_b200.RvcInfer(paths["data"], device=local_rank, noise_seed=rank * 1000 + i)
            e2.load_contentvec(2); e2.load_f0(1); e2.load_model(paths["model"]); e2.load_index(paths["index"], 0.5)
            engs.append(e2)
        outs = [torch.empty(out_len, dtype=torch.float32, device=dev) for _ in engs]
        ksteps = max(20, args.steps // 4)

        def step_all(i):
            for e_, o_ in zip(engs, outs):
                e_.infer_ptr(pcm_dev.data_ptr() + 4 * (i % total) * sf, n16k, sf, 12, skip, R, o_.data_ptr(), out_len, True)

        for i in range(5):
            step_all(i)
        for e_ in engs:
            e_.sync()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for i in range(ksteps):
            step_all(5 + i)
        for e_ in engs:
            e_.sync()
        torch.cuda.synchronize(dev)
        ms_t = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([ms_t], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_t = float(t.item())
        multi = {"streams_per_gpu": args.streams, "value": world * args.streams * ksteps / ms_t, "unit": UNIT,
                 "ms_per_round": ms_t / ksteps * 1e3,
                 "timing": "host wall clock around K rounds of S async windows, device synchronised on both sides",
                 "note": "plans built without persistent chains while several contexts share the device"}
        for e_ in engs[1:]:
            e_.close()

    # ---- roofline of the dominant kernel (rank 0): per-op device times via CUDA events ----------
    roof = None
    roof_extra = []
    prof = None
    if rank == 0:
        pk = peaks()
        step_dev(0); eng.sync()         # back on the single-stream plan (the 8-stream section above runs without chains)
        chains = eng.profile_chains()   # per-phase device times of the persistent chain kernels in that window
        prof = eng.profile_ops(10)
        step_us = dev_ms / args.steps * 1e3

        def family(o):
            if o["kind"] != "gemm":
                return o["kind"]
            v = o.get("variant", 0)
            return "umma_gemm_kernel(tcgen05 2xFP16 split)" if v >= 5 else ("gemm_v2_kernel(fp32 split-K)" if v >= 1 else "gemm_f32_kernel")

        fam = {}
        for o in prof:
            if o.get("chain", -1) >= 0:
                # executed inside a persistent chain kernel: its work counts there, its time is the chain's own clock
                f = fam.setdefault("chain_kernel(fp32, persistent)", {"us": 0.0, "flops": 0.0, "bytes": 0.0, "n": 0})
                f["flops"] += o["flops"]; f["bytes"] += o["wbytes"] + o["iobytes"]
                continue
            f = fam.setdefault(family(o), {"us": 0.0, "flops": 0.0, "bytes": 0.0, "n": 0})
            f["us"] += o["us"]; f["flops"] += o["flops"]; f["bytes"] += o["wbytes"] + o["iobytes"]; f["n"] += 1
        if chains:
            f = fam["chain_kernel(fp32, persistent)"]
            f["us"] = sum(ph["us"] for c in chains for ph in c["phases"]); f["n"] = len(chains)

        def roof_of(name, f):
            tfl = f["flops"] / (f["us"] * 1e-6) / 1e12
            gbs = f["bytes"] / (f["us"] * 1e-6) / 1e9
            if name.startswith("umma"):
                r = {"bound": "tensor", "achieved": tfl, "peak": pk["bf16"], "unit": "TFLOP/s", "frac": tfl / pk["bf16"]}
            else:
                r = {"bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"]}
            r.update({"kernel": name, "launches_per_step": f["n"], "kernel_us_per_step": f["us"], "avg_launch_us": f["us"] / f["n"],
                      "share_of_step_device_time": f["us"] / sum(x["us"] for x in fam.values()), "peak_source": pk["src"]})
            return r

        order = sorted(fam.items(), key=lambda kv: -kv[1]["us"])
        roof = roof_of(*order[0])
        # dram__bytes_read+write of one launch of this kernel under `ncu --set full` (profiles/r01_e_prof_umma.md: grid
        # (4,2,8) = a HiFiGAN stage-0 ResBlock conv, M=210 N=256 K=1792: fp16 hi + scaled-lo weight planes = 4 N K = 1.8 MB
        # ... 2.9 MB for K=2816, + the fp32 A rows) - at the algorithmic figure, half of the 3xTF32 planes (6.26 MB, r01_d)
        roof["traffic"] = 3.38e6 if order[0][0].startswith("umma") else None
        roof["note"] = ("achieved = sum of algorithmic bytes (weights + activations) or flops (2MNK) of this kernel's launches in one "
                        "window / sum of their device times; each op timed as a 10-launch CUDA graph between events on the "
                        "engine stream (rvc_profile_ops); ncu captures: profiles/")
        for name, f in order[1:5]:
            roof_extra.append(roof_of(name, f))
        for r in roof_extra:
            if r["kernel"] == "knn_scan":
                r["traffic"] = 123.0e6  # dram__bytes_read of profiles/r01_d_prof_knn2.md (one-CTA-per-SM variant; algorithmic 122.9 MB)
        if args.profile_ops:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            json.dump({"step_us": step_us, "ops": prof}, open(os.path.join(ROOT, "gpurun_out", "profile_ops.json"), "w"))

    cpu_b = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fps, ms, cores, threads = cpu_reference(40, 2, paths, bounded_seconds=25.0)
        cpu_b = {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
                 "sample": f"{len(ms)} consecutive windows of the same workload (oracle/, torch-CPU fp32 + numpy)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "streams_per_gpu": 1, "parallelism": f"{world} independent stream(s), 1 per GPU",
                       "l2": "weights (850 MB fp32) + index (123 MB) exceed the 126 MB L2: every window re-streams them",
                       "realtime_factor": value * 0.16 / world},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": n16k * 4, "d2h_bytes_per_step": out_len * 4},
            "p50_ms": float(np.percentile(lat, 50)), "p99_ms": float(np.percentile(lat, 99)),
            "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roof, "cpu_baseline": cpu_b,
            "multi_stream": multi, "roofline_other_kernels": roof_extra,
        }
        print(json.dumps(line))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
