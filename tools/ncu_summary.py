"""Summarises gpurun_out/*.ncu-rep + launches.csv into profiles/ (text, tracked).  Usage:
python tools/ncu_summary.py <tag>"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")
KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "smsp__cycles_active.avg", "sm__inst_executed.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def launches(tag, last_n=0):
    path = os.path.join(GO, "launches.csv")
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if not l.startswith("==")]
    tot = collections.defaultdict(float); cnt = collections.Counter()
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    if last_n > 0:
        rows = rows[-last_n:]   # exactly the last window (tools/ncu_window.py prints its launch count)
    for row in rows:
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        tot[name] += v; cnt[name] += 1
    s = sum(tot.values())
    with open(os.path.join(OUT, f"{tag}_launches_summary.md"), "w") as f:
        f.write(f"# ncu launch list, one graph-replayed window ({sum(cnt.values())} kernels, {s:.0f} us serialised, cold cache)\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/ncu_window.py` (last window of the run)\n\n")
        f.write("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write(f"| `{k}` | {cnt[k]} | {v:.1f} | {100 * v / s:.1f}% | {v / cnt[k]:.1f} |\n")
    os.replace(path, os.path.join(OUT, f"{tag}_launches.csv"))


def full(tag, rep):
    path = os.path.join(GO, rep + ".ncu-rep")
    if not os.path.exists(path):
        return
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(OUT, f"{tag}_{rep}.md"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on : {rep}\n\n")
        for r in rows[2:]:
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"- {k}: {r[i]} {units[i]}\n")
            f.write("\n")
    if os.environ.get("RAW_CSV"):   # the full raw page (200+ KB per capture) only on request
        with open(os.path.join(OUT, f"{tag}_{rep}_raw.csv"), "w") as f:
            f.write(out)


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(OUT, exist_ok=True)
    launches(tag, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    import glob
    for path in sorted(glob.glob(os.path.join(GO, "prof_*.ncu-rep"))):
        full(tag, os.path.basename(path)[:-len(".ncu-rep")])
