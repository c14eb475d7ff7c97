#!/usr/bin/env python3
"""Turn an .ncu-rep capture and a launch-list CSV into the small tracked summaries under profiles/.
usage: tools/ncu_summary.py <prof.ncu-rep> <launches.csv> <round tag, e.g. r01>"""
import collections
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

rep, launches, tag = sys.argv[1], sys.argv[2], sys.argv[3]
out = Path(__file__).resolve().parent.parent / "profiles"
out.mkdir(exist_ok=True)

# --- launch list: per-kernel share of the captured region
lines = [l for l in open(launches) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict()
for x in rows:
    name = x["Kernel Name"].split("(")[0].replace("void ", "")
    agg.setdefault(name, []).append(float(x["Metric Value"].replace(",", "")) / 1000.0)
tot = sum(sum(v) for v in agg.values())
with open(out / f"{tag}_launch_shares.csv", "w") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "launches", "mean_us", "min_us", "max_us", "share_of_captured_gpu_time"])
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        w.writerow([k, len(v), f"{sum(v)/len(v):.2f}", f"{min(v):.2f}", f"{max(v):.2f}", f"{sum(v)/tot:.4f}"])
(out / f"{tag}_launches.csv").write_text("".join(lines))

# --- full capture: the metrics the roofline arithmetic needs
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hdr, units = r[0], r[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
idx = [(w, hdr.index(w)) for w in want if w in hdr]
traffic = {}
with open(out / f"{tag}_ncu_full_top_kernels.csv", "w") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [row[hdr.index("Kernel Name")].split("(")[0].replace("void ", "") for row in r[2:]])
    for name, i in idx[1:]:
        w.writerow([name, units[i]] + [row[i] for row in r[2:]])
for row in r[2:]:
    kn = row[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").split("<")[0]

    def val(m):
        i = hdr.index(m)
        v = float(row[i].replace(",", ""))
        u = units[i].lower()
        return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1, "tbyte": 1e12}.get(u, 1)
    rec = {"dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
           "duration_us_under_ncu": float(row[hdr.index("gpu__time_duration.sum")].replace(",", ""))}
    # several launches of a kernel in the capture: keep the one that moved most (the full-work launch; the incremental
    # M-step launches of later iterations stream only the rows that changed class)
    if kn not in traffic or rec["dram_bytes_per_launch"] > traffic[kn]["dram_bytes_per_launch"]:
        traffic[kn] = rec
(out / f"{tag}_traffic.json").write_text(json.dumps(traffic, indent=1))
print(json.dumps(traffic, indent=1))
