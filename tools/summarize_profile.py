"""Turns gpurun_out/launches.csv (+ prof_gemm.ncu-rep) into the committed summaries under profiles/."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
lines = []

rows = list(csv.reader(open(os.path.join(ROOT, "gpurun_out", "launches.csv"))))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    if len(r) <= vi:
        continue
    name = re.sub(r"^void ", "", re.sub(r"\(.*", "", r[ki]))
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
lines.append(f"# {tag}: ncu launch list of `python bench.py --steps 1 --warmup 3` (window of {sum(v[0] for v in agg.values())} "
             f"launches ~ one training step; cold-cache, serialised: compare SHARES)\n")
lines.append("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    lines.append(f"| `{k[:80]}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |")
lines.append(f"| **total** | {sum(v[0] for v in agg.values())} | {tot:.1f} | 100% |\n")

rep = os.path.join(ROOT, "gpurun_out", "prof_gemm.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, units = rr[0], rr[1]
    want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size"]
    idx = [h.index(w) for w in want if w in h]
    lines.append(f"# {tag}: `ncu --set full --clock-control none` of splice_gemm_kernel (tcgen05 GEMM), per launch\n")
    lines.append("| kernel | " + " | ".join(f"{h[i]} [{units[i]}]" for i in idx) + " |")
    lines.append("|---|" + "---:|" * len(idx))
    for r in rr[2:]:
        lines.append("| `" + re.sub(r"\(.*", "", r[h.index("Kernel Name")])[:40] + "` | " + " | ".join(r[i] for i in idx) + " |")
    lines.append("")
bench = os.path.join(ROOT, "gpurun_out", "bench.log")
if os.path.exists(bench):
    try:
        d = json.loads(open(bench).read().strip().splitlines()[-1])
        lines.append(f"# {tag}: bench line (plain run, no profiler)\n\n```json\n{json.dumps(d, indent=1)}\n```\n")
    except Exception:
        pass
open(os.path.join(out_dir, f"{tag}_summary.md"), "w").write("\n".join(lines) + "\n")
print("wrote", os.path.join(out_dir, f"{tag}_summary.md"))
