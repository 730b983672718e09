"""Turns the ncu output under gpurun_out/ into the committed summaries under profiles/.

  gpurun_out/launches.csv        ncu --metrics gpu__time_duration.sum --clock-control none --csv of
                                 `python tools/profile_step.py --warmup 14 --steps 4` (one natural-gradient period)
  gpurun_out/ncu_launches.log    that run's stdout (JSON with the launches per step)
  gpurun_out/prof_gemm.ncu-rep   ncu --set full of splice_gemm_kernel launches of one light step
  gpurun_out/prof_den.ncu-rep    ncu --set full of the denominator frame kernels
  gpurun_out/bench.log           the plain bench line
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G = os.path.join(ROOT, sys.argv[2]) if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out")
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
lines = []


def short(name):
    return re.sub(r"^void ", "", re.sub(r"\(.*", "", name)).replace("tdnnf::", "").replace("<unnamed>::", "")


rows = list(csv.reader(open(os.path.join(G, "launches.csv"))))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
L = []
for r in data:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    L.append((short(r[ki]), v, r[gi], r[bi]))
with open(os.path.join(out_dir, f"{tag}_launches.csv"), "w") as f:
    f.write("id,kernel,grid,block,duration_us\n")
    for i, (n, v, g, b) in enumerate(L):
        f.write(f'{i},"{n}","{g}","{b}",{v:.3f}\n')
per_step = None
try:
    per_step = json.loads(open(os.path.join(G, "ncu_launches.log")).read().strip().splitlines()[-1])["launches"]
except Exception:
    pass
names = ["refresh (H, L, J, K with 3 planes)", "finish (W update of the refresh)", "plain", "plain"]
lines.append(f"# {tag}: ncu launch list of one natural-gradient period (`python tools/profile_step.py --warmup 14 --steps 4`, "
             f"{len(L)} launches; durations are cold-cache and serialised: compare SHARES)\n")
bounds = [0]
extra = len(L) - sum(per_step) if per_step else -1  # torch's own kernels (a few per step) are not counted by the context
if per_step and 0 <= extra <= 8 * len(per_step) and extra % len(per_step) == 0:
    for c in per_step:
        bounds.append(bounds[-1] + c + extra // len(per_step))
else:
    bounds.append(len(L))
for s in range(len(bounds) - 1):
    seg = L[bounds[s]:bounds[s + 1]]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v, _, _ in seg:
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    label = names[s] if len(bounds) == 5 else "all"
    lines.append(f"## step {s}: {label} — {len(seg)} launches, {tot / 1e3:.2f} ms of kernel time\n")
    lines.append("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:22]:
        lines.append(f"| `{k[:70]}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |")
    lines.append("")


def full_table(rep, title, want):
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, units = rr[0], rr[1]
    idx = [h.index(w) for w in want if w in h]
    lines.append(f"# {tag}: {title}\n")
    lines.append("| kernel | grid | " + " | ".join(f"{h[i]} [{units[i]}]" for i in idx) + " |")
    lines.append("|---|---|" + "---:|" * len(idx))
    for r in rr[2:]:
        lines.append("| `" + short(r[h.index("Kernel Name")])[:44] + "` | " + r[h.index("Grid Size")] + " | " +
                     " | ".join(r[i] for i in idx) + " |")
    lines.append("")


full_table(os.path.join(G, "prof_gemm.ncu-rep"),
           "`ncu --set full --clock-control none` of splice_gemm_kernel launches of a plain step (round 2 final: 16 consecutive launches of "
           "forward and backward: CTA-pair kernels `<BN, 2, 2, false, true>`, natural-gradient products, the MN-major parameter gradient "
           "`<160, 2, 2, true>`), per launch",
           ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
            "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread"])
for den_rep in ("prof_den.ncu-rep", "prof_den_a.ncu-rep", "prof_den_b.ncu-rep"):
  full_table(os.path.join(G, den_rep), f"`ncu --set full` of denominator kernels ({den_rep}; N = 16384, S = 64), per launch",
            ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
             "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
             "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread"])
bench = os.path.join(G, "bench.log")
if os.path.exists(bench):
    try:
        d = json.loads(open(bench).read().strip().splitlines()[-1])
        lines.append(f"# {tag}: bench line (plain run, no profiler)\n\n```json\n{json.dumps(d, indent=1)}\n```\n")
    except Exception:
        pass
open(os.path.join(out_dir, f"{tag}_summary.md"), "w").write("\n".join(lines) + "\n")
print("wrote", os.path.join(out_dir, f"{tag}_summary.md"))
