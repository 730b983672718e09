"""profiles/gemm_traffic.json from an `ncu --set full` capture of splice_gemm_kernel launches: mean dram__bytes_read.sum +
dram__bytes_write.sum per launch (the `traffic` field of bench.py's roofline object is read from this file at run time).

  python tools/gemm_traffic.py gpurun_out/prof_gemm.ncu-rep "r02 capture of ..."
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
note = sys.argv[2] if len(sys.argv) > 2 else os.path.basename(rep)
min_us = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0  # only launches at least this long (the TdnnDARTSV3 GEMMs, not the skinny products)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
per = []
for r in rows[2:]:
    if "splice_gemm_kernel" not in r[col["Kernel Name"]]:
        continue
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(r[col[m]].replace(",", "")) * scale[units[col[m]]]
    per.append(dict(kernel=r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", ""), dram_bytes=tot,
                    duration_us=float(r[col["gpu__time_duration.sum"]].replace(",", "")) * {"us": 1.0, "ns": 1e-3, "ms": 1e3}[units[col["gpu__time_duration.sum"]]],
                    tensor_active_pct=float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]])))
per = [p for p in per if p["duration_us"] >= min_us]
out = dict(source=note, launches=len(per), mean_dram_bytes_per_launch=sum(p["dram_bytes"] for p in per) / max(1, len(per)), per_launch=per)
json.dump(out, open(os.path.join(ROOT, "profiles", "gemm_traffic.json"), "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "per_launch"}))
