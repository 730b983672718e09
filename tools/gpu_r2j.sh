#!/bin/bash
# Round-2 GPU check J: ncu durations of the natural-gradient helper kernels only (one period).
mkdir -p gpurun_out
CMD="python tools/profile_step.py --warmup 14 --steps 4"
ncu --profile-from-start off -k regex:'ng_|copy_blocks|stack_|mat_axpy_dev' --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ng.csv $CMD > gpurun_out/ncu_launches_ng.log 2>&1
echo "rc=$?"; python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches_ng.csv')))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hi]; ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
agg = collections.defaultdict(list)
for r in rows[hi + 1:]:
    if len(r) > vi:
        v = float(r[vi].replace(',', '')); v = v / 1e3 if r[ui] == 'ns' else v
        agg[r[ki][:60]].append(v)
for k, v in agg.items():
    print(k, len(v), 'mean us', round(sum(v) / len(v), 2), 'min', round(min(v), 2), 'max', round(max(v), 2))
PY
