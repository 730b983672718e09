#!/bin/bash
# Round-2 GPU check C: whole-step parity against the CPU reference, the natural-gradient changes, a bench line with the
# measured CPU baseline.
mkdir -p gpurun_out
for f in step_parity ng darts tdnn_plain supernet; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q --tb=short --maxfail=12 -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "exit=$?" >> gpurun_out/test_$f.log
  echo "== $f: $(tail -2 gpurun_out/test_$f.log | tr '\n' ' ')"
done
grep -n "^E " gpurun_out/test_step_parity.log | head -20
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench exit=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['gemm_ms_per_step'], d['roofline']['skinny_ng_gemm_ms_per_step'], d['den']['ms'])
print(d['cpu_baseline'])
PY
tail -3 gpurun_out/bench.err
nproc; free -g | head -2
