#!/bin/bash
# Round-2 GPU check E: operand planes kept / produced (tests + bench), then the 1-GPU den sweep.
mkdir -p gpurun_out
for f in neighbours step_parity darts ng supernet nnet3 tdnn_plain fullsize; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q --tb=short --maxfail=12 -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "exit=$?" >> gpurun_out/test_$f.log
  echo "== $f: $(tail -2 gpurun_out/test_$f.log | tr '\n' ' ')"
done
grep -n "^E   " gpurun_out/test_neighbours.log gpurun_out/test_step_parity.log | cut -c1-600 | head -12
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench exit=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['gemm_ms_per_step'], d['roofline']['skinny_ng_gemm_ms_per_step'], d['den']['ms'])
PY
tail -3 gpurun_out/bench.err
timeout 600 python tools/profile_step.py --warmup 14 --steps 4 --phases > gpurun_out/steps_planes.json 2>gpurun_out/steps_planes.err; cat gpurun_out/steps_planes.json
