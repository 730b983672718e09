#!/bin/bash
# Round-2 GPU check G: CTA-pair (cta_group::2) GEMM kernels: parity tests, then the step with the pair kernels on / off.
mkdir -p gpurun_out
for f in fullsize darts step_parity supernet ng; do
  timeout 600 python -m pytest tests/test_gpu_$f.py -m gpu -q --tb=short --maxfail=8 -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "exit=$?" >> gpurun_out/test_$f.log
  echo "== $f: $(tail -2 gpurun_out/test_$f.log | tr '\n' ' ')"
done
grep -n "^E   \|tdnnf:" gpurun_out/test_fullsize.log gpurun_out/test_darts.log gpurun_out/test_step_parity.log | cut -c1-400 | head -20
for pair in 1 0; do
  TDNNF_GEMM_PAIR=$pair timeout 600 python tools/profile_step.py --warmup 14 --steps 4 --phases > gpurun_out/steps_pair$pair.json 2>gpurun_out/steps_pair$pair.err; echo "pair=$pair: $(cat gpurun_out/steps_pair$pair.json | cut -c1-400)"
done
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench exit=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['gemm_ms_per_step'], d['roofline']['skinny_ng_gemm_ms_per_step'], d['den']['ms'])
PY
tail -3 gpurun_out/bench.err
