#!/bin/bash
# Round-2 GPU check K: NG tests + parity, NG helper kernel durations, steps.
mkdir -p gpurun_out
for f in ng step_parity fullsize; do
  timeout 600 python -m pytest tests/test_gpu_$f.py -m gpu -q --tb=short --maxfail=8 -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "== $f: $(tail -1 gpurun_out/test_$f.log)"
done
grep -n "^E   \|tdnnf:" gpurun_out/test_ng.log gpurun_out/test_step_parity.log | cut -c1-300 | head
timeout 600 python tools/profile_step.py --warmup 14 --steps 4 --phases > gpurun_out/steps.json 2>gpurun_out/steps.err; cut -c1-200 gpurun_out/steps.json
bash tools/gpu_r2j.sh
