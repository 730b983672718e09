#!/bin/bash
# Repeat one test with programmatic dependent launch on / off (flake triage).
T=${1:-tests/test_gpu_supernet.py::test_bottleneck_search_step}
N=${2:-10}
for pdl in 1 0; do
  f=0
  for i in $(seq 1 $N); do
    TDNNF_PDL=$pdl timeout 300 python -m pytest "$T" -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/rep_${pdl}_$i.log 2>&1 || { f=$((f+1)); grep -h "Max relative\|Mismatched\|^E  " gpurun_out/rep_${pdl}_$i.log | head -3; }
  done
  echo "PDL=$pdl: $f failures of $N"
done
