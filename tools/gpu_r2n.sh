#!/bin/bash
# Sign-tie or race?  The strict form of the trajectory comparison (1e-5 of max |alpha|) on other inputs with PDL on.
T="tests/test_gpu_supernet.py::test_bottleneck_search_step"
for seed in 0 1 2 3; do
  f=0
  for i in 1 2 3 4 5 6; do
    TDNNF_TEST_STRICT=1 TDNNF_TEST_INPUT_SEED=$seed timeout 300 python -m pytest "$T" -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/tie_${seed}_$i.log 2>&1 || f=$((f+1))
  done
  echo "input seed $seed (PDL on, strict): $f failures of 6"
done
