#!/bin/bash
# Which file's kernels need their programmatic dependent launch switched off for the flaky test to pass?
T="tests/test_gpu_supernet.py::test_bottleneck_search_step"
N=${1:-8}
for off in mixing neighbours splice_gemm den. num ng context; do
  f=0
  for i in $(seq 1 $N); do
    TDNNF_PDL_OFF=$off timeout 300 python -m pytest "$T" -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/bis_${off}_$i.log 2>&1 || f=$((f+1))
  done
  echo "PDL off in $off: $f failures of $N"
done
