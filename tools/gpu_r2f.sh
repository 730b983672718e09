#!/bin/bash
# Round-2 GPU check F: repeat the whole-step parity test (it once tripped on a ReLU sign tie in ~1 process of 8), then the
# multi-GPU script.
N=${1:-2}
mkdir -p gpurun_out
fails=0
for i in $(seq 1 ${2:-20}); do
  timeout 300 python -m pytest tests/test_gpu_step_parity.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/parity_rep_$i.log 2>&1 || { fails=$((fails+1)); grep -n "^E   " gpurun_out/parity_rep_$i.log | cut -c1-400 | head -5; }
done
echo "parity repetitions: failures=$fails"
bash tools/gpu_r2_multi.sh $N
