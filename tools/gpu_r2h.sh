#!/bin/bash
# Round-2 GPU check H: ncu evidence for the current step: launch list of one natural-gradient period and --set full of the
# GEMM launches of a backward pass (CTA-pair kernels included).
mkdir -p gpurun_out
CMD="python tools/profile_step.py --warmup 14 --steps 4"
$CMD > gpurun_out/steps.json 2> gpurun_out/steps.err && cat gpurun_out/steps.json &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
CMD2="python tools/profile_step.py --warmup 16 --steps 1"
$CMD2 > gpurun_out/plain2.log 2>&1 &&
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:splice_gemm -s 62 -c 14 -o gpurun_out/prof_gemm_pair -f $CMD2 > gpurun_out/ncu_gemm.log 2>&1
echo "gemm full rc=$?"; ls -la gpurun_out/*.ncu-rep
