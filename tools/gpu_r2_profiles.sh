#!/bin/bash
# Round-2 ncu evidence (after the bench command has exited 0 without ncu): launch list of one natural-gradient period,
# --set full of 16 consecutive GEMM launches of a plain step, --set full of the denominator frame kernels.
mkdir -p gpurun_out
CMD="python tools/profile_step.py --warmup 14 --steps 4"
$CMD > gpurun_out/steps.json 2> gpurun_out/steps.err && cat gpurun_out/steps.json &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
CMD2="python tools/profile_step.py --warmup 16 --steps 1"
$CMD2 > gpurun_out/plain2.log 2>&1 &&
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:splice_gemm -s 40 -c 16 -o gpurun_out/prof_gemm -f $CMD2 > gpurun_out/ncu_gemm.log 2>&1
echo "gemm full rc=$?"
$CMD2 > gpurun_out/plain3.log 2>&1 &&
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:den_ -s 10 -c 3 -o gpurun_out/prof_den_a -f $CMD2 > gpurun_out/ncu_den.log 2>&1
echo "den alpha full rc=$?"; ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
