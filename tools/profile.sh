#!/bin/bash
# ncu evidence for profiles/: (1) launch list of one bench step, (2) --set full of the GEMM kernel.
# Each ncu command is preceded by the identical plain command exiting 0 (B200_PROFILING.md).
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 3250 -c 800 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:splice_gemm -s 90 -c 6 -o gpurun_out/prof_gemm -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; ls -la gpurun_out/*.ncu-rep
