#!/bin/bash
# Round-2 GPU check I: ncu launch list of one natural-gradient period (no --set full).
mkdir -p gpurun_out
CMD="python tools/profile_step.py --warmup 14 --steps 4"
$CMD > gpurun_out/steps.json 2> gpurun_out/steps.err && cat gpurun_out/steps.json &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
