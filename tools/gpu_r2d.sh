#!/bin/bash
# Round-2 GPU check D: whole-step parity, then the ncu evidence (launch list of one natural-gradient period; --set full of the
# GEMM launches of a backward pass incl. the MN-major parameter gradient; --set full of the denominator frame kernels).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_step_parity.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/test_step_parity.log 2>&1
echo "== step_parity: $(tail -2 gpurun_out/test_step_parity.log | tr '\n' ' ')"; grep -n "^E   Assert\|^E   assert not" gpurun_out/test_step_parity.log | cut -c1-1500 | head -6
CMD="python tools/profile_step.py --warmup 14 --steps 4"
$CMD > gpurun_out/steps.json 2> gpurun_out/steps.err && cat gpurun_out/steps.json &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
CMD2="python tools/profile_step.py --warmup 16 --steps 1"
$CMD2 > gpurun_out/plain2.log 2>&1 &&
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:splice_gemm -s 62 -c 14 -o gpurun_out/prof_gemm -f $CMD2 > gpurun_out/ncu_gemm.log 2>&1
echo "gemm full rc=$?"
$CMD2 > gpurun_out/plain3.log 2>&1 &&
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:den_ -s 10 -c 3 -o gpurun_out/prof_den_a -f $CMD2 > gpurun_out/ncu_den.log 2>&1
echo "den alpha full rc=$?"
$CMD2 > gpurun_out/plain4.log 2>&1 &&
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:den_beta -s 10 -c 2 -o gpurun_out/prof_den_b -f $CMD2 > gpurun_out/ncu_den_b.log 2>&1
echo "den beta full rc=$?"; ls -la gpurun_out/*.ncu-rep
