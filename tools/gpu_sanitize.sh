#!/bin/bash
# compute-sanitizer over the small GPU parity tests (memcheck, then racecheck and synccheck on the kernels that use shared
# memory / clusters).  NOT run in round 2: the round's GPU budget went to parity, bench and ncu; this is the first thing to
# spend GPU minutes on next.  Slow (10-50x): bounded to the small-shape tests; summaries land in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_sanitize.sh'
set -u
mkdir -p gpurun_out
run() {  # tool, log name, pytest args
  local tool=$1 log=$2; shift 2
  timeout 600 compute-sanitizer --tool "$tool" --error-exitcode 9 --print-limit 20 \
    python -m pytest "$@" -m gpu -x -q -p no:cacheprovider > "gpurun_out/sanitize_${log}.log" 2>&1
  echo "$tool $log: exit $? ($(grep -c 'ERROR SUMMARY' gpurun_out/sanitize_${log}.log) summaries, $(grep -h 'ERROR SUMMARY' gpurun_out/sanitize_${log}.log | tail -1))"
}
run memcheck mixing tests/test_gpu_mixing.py tests/test_gpu_dropout.py
run memcheck den tests/test_gpu_den.py::test_den_parity tests/test_gpu_den.py::test_numerator_parity
run memcheck darts tests/test_gpu_darts.py -k "adds_mode or mn_major or project"
run racecheck den tests/test_gpu_den.py::test_den_parity tests/test_gpu_den.py::test_numerator_parity
run racecheck ng tests/test_gpu_ng.py -k "precondition"
run synccheck darts tests/test_gpu_darts.py -k "adds_mode or mn_major"
