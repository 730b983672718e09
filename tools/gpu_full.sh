#!/bin/bash
# Full GPU check: every -m gpu test (one process per file so that a fault cannot poison the rest),
# smoke, a bench line, the ncu launch list and one --set full capture of the GEMM kernel.
mkdir -p gpurun_out
for f in darts mixing den neighbours ng nnet3 tdnn_plain dropout bottleneck_block fullsize supernet; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q --tb=short --maxfail=12 -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "exit=$?" >> gpurun_out/test_$f.log
  echo "== $f: $(tail -2 gpurun_out/test_$f.log | tr '\n' ' ')"
done
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench exit=$?"; tail -c 1500 gpurun_out/bench.log
