#!/bin/bash
# Round-2 GPU check A: the new slice denominator first (own process, short timeout), then every -m gpu test (one process
# per file), smoke, the den sweep (variants) and a short bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 600 python -m pytest tests/test_gpu_den.py -m gpu -q --tb=short --maxfail=8 -p no:cacheprovider > gpurun_out/test_den.log 2>&1
echo "exit=$?" >> gpurun_out/test_den.log; echo "== den: $(tail -3 gpurun_out/test_den.log | tr '\n' ' ')"
for f in train_glue darts mixing neighbours ng nnet3 tdnn_plain dropout bottleneck_block fullsize supernet; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q --tb=short --maxfail=12 -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "exit=$?" >> gpurun_out/test_$f.log
  echo "== $f: $(tail -2 gpurun_out/test_$f.log | tr '\n' ' ')"
done
timeout 600 python -m pytest tests/test_zz_gpu_reference_compat.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/test_zz.log 2>&1; echo "== zz: $(tail -1 gpurun_out/test_zz.log)"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python tools/den_sweep.py --quick --variants > gpurun_out/den_sweep_quick.md 2>gpurun_out/den_sweep_quick.err; echo "sweep exit=$?"; cat gpurun_out/den_sweep_quick.md
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench exit=$?"; tail -c 3000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
