#!/bin/bash
# Round-2 GPU check B: changed tests, bench lines (search, bottleneck with the mixing-kernel table).
mkdir -p gpurun_out
for f in train_glue supernet bottleneck_block den; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q --tb=short --maxfail=12 -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "exit=$?" >> gpurun_out/test_$f.log
  echo "== $f: $(tail -2 gpurun_out/test_$f.log | tr '\n' ' ')"
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench exit=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['gemm_ms_per_step'], d['roofline']['skinny_ng_gemm_ms_per_step'], d['den']['ms'], d['cpu_baseline']['value'])
PY
tail -3 gpurun_out/bench.err
timeout 900 python bench.py --mode bottleneck --steps 10 --warmup 3 > gpurun_out/bench_bottleneck.log 2>gpurun_out/bench_bottleneck.err; echo "bottleneck exit=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_bottleneck.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
for m in d['mixing_kernels']: print(m)
PY
tail -3 gpurun_out/bench_bottleneck.err
timeout 600 python bench.py --mode bottleneck --unfused-mask --steps 5 --warmup 3 > gpurun_out/bench_bottleneck_unfused.log 2>gpurun_out/bench_bottleneck_unfused.err; echo "unfused exit=$?"; tail -c 400 gpurun_out/bench_bottleneck_unfused.log | head -c 400; tail -3 gpurun_out/bench_bottleneck_unfused.err
