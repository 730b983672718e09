#!/bin/bash
# Round-2 final GPU check: the whole -m gpu suite exactly as the driver runs it, smoke, the default bench line, the
# bottleneck-search and manual bench lines, the reference arm.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1; echo "pytest -m gpu exit=$? $(tail -1 gpurun_out/gpu_tests.log)"
grep -n "^E   \|FAILED\|Error" gpurun_out/gpu_tests.log | cut -c1-300 | head
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$? $(tail -1 gpurun_out/smoke.log | cut -c1-200)"
timeout 900 python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench exit=$?"; tail -c 600 gpurun_out/bench.log
timeout 600 python bench.py --mode bottleneck --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bottleneck.log 2>gpurun_out/bench_bottleneck.err; echo "bottleneck exit=$?"
timeout 600 python bench.py --mode manual --chunks 128 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_manual.log 2>gpurun_out/bench_manual.err; echo "manual exit=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>gpurun_out/bench_ref.err; echo "reference exit=$?"
python - <<'PY'
import json
for f in ("bench", "bench_bottleneck", "bench_manual", "bench_ref"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.log") if l.startswith("{")][-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", (d.get("e2e") or {}).get("value"), "roofline", (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "unreadable", e)
PY
