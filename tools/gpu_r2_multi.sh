#!/bin/bash
# Multi-GPU check (run with gpurun --gpus N): the bench at N ranks with the data-parallel self-check (dp_check in the JSON
# line), then the denominator sweep under torch.distributed.run.  BUCKETS=1 adds the overlapped-bucket variant.
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $RUN bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err; echo "bench N=$N exit=$?"
if [ "${BUCKETS:-0}" = "1" ]; then
  timeout 900 $RUN bench.py --gpus $N --steps 10 --warmup 3 --dp-buckets 4 --no-dp-check --no-cpu-baseline > gpurun_out/bench_${N}gpu_b4.log 2> gpurun_out/bench_${N}gpu_b4.err; echo "bench N=$N buckets=4 exit=$?"
fi
python - <<PY
import json, os
for f in ("gpurun_out/bench_${N}gpu.log", "gpurun_out/bench_${N}gpu_b4.log"):
    if not os.path.exists(f):
        continue
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, {k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "e2e", d["e2e"]["value"], "allreduce", d.get("allreduce"), "dp_check", d.get("dp_check"))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 gpurun_out/bench_${N}gpu.err
timeout 600 $RUN tools/den_sweep.py --quick > gpurun_out/den_sweep_${N}gpu.md 2> gpurun_out/den_sweep_${N}gpu.err; echo "den sweep exit=$?"; cat gpurun_out/den_sweep_${N}gpu.md
