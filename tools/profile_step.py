"""Steady-state step timing / profiling window for the bench workload.

  python tools/profile_step.py [--warmup 14] [--steps 4]
      per-step CUDA-event times (ms) after `warmup` steps (the natural-gradient Fisher estimates refresh on each of
      the first 10 calls, then on every 4th: 14 warm-up steps put the window on one full period, first step = refresh).
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X \
      python tools/profile_step.py --steps 4
      the same window bracketed by cudaProfilerStart/Stop, for the launch list.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tdnnf_nas_b200.supernet import Supernet, SupernetConfig

ap = argparse.ArgumentParser()
ap.add_argument("--warmup", type=int, default=14)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--mode", default="search")
ap.add_argument("--chunks", type=int, default=64)
ap.add_argument("--no-xent", action="store_true")
ap.add_argument("--phases", action="store_true", help="also time forward / objective / backward / update separately")
args = ap.parse_args()

net = Supernet(SupernetConfig(mode=args.mode, num_seqs=args.chunks, l2_regularize=0.01 if args.mode == "manual" else 0.0,
                              xent=not args.no_xent), device=0)
net.x.copy_(net.make_input(0))
for _ in range(args.warmup):
    net.step(None)
torch.cuda.synchronize()
torch.cuda.profiler.start()
times, launches = [], []
for _ in range(args.steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = net.ctx.launches
    e0.record()
    net.step(None)
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
    launches.append(net.ctx.launches - l0)
torch.cuda.profiler.stop()
out = dict(step_ms=times, launches=launches)
if args.phases:
    ph = []
    for _ in range(args.steps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        net.fwd_plan.run()
        ev[1].record()
        net.objective.compute(net.head["out"], net.head["d_out"])
        ev[2].record()
        net.bwd_plan.run()
        ev[3].record()
        if net.cfg.l2_regularize != 0.0:
            net._apply_l2_regularization()
        net._update_with_max_change()
        net._after_update()
        ev[4].record()
        torch.cuda.synchronize()
        ph.append(dict(zip(["fwd", "objf", "bwd", "update"], [ev[i].elapsed_time(ev[i + 1]) for i in range(4)])))
    out["phases_ms"] = ph
print(json.dumps(out))
net.close()
