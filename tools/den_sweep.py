"""BASELINE.json configs[4]: denominator forward-backward sweep over den-graph sizes and chunk lengths, at 1 GPU or under
`python -m torch.distributed.run --nproc-per-node N` (every rank runs its own shard of S sequences: the path shards over
sequences with no collective, so the time is the max over ranks and the bytes add up).

Prints a markdown table (for profiles/): ms, algorithmic GB/s and fraction of the measured HBM peak (SURVEY 8d bytes), and the
L2 row-visit rate that actually bounds the recursion.

  python tools/den_sweep.py                      full sweep, default kernels
  python tools/den_sweep.py --quick              N = 16384, S = 64 / 128, T = 50 only
  python tools/den_sweep.py --variants           each row with: slices (2 parts), slices (1 part), per-frame kernels
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tdnnf_nas_b200 import capi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
ap.add_argument("--variants", action="store_true")
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peaks = json.load(open(pk)) if os.path.exists(pk) else {"hbm_gbs": 6650.0}
ctx = capi.Context(local)
ctx.use_current_stream()
P = 6008
VARIANTS = [("default", {})]
if args.variants:
    VARIANTS = [("slices/2", dict(TDNNF_DEN_PATH="slices", TDNNF_DEN_PARTS="2")), ("slices/1", dict(TDNNF_DEN_PATH="slices", TDNNF_DEN_PARTS="1")),
                ("frames", dict(TDNNF_DEN_PATH="frames"))]
if rank == 0:
    print(f"GPUs: {world} (each rank runs `seqs` sequences of its own; ms = max over ranks; GB/s summed over ranks)")
    print("| states | arcs | seqs/GPU | T | kernels | ms | alg GB/s | frac of HBM peak | L2 row visits TB/s |")
    print("|---:|---:|---:|---:|---|---:|---:|---:|---:|")
sizes = (16384,) if args.quick else (8192, 16384, 32768)
frames = (50,) if args.quick else (17, 34, 50, 67, 100)
for N in sizes:
    graph = synth.make_den_graph(N, P, 16.0, seed=7)
    dg = capi.DenGraph(ctx, graph)
    A = graph["num_arcs"]
    for S in (64, 128):
        for T in frames:
            x = torch.randn(T * S, P, device="cuda")
            d = torch.zeros_like(x)
            for name, env in VARIANTS:
                for k in ("TDNNF_DEN_PATH", "TDNNF_DEN_PARTS", "TDNNF_DEN_CLUSTER"):
                    os.environ.pop(k, None)
                os.environ.update(env)
                dc = capi.DenominatorComputation(ctx, dg, S, T, 0.1)
                desc = dc.describe()
                for _ in range(2):
                    dc.forward(x); dc.backward(-1.0, d)
                e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                e0.record()
                for _ in range(args.reps):
                    dc.forward(x); dc.backward(-1.0, d)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.reps
                if world > 1:
                    t = torch.tensor([ms], device="cuda")
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms = float(t.item())
                alg = world * (4.0 * S * (2 * (T + 1) * (N + 1) + 3 * P * T) + 24.0 * A)
                # row visits out of L2 per arc, sequence and frame-pair: per-frame kernels alpha + E, beta + E + posterior atomics (5);
                # slice kernels alpha, beta + posterior atomics (3), E comes from shared memory
                visits = world * (3.0 if desc["path"] == "slices" else 5.0) * 4.0 * A * S * T
                kern = desc["path"] + (f" (cluster {desc['cluster']}, {desc['parts']} parts, {desc['ctas']} CTAs)" if desc["path"] == "slices" else "")
                if rank == 0:
                    print(f"| {N} | {A} | {S} | {T} | {kern} | {ms:.3f} | {alg / ms / 1e6:.0f} | {alg / ms / 1e6 / peaks['hbm_gbs'] / world:.3f} | "
                          f"{visits / ms / 1e9:.2f} |", flush=True)
                dc.close()
            del x, d
    dg.close()
if world > 1:
    dist.destroy_process_group()
