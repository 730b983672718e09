"""BASELINE.json configs[4]: denominator forward-backward sweep over den-graph sizes and chunk lengths.
Prints a markdown table (for profiles/): ms, algorithmic GB/s and fraction of the measured HBM peak,
and the L2 gather rate that actually bounds the recursion."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tdnnf_nas_b200 import capi, synth  # noqa: E402

peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
ctx = capi.Context(0)
ctx.use_current_stream()
P = 6008
print("| states | arcs | seqs | T | ms | alg GB/s | frac of HBM peak | L2 gather TB/s |")
print("|---:|---:|---:|---:|---:|---:|---:|---:|")
for N in (8192, 16384, 32768):
    graph = synth.make_den_graph(N, P, 16.0, seed=7)
    dg = capi.DenGraph(ctx, graph)
    A = graph["num_arcs"]
    for S in (64, 128):
        for T in (17, 34, 50, 67, 100):
            dc = capi.DenominatorComputation(ctx, dg, S, T, 0.1)
            x = torch.randn(T * S, P, device="cuda")
            d = torch.zeros_like(x)
            for _ in range(2):
                dc.forward(x); dc.backward(-1.0, d)
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                dc.forward(x); dc.backward(-1.0, d)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            alg = 4.0 * S * (2 * (T + 1) * (N + 1) + 3 * P * T) + 24.0 * A
            gather = 16.0 * A * S * T
            print(f"| {N} | {A} | {S} | {T} | {ms:.3f} | {alg / ms / 1e6:.0f} | {alg / ms / 1e6 / peaks['hbm_gbs']:.3f} | {gather / ms / 1e9:.2f} |")
            dc.close()
            del x, d
    dg.close()
