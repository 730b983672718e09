"""Per-op timing of the hot kernels at supernet shapes (CUDA events; not the contract bench)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tdnnf_nas_b200 import capi, synth

ctx = capi.Context(0); ctx.use_current_stream()
dev = "cuda"

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

res = {}
S = 64
for name, din, dout, offsets, t_out in [("linear_1536_160", 1536, 160, list(range(-6, 1)), 310), ("affine_160_1536", 160, 1536, list(range(0, 7)), 304)]:
    n = 7
    t_in = t_out + 6
    rs, ro = synth.regular_row_offsets(offsets, min(offsets), 0, S, 1, 1)
    in_rows, out_rows = t_in * S, t_out * S
    x = torch.randn(in_rows, din, device=dev); W = torch.randn(dout, n * din, device=dev) * 0.01
    od = torch.randn(out_rows, dout, device=dev); weff = torch.full((n,), 0.3, device=dev)
    out = torch.zeros(out_rows, dout, device=dev); ind = torch.zeros(in_rows, din, device=dev)
    dW = torch.zeros(dout, n * din, device=dev); db = torch.zeros(dout, device=dev); s = torch.zeros(n, device=dev)
    bias = torch.zeros(dout, device=dev)
    flop = 2.0 * out_rows * n * din * dout
    t = timeit(lambda: ctx.darts_propagate(x, out, W, bias, 2, weff, ro, 1))
    res[name + "_fwd"] = dict(ms=t, tflops_fp32eq=flop / t / 1e9, tflops_bf16_raw=3 * flop / t / 1e9)
    t = timeit(lambda: ctx.darts_backprop_data(od, ind, W, weff, ro, 1))
    res[name + "_dgrad"] = dict(ms=t, tflops_fp32eq=flop / t / 1e9, tflops_bf16_raw=3 * flop / t / 1e9)
    # parameter gradient, both operand forms (the pre-passes are inside the call: nothing is cached here)
    for label, min_rows in (("wgrad_mn_major", 1), ("wgrad_transposed_planes", -1)):
        ctx.set_wgrad_mn_min_rows(min_rows)
        t = timeit(lambda: ctx.darts_backprop_params(x, od, W, dW, db, weff, ro, 1, 1e-3, s))
        res[name + "_" + label] = dict(ms=t, tflops_fp32eq=flop / t / 1e9, tflops_bf16_raw=3 * flop / t / 1e9)
    ctx.set_wgrad_mn_min_rows(512)

for N, Sd, T in [(8192, 64, 50), (16384, 64, 50), (32768, 128, 50)]:
    P = 6008
    graph = synth.make_den_graph(N, P, 16.0, seed=7)
    dg = capi.DenGraph(ctx, graph)
    dc = capi.DenominatorComputation(ctx, dg, Sd, T, 0.1)
    xo = torch.randn(T * Sd, P, device=dev)
    deriv = torch.zeros_like(xo)
    def fb():
        dc.forward(xo); dc.backward(-1.0, deriv)
    t = timeit(fb, iters=3, warm=1)
    A = graph["num_arcs"]
    bytes_alg = 4.0 * Sd * (2 * (T + 1) * (N + 1) + 3 * P * T) + 12.0 * A * 2
    res[f"den_N{N}_S{Sd}_T{T}"] = dict(ms=t, arcs=A, alg_GBps=bytes_alg / t / 1e6, arc_visits_per_s=2.0 * A * Sd * T / t * 1e3)
    dc.close(); dg.close()
print(json.dumps(res, indent=1))
