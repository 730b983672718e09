"""Generates tests/golden/*.npz from the CPU oracle (seeded).  The reference ships no golden vectors and
cannot be run here, so these pin the ORACLE (drift detection across rounds), not the reference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tdnnf_nas_b200 import synth  # noqa: E402

out_dir = os.path.join(ROOT, "tests", "golden")
os.makedirs(out_dir, exist_ok=True)

# ---- TdnnDARTSV3 forward/backward, Gumbel mode, 5 offsets, row stride 1 and 3
for name, offsets, stride in [("darts_gumbel_s1", [0, 1, 2, 3, 4], 1), ("darts_softmax_s3", [-4, -3, -2, -1, 0], 3)]:
    g = np.random.default_rng(42)
    n, din, dout, S, t_out = len(offsets), 24, 16, 4, 6
    t0 = min(offsets)
    n_t_in = (t_out - 1) * stride + max(offsets) - t0 + 1
    n_t_in = stride * ((n_t_in + stride - 1) // stride)
    rs, ro = synth.regular_row_offsets(offsets, t0, 0, S, 1, stride)
    x = g.standard_normal((n_t_in * S, din)).astype(np.float32)
    W = (g.standard_normal((dout, n * din)) / np.sqrt(n * din)).astype(np.float32)
    bp = g.standard_normal(n + dout).astype(np.float32)
    od = (g.standard_normal((t_out * S, dout)) / (t_out * S)).astype(np.float32)
    ug = g.uniform(0.1, 0.9, n).astype(np.float32)
    flags = (O.USE_GUMBEL | O.UPDATE_ALPHA) if "gumbel" in name else O.UPDATE_ALPHA
    temp, lr = 0.6, 0.02
    out, coef = O.tdnn_propagate(offsets, flags, temp, W, bp, x, t_out * S, ro, stride, ug)
    ind, dW, db = np.zeros_like(x), np.zeros_like(W), np.zeros_like(bp)
    s = O.tdnn_backprop(offsets, flags, temp, W, x, od, coef, ro, stride, lr, in_deriv=ind, dW=dW, dbias=db)
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), offsets=np.array(offsets), stride=stride, flags=flags,
                        temp=temp, lr=lr, row_offsets=np.array(ro), x=x, W=W, bias_params=bp, out_deriv=od, u_gumbel=ug,
                        out=out, coef=coef, in_deriv=ind, dW=dW, dbias=db, s=s)

# ---- Gumbel softmax flops
g = np.random.default_rng(7)
x = g.standard_normal((12, 8)).astype(np.float32)
u = g.uniform(0.1, 0.9, 8).astype(np.float32)
p = O.softmax_flops_fwd(x, u, 0.5)
od = g.standard_normal((12, 8)).astype(np.float32)
ind, od_after = O.softmax_flops_bwd(p, od, 0.1, True, 0.5)
np.savez_compressed(os.path.join(out_dir, "gumbel_softmax_flops.npz"), x=x, u=u, temp=0.5, scale=0.1, p=p, out_deriv=od,
                    in_deriv=ind, out_deriv_after=od_after)

# ---- denominator
graph = synth.make_den_graph(24, 9, 3.0, seed=2)
g = np.random.default_rng(9)
S, T = 3, 4
xo = g.standard_normal((T * S, 9)).astype(np.float32)
lp, deriv, ok = O.den_forward_backward(graph, xo, S, T, 0.1, deriv_weight=-1.0)
np.savez_compressed(os.path.join(out_dir, "den_small.npz"), nnet_output=xo, S=S, T=T, leaky=0.1, logprob=lp, deriv=deriv,
                    ok=ok, **{"g_" + k: np.asarray(v) for k, v in graph.items()})
print("wrote", sorted(os.listdir(out_dir)))

# ---- stock TdnnComponent (natural-gradient update over two minibatches) and ConstrainOrthonormal
only_new = "--only-new" in sys.argv
g = np.random.default_rng(21)
offsets, stride, S, t_out, din, dout = [-3, 0], 1, 4, 7, 24, 10
n = len(offsets)
rs, ro = synth.regular_row_offsets(offsets, min(offsets), 0, S, 1, stride)
n_t_in = t_out + 3
W = (g.standard_normal((dout, n * din)) / np.sqrt(n * din)).astype(np.float32)
b = g.standard_normal(dout).astype(np.float32)
xs = [g.standard_normal((n_t_in * S, din)).astype(np.float32) for _ in range(2)]
ods = [(g.standard_normal((t_out * S, dout)) / (t_out * S)).astype(np.float32) for _ in range(2)]
out = O.plain_tdnn_propagate(W, b, xs[0], t_out * S, ro, stride)
ng_in, ng_out = O.NaturalGradient(8, 4, 2000.0, 4.0), O.NaturalGradient(5, 4, 2000.0, 4.0)
ind, dW, db = np.zeros_like(xs[0]), np.zeros_like(W), np.zeros_like(b)
scales = np.zeros((2, 2), np.float32)
for k in range(2):
    O.plain_tdnn_backprop(W, xs[k], ods[k], ro, stride, 0.02, in_deriv=ind if k == 0 else None, dW=dW, dbias=db,
                          natural_gradient=True, ng_in=ng_in, ng_out=ng_out, scales=scales[k])
np.savez_compressed(os.path.join(out_dir, "plain_tdnn_ng.npz"), offsets=np.array(offsets), row_offsets=np.array(ro), W=W, bias=b,
                    x0=xs[0], x1=xs[1], od0=ods[0], od1=ods[1], lr=0.02, rank_in=8, rank_out=5, out=out, in_deriv=ind, dW=dW,
                    dbias=db, scales=scales)
q, _ = np.linalg.qr(g.standard_normal((40, 12)))
M = (0.8 * q.T + 0.3 / np.sqrt(40) * g.standard_normal((12, 40))).astype(np.float32)
fl, fx = M.copy(), M.copy()
info_fl = O.constrain_orthonormal(fl, -1.0)
info_fx = O.constrain_orthonormal(fx, 1.0)
np.savez_compressed(os.path.join(out_dir, "constrain_orthonormal.npz"), M=M, floating=fl, info_floating=info_fl, fixed=fx,
                    info_fixed=info_fx)
print("wrote", sorted(os.listdir(out_dir)))
