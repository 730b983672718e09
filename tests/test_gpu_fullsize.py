"""GPU checks at BASELINE.json's FULL sizes (configs[2]: 64 chunks x 150 frames, TdnnDARTSV3 1536<->160, 7 offsets), where
the CPU oracle would take minutes: a float64 numpy reference on RANDOM SAMPLES of the outputs (rows of the forward
pass and of the data gradient, columns of the parameter gradient, all s_i), plus size-independent properties
(linearity in the input, additivity over sequence shards = the data-parallel contract of SURVEY 8e)."""
import numpy as np
import pytest

from tests.util import rel_err

pytestmark = pytest.mark.gpu

# (name, in_dim, out_dim, offsets, input frames): the widest layers of the supernet (R_out 19 840 / 19 456)
LAYERS = [
    ("tdnnf2.linear", 1536, 160, list(range(-6, 1)), 316),
    ("tdnnf2.affine", 160, 1536, list(range(0, 7)), 310),
]


def _layer(name, in_dim, out_dim, offsets, frames_in, S=64, seed=0):
    from tdnnf_nas_b200 import synth

    g = np.random.default_rng(seed)
    n = len(offsets)
    t_out = frames_in - (n - 1)
    _, ro = synth.regular_row_offsets(offsets, min(offsets), 0, S, 1, 1)
    x = g.standard_normal((frames_in * S, in_dim)).astype(np.float32)
    W = (g.standard_normal((out_dim, n * in_dim)) / np.sqrt(n * in_dim)).astype(np.float32)
    bias = g.standard_normal(out_dim).astype(np.float32)
    w = g.uniform(0.05, 1.0, n).astype(np.float32)
    od = (g.standard_normal((t_out * S, out_dim)) / (t_out * S)).astype(np.float32)
    return g, n, t_out * S, ro, x, W, bias, w, od


@pytest.mark.parametrize("layer", LAYERS, ids=[l[0] for l in LAYERS])
def test_fullsize_gemms_on_samples(ctx, layer):
    import torch

    name, in_dim, out_dim, offsets, frames_in = layer
    g, n, out_rows, ro, x, W, bias, w, od = _layer(*layer)
    xd, Wd, odd = torch.from_numpy(x).cuda(), torch.from_numpy(W).cuda(), torch.from_numpy(od).cuda()
    wd, bd = torch.from_numpy(w).cuda(), torch.from_numpy(bias).cuda()
    x64, W64, od64, w64 = x.astype(np.float64), W.astype(np.float64), od.astype(np.float64), w.astype(np.float64)

    # ---- Propagate: 512 random output rows (every m-tile position class: first / last rows included)
    out = torch.empty((out_rows, out_dim), device="cuda")
    ctx.darts_propagate(xd, out, Wd, bd, 2, wd, ro, 1)
    rows = np.unique(np.concatenate([[0, 1, 127, 128, out_rows - 1], g.integers(0, out_rows, 512)]))
    ref = np.tile(bias.astype(np.float64), (len(rows), 1))
    for i in range(n):
        ref += w64[i] * x64[rows + ro[i]] @ W64[:, i * in_dim:(i + 1) * in_dim].T
    assert rel_err(out[torch.from_numpy(rows).cuda()].cpu().numpy(), ref) < 1e-4
    # linearity: Propagate(2 x) - bias == 2 (Propagate(x) - bias) to rounding (size-independent property)
    out2 = torch.empty_like(out)
    ctx.darts_propagate(2.0 * xd, out2, Wd, bd, 2, wd, ro, 1)
    assert rel_err((out2 - bd).cpu().numpy(), 2.0 * (out - bd).cpu().numpy()) < 2e-5

    # ---- data gradient: 512 random input rows; in_deriv[r] = sum_i w_i out_deriv[r - off_i] W_i over valid rows
    in_deriv = torch.zeros_like(xd)
    ctx.darts_backprop_data(odd, in_deriv, Wd, wd, ro, 1)
    in_rows = x.shape[0]
    rows = np.unique(np.concatenate([[0, in_rows - 1], g.integers(0, in_rows, 512)]))
    ref = np.zeros((len(rows), in_dim))
    for i in range(n):
        k = rows - ro[i]
        ok = (k >= 0) & (k < out_rows)
        ref[ok] += w64[i] * od64[k[ok]] @ W64[:, i * in_dim:(i + 1) * in_dim]
    assert rel_err(in_deriv[torch.from_numpy(rows).cuda()].cpu().numpy(), ref) < 1e-3

    # ---- parameter gradient: 96 random columns of dW (K = all 19k rows), the bias gradient and every s_i
    lr = 0.5
    dW = torch.zeros_like(Wd)
    db = torch.zeros(out_dim, device="cuda")
    s = torch.zeros(n, device="cuda")
    ctx.darts_backprop_params(xd, odd, Wd, dW, db, wd, ro, 1, lr, s)
    cols = np.unique(g.integers(0, n * in_dim, 96))
    ref = np.zeros((out_dim, len(cols)))
    for j, c in enumerate(cols):
        i, cc = divmod(int(c), in_dim)
        ref[:, j] = lr * w64[i] * od64.T @ x64[ro[i]: ro[i] + out_rows, cc]
    assert rel_err(dW[:, torch.from_numpy(cols).cuda()].cpu().numpy(), ref) < 1e-3
    assert rel_err(db.cpu().numpy(), lr * od64.sum(0)) < 1e-3
    s_ref = np.array([((x64[ro[i]: ro[i] + out_rows] @ W64[:, i * in_dim:(i + 1) * in_dim].T) * od64).sum() for i in range(n)])
    assert np.abs(s.cpu().numpy() - s_ref).max() < 1e-3 * np.abs(s_ref).max()

    # ---- additivity over sequence shards: gradients of the 64-sequence minibatch == sum over two
    # 32-sequence shards (rows are t-major, sequence fastest: shard g takes sequences [32 g, 32 g + 32))
    S, half = 64, 32
    from tdnnf_nas_b200 import synth

    _, ro_h = synth.regular_row_offsets(offsets, min(offsets), 0, half, 1, 1)
    dW_sum = torch.zeros_like(Wd)
    for sh in range(2):
        xs = xd.view(-1, S, in_dim)[:, sh * half:(sh + 1) * half].reshape(-1, in_dim).contiguous()
        ods = odd.view(-1, S, out_dim)[:, sh * half:(sh + 1) * half].reshape(-1, out_dim).contiguous()
        ctx.darts_backprop_params(xs, ods, None, dW_sum, None, wd, ro_h, 1, lr, None)
    assert rel_err(dW_sum.cpu().numpy(), dW.cpu().numpy()) < 2e-5


def test_fullsize_denominator_properties(ctx):
    """configs[4] at the bench size (16 384 states, 6008 pdfs, 64 sequences, 50 frames): the posteriors of every frame of
    every sequence sum to one, and the log-probability is additive over sequence shards."""
    import torch

    from tdnnf_nas_b200 import capi, synth

    N, P, S, T = 16384, 6008, 64, 50
    graph = synth.make_den_graph(N, P, 16.0, seed=5)
    dg = capi.DenGraph(ctx, graph)
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn((T * S, P), device="cuda", generator=g).clamp_(-30, 30)

    def run(xm, seqs):
        den = capi.DenominatorComputation(ctx, dg, seqs, T, 0.1)
        lp = den.forward(xm)
        d = torch.zeros_like(xm)
        ok = den.backward(-1.0, d)
        den.close()
        return lp, d, ok

    lp, deriv, ok = run(x, S)
    assert ok and np.isfinite(lp)
    occ = -deriv.sum(dim=1)  # deriv = -posterior
    assert float((occ - 1.0).abs().max()) < 2e-3
    assert float(deriv.max()) <= 1e-6  # posteriors are non-negative
    halves = []
    for sh in range(2):
        xs = x.view(T, S, P)[:, sh * 32:(sh + 1) * 32].reshape(-1, P).contiguous()
        halves.append(run(xs, 32))
    assert abs(halves[0][0] + halves[1][0] - lp) < 1e-4 * abs(lp)
    d_cat = torch.cat([h[1].view(T, 32, P) for h in halves], dim=1).reshape(-1, P)  # (T, 64, P): shard 0 then shard 1
    assert rel_err(d_cat.cpu().numpy(), deriv.cpu().numpy()) < 1e-3
    dg.close()
