"""CPU tests of the oracle itself (no GPU): the C++ restatement must agree with an independent
float64 numpy restatement and with finite differences of the forward pass."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import np_ref as R
from tests.util import rel_err

MODES = [0, 1, 2, 4, 5, 1 | 8, 2 | 8, 0 | 16]


def _case(n, din, dout, S, t_out, offsets, row_stride, seed):
    from tdnnf_nas_b200 import synth

    g = np.random.default_rng(seed)
    t0 = min(offsets)
    n_t_in = (t_out - 1) * row_stride + max(offsets) - t0 + 1
    n_t_in = row_stride * ((n_t_in + row_stride - 1) // row_stride)
    rs, ro = synth.regular_row_offsets(offsets, t0, 0, S, 1, row_stride)
    x = g.standard_normal((n_t_in * S, din)).astype(np.float32)
    W = (g.standard_normal((dout, n * din)) / np.sqrt(n * din)).astype(np.float32)
    bp = g.standard_normal(n + dout).astype(np.float32)
    od = (g.standard_normal((t_out * S, dout)) / (t_out * S)).astype(np.float32)
    return g, x, W, bp, od, ro


@pytest.mark.parametrize("flags", MODES)
@pytest.mark.parametrize("offsets,row_stride", [([0, 1, 2, 3], 1), ([-3, -2, -1, 0], 1), ([0, 3, 6], 3), ([-2, -1], 1)])
def test_oracle_matches_numpy(flags, offsets, row_stride):
    n = len(offsets)
    g, x, W, bp, od, ro = _case(n, 24, 20, 3, 11, offsets, row_stride, seed=flags * 10 + n)
    T = 0.7
    ug = g.uniform(0.05, 0.95, n).astype(np.float32)
    uu = float(g.uniform())
    out_rows = od.shape[0]
    share = O.share_index(offsets)
    out, coef = O.tdnn_propagate(offsets, flags, T, W, bp, x, out_rows, ro, row_stride, ug, uu)
    c = R.coef(bp[:n], flags, T, ug, uu)
    np.testing.assert_allclose(coef, c, rtol=3e-6)
    w = R.weights(c, flags, share)
    ref = R.propagate(W, bp[n:] if offsets[1] > 0 else None, x, out_rows, ro, row_stride, w)
    assert rel_err(out, ref) < 2e-6

    lr = 0.05
    ind = np.zeros_like(x)
    dW = np.zeros_like(W)
    db = np.zeros(n + W.shape[0], np.float32)
    s = O.tdnn_backprop(offsets, flags, T, W, x, od, coef, ro, row_stride, lr, in_deriv=ind, dW=dW, dbias=db)
    ind_r, dW_r, db_r, s_r, da_r = R.backprop(W, x, od, ro, row_stride, w, c, flags, T, share, lr)
    assert rel_err(ind, ind_r) < 5e-6
    assert rel_err(dW, dW_r) < 5e-6
    assert rel_err(db[n:], db_r) < 5e-6
    assert rel_err(s, s_r) < 1e-4
    assert np.abs(db[:n] - da_r).max() <= 1e-4 * (np.abs(da_r).max() + 1e-30)


def test_share_index_quirk():
    # Q1: uninitialised share_offset_index when time_offsets[1] == 0 or n == 1 -> the oracle refuses
    assert O.share_index([0, 1, 2]) == 0
    assert O.share_index([-2, -1, 0]) == 2
    assert O.share_index([-1, 0]) == -1
    assert O.share_index([0]) == -1


def test_alpha_gradient_is_derivative_of_objective():
    """Finite differences: d/d(alpha) <out(alpha), od> equals the alpha delta / (5*lr) in softmax mode
    (A.4/A.5), which pins the Jacobian algebra of tdnn.cc:541-560 independently of any restatement."""
    offsets = [0, 1, 2, 3, 4]
    n = len(offsets)
    g, x, W, bp, od, ro = _case(n, 16, 12, 2, 9, offsets, 1, seed=5)
    x64, W64, od64 = x.astype(np.float64), W.astype(np.float64), od.astype(np.float64)

    def objective(alpha):
        c = R.coef(alpha, 0, 1.0)
        w = R.weights(c, 0, 0)
        return float(np.sum(R.propagate(W64, bp[n:].astype(np.float64), x64, od.shape[0], ro, 1, w) * od64))

    a0 = bp[:n].astype(np.float64)
    fd = np.zeros(n)
    for i in range(n):
        e = np.zeros(n)
        e[i] = 1e-5
        fd[i] = (objective(a0 + e) - objective(a0 - e)) / 2e-5
    _, coef = O.tdnn_propagate(offsets, 0, 1.0, W, bp, x, od.shape[0], ro, 1)
    db = np.zeros(n + W.shape[0], np.float32)
    lr = 0.01
    O.tdnn_backprop(offsets, 0, 1.0, W, x, od, coef, ro, 1, lr, dW=np.zeros_like(W), dbias=db)
    np.testing.assert_allclose(db[:n] / (5 * lr), fd, rtol=2e-3, atol=1e-7)


def test_data_gradient_finite_difference():
    offsets = [-2, -1, 0]
    n = 3
    g, x, W, bp, od, ro = _case(n, 10, 8, 2, 7, offsets, 1, seed=9)
    _, coef = O.tdnn_propagate(offsets, 1, 0.5, W, bp, x, od.shape[0], ro, 1, u_gumbel=np.full(n, 0.3, np.float32))
    ind = np.zeros_like(x)
    O.tdnn_backprop(offsets, 1, 0.5, W, x, od, coef, ro, 1, 0.0, in_deriv=ind)
    d = g.standard_normal(x.shape).astype(np.float32)
    eps = 1e-2
    op, _ = O.tdnn_propagate(offsets, 1, 0.5, W, bp, x + eps * d, od.shape[0], ro, 1, u_gumbel=np.full(n, 0.3, np.float32))
    om, _ = O.tdnn_propagate(offsets, 1, 0.5, W, bp, x - eps * d, od.shape[0], ro, 1, u_gumbel=np.full(n, 0.3, np.float32))
    fd = float(np.sum((op.astype(np.float64) - om) * od)) / (2 * eps)
    an = float(np.sum(ind.astype(np.float64) * d))
    assert abs(fd - an) <= 2e-3 * abs(an)


def test_softmax_flops_oracle_vs_numpy():
    g = np.random.default_rng(0)
    x = g.standard_normal((50, 8)).astype(np.float32)
    u = g.uniform(0.1, 0.9, 8).astype(np.float32)
    T = 0.4
    p = O.softmax_flops_fwd(x, u, T)
    z = (x.astype(np.float64) - np.log(-np.log(u.astype(np.float64)))) / T
    pr = np.exp(z - z.max(1, keepdims=True))
    pr /= pr.sum(1, keepdims=True)
    assert rel_err(p, np.maximum(pr, 1e-20)) < 1e-5
    od = g.standard_normal((50, 8)).astype(np.float32)
    ind, od_after = O.softmax_flops_bwd(p, od, 0.3, True, T)
    f = np.array([-25, -50, -80, -100, -120, -160, -200, -240], dtype=np.float64)
    e = od + 0.3 / 50 / 8 * f
    ref = (pr * e - pr * (pr * e).sum(1, keepdims=True)) / T
    assert rel_err(ind, ref) < 1e-5
    assert rel_err(od_after, e) < 1e-6
    # in place: same in_deriv
    ind2, _ = O.softmax_flops_bwd(p, od, 0.3, True, T, in_place=True)
    assert np.array_equal(ind, ind2)
