"""CPU tests of the oracle's stock-TdnnComponent and ConstrainOrthonormal restatements (BASELINE configs[1]):
an independent float64 numpy restatement written from the equations, plus the properties the update is designed to
have (Povey et al. 2018, sections 2.2-2.3: quadratic convergence to a semi-orthogonal matrix, and with the floating
scale an update orthogonal to M)."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import np_ref as R
from tests.util import rel_err


def _case(n, din, dout, S, t_out, offsets, row_stride, seed):
    from tdnnf_nas_b200 import synth

    g = np.random.default_rng(seed)
    t0 = min(offsets)
    n_t_in = (t_out - 1) * row_stride + max(offsets) - t0 + 1
    n_t_in = row_stride * ((n_t_in + row_stride - 1) // row_stride)
    _, ro = synth.regular_row_offsets(offsets, t0, 0, S, 1, row_stride)
    x = g.standard_normal((n_t_in * S, din)).astype(np.float32)
    W = (g.standard_normal((dout, n * din)) / np.sqrt(n * din)).astype(np.float32)
    b = g.standard_normal(dout).astype(np.float32)
    od = (g.standard_normal((t_out * S, dout)) / (t_out * S)).astype(np.float32)
    return g, x, W, b, od, ro


@pytest.mark.parametrize("use_bias", [True, False])
@pytest.mark.parametrize("offsets,row_stride", [([-3, 0], 1), ([0, 3], 1), ([0], 1), ([0, 3], 3), ([-1, 0, 1], 1)])
def test_plain_tdnn_matches_numpy(offsets, row_stride, use_bias):
    n = len(offsets)
    g, x, W, b, od, ro = _case(n, 24, 20, 3, 11, offsets, row_stride, seed=n * 7 + row_stride)
    out_rows = od.shape[0]
    ones = np.ones(n)
    pre = g.standard_normal((out_rows, W.shape[0])).astype(np.float32)
    out = O.plain_tdnn_propagate(W, b if use_bias else None, x, out_rows, ro, row_stride, out=pre.copy())
    ref = R.propagate(W, b if use_bias else None, x, out_rows, ro, row_stride, ones)
    if not use_bias:
        ref = ref + pre  # kPropagateAdds
    assert rel_err(out, ref) < 2e-6

    lr = 0.05
    ind = g.standard_normal(x.shape).astype(np.float32)
    ind0 = ind.copy()
    dW = np.zeros_like(W)
    db = np.zeros(W.shape[0], np.float32) if use_bias else None
    O.plain_tdnn_backprop(W, x, od, ro, row_stride, lr, in_deriv=ind, dW=dW, dbias=db, natural_gradient=False)
    # float64: in_deriv_i += out_deriv W_i ; dW_i = lr out_deriv^T X_i ; dbias = lr colsum(out_deriv)
    din = W.shape[1] // n
    ind_r = ind0.astype(np.float64)
    dW_r = np.zeros(W.shape)
    for i, o in enumerate(ro):
        sl = slice(o, o + out_rows * row_stride, row_stride)
        ind_r[sl][:out_rows] += od.astype(np.float64) @ W[:, i * din:(i + 1) * din].astype(np.float64)
        dW_r[:, i * din:(i + 1) * din] = lr * od.astype(np.float64).T @ x[sl][:out_rows].astype(np.float64)
    assert rel_err(ind, ind_r) < 2e-6
    assert rel_err(dW, dW_r) < 2e-6
    if use_bias:
        assert rel_err(db, lr * od.astype(np.float64).sum(0)) < 2e-6

    # natural-gradient path with both preconditioners = identity is the simple update
    dW2 = np.zeros_like(W)
    db2 = np.zeros(W.shape[0], np.float32) if use_bias else None
    O.plain_tdnn_backprop(W, x, od, ro, row_stride, lr, dW=dW2, dbias=db2, natural_gradient=True)
    assert rel_err(dW2, dW_r) < 2e-6
    if use_bias:
        assert rel_err(db2, db) < 2e-6


@pytest.mark.parametrize("use_bias", [True, False])
def test_plain_tdnn_natural_gradient_matches_float64(use_bias):
    """UpdateNaturalGradient of the stock TdnnComponent against a float64 restatement from the equations: the input
    operand is [X_1 | ... | X_n] with the column of ones appended ONLY when the component has a bias
    (tdnn.cc:477-478, bias_params_.Dim() != 0; the `linear` halves of the TDNN-F layers have none)."""
    offsets, row_stride, n, din, dout, S, t_out, lr = [-3, 0], 1, 2, 24, 18, 6, 20, 0.05
    g, x, W, b, od, ro = _case(n, din, dout, S, t_out, offsets, row_stride, seed=11)
    out_rows = od.shape[0]
    rank_in, rank_out = 7, 5
    ng_in, ng_out = O.NaturalGradient(rank_in, 4, 2000.0, 4.0), O.NaturalGradient(rank_out, 4, 2000.0, 4.0)
    ref_in, ref_out = R.NaturalGradientF64(rank_in, 4, 2000.0, 4.0), R.NaturalGradientF64(rank_out, 4, 2000.0, 4.0)
    dW = np.zeros_like(W)
    db = np.zeros(dout, np.float32) if use_bias else None
    dW_r, db_r = np.zeros(W.shape), np.zeros(dout)
    for step in range(3):
        xs = (x + 0.1 * step * g.standard_normal(x.shape)).astype(np.float32)
        ods = (od * (1.0 + 0.2 * step)).astype(np.float32)
        O.plain_tdnn_backprop(W, xs, ods, ro, row_stride, lr, dW=dW, dbias=db, natural_gradient=True, ng_in=ng_in, ng_out=ng_out)
        cols = [xs[o: o + out_rows * row_stride: row_stride][:out_rows].astype(np.float64) for o in ro]
        if use_bias:
            cols.append(np.ones((out_rows, 1)))
        X = np.concatenate(cols, axis=1)
        assert X.shape[1] == n * din + (1 if use_bias else 0)
        Xh, s_in = ref_in.precondition(X)
        Dh, s_out = ref_out.precondition(ods.astype(np.float64))
        G = lr * s_in * s_out * Dh.T @ Xh
        dW_r += G[:, : n * din]
        if use_bias:
            db_r += G[:, n * din]
        assert rel_err(dW, dW_r) < 5e-4, (step, rel_err(dW, dW_r))
        if use_bias:
            assert rel_err(db, db_r) < 5e-4
    assert ng_in.state()["D"] == n * din + (1 if use_bias else 0)


def test_plain_tdnn_is_darts_with_unit_weights():
    """TdnnDARTSV3 in uniform-sample mode with the sampled slot == the shared slot is a single-offset TdnnComponent; with
    free-select and sigmoid(alpha) -> 1 it is the all-offsets TdnnComponent (the two classes share every GEMM)."""
    offsets = [0, 1, 2]
    n = len(offsets)
    g, x, W, b, od, ro = _case(n, 16, 12, 2, 9, offsets, 1, seed=3)
    out_rows = od.shape[0]
    bp = np.concatenate([np.full(n, 40.0, np.float32), b])  # sigmoid(40) == 1 in fp32
    darts, _ = O.tdnn_propagate(offsets, 2, 1.0, W, bp, x, out_rows, ro, 1)
    plain = O.plain_tdnn_propagate(W, b, x, out_rows, ro, 1)
    np.testing.assert_allclose(darts, plain, rtol=0, atol=0)


def _np_constrain(M, scale):
    M = M.astype(np.float64)
    P = M @ M.T
    speed = 0.125
    if scale < 0:
        tr, tr2 = np.trace(P), (P * P).sum()
        scale = np.sqrt(tr2 / tr)
        ratio = tr2 * P.shape[0] / tr ** 2
        if ratio > 1.02:
            speed *= 0.5
            if ratio > 1.1:
                speed *= 0.5
    Q = P - scale ** 2 * np.eye(P.shape[0])
    return M - 4.0 * (speed / scale ** 2) * Q @ M, scale


@pytest.mark.parametrize("scale", [-1.0, 1.0, 0.5])
@pytest.mark.parametrize("rows,cols,spread", [(16, 48, 0.05), (24, 24, 0.3), (160, 320, 1.0)])
def test_constrain_orthonormal_matches_numpy(rows, cols, spread, scale):
    g = np.random.default_rng(rows + cols)
    q, _ = np.linalg.qr(g.standard_normal((cols, rows)))
    target = abs(scale) if scale > 0 else 0.8
    M = (target * q.T + spread / np.sqrt(cols) * g.standard_normal((rows, cols))).astype(np.float32)
    ref, s_used = _np_constrain(M, scale)
    out = M.copy()
    info = O.constrain_orthonormal(out, scale)
    assert rel_err(out, ref) < 3e-6
    assert abs(info[0] - s_used) < 1e-5 * s_used


def test_constrain_orthonormal_properties():
    g = np.random.default_rng(0)
    rows, cols = 20, 60
    q, _ = np.linalg.qr(g.standard_normal((cols, rows)))
    M = (q.T + 0.02 * g.standard_normal((rows, cols))).astype(np.float32)
    # floating scale: the update is orthogonal to M, tr(M X^T) = 0
    out = M.copy()
    O.constrain_orthonormal(out, -1.0)
    X = out.astype(np.float64) - M
    assert abs((M * X).sum()) < 1e-5 * np.linalg.norm(M) * np.linalg.norm(X)
    # fixed scale 1: quadratic convergence of ||M M^T - I||_F
    errs = []
    cur = M.copy()
    for _ in range(4):
        errs.append(np.linalg.norm(cur.astype(np.float64) @ cur.T - np.eye(rows)))
        O.constrain_orthonormal(cur, 1.0)
    assert errs[1] < 0.2 * errs[0] and errs[2] < 0.05 * errs[1]
    assert np.linalg.norm(cur.astype(np.float64) @ cur.T - np.eye(rows)) < 1e-5
    with pytest.raises(RuntimeError):
        O.constrain_orthonormal(M.copy(), 0.0)
