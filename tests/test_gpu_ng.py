"""GPU parity of OnlineNaturalGradient (csrc/nnet3/natural_gradient.cc: implicit operands, tcgen05 GEMMs, lazy
host eigen-update) against the oracle (oracle/oracle_ng.inc: materialised, CPU), over enough calls to pass the
initialisation, the ten initial updates and the periodic ones."""
import numpy as np
import pytest

from tests.util import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture()
def nn(ctx):
    from tdnnf_nas_b200 import nnet3

    nnet3.set_context(ctx)
    nnet3.set_rand_seed(4242)
    nnet3.set_ng_identity(False)
    return nnet3


def _data(g, N, D, k=6):
    basis = g.standard_normal((k, D))
    return ((g.standard_normal((N, k)) * np.linspace(3.0, 1.0, k)) @ basis + 0.3 * g.standard_normal((N, D))).astype(np.float32)


@pytest.mark.parametrize("N,D,rank,period", [(512, 160, 80, 4), (300, 41, 8, 1), (1000, 1537, 20, 4), (64, 33, 10, 4),
                                             (700, 9, 40, 1)])
def test_precondition_directions_vs_oracle(nn, N, D, rank, period):
    import torch

    from oracle import oracle as O

    g = np.random.default_rng(N + D)
    ng = nn.NaturalGradient(rank, period, 2000.0, 4.0)
    orc = O.NaturalGradient(rank, period, 2000.0, 4.0)
    for step in range(15):
        X = _data(g, N, D)
        Xo = X.copy()
        s_o = orc.precondition(Xo)
        Xd = torch.from_numpy(X).cuda()
        s_g = ng.precondition(Xd)
        e = rel_err(Xd.cpu().numpy(), Xo)
        assert e < 1e-3, (step, e)
        assert abs(s_g - s_o) / s_o < 1e-3, (step, s_g, s_o)
    sg, so = ng.state(), orc.state()
    assert sg["t"] == so["t"] == 15 and sg["rank"] == so["rank"] == min(rank, D - 1) and sg["D"] == D
    # the internal state is a fixed-point iteration of fp32 quantities (eigenvalues spanning 1e2..1e3): it is
    # allowed to drift a little more than the observable outputs (X_hat and scale above, 1e-3)
    assert abs(sg["rho"] - so["rho"]) / so["rho"] < 2e-2
    assert rel_err(np.sort(sg["d"]), np.sort(so["d"])) < 2e-2
    # rows of W_t are defined up to sign: compare the projector
    assert rel_err(sg["W"].T.astype(np.float64) @ sg["W"], so["W"].T.astype(np.float64) @ so["W"]) < 2e-2


def test_frozen_and_dim_one(nn):
    import torch

    g = np.random.default_rng(5)
    ng = nn.NaturalGradient(5, 1, 2000.0, 4.0)
    one = torch.ones((10, 1), device="cuda")
    assert ng.precondition(one) == 1.0 and torch.all(one == 1) and ng.state()["t"] == 0
    ng = nn.NaturalGradient(5, 1, 2000.0, 4.0)
    for _ in range(3):
        ng.precondition(torch.from_numpy(_data(g, 100, 20)).cuda())
    ng.freeze(True)
    W0 = ng.state()["W"].copy()
    ng.precondition(torch.from_numpy(_data(g, 100, 20)).cuda())
    assert np.array_equal(ng.state()["W"], W0)


@pytest.mark.parametrize("mode_cfg,flags", [
    ("use-gumbel=true use-entropy=false free-select=false update-alpha=true update-theta=false uniform-sample=false Temp-Proportion=0.5", 1 | 16),
    ("use-gumbel=false use-entropy=false free-select=false update-alpha=false update-theta=true uniform-sample=true", 4),
])
@pytest.mark.parametrize("din,dout,offsets", [(192, 40, list(range(-6, 1))), (40, 192, list(range(0, 7)))])
def test_darts_natural_gradient_update_over_steps(nn, mode_cfg, flags, din, dout, offsets):
    """TdnnDARTSV3Component::UpdateNaturalGradient (tdnn.cc:457-626) over 13 minibatches: the delta of every step
    (theta, bias tail, alpha) against the oracle that materialises in_value_temp and preconditions it."""
    import torch

    from oracle import oracle as O

    n, S = len(offsets), 16
    cfg = (f"input-dim={din} output-dim={dout} time-offsets={','.join(map(str, offsets))} learning-rate=0.02 {mode_cfg} "
           "rank-in=20 rank-out=30")
    comp = nn.Component.new("TdnnDARTSV3Component", cfg)
    g = np.random.default_rng(din)
    v = comp.vectorize()
    v[dout * n * din: dout * n * din + n] = g.standard_normal(n).astype(np.float32)
    comp.unvectorize(v)
    W, bias_params = v[: dout * n * din].reshape(dout, n * din).copy(), v[dout * n * din:].copy()
    t_out = list(range(0, 24))
    t_in = list(range(min(offsets), t_out[-1] + max(offsets) + 1))
    inp = [(s, t, 0) for t in t_in for s in range(S)]
    outp = [(s, t, 0) for t in t_out for s in range(S)]
    idx = comp.precompute_indexes(inp, outp)
    row_stride, row_offsets = idx.row_stride_and_offsets()
    delta = comp.copy()
    ng_in, ng_out = O.NaturalGradient(20, 4, 2000.0, 4.0), O.NaturalGradient(30, 4, 2000.0, 4.0)
    temp = comp.temp_proportion()
    worst = 0.0
    for step in range(13):
        x = _data(g, len(inp), din, k=4)
        od = (_data(g, len(outp), dout, k=3) / len(outp)).astype(np.float32)
        xd, odd = torch.from_numpy(x).cuda(), torch.from_numpy(od).cuda()
        out = torch.zeros((len(outp), dout), device="cuda")
        c0 = nn.get_rand_counter()
        memo = comp.propagate(idx, xd, out)
        nn.set_rand_counter(c0)
        u_g = [nn.rand_uniform() for _ in range(n)] if flags & 1 else None
        u_u = nn.rand_uniform() if flags & 4 else 0.0
        _, coef_ref = O.tdnn_propagate(offsets, flags, temp, W, bias_params, x, len(outp), row_offsets, row_stride, u_g, u_u)
        delta.scale(0.0)
        comp.backprop(idx, xd, None, odd, memo, delta, None)
        comp.delete_memo(memo)
        dW_ref = np.zeros_like(W)
        db_ref = np.zeros(n + dout, np.float32)
        scales = np.zeros(2, np.float32)
        O.tdnn_backprop(offsets, flags, temp, W, x, od, coef_ref, row_offsets, row_stride, delta.learning_rate(),
                        dW=dW_ref, dbias=db_ref, ng_in=ng_in, ng_out=ng_out, scales=scales)
        dv = delta.vectorize()
        dW, db = dv[: dout * n * din].reshape(dout, n * din), dv[dout * n * din:]
        e_w, e_b = rel_err(dW, dW_ref), rel_err(db[n:], db_ref[n:])
        worst = max(worst, e_w, e_b)
        assert e_w < 1e-3 and e_b < 1e-3, (step, e_w, e_b, scales)
        sc = np.abs(db_ref[:n]).max()
        if sc > 0:
            assert np.abs(db[:n] - db_ref[:n]).max() / sc < 5e-3
    print(f"worst delta rel err over 13 steps: {worst:.2e}")
    # the preconditioners' own state stayed in step with the oracle's
    for which, orc in ((0, ng_in), (1, ng_out)):
        sg, so = delta.preconditioner(which).state(), orc.state()
        assert sg["t"] == so["t"] == 13
        assert abs(sg["rho"] - so["rho"]) / so["rho"] < 1e-2
        assert rel_err(sg["W"].T.astype(np.float64) @ sg["W"], so["W"].T.astype(np.float64) @ so["W"]) < 1e-2


def test_onehot_natural_gradient(nn):
    """OnehotFunctionComponent::Backprop with use-natural-gradient=true (simple.cc:9536-9546): the preconditioned
    row sum, against the oracle's preconditioner applied to a copy of out_deriv."""
    import torch

    from oracle import oracle as O

    g = np.random.default_rng(11)
    R = 256
    oh = nn.Component.new("OnehotFunctionComponent", "input-dim=220 output-dim=8 is-updatable=true use-natural-gradient=true learning-rate=0.1")
    d = oh.copy()
    orc = O.NaturalGradient(40, 1, 2000.0, 4.0)  # the class defaults (rank 40 -> clipped to 7)
    for step in range(5):
        od = _data(g, R, 8, k=2)
        d.scale(0.0)
        oh.backprop(None, None, None, torch.from_numpy(od).cuda(), None, d, None)
        c = od.copy()
        scale = orc.precondition(c)
        ref = 0.1 * scale * c.astype(np.float64).sum(0)
        assert rel_err(d.vectorize(), ref) < 1e-3, step
