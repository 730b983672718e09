"""GPU tests of the host glue a C++ trainer calls through the C ABI: UpdateNnetWithMaxChange
(nnet-utils.cc:2085-2175), ApplyL2Regularization (:2223-2245), PenalizeOutOfRange / ComputeChainObjfAndDeriv with the
out-of-range penalty, each against a numpy restatement of the reference lines."""
import ctypes as C

import numpy as np
import pytest

from tests.util import rel_err

pytestmark = pytest.mark.gpu


def _ref_max_change(deltas, groups, max_change, max_param_change, max_change_scale, scale):
    """nnet-utils.cc:2085-2175 in float32 (BaseFloat).  Returns (applied, per-group factor * scale)."""
    f32 = np.float32
    ng = len(max_change)
    dots = np.zeros(ng, np.float64)
    for d, g in zip(deltas, groups):
        dots[g] += (d.astype(np.float64) ** 2).sum()
    factors = np.ones(ng, f32)
    pds = f32(0)
    for g in range(ng):
        dp = f32(dots[g])
        if max_change[g] != 0 and np.sqrt(dp) * abs(f32(scale)) > f32(max_change[g]) * f32(max_change_scale):
            factors[g] = f32(max_change[g]) * f32(max_change_scale) / (np.sqrt(dp) * abs(f32(scale)))
        pds += factors[g] * factors[g] * dp
    pd = np.sqrt(pds) * abs(f32(scale))
    scale = f32(scale)
    if max_param_change != 0 and pd > f32(max_param_change) * f32(max_change_scale):
        if not np.isfinite(pd):
            return False, np.zeros(ng, f32)
        scale = scale * f32(max_param_change) * f32(max_change_scale) / pd
    return True, factors * scale


def _table(ctx, models, deltas, groups, max_change):
    from tdnnf_nas_b200 import capi

    bufs = []
    for m, d, g in zip(models, deltas, groups):
        mp, r, c, ms = capi._mat(m)
        dp, _, _, ds = capi._mat(d)
        bufs.append((mp, ms, dp, ds, r, c, g))
    return capi.ParamTable(ctx, bufs, max_change)


@pytest.mark.parametrize("case", ["none", "per_component", "global", "both", "scaled"])
def test_update_with_max_change(ctx, case):
    import torch

    g = np.random.default_rng(5)
    shapes = [(160, 1536 * 2 + 3), (1, 167), (1536, 160), (1, 1536), (256, 1536)]
    groups = [0, 0, 1, 1, 2]
    mags = dict(none=[1e-4, 1e-4, 1e-4, 1e-4, 1e-4], per_component=[1e-2, 1e-4, 1e-5, 1e-5, 1e-5], **{"global": [2e-3] * 5},
                both=[5e-2, 1e-2, 2e-3, 2e-3, 3e-3], scaled=[3e-3] * 5)[case]
    max_change = [1.5, 1.5, 1.5] if case == "global" else [0.75, 0.75, 0.0 if case == "both" else 1.5]
    scale, mcs = (0.5, 0.7) if case == "scaled" else (1.0, 1.0)
    models = [g.standard_normal(s).astype(np.float32) for s in shapes]
    deltas = [(g.standard_normal(s) * m).astype(np.float32) for s, m in zip(shapes, mags)]
    # strided views (CuMatrix pitch): a column slice of a wider buffer
    md = [torch.from_numpy(np.pad(m, ((0, 0), (0, 5)))).cuda()[:, : m.shape[1]] for m in models]
    dd_full = [torch.from_numpy(np.pad(d, ((0, 0), (0, 3)), constant_values=7.0)).cuda() for d in deltas]
    dd = [f[:, : d.shape[1]] for f, d in zip(dd_full, deltas)]
    tab = _table(ctx, md, dd, groups, max_change)
    applied = tab.update_with_max_change(2.0, mcs, scale, momentum=0.0)
    ok_ref, f_ref = _ref_max_change(deltas, groups, max_change, 2.0, mcs, scale)
    assert applied and ok_ref
    np.testing.assert_allclose(np.array(list(tab.factors)), f_ref, rtol=2e-6)
    if case in ("per_component", "both"):
        assert tab.num_per_component[0] == 1
    if case in ("global", "both"):
        assert tab.num_global.value == 1
    for m, d, mdv, ddv, full, gi in zip(models, deltas, md, dd, dd_full, groups):
        assert rel_err(mdv.cpu().numpy(), m + f_ref[gi] * d) < 1e-6
        assert torch.all(ddv == 0)
        assert torch.all(full[:, d.shape[1]:] == 7.0)  # pitch padding untouched


def test_update_with_max_change_infinite_delta_leaves_model(ctx):
    """an infinite delta must not poison the model (0 * inf = NaN): the reference returns false and the
    trainer then zeroes the delta with ScaleNnet(0.0)."""
    import torch

    g = np.random.default_rng(1)
    model = torch.from_numpy(g.standard_normal((40, 33)).astype(np.float32)).cuda()
    model2 = torch.from_numpy(g.standard_normal((1, 40)).astype(np.float32)).cuda()
    before, before2 = model.clone(), model2.clone()
    delta = torch.from_numpy(g.standard_normal((40, 33)).astype(np.float32)).cuda()
    delta[3, 4] = float("inf")
    delta2 = torch.ones((1, 40), device="cuda")
    tab = _table(ctx, [model, model2], [delta, delta2], [0, 1], [0.75, 0.75])
    assert tab.update_with_max_change(2.0) is False
    assert torch.equal(model, before) and torch.equal(model2, before2)
    assert torch.all(delta == 0) and torch.all(delta2 == 0)
    assert tab.num_global.value == 0
    # the same without a per-component max-change: the sum is +inf, the case the reference itself refuses (:2147-2150)
    delta = torch.from_numpy(g.standard_normal((40, 33)).astype(np.float32)).cuda()
    delta[0, 0] = float("inf")
    delta2 = torch.ones((1, 40), device="cuda")
    tab = _table(ctx, [model, model2], [delta, delta2], [0, 1], [0.0, 0.0])
    assert tab.update_with_max_change(2.0) is False
    assert torch.equal(model, before) and torch.equal(model2, before2) and torch.all(delta == 0)
    # momentum: delta *= momentum after a normal step
    delta = torch.full((40, 33), 1e-3, device="cuda")
    delta2 = torch.full((1, 40), 1e-3, device="cuda")
    tab = _table(ctx, [model, model2], [delta, delta2], [0, 1], [0.75, 0.75])
    assert tab.update_with_max_change(2.0, momentum=0.5)
    assert torch.allclose(delta, torch.full_like(delta, 5e-4)) and torch.allclose(model, before + 1e-3)


def test_apply_l2_regularization(ctx):
    import torch

    g = np.random.default_rng(2)
    shapes, groups = [(30, 50), (1, 30), (20, 30)], [0, 0, 1]
    models = [g.standard_normal(s).astype(np.float32) for s in shapes]
    deltas = [g.standard_normal(s).astype(np.float32) for s in shapes]
    md, dd = [torch.from_numpy(m).cuda() for m in models], [torch.from_numpy(d).cuda() for d in deltas]
    tab = _table(ctx, md, dd, groups, [0.75, 0.75])
    lrate, l2 = [2.5e-4, 1e-3], [0.01, 0.0]
    tab.apply_l2_regularization(lrate, l2, 64.0)
    for m, d, mdv, ddv, gi in zip(models, deltas, md, dd, groups):
        sc = np.float32(-2.0 * 64.0 * lrate[gi] * l2[gi])
        assert rel_err(ddv.cpu().numpy(), d + sc * m) < 1e-6
        assert torch.equal(mdv, torch.from_numpy(m).cuda())


@pytest.mark.parametrize("step,offset", [(1, 0), (4, 0), (4, 3), (5, 2)])
def test_penalize_out_of_range(ctx, step, offset):
    import torch

    g = np.random.default_rng(step + offset)
    rows, cols = 131, 77
    x = (g.standard_normal((rows, cols)) * 25.0).astype(np.float32)
    d0 = g.standard_normal((rows, cols)).astype(np.float32)
    xd = torch.from_numpy(np.pad(x, ((0, 0), (0, 4)))).cuda()[:, :cols]
    dd = torch.from_numpy(d0.copy()).cuda()
    ctx.penalize_out_of_range(xd, dd, 30.0, 0.02 * step, step, offset)
    ref = d0.copy()
    sub = x[offset::step]
    ref[offset::step] -= np.float32(0.02 * step) * (np.where(sub > 30, sub - 30, 0) + np.where(sub < -30, sub + 30, 0))
    assert (np.abs(x) > 30).sum() > 100
    np.testing.assert_allclose(dd.cpu().numpy(), ref, rtol=1e-6, atol=1e-6)


def test_chain_objf_out_of_range_penalty(ctx):
    """ComputeChainObjfAndDeriv with outputs beyond +-30: the objective uses the clamped exp (as the denominator does),
    the derivative gains -2 * reg * step * (x -+ 30) on the sub-sampled rows."""
    import torch

    from oracle import oracle as O
    from tdnnf_nas_b200 import capi, chain, synth

    S, P, T, N = 8, 40, 9, 120
    dgraph = synth.make_den_graph(N, P, 5.0, seed=4)
    ngraph = synth.make_num_graphs(S, P, T, seed=8)
    g = np.random.default_rng(3)
    x = (g.standard_normal((T * S, P)) * 12.0).astype(np.float32)
    x[5, 7], x[12, 3], x[13, 3] = 34.0, -33.0, 31.0
    den_lp, den_d, _ = O.den_forward_backward(dgraph, x, S, T, 0.1, deriv_weight=-1.0)
    num_lp, num_d, _ = O.num_forward_backward(ngraph, x, T, deriv_weight=1.0)
    dg, ng = capi.DenGraph(ctx, dgraph), capi.NumeratorGraph(ctx, ngraph)
    opts = chain.ChainTrainingOptions(out_of_range_regularize=0.01, oor_row_step=4)
    obj = chain.ChainObjective(ctx, dg, ng, S, T, opts)
    xd = torch.from_numpy(x).cuda()
    deriv = torch.zeros_like(xd)
    objf, l2, weight = obj.compute(xd, deriv, oor_row_offset=1)
    assert objf == pytest.approx(num_lp - den_lp, rel=1e-4) and l2 == 0.0
    ref = num_d + den_d
    sub = x[1::4]
    ref[1::4] -= np.float32(2 * 0.01 * 4) * (np.where(sub > 30, sub - 30, 0) + np.where(sub < -30, sub + 30, 0))
    assert abs(ref[5, 7] - (num_d + den_d)[5, 7]) > 0.1 and ref[12, 3] == (num_d + den_d)[12, 3]  # row 12 not sampled, 5 and 13 are
    assert rel_err(deriv.cpu().numpy(), ref) < 1e-3
    obj.close(); ng.close(); dg.close()


def test_print_log_alpha_prints_the_models_alpha(ctx, capfd):
    """tdnn.cc:571 prints bias_params_temp_ (the MODEL's log-alpha) every minibatch; scripts grep it from the logs.
    The delta component's alpha (x lr x 10000) must not be what is printed."""
    import torch

    from tdnnf_nas_b200 import nnet3

    nnet3.set_context(ctx)
    nnet3.set_rand_seed(9)
    offsets, din, dout, S, t_out = [0, 1, 2], 32, 24, 4, 10
    n = len(offsets)
    comp = nnet3.Component.new(
        "TdnnDARTSV3Component",
        f"input-dim={din} output-dim={dout} time-offsets=0,1,2 use-gumbel=false use-entropy=false free-select=false "
        "update-alpha=true update-theta=true uniform-sample=false learning-rate=0.01")
    v = comp.vectorize()
    alpha = np.array([0.25, -1.5, 0.75], np.float32)
    v[dout * n * din: dout * n * din + n] = alpha
    comp.unvectorize(v)
    inp = [(s, t, 0) for t in range(t_out + 2) for s in range(S)]
    outp = [(s, t, 0) for t in range(t_out) for s in range(S)]
    idx = comp.precompute_indexes(inp, outp)
    g = np.random.default_rng(0)
    x = torch.from_numpy(g.standard_normal((len(inp), din)).astype(np.float32)).cuda()
    od = torch.from_numpy(g.standard_normal((len(outp), dout)).astype(np.float32)).cuda()
    out = torch.zeros((len(outp), dout), device="cuda")
    delta = comp.copy()
    delta.scale(0.0)
    memo = comp.propagate(idx, x, out)
    nnet3.set_print_log_alpha(True)
    try:
        capfd.readouterr()
        comp.backprop(idx, x, None, od, memo, delta, None)
        torch.cuda.synchronize()
        printed = capfd.readouterr().out
    finally:
        nnet3.set_print_log_alpha(False)
        comp.delete_memo(memo)
    assert "log_alpha" in printed
    vals = [float(t) for t in printed.split("[")[1].split("]")[0].split()]
    np.testing.assert_allclose(vals, alpha, rtol=1e-5)
    d_alpha = delta.vectorize()[dout * n * din: dout * n * din + n]
    assert np.abs(d_alpha).max() > 0 and not np.allclose(d_alpha, alpha, rtol=1e-2)
