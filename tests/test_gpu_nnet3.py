"""GPU tests of the nnet3 component mirror, driven like NnetComputer would (config line / model
stream -> PrecomputeIndexes -> Propagate -> Backprop(to_update)) and checked against the oracle."""
import numpy as np
import pytest

from tests.util import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture()
def nn(ctx):
    from tdnnf_nas_b200 import nnet3

    nnet3.set_context(ctx)
    nnet3.set_rand_seed(777)
    nnet3.set_ng_identity(False)
    yield nnet3
    nnet3.set_ng_identity(False)


def _grid(S, t_in, t_out):
    return [(n, t, 0) for t in t_in for n in range(S)], [(n, t, 0) for t in t_out for n in range(S)]


def _params(comp, n, din, dout):
    v = comp.vectorize()
    W = v[: dout * n * din].reshape(dout, n * din).copy()
    return W, v[dout * n * din:].copy()


@pytest.mark.parametrize("mode_cfg,flags", [
    ("use-gumbel=false use-entropy=false free-select=false update-alpha=false update-theta=true uniform-sample=true", 4),
    ("use-gumbel=true use-entropy=false free-select=false update-alpha=true update-theta=false uniform-sample=false Temp-Proportion=0.5", 1 | 16),
    ("use-gumbel=false use-entropy=false free-select=false update-alpha=true update-theta=true uniform-sample=false", 16),
    ("use-gumbel=false use-entropy=true free-select=true update-alpha=false update-theta=true uniform-sample=false", 2 | 8),
])
@pytest.mark.parametrize("offsets,subsample", [(list(range(0, 7)), 1), (list(range(-6, 1)), 1), (list(range(0, 7)), 3)])
@pytest.mark.parametrize("natural_gradient", [False, True], ids=["raw", "ng"])
def test_tdnn_darts_component_vs_oracle(nn, mode_cfg, flags, offsets, subsample, natural_gradient):
    import torch

    from oracle import oracle as O

    # "raw": both preconditioners forced to the identity (the un-preconditioned gradient G and s_i);
    # "ng": OnlineNaturalGradient as in the reference (rank-in 20; rank-out min(80, (D_out+1)/2) = 24, tdnn.cc:196-201)
    nn.set_ng_identity(not natural_gradient)
    n, din, dout, S = len(offsets), 64, 48, 8
    cfg = f"input-dim={din} output-dim={dout} time-offsets={','.join(map(str, offsets))} learning-rate=0.02 {mode_cfg}"
    comp = nn.Component.new("TdnnDARTSV3Component", cfg)
    assert comp.type() == "TdnnDARTSV3Component" and comp.input_dim() == din and comp.output_dim() == dout
    assert comp.properties() == (nn.kUpdatableComponent | nn.kReordersIndexes | nn.kBackpropAdds |
                                 nn.kBackpropNeedsInput | nn.kUsesMemo)
    # give alpha non-trivial values (InitFromConfig zeroes them, tdnn.cc:176)
    v = comp.vectorize()
    g = np.random.default_rng(len(mode_cfg) + subsample)
    v[dout * n * din: dout * n * din + n] = g.standard_normal(n).astype(np.float32)
    comp.unvectorize(v)
    W, bias_params = _params(comp, n, din, dout)

    t_out = list(range(0, 30, subsample))
    t_in = list(range(min(offsets), t_out[-1] + max(offsets) + 1))
    inp, outp = comp.reorder_indexes(*_grid(S, t_in, t_out))
    idx = comp.precompute_indexes(inp, outp)
    row_stride, row_offsets = idx.row_stride_and_offsets()
    assert row_stride == subsample
    in_rows, out_rows = len(inp), len(outp)

    x = g.standard_normal((in_rows, din)).astype(np.float32)
    od = (g.standard_normal((out_rows, dout)) / out_rows).astype(np.float32)
    xd, odd = torch.from_numpy(x).cuda(), torch.from_numpy(od).cuda()
    out = torch.full((out_rows, dout), 3.0, device="cuda")

    c0 = nn.get_rand_counter()
    memo = comp.propagate(idx, xd, out)
    c1 = nn.get_rand_counter()
    # replay the component's draws: n Gumbel uniforms (if use-gumbel) then one uniform (if uniform-sample)
    nn.set_rand_counter(c0)
    u_g = [nn.rand_uniform() for _ in range(n)] if flags & 1 else None
    u_u = nn.rand_uniform() if flags & 4 else 0.0
    assert nn.get_rand_counter() == c1
    temp = comp.temp_proportion()
    out_ref, coef_ref = O.tdnn_propagate(offsets, flags, temp, W, bias_params, x, out_rows, row_offsets, row_stride,
                                         u_g, u_u)
    assert rel_err(out.cpu().numpy(), out_ref) < 1e-4

    delta = comp.copy()
    delta.scale(0.0)
    in_deriv = torch.zeros((in_rows, din), device="cuda")
    comp.backprop(idx, xd, None, odd, memo, delta, in_deriv)
    comp.delete_memo(memo)
    ind_ref = np.zeros_like(x)
    dW_ref = np.zeros_like(W)
    db_ref = np.zeros(n + dout, np.float32)
    ng_in = O.NaturalGradient(20, 4, 2000.0, 4.0) if natural_gradient else None
    ng_out = O.NaturalGradient(24, 4, 2000.0, 4.0) if natural_gradient else None
    O.tdnn_backprop(offsets, flags, temp, W, x, od, coef_ref, row_offsets, row_stride, delta.learning_rate(),
                    in_deriv=ind_ref, dW=dW_ref, dbias=db_ref, ng_in=ng_in, ng_out=ng_out)
    dW, db = _params(delta, n, din, dout)
    assert rel_err(in_deriv.cpu().numpy(), ind_ref) < 1e-3
    assert rel_err(dW, dW_ref) < 1e-3
    assert rel_err(db[n:], db_ref[n:]) < 1e-3
    scale = np.abs(db_ref[:n]).max()
    if scale > 0:
        assert np.abs(db[:n] - db_ref[:n]).max() / scale < 5e-3
    else:
        assert np.all(db[:n] == 0)
    # model.Add(1.0, delta): the parameter step; DotProduct sees alpha too (quirk Q5)
    before = comp.vectorize()
    comp.add(1.0, delta)
    np.testing.assert_allclose(comp.vectorize(), before + delta.vectorize(), rtol=1e-6, atol=1e-7)
    dv = delta.vectorize().astype(np.float64)
    assert delta.dot_product(delta) == pytest.approx(float(dv @ dv), rel=1e-4)


def test_tdnn_darts_io_roundtrip_and_sed_surgery(nn):
    cfg = ("input-dim=20 output-dim=12 time-offsets=-2,-1,0 use-gumbel=false use-entropy=false free-select=false "
           "update-alpha=false update-theta=true uniform-sample=true learning-rate-factor=0.5 max-change=0.75 "
           "l2-regularize=0.01 orthonormal-constraint=-1.0")
    comp = nn.Component.new("TdnnDARTSV3Component", cfg)
    txt = comp.write(False)
    toks = txt.decode().split()
    # token order of tdnn.cc:659-700 / itf.cc:390-414
    order = ["<TdnnDARTSV3Component>", "<LearningRateFactor>", "<MaxChange>", "<L2Regularize>", "<LearningRate>",
             "<use-gumbel>", "<use-entropy>", "<free-select>", "<update-alpha>", "<update-theta>", "<uniform-sample>",
             "<Temp-Proportion>", "<TimeOffsets>", "<LinearParams>", "<BiasParams>", "<OrthonormalConstraint>",
             "<UseNaturalGradient>", "<NumSamplesHistory>", "<AlphaInOut>", "<RankInOut>", "</TdnnDARTSV3Component>"]
    pos = [toks.index(t) for t in order]
    assert pos == sorted(pos)
    assert b"<use-gumbel> F <use-entropy> F <free-select> F <update-alpha> F <update-theta> T <uniform-sample> T" in txt
    # binary round trip is exact; text round trip is exact after one pass (6 significant digits like Kaldi)
    binary = comp.write(True)
    back = nn.Component.read(binary, True)
    assert back.write(True) == binary
    np.testing.assert_array_equal(back.vectorize(), comp.vectorize())
    t2 = nn.Component.read(txt, False)
    assert t2.write(False) == txt
    # the recipes' sed surgery on the text model (run_TDNN_DARTSV3_fbk_stride_cvupdate.sh:129-134)
    edited = txt.replace(b"<use-gumbel> F", b"<use-gumbel> T").replace(b"<update-alpha> F", b"<update-alpha> T") \
                .replace(b"<update-theta> T", b"<update-theta> F").replace(b"<uniform-sample> T", b"<uniform-sample> F")
    s = nn.Component.read(edited, False)
    assert b"<use-gumbel> T" in s.write(False) and b"<uniform-sample> F" in s.write(False)
    # legacy <Alpha> token (tdnn.cc:733-746)
    legacy = txt.replace(b"<AlphaInOut> 4 4", b"<Alpha> 4")
    assert b"<AlphaInOut> 4 4" in nn.Component.read(legacy, False).write(False)
    assert "rank-in=20" in comp.info() and "time-offsets=-2,-1,0" in comp.info()


def test_set_temperature_proportion_directive(nn):
    a = nn.Component.new("TdnnDARTSV3Component", "input-dim=8 output-dim=8 time-offsets=0,1")
    b = nn.Component.new("GumbelSoftmaxFlopsComponent", "dim=8 scale=0.1 temp-proportion=1.0")
    c = nn.Component.new("CopyNComponent", "input-dim=1 output-dim=25")
    comps = [("tdnnf2.linear", a), ("tdnnf2.softmax", b), ("tdnnf2.copyn", c)]
    nn.apply_edits(nn.temperature_edit_string(1, 4), comps)
    assert a.temp_proportion() == pytest.approx(0.7575) and b.temp_proportion() == pytest.approx(0.7575)
    nn.apply_edits("set-temperature-proportion name=*.softmax proportion=0.25", comps)
    assert a.temp_proportion() == pytest.approx(0.7575) and b.temp_proportion() == pytest.approx(0.25)
    assert b"<TempProportion> 0.25" in b.write(False)


def test_undefined_reference_behaviour_is_an_error(nn):
    import torch

    # quirk Q1: time_offsets[1] == 0 leaves share_offset_index uninitialised in the reference
    comp = nn.Component.new("TdnnDARTSV3Component", "input-dim=8 output-dim=8 time-offsets=-1,0")
    inp, outp = _grid(2, range(-1, 4), range(0, 4))
    idx = comp.precompute_indexes(inp, outp)
    with pytest.raises(nn.Nnet3Error, match="share_offset_index"):
        comp.propagate(idx, torch.zeros((len(inp), 8), device="cuda"), torch.zeros((len(outp), 8), device="cuda"))
    # quirk Q3: use-bias=false cannot work
    nb = nn.Component.new("TdnnDARTSV3Component", "input-dim=8 output-dim=8 time-offsets=0,1 use-bias=false")
    assert nb.properties() & nn.kPropagateAdds
    idx = nb.precompute_indexes(*_grid(2, range(0, 5), range(0, 4)))
    with pytest.raises(nn.Nnet3Error, match="use-bias"):
        nb.propagate(idx, torch.zeros((10, 8), device="cuda"), torch.zeros((8, 8), device="cuda"))
    with pytest.raises(nn.Nnet3Error):  # bad initializer (tdnn.cc:124-132)
        nn.Component.new("TdnnDARTSV3Component", "input-dim=8 output-dim=8 time-offsets=0,0")


def test_gumbel_softmax_flops_copyn_onehot_components(nn):
    import torch

    from oracle import oracle as O

    g = np.random.default_rng(3)
    R = 640
    x = g.standard_normal((R, 8)).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    gs = nn.Component.new("GumbelSoftmaxFlopsComponent", "dim=8 scale=0.1 temp-proportion=0.4")
    assert gs.properties() == (nn.kBackpropInPlace | nn.kSimpleComponent | nn.kBackpropNeedsInput |
                               nn.kBackpropNeedsOutput | nn.kRandomComponent)
    c0 = nn.get_rand_counter()
    p = torch.zeros_like(xd)
    gs.propagate(None, xd, p)
    nn.set_rand_counter(c0)
    u = [nn.rand_uniform() for _ in range(8)]
    p_ref = O.softmax_flops_fwd(x, u, 0.4)
    assert rel_err(p.cpu().numpy(), p_ref) < 1e-4
    od = (g.standard_normal((R, 8)) / R).astype(np.float32)
    odd = torch.from_numpy(od).cuda()
    gs.backprop(None, xd, p, odd, None, None, odd)  # in place
    ind_ref, _ = O.softmax_flops_bwd(p_ref, od, 0.1, True, 0.4)
    assert rel_err(odd.cpu().numpy(), ind_ref) < 1e-3
    assert gs.write(False) == b"<GumbelSoftmaxFlopsComponent> <Dim> 8 <Scale> 0.1 <TempProportion> 0.4 </GumbelSoftmaxFlopsComponent> "
    assert nn.Component.read(gs.write(True), True).write(False) == gs.write(False)

    sm = nn.Component.new("SoftmaxFlopsComponent", "dim=8 scale=0.001")
    p2 = torch.zeros_like(xd)
    sm.propagate(None, xd, p2)
    assert rel_err(p2.cpu().numpy(), O.softmax_flops_fwd(x)) < 1e-4
    assert sm.write(False) == b"<SoftmaxFlopsComponent> <Dim> 8 <Scale> 0.001 </SoftmaxFlopsComponent> "

    cn = nn.Component.new("CopyNComponent", "input-dim=1 output-dim=30 scale=1.0")
    assert cn.properties() == (nn.kSimpleComponent | nn.kPropagateAdds | nn.kBackpropAdds)
    col = torch.from_numpy(x[:, :1].copy()).cuda()
    o = torch.zeros((R, 30), device="cuda")
    cn.propagate(None, col, o)
    assert np.array_equal(o.cpu().numpy(), np.tile(x[:, :1], (1, 30)))
    with pytest.raises(nn.Nnet3Error):
        nn.Component.new("CopyNComponent", "input-dim=4 output-dim=30")
    assert cn.write(False) == b"<CopyNComponent> <InputDim> 1 <OutputDim> 30 <Scale> 1 </CopyNComponent> "

    oh = nn.Component.new("OnehotFunctionComponent", "input-dim=220 output-dim=8 is-updatable=true use-natural-gradient=false learning-rate=0.1")
    c0 = nn.get_rand_counter()
    oo = torch.zeros((R, 8), device="cuda")
    oh.propagate(None, torch.zeros((R, 220), device="cuda"), oo)
    nn.set_rand_counter(c0)
    assert np.array_equal(oo.cpu().numpy(), O.onehot_fwd(R, 8, nn.rand_uniform()))
    d = oh.copy()
    d.scale(0.0)
    oh.backprop(None, None, None, odd, None, d, None)
    np.testing.assert_allclose(d.vectorize(), 0.1 * odd.cpu().numpy().sum(0), rtol=1e-4, atol=1e-8)
    back = nn.Component.read(oh.write(False), False)
    assert back.write(False) == oh.write(False)
    # ConstantFunctionComponent: the reference's x5 non-NG update (simple.cc:2636)
    cf = nn.Component.new("ConstantFunctionComponent", "input-dim=220 output-dim=8 is-updatable=true use-natural-gradient=false learning-rate=0.1 output-mean=0.5")
    co = torch.zeros((R, 8), device="cuda")
    cf.propagate(None, torch.zeros((R, 220), device="cuda"), co)
    assert torch.all(co == 0.5)
    d = cf.copy()
    d.scale(0.0)
    cf.backprop(None, None, None, odd, None, d, None)
    np.testing.assert_allclose(d.vectorize(), 5 * 0.1 * odd.cpu().numpy().sum(0), rtol=1e-4, atol=1e-8)


def test_batchnorm_test_component_from_model_text(nn):
    import torch

    from oracle import oracle as O

    g = np.random.default_rng(5)
    D = 96
    mean = g.standard_normal(D).astype(np.float32)
    var = g.uniform(0.2, 3.0, D).astype(np.float32)
    fmt = lambda v: " ".join(repr(float(x)) for x in v)
    # what `sed s/BatchNormComponent/BatchNormTestComponent/` leaves in the text model (norm.cc:956-982)
    txt = (f"<BatchNormTestComponent> <Dim> {D} <BlockDim> {D} <Epsilon> 0.001 <TargetRms> 1 <TestMode> T <Count> 5000 "
           f"<StatsMean>  [ {fmt(mean)} ]\n<StatsVar>  [ {fmt(var)} ]\n</BatchNormTestComponent> ").encode()
    bn = nn.Component.read(txt, False)
    assert bn.properties() == (nn.kSimpleComponent | nn.kBackpropNeedsOutput | nn.kPropagateInPlace | nn.kBackpropInPlace)
    count = 5000.0
    scale, offset = O.bn_test_derived(mean.astype(np.float64) * count,
                                      (var.astype(np.float64) + mean.astype(np.float64) ** 2) * count, count, 1e-3, 1.0)
    x = g.standard_normal((300, D)).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    bn.propagate(None, xd, xd)  # in place
    assert rel_err(xd.cpu().numpy(), O.scale_offset_rows(x, scale, offset)) < 1e-5
    od = torch.from_numpy(x).cuda()
    bn.backprop(None, None, xd, od, None, None, od)
    assert rel_err(od.cpu().numpy(), O.scale_offset_rows(x, scale, None)) < 1e-5
    out = bn.write(False)
    assert out.startswith(b"<BatchNormTestComponent> <Dim> 96 <BlockDim> 96 <Epsilon> 0.001 <TargetRms> 1 <TestMode> T <Count> 5000 <StatsMean>")
    # Write() turns the sums back into mean / variance in fp32 (norm.cc:968-975), so a round trip is
    # value-stable to rounding, not byte-stable -- in the reference as well.
    again = nn.Component.read(bn.write(True), True)
    assert len(again.write(True)) == len(bn.write(True))
    y1, y2 = torch.from_numpy(x).cuda(), torch.from_numpy(x).cuda()
    bn.propagate(None, y1, y1)
    again.propagate(None, y2, y2)
    assert rel_err(y2.cpu().numpy(), y1.cpu().numpy()) < 1e-6
    # not in test mode: undefined in the reference (quirk Q8) -> error here
    bn.set_test_mode(False)
    with pytest.raises(nn.Nnet3Error, match="test mode"):
        bn.propagate(None, xd, xd)


@pytest.mark.parametrize("D,block,R", [(96, 96, 200), (64, 16, 50)])
def test_batchnorm_component_training_mode(nn, D, block, R):
    """BatchNormComponent (nnet-normalize-component.cc:401-589): training-mode Propagate / Backprop against float64 numpy
    of the BATCHNORM_MATH comment (:321-398), StoreStats over two minibatches, the on-disk statistics, test mode, and
    equality with BatchNormTestComponent once the model text is `sed`-ed as the search recipe does."""
    import torch

    g = np.random.default_rng(D + R)
    eps, rms = 1e-3, 0.7
    bn = nn.Component.new("BatchNormComponent", f"dim={D} block-dim={block} epsilon={eps} target-rms={rms}")
    ratio = D // block
    pooled_sum, pooled_sumsq, count = np.zeros(block), np.zeros(block), 0
    for step in range(2):
        x = (g.standard_normal((R, D)) * (1 + step) + 0.3).astype(np.float32)
        zp = g.standard_normal((R, D)).astype(np.float32)
        xr, zpr = x.astype(np.float64).reshape(R * ratio, block), zp.astype(np.float64).reshape(R * ratio, block)
        mean, uvar = xr.mean(0), (xr ** 2).mean(0)
        scale = rms * (np.maximum(uvar - mean ** 2, 0) + eps) ** -0.5
        z = (xr - mean) * scale
        vdm = -1.0 / (rms * rms) * (zpr * z).mean(0) * scale
        xp = scale * (zpr - zpr.mean(0)) + z * vdm
        xd = torch.from_numpy(x).cuda()
        out = torch.empty_like(xd)
        memo = bn.propagate(None, xd, out)
        assert memo is not None
        assert rel_err(out.cpu().numpy(), z.reshape(R, D)) < 1e-5
        bn.store_stats(None, out, memo)
        ind = torch.from_numpy(zp).cuda()  # in place (kBackpropInPlace)
        bn.backprop(None, None, out, ind, memo, None, ind)
        bn.delete_memo(memo)
        assert rel_err(ind.cpu().numpy(), xp.reshape(R, D)) < 1e-4
        pooled_sum += xr.sum(0)
        pooled_sumsq += (xr ** 2).sum(0)
        count += R * ratio
    assert bn.bn_count() == count
    txt = bn.write(False).decode()
    mean_txt = np.array(txt.split("<StatsMean>")[1].split("[")[1].split("]")[0].split(), dtype=np.float64)
    var_txt = np.array(txt.split("<StatsVar>")[1].split("[")[1].split("]")[0].split(), dtype=np.float64)
    pm = pooled_sum / count
    np.testing.assert_allclose(mean_txt, pm, rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(var_txt, pooled_sumsq / count - pm ** 2, rtol=2e-4)
    # a copy carries the statistics; Scale(0) clears them; Add restores
    cp = bn.copy()
    assert cp.bn_count() == count
    cp.scale(0.0)
    assert cp.bn_count() == 0
    cp.add(1.0, bn)
    assert cp.bn_count() == count
    # test mode: the affine map from the pooled statistics, no memo, no kStoresStats
    bn.set_test_mode(True)
    assert bn.properties() & (nn.kUsesMemo | nn.kStoresStats) == 0
    x = g.standard_normal((R, D)).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    out = torch.empty_like(xd)
    assert bn.propagate(None, xd, out) is None
    sc = rms * (np.maximum(pooled_sumsq / count - pm ** 2, 0) + eps) ** -0.5
    ref = ((x.astype(np.float64).reshape(R * ratio, block) - pm) * sc).reshape(R, D)
    assert rel_err(out.cpu().numpy(), ref) < 1e-4
    # `sed s/BatchNormComponent/BatchNormTestComponent/` on the text model (the search-stage recipe) gives the same map
    test_comp = nn.Component.read(bn.write(False).replace(b"BatchNormComponent", b"BatchNormTestComponent"), False)
    assert test_comp.type() == "BatchNormTestComponent"
    out2 = torch.empty_like(xd)
    test_comp.propagate(None, xd, out2)
    assert rel_err(out2.cpu().numpy(), out.cpu().numpy()) < 1e-5  # the text form rounds the statistics to ~7 digits
    # round trip of the component itself; ZeroStats is a no-op in test mode (norm.cc:668-678)
    back = nn.Component.read(bn.write(True), True)
    assert back.type() == "BatchNormComponent" and back.write(False) == bn.write(False)
    bn.zero_stats()
    assert bn.bn_count() == count
    bn.set_test_mode(False)
    bn.zero_stats()
    assert bn.bn_count() == 0


def test_rectified_linear_component_stats_and_self_repair(nn):
    """RectifiedLinearComponent (nnet-simple-component.cc:958-1094, nnet-component-itf.cc:433-481): Propagate, StoreStats
    (value / derivative sums), Backprop with the self-repair term and the out_deriv statistics, replaying the component's
    RandInt / RandUniform draws; then the text form, a round trip, Scale / Add / ZeroStats."""
    import torch

    g = np.random.default_rng(12)
    R, D = 300, 64
    scale = 0.01
    comp = nn.Component.new("RectifiedLinearComponent", f"dim={D} self-repair-scale={scale}")
    x = g.standard_normal((R, D)).astype(np.float32)
    x[:, :5] = -np.abs(x[:, :5]) - 0.1       # dead units: derivative average 0 <= 0.05
    x[:, 5:9] = np.abs(x[:, 5:9]) + 0.1      # always-on units: derivative average 1 > 0.95
    xd = torch.from_numpy(x).cuda()
    out = torch.empty_like(xd)
    assert comp.propagate(None, xd, out) is None
    ref_out = np.maximum(x, 0)
    assert np.array_equal(out.cpu().numpy(), ref_out)
    # StoreStats: the first call always stores (count == 0), whatever RandInt(0, 1) says
    comp.store_stats(None, out, None)
    txt = comp.write(False).decode()
    vec = lambda tok: np.array(txt.split(tok)[1].split("[")[1].split("]")[0].split(), dtype=np.float64)
    np.testing.assert_allclose(vec("<ValueAvg>"), ref_out.mean(0), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(vec("<DerivAvg>"), (ref_out > 0).mean(0), rtol=1e-6)
    assert f"<Count> {R} " in txt
    # a second StoreStats is skipped when RandInt(0, 1) == 0
    c0 = nn.get_rand_counter()
    skip = nn.rand_int(0, 1) == 0
    nn.set_rand_counter(c0)
    comp.store_stats(None, out, None)
    assert (f"<Count> {R} " in comp.write(False).decode()) == skip
    # Backprop into a delta copy: self-repair uses the MODEL's statistics, the counters go to the delta
    delta = comp.copy()
    delta.zero_stats()
    od = g.standard_normal((R, D)).astype(np.float32)
    odd = torch.from_numpy(od).cuda()
    ind = torch.zeros_like(xd)
    c0 = nn.get_rand_counter()
    repaired = not (nn.rand_uniform() > 0.5)     # RepairGradients: "if (RandUniform() > repair_probability) return"
    nn.rand_int(0, 3)                            # StoreBackpropStats draws, then stores anyway (oderiv_count == 0)
    c1 = nn.get_rand_counter()
    nn.set_rand_counter(c0)
    comp.backprop(None, None, out, odd, None, delta, ind)
    assert nn.get_rand_counter() == c1
    ref = (ref_out > 0) * od
    if repaired:
        ref = ref.copy()
        ref[:, :5] += scale / 0.5                # stats <= lower threshold: push the derivative up
        ref[:, 5:9] -= scale / 0.5               # stats > upper threshold: push it down
    np.testing.assert_allclose(ind.cpu().numpy(), ref, rtol=1e-6, atol=1e-7)
    dtxt = delta.write(False).decode()
    assert f"<NumDimsProcessed> {D if repaired else 0} " in dtxt and f"<NumDimsSelfRepaired> {9 if repaired else 0} " in dtxt
    dvec = lambda tok: np.array(dtxt.split(tok)[1].split("[")[1].split("]")[0].split(), dtype=np.float64)
    np.testing.assert_allclose(dvec("<OderivRms>"), np.sqrt((od.astype(np.float64) ** 2).mean(0)), rtol=1e-5)
    assert f"<OderivCount> {R} " in dtxt
    # without to_update: the plain derivative, no draws
    c0 = nn.get_rand_counter()
    ind2 = torch.zeros_like(xd)
    comp.backprop(None, None, out, odd, None, None, ind2)
    assert nn.get_rand_counter() == c0
    np.testing.assert_allclose(ind2.cpu().numpy(), (ref_out > 0) * od, rtol=1e-6)
    # I/O round trip (text and binary), Add / Scale
    back = nn.Component.read(comp.write(True), True)
    assert back.write(False) == comp.write(False)
    assert nn.Component.read(comp.write(False), False).write(False) == comp.write(False)
    acc = comp.copy()
    acc.add(1.0, comp)
    acc.scale(0.5)
    assert acc.write(False) == comp.write(False)
    acc.zero_stats()
    assert "<Count> 0 " in acc.write(False).decode()
