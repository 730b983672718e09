"""GPU tests of the stock TdnnComponent (BASELINE configs[1]: the manual TDNN-F 7q system and every architecture the
search emits are built from it) and of ConstrainOrthonormal, against the oracle."""
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
    yield nnet3
    nnet3.set_ng_identity(False)


def _grid(S, t_in, t_out):
    return [(n, t, 0) for t in t_in for n in range(S)], [(n, t, 0) for t in t_out for n in range(S)]


@pytest.mark.parametrize("offsets,subsample", [([-3, 0], 1), ([0, 3], 1), ([0], 1), ([0, 3], 3), ([-6, 0], 1)])
@pytest.mark.parametrize("use_bias", [True, False], ids=["bias", "nobias"])
@pytest.mark.parametrize("update", ["simple", "ng"])
def test_tdnn_component_vs_oracle(nn, offsets, subsample, use_bias, update):
    import torch

    from oracle import oracle as O

    n, din, dout, S = len(offsets), 96, 40, 8
    cfg = (f"input-dim={din} output-dim={dout} time-offsets={','.join(map(str, offsets))} learning-rate=0.02 "
           f"use-bias={'true' if use_bias else 'false'} use-natural-gradient={'true' if update == 'ng' else 'false'}")
    comp = nn.Component.new("TdnnComponent", cfg)
    assert comp.type() == "TdnnComponent" and comp.input_dim() == din and comp.output_dim() == dout
    want = nn.kUpdatableComponent | nn.kReordersIndexes | nn.kBackpropAdds | nn.kBackpropNeedsInput
    if not use_bias:
        want |= nn.kPropagateAdds
    assert comp.properties() == want  # no kUsesMemo: the stock class has no memo
    assert comp.num_parameters() == dout * n * din + (dout if use_bias else 0)
    v = comp.vectorize()
    W = v[: dout * n * din].reshape(dout, n * din).copy()
    b = v[dout * n * din:].copy() if use_bias else None

    t_out = list(range(0, 30, subsample))
    t_in = list(range(min(offsets), t_out[-1] + max(offsets) + 1))
    inp, outp = comp.reorder_indexes(*_grid(S, t_in, t_out))
    idx = comp.precompute_indexes(inp, outp)
    row_stride, row_offsets = idx.row_stride_and_offsets()
    in_rows, out_rows = len(inp), len(outp)
    g = np.random.default_rng(n + subsample)
    x = g.standard_normal((in_rows, din)).astype(np.float32)
    od = (g.standard_normal((out_rows, dout)) / out_rows).astype(np.float32)
    pre = g.standard_normal((out_rows, dout)).astype(np.float32)
    xd, odd, out = torch.from_numpy(x).cuda(), torch.from_numpy(od).cuda(), torch.from_numpy(pre).cuda()

    c0 = nn.get_rand_counter()
    memo = comp.propagate(idx, xd, out)
    assert nn.get_rand_counter() == c0 and not memo  # deterministic, no memo
    out_ref = O.plain_tdnn_propagate(W, b, x, out_rows, row_offsets, row_stride, out=pre.copy())
    assert rel_err(out.cpu().numpy(), out_ref) < 1e-4

    delta = comp.copy()
    delta.scale(0.0)
    in0 = g.standard_normal((in_rows, din)).astype(np.float32) * 1e-3
    in_deriv = torch.from_numpy(in0).cuda()  # kBackpropAdds
    comp.backprop(idx, xd, None, odd, memo, delta, in_deriv)
    ind_ref = in0.copy()
    dW_ref = np.zeros_like(W)
    db_ref = np.zeros(dout, np.float32) if use_bias else None
    ng = update == "ng"
    ng_in = O.NaturalGradient(min(20, (n * din + 1) // 2), 4, 2000.0, 4.0) if ng else None
    ng_out = O.NaturalGradient(min(80, (dout + 1) // 2), 4, 2000.0, 4.0) if ng else None
    O.plain_tdnn_backprop(W, x, od, row_offsets, row_stride, delta.learning_rate(), in_deriv=ind_ref, dW=dW_ref, dbias=db_ref,
                          natural_gradient=ng, ng_in=ng_in, ng_out=ng_out)
    dv = delta.vectorize()
    assert rel_err(in_deriv.cpu().numpy(), ind_ref) < 1e-3
    assert rel_err(dv[: dout * n * din].reshape(dout, n * din), dW_ref) < 1e-3
    if use_bias:
        assert rel_err(dv[dout * n * din:], db_ref) < 1e-3
    # a second minibatch through the same preconditioners (t = 1: the first real Fisher update)
    if ng:
        x2 = g.standard_normal((in_rows, din)).astype(np.float32)
        od2 = (g.standard_normal((out_rows, dout)) / out_rows).astype(np.float32)
        comp.backprop(idx, torch.from_numpy(x2).cuda(), None, torch.from_numpy(od2).cuda(), None, delta, None)
        O.plain_tdnn_backprop(W, x2, od2, row_offsets, row_stride, delta.learning_rate(), dW=dW_ref, dbias=db_ref,
                              natural_gradient=True, ng_in=ng_in, ng_out=ng_out)
        dv = delta.vectorize()
        assert rel_err(dv[: dout * n * din].reshape(dout, n * din), dW_ref) < 1e-3
    before = comp.vectorize()
    comp.add(1.0, delta)
    np.testing.assert_allclose(comp.vectorize(), before + delta.vectorize(), rtol=1e-6, atol=1e-7)


def test_tdnn_component_io_and_errors(nn):
    cfg = ("input-dim=20 output-dim=12 time-offsets=-3,0 learning-rate-factor=0.5 max-change=0.75 l2-regularize=0.01 "
           "orthonormal-constraint=-1.0 rank-in=7 rank-out=5")
    comp = nn.Component.new("TdnnComponent", cfg)
    assert comp.orthonormal_constraint() == -1.0
    txt = comp.write(False)
    toks = txt.decode().split()
    order = ["<TdnnComponent>", "<LearningRateFactor>", "<MaxChange>", "<L2Regularize>", "<LearningRate>", "<TimeOffsets>",
             "<LinearParams>", "<BiasParams>", "<OrthonormalConstraint>", "<UseNaturalGradient>", "<NumSamplesHistory>",
             "<AlphaInOut>", "<RankInOut>", "</TdnnComponent>"]
    pos = [toks.index(t) for t in order]
    assert pos == sorted(pos)
    assert b"<RankInOut> 7 5" in txt and b"use-gumbel" not in txt
    binary = comp.write(True)
    back = nn.Component.read(binary, True)
    assert back.type() == "TdnnComponent" and back.write(True) == binary
    np.testing.assert_array_equal(back.vectorize(), comp.vectorize())
    assert nn.Component.read(txt, False).write(False) == txt
    assert "orthonormal-constraint=-1" in comp.info() and "time-offsets=-3,0" in comp.info()
    cp = comp.copy()
    assert cp.type() == "TdnnComponent" and cp.write(True) == binary
    with pytest.raises(nn.Nnet3Error, match="Could not process"):
        nn.Component.new("TdnnComponent", "input-dim=4 output-dim=4 time-offsets=0 use-gumbel=true")
    with pytest.raises(nn.Nnet3Error, match="time-offsets"):
        nn.Component.new("TdnnComponent", "input-dim=4 output-dim=4")


def _near_semi_orthogonal(g, rows, cols, scale, spread):
    k, m = min(rows, cols), max(rows, cols)
    q, _ = np.linalg.qr(g.standard_normal((m, k)))
    a = scale * q.T + spread / np.sqrt(m) * g.standard_normal((k, m))
    return (a if rows <= cols else a.T).astype(np.float32)


@pytest.mark.parametrize("scale", [-1.0, 1.0, 0.5])
@pytest.mark.parametrize("rows,cols,spread,pad", [(160, 3072, 0.3, 0), (160, 1536, 0.05, 0), (24, 24, 0.3, 3), (37, 130, 1.0, 5),
                                                  (256, 1536, 0.3, 0), (300, 48, 0.3, 2), (512, 700, 0.2, 0)])
def test_constrain_orthonormal_vs_oracle(ctx, rows, cols, spread, pad, scale):
    import torch

    from oracle import oracle as O

    g = np.random.default_rng(rows * 7 + cols)
    M = _near_semi_orthogonal(g, rows, cols, abs(scale) if scale > 0 else 0.8, spread)
    buf = torch.full((rows, cols + pad), 7.0, device="cuda")
    buf[:, :cols] = torch.from_numpy(M).cuda()
    view = buf[:, :cols]
    info = torch.zeros(4, device="cuda")
    ctx.constrain_orthonormal(view, scale, info)
    # the oracle takes rows <= cols (the reference transposes otherwise, utils.cc:1067-1074)
    ref = M.copy() if rows <= cols else np.ascontiguousarray(M.T)
    info_ref = O.constrain_orthonormal(ref, scale)
    if rows > cols:
        ref = ref.T
    got = view.cpu().numpy()
    # compare the UPDATE (M_new - M), which is what the kernels compute, as well as the result
    assert rel_err(got, ref) < 2e-6
    assert rel_err(got - M, ref - M) < 1e-3
    np.testing.assert_allclose(info.cpu().numpy()[:3], info_ref[:3], rtol=2e-5)
    assert abs(float(info[3]) - info_ref[3]) < 1e-3 * max(info_ref[3], 1e-2)
    if pad:
        assert bool((buf[:, cols:] == 7.0).all())  # pitch padding untouched


def test_constrain_orthonormal_over_components(nn, ctx):
    """utils.cc:1037-1077: only TdnnComponents with a constraint, one RandInt(0,3) draw per constrained component."""
    import torch

    lin = [nn.Component.new("TdnnComponent", f"input-dim=64 output-dim=16 time-offsets=-1,0 orthonormal-constraint=-1.0 use-bias=false")
           for _ in range(6)]
    free = nn.Component.new("TdnnComponent", "input-dim=16 output-dim=64 time-offsets=0,1")
    darts = nn.Component.new("TdnnDARTSV3Component", "input-dim=64 output-dim=16 time-offsets=-1,0 orthonormal-constraint=-1.0")
    comps = [lin[0], free, lin[1], darts] + lin[2:]
    before = [c.vectorize() for c in comps]
    c0 = nn.get_rand_counter()
    k = nn.constrain_orthonormal(comps)
    c1 = nn.get_rand_counter()
    nn.set_rand_counter(c0)
    draws = [nn.rand_int(0, 3) for _ in range(6)]
    assert nn.get_rand_counter() == c1  # exactly one draw per constrained TdnnComponent
    chosen = [d == 0 for d in draws]
    assert k == sum(chosen)
    from oracle import oracle as O

    j = 0
    for c, b in zip(comps, before):
        after = c.vectorize()
        if c in lin:
            if chosen[j]:
                ref = b.reshape(16, 128).copy()
                O.constrain_orthonormal(ref, -1.0)
                assert rel_err(after.reshape(16, 128), ref) < 2e-6 and not np.array_equal(after, b)
            else:
                np.testing.assert_array_equal(after, b)
            j += 1
        else:
            np.testing.assert_array_equal(after, b)  # unconstrained TdnnComponent, and TdnnDARTSV3Component (not covered)
    # repeated application drives M M^T to scale^2 I
    for _ in range(100):
        nn.constrain_orthonormal(lin)
    for c in lin:
        M = c.vectorize().reshape(16, 128).astype(np.float64)
        P = M @ M.T
        s2 = np.trace(P) / 16
        assert np.linalg.norm(P - s2 * np.eye(16)) < 1e-4 * s2 * 4
