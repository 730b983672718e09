"""GPU parity of the chain denominator forward-backward against the CPU oracle, plus the
size-independent invariants at larger sizes (posteriors sum to 1 per frame per sequence)."""
import numpy as np
import pytest

from tests.util import rel_err

pytestmark = pytest.mark.gpu


def _run_gpu(ctx, graph, x, S, T, leaky, w, want_path=None, pad=0, repeat=1):
    import torch

    from tdnnf_nas_b200 import capi

    dg = capi.DenGraph(ctx, graph)
    dc = capi.DenominatorComputation(ctx, dg, S, T, leaky)
    if want_path is not None:
        assert dc.describe()["path"] == want_path, dc.describe()
    P = x.shape[1]
    xd = torch.from_numpy(np.pad(x, ((0, 0), (0, pad)), constant_values=99.0)).cuda()[:, :P]  # stride > cols like a CuMatrix pitch
    for _ in range(repeat):  # a second call reuses every workspace
        lp = dc.forward(xd)
        deriv = torch.zeros((x.shape[0], P + pad), device="cuda")[:, :P]
        ok = dc.backward(w, deriv)
    torch.cuda.synchronize()
    out = deriv.cpu().numpy()
    dc.close()
    dg.close()
    return lp, out, ok


@pytest.mark.parametrize("N,P,S,T,deg", [(40, 12, 3, 5, 3.0), (300, 50, 32, 9, 6.0), (500, 200, 64, 12, 8.0),
                                          (1000, 300, 128, 7, 16.0), (257, 33, 20, 4, 5.0)])
def test_den_parity(ctx, monkeypatch, N, P, S, T, deg):
    """The per-frame kernels (any num_seqs)."""
    from oracle import oracle as O
    from tdnnf_nas_b200 import synth

    monkeypatch.setenv("TDNNF_DEN_PATH", "frames")
    graph = synth.make_den_graph(N, P, deg, seed=N)
    g = np.random.default_rng(N + S)
    x = np.clip(g.standard_normal((T * S, P)) * 2.0, -30, 30).astype(np.float32)
    x[0, 0] = 40.0   # exercises the +-30 clamp (ApplyExpLimited)
    x[1, 1] = -45.0
    lp_ref, d_ref, ok_ref = O.den_forward_backward(graph, x, S, T, 0.1, deriv_weight=-1.0)
    lp, d, ok = _run_gpu(ctx, graph, x, S, T, 0.1, -1.0)
    assert ok and ok_ref
    assert abs(lp - lp_ref) <= 1e-4 * abs(lp_ref), (lp, lp_ref)
    assert rel_err(d, d_ref) < 1e-3
    # posterior mass: -deriv sums to 1 per (t, s)
    np.testing.assert_allclose(-d.sum(axis=1), 1.0, rtol=2e-3)


@pytest.mark.parametrize("parts", [1, 2])
@pytest.mark.parametrize("N,P,S,T,deg,cluster", [(40, 12, 8, 5, 3.0, 1), (300, 50, 32, 9, 6.0, 16), (500, 203, 64, 12, 8.0, 16),
                                                  (1000, 300, 128, 7, 16.0, 9), (257, 33, 16, 4, 5.0, 3), (2000, 601, 24, 6, 12.0, 8),
                                                  (16, 7, 8, 3, 2.0, 4), (700, 90, 40, 1, 9.0, 2)])
def test_den_slice_path_parity(ctx, monkeypatch, N, P, S, T, deg, cluster, parts):
    """The sequence-slice cluster kernels (den_slices.cu; opt-in, TDNNF_DEN_PATH=slices), on small shapes: cluster sizes 1..16 (more CTAs than warp-tasks included), E staged in one or two pdf ranges, pdf counts
    that are not multiples of 4, state counts that are not multiples of 16, strided matrices, T = 1, two calls."""
    from oracle import oracle as O
    from tdnnf_nas_b200 import synth

    monkeypatch.setenv("TDNNF_DEN_PATH", "slices")
    monkeypatch.setenv("TDNNF_DEN_PARTS", str(parts))
    monkeypatch.setenv("TDNNF_DEN_CLUSTER", str(cluster))
    graph = synth.make_den_graph(N, P, deg, seed=N)
    g = np.random.default_rng(N + S)
    x = np.clip(g.standard_normal((T * S, P)) * 2.0, -30, 30).astype(np.float32)
    x[0, 0] = 40.0
    x[1, 1] = -45.0
    lp_ref, d_ref, ok_ref = O.den_forward_backward(graph, x, S, T, 0.1, deriv_weight=-1.0)
    lp, d, ok = _run_gpu(ctx, graph, x, S, T, 0.1, -1.0, want_path="slices", pad=3, repeat=2)
    assert ok and ok_ref
    assert abs(lp - lp_ref) <= 1e-4 * abs(lp_ref), (lp, lp_ref)
    assert rel_err(d, d_ref) < 1e-3
    np.testing.assert_allclose(-d.sum(axis=1), 1.0, rtol=2e-3)


def test_den_slice_path_matches_frame_path_at_full_size(ctx, monkeypatch):
    """BASELINE configs[4] size (16 384 states, 6008 pdfs, 64 sequences): the opt-in slice path agrees with the default
    per-frame kernels (different summation orders: 1e-5 / 1e-4)."""
    from tdnnf_nas_b200 import synth

    N, P, S, T = 16384, 6008, 64, 17
    graph = synth.make_den_graph(N, P, 16.0, seed=12)
    g = np.random.default_rng(1)
    x = np.clip(g.standard_normal((T * S, P)), -30, 30).astype(np.float32)
    monkeypatch.setenv("TDNNF_DEN_PATH", "slices")
    lp_s, d_s, ok_s = _run_gpu(ctx, graph, x, S, T, 0.1, -1.0, want_path="slices")
    monkeypatch.delenv("TDNNF_DEN_PATH", raising=False)
    lp_f, d_f, ok_f = _run_gpu(ctx, graph, x, S, T, 0.1, -1.0, want_path="frames")
    assert ok_s and ok_f
    assert abs(lp_s - lp_f) <= 1e-5 * abs(lp_f), (lp_s, lp_f)
    assert rel_err(d_s, d_f) < 1e-4
    np.testing.assert_allclose(-d_s.sum(axis=1), 1.0, rtol=2e-3)


@pytest.mark.parametrize("N,P,S,T,deg", [(300, 50, 32, 9, 6.0), (1000, 300, 128, 7, 16.0), (257, 33, 21, 4, 5.0)])
def test_den_resident_path_parity(ctx, monkeypatch, N, P, S, T, deg):
    """The opt-in cluster-resident kernels (TDNNF_DEN_RESIDENT=1; V = 4 / 4 / 1 sequences per cluster here)."""
    from oracle import oracle as O
    from tdnnf_nas_b200 import synth

    monkeypatch.setenv("TDNNF_DEN_RESIDENT", "1")
    graph = synth.make_den_graph(N, P, deg, seed=N)
    g = np.random.default_rng(N + S)
    x = np.clip(g.standard_normal((T * S, P)) * 2.0, -30, 30).astype(np.float32)
    lp_ref, d_ref, ok_ref = O.den_forward_backward(graph, x, S, T, 0.1, deriv_weight=-1.0)
    lp, d, ok = _run_gpu(ctx, graph, x, S, T, 0.1, -1.0)
    assert ok and ok_ref
    assert abs(lp - lp_ref) <= 1e-4 * abs(lp_ref), (lp, lp_ref)
    assert rel_err(d, d_ref) < 1e-3


def test_den_invariants_large(ctx):
    """Switchboard-shaped sizes the oracle would take minutes on: check invariants instead."""
    from tdnnf_nas_b200 import synth

    N, P, S, T = 8192, 6008, 64, 34
    graph = synth.make_den_graph(N, P, 16.0, seed=11)
    g = np.random.default_rng(0)
    x = np.clip(g.standard_normal((T * S, P)), -30, 30).astype(np.float32)
    lp, d, ok = _run_gpu(ctx, graph, x, S, T, 0.1, -1.0)
    assert ok and np.isfinite(lp)
    np.testing.assert_allclose(-d.sum(axis=1), 1.0, rtol=2e-3)
    assert (d <= 1e-6).all()
    # shifting every output of a frame by a constant c changes the log-prob by exactly S*c per frame
    x2 = x.copy()
    x2[: S] += 0.5
    lp2, _, _ = _run_gpu(ctx, graph, x2, S, T, 0.1, -1.0)
    assert abs((lp2 - lp) - 0.5 * S) < 1e-3 * abs(lp) + 1e-2


@pytest.mark.parametrize("S,P,T", [(4, 9, 6), (64, 300, 50), (7, 40, 17)])
def test_numerator_parity(ctx, S, P, T):
    import torch

    from oracle import oracle as O
    from tdnnf_nas_b200 import capi, synth

    graph = synth.make_num_graphs(S, P, T, seed=S)
    g = np.random.default_rng(S)
    x = (g.standard_normal((T * S, P)) * 2).astype(np.float32)
    lp_ref, d_ref, ok_ref = O.num_forward_backward(graph, x, T, deriv_weight=1.0)
    ng = capi.NumeratorGraph(ctx, graph)
    xd = torch.from_numpy(x).cuda()
    deriv = torch.zeros_like(xd)
    lp, ok = ng.forward_backward(xd, T, 1.0, deriv)
    assert ok and ok_ref
    assert abs(lp - lp_ref) <= 1e-4 * abs(lp_ref)
    assert rel_err(deriv.cpu().numpy(), d_ref) < 1e-3
    np.testing.assert_allclose(deriv.cpu().numpy().sum(1), 1.0, rtol=1e-3)
    ng.close()


def test_chain_objf_and_deriv(ctx):
    """ComputeChainObjfAndDeriv (chain.py) == oracle numerator - oracle denominator, derivative included; the
    failure path (no numerator path) returns -10 per frame with a zero derivative, as in Kaldi."""
    import torch

    from oracle import oracle as O
    from tdnnf_nas_b200 import capi, chain, synth

    S, P, T, N = 16, 60, 12, 200
    dgraph = synth.make_den_graph(N, P, 5.0, seed=4)
    ngraph = synth.make_num_graphs(S, P, T, seed=8)
    g = np.random.default_rng(2)
    x = g.standard_normal((T * S, P)).astype(np.float32)
    den_lp, den_d, _ = O.den_forward_backward(dgraph, x, S, T, 0.1, deriv_weight=-1.0)
    num_lp, num_d, _ = O.num_forward_backward(ngraph, x, T, deriv_weight=1.0)
    dg, ng = capi.DenGraph(ctx, dgraph), capi.NumeratorGraph(ctx, ngraph)
    obj = chain.ChainObjective(ctx, dg, ng, S, T, chain.ChainTrainingOptions(l2_regularize=1e-3))
    xd = torch.from_numpy(x).cuda()
    deriv = torch.full_like(xd, 3.0)
    objf, l2, weight = obj.compute(xd, deriv)
    assert weight == S * T
    assert objf == pytest.approx(num_lp - den_lp, rel=1e-4)
    assert l2 == pytest.approx(-0.5 * 1e-3 * float((x.astype(np.float64) ** 2).sum()), rel=1e-4)
    assert rel_err(deriv.cpu().numpy(), num_d + den_d - 1e-3 * x) < 1e-3
    # xent branch: the numerator posteriors come back on their own (targets of the output-xent node), the total
    # derivative is unchanged, and the cross-entropy objective / scaled derivative follow NnetChainTrainer
    deriv_x = torch.full_like(xd, -1.0)
    xent = torch.full_like(xd, 9.0)
    objf_x, _, _ = obj.compute(xd, deriv_x, xent)
    assert objf_x == pytest.approx(objf, rel=1e-6)
    assert rel_err(xent.cpu().numpy(), num_d) < 1e-3
    assert rel_err(deriv_x.cpu().numpy(), deriv.cpu().numpy()) < 1e-6
    xo = torch.log_softmax(torch.from_numpy(g.standard_normal((T * S, P)).astype(np.float32)).cuda(), dim=1)
    ref_xent_objf = float((xo.double().cpu().numpy() * num_d.astype(np.float64)).sum())
    assert obj.xent_objf_and_deriv(xo, xent) == pytest.approx(ref_xent_objf, rel=1e-4)
    assert rel_err(xent.cpu().numpy(), 0.1 * num_d) < 1e-3  # xent_regularize = 0.1
    # T shorter than the phone strings: no numerator path
    bad = capi.NumeratorGraph(ctx, synth.make_num_graphs(S, P, T, seed=8, min_phones=T, max_phones=T))
    obj2 = chain.ChainObjective(ctx, dg, bad, S, 4)
    x2 = torch.from_numpy(x[: 4 * S].copy()).cuda()
    d2 = torch.full_like(x2, 3.0)
    x2ent = torch.full_like(x2, 3.0)
    objf2, _, w2 = obj2.compute(x2, d2, x2ent)
    assert objf2 == -10.0 * w2 and torch.all(d2 == 0) and torch.all(x2ent == 0)
    for o in (obj, obj2):
        o.close()
    ng.close(); bad.close(); dg.close()


def test_numerator_supervision_update(ctx):
    """tdnnf_num_graph_update: the next minibatch's numerator FSTs uploaded into the same handle give what a freshly
    created graph gives -- smaller and larger supervisions than the first one (capacity growth included)."""
    import torch

    from tdnnf_nas_b200 import capi, synth

    S, P, T = 6, 40, 14
    g = np.random.default_rng(3)
    x = torch.from_numpy((g.standard_normal((T * S, P)) * 2).astype(np.float32)).cuda()
    first = synth.make_num_graphs(S, P, T, seed=1, min_phones=3, max_phones=5)
    handle = capi.NumeratorGraph(ctx, first)
    for seed, lo, hi in ((2, 3, 4), (3, 8, 12), (4, 3, 12), (5, 10, 12)):
        sup = synth.make_num_graphs(S, P, T, seed=seed, min_phones=lo, max_phones=hi)
        nbytes = handle.update(capi.NumeratorGraph.host_arrays(sup))
        assert nbytes > 0
        d_upd, d_new = torch.zeros_like(x), torch.zeros_like(x)
        lp_upd, ok_upd = handle.forward_backward(x, T, 1.0, d_upd)
        fresh = capi.NumeratorGraph(ctx, sup)
        lp_new, ok_new = fresh.forward_backward(x, T, 1.0, d_new)
        fresh.close()
        assert ok_upd and ok_new and lp_upd == pytest.approx(lp_new, rel=1e-6)
        assert rel_err(d_upd.cpu().numpy(), d_new.cpu().numpy()) < 1e-6
    handle.close()
