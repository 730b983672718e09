"""GPU parity of the chain denominator forward-backward against the CPU oracle, plus the
size-independent invariants at larger sizes (posteriors sum to 1 per frame per sequence)."""
import numpy as np
import pytest

from tests.util import rel_err

pytestmark = pytest.mark.gpu


def _run_gpu(ctx, graph, x, S, T, leaky, w):
    import torch

    from tdnnf_nas_b200 import capi

    dg = capi.DenGraph(ctx, graph)
    dc = capi.DenominatorComputation(ctx, dg, S, T, leaky)
    xd = torch.from_numpy(x).cuda()
    lp = dc.forward(xd)
    deriv = torch.zeros_like(xd)
    ok = dc.backward(w, deriv)
    torch.cuda.synchronize()
    out = deriv.cpu().numpy()
    dc.close()
    dg.close()
    return lp, out, ok


@pytest.mark.parametrize("N,P,S,T,deg", [(40, 12, 3, 5, 3.0), (300, 50, 32, 9, 6.0), (500, 200, 64, 12, 8.0),
                                          (1000, 300, 128, 7, 16.0), (257, 33, 20, 4, 5.0)])
def test_den_parity(ctx, N, P, S, T, deg):
    from oracle import oracle as O
    from tdnnf_nas_b200 import synth

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
