"""CPU tests of the oracle's chain objective pieces: the denominator invariants and a brute-force
enumeration of every path of tiny numerator FSTs."""
import itertools

import numpy as np
import pytest

from oracle import oracle as O
from tdnnf_nas_b200 import synth


def test_numerator_oracle_vs_path_enumeration():
    P, T, S = 7, 5, 4
    graph = synth.make_num_graphs(S, P, T, seed=3, min_phones=2, max_phones=4)
    g = np.random.default_rng(0)
    x = g.standard_normal((T * S, P)).astype(np.float32)
    lp, deriv, ok = O.num_forward_backward(graph, x, T, deriv_weight=1.0)
    assert ok
    so, fr = graph["state_offsets"], graph["fwd_ranges"]
    total = 0.0
    post = np.zeros((T * S, P))
    for s in range(S):
        s0, ns = so[s], so[s + 1] - so[s]
        arcs = {h: list(range(fr[s0 + h][0], fr[s0 + h][1])) for h in range(ns)}
        paths = []  # (logprob, [(t, pdf)])

        def walk(h, t, lpacc, occ):
            if t == T:
                f = graph["final_logprob"][s0 + h]
                if f > -1e29:
                    paths.append((lpacc + f, occ))
                return
            for a in arcs[h]:
                pdf = graph["arc_pdf"][a]
                walk(graph["arc_state"][a] - s0, t + 1, lpacc + graph["arc_logprob"][a] + x[t * S + s, pdf], occ + [(t, pdf)])

        walk(0, 0, 0.0, [])
        assert paths
        m = max(p for p, _ in paths)
        z = m + np.log(sum(np.exp(p - m) for p, _ in paths))
        total += z
        for p, occ in paths:
            for t, pdf in occ:
                post[t * S + s, pdf] += np.exp(p - z)
    assert lp == pytest.approx(total, rel=1e-6)
    np.testing.assert_allclose(deriv, post, rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(deriv.sum(1), 1.0, rtol=1e-5)  # one pdf per frame per sequence


def test_numerator_without_a_path_is_flagged():
    graph = synth.make_num_graphs(2, 5, 3, seed=1, min_phones=3, max_phones=3)
    x = np.zeros((2 * 2, 5), np.float32)  # T = 2 < 3 phones: no complete path
    lp, deriv, ok = O.num_forward_backward(graph, x, 2, deriv_weight=1.0)
    assert not ok and np.all(deriv == 0)


@pytest.mark.parametrize("N,P,S,T", [(30, 8, 3, 6), (120, 20, 5, 4)])
def test_denominator_invariants(N, P, S, T):
    graph = synth.make_den_graph(N, P, 4.0, seed=N)
    g = np.random.default_rng(1)
    x = g.standard_normal((T * S, P)).astype(np.float32)
    lp, d, ok = O.den_forward_backward(graph, x, S, T, 0.1, deriv_weight=-1.0)
    assert ok
    np.testing.assert_allclose(-d.sum(axis=1), 1.0, rtol=1e-4)       # posteriors sum to one per (t, s)
    # d(logprob)/dx equals the posterior: finite differences along a random direction
    v = g.standard_normal(x.shape).astype(np.float32)
    eps = 1e-2
    lp_p, _, _ = O.den_forward_backward(graph, x + eps * v, S, T, 0.1)
    lp_m, _, _ = O.den_forward_backward(graph, x - eps * v, S, T, 0.1)
    assert (lp_p - lp_m) / (2 * eps) == pytest.approx(float(-(d * v).sum()), rel=2e-2)
