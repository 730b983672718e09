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


def _den_logprob_f64(graph, x, S, T, leaky):
    """Independent float64 statement of the leaky-HMM denominator log-probability, dense matrices:
    alpha_0 = init;  alpha'_t = alpha_t + leaky * init * sum(alpha_t);  alpha_{t+1} = alpha'_t . (sum_pdf e_t[pdf] A[pdf]);
    log p = sum_s log sum(alpha'_T)."""
    N, P = int(graph["num_states"]), int(graph["num_pdfs"])
    A = np.zeros((P, N, N))
    fr = graph["fwd_ranges"]
    for h in range(N):
        for a in range(fr[h][0], fr[h][1]):
            A[graph["pdf"][a], h, graph["state"][a]] += float(graph["prob"][a])
    init = graph["init"].astype(np.float64)
    total = 0.0
    for s in range(S):
        alpha = init.copy()
        for t in range(T):
            ad = alpha + leaky * init * alpha.sum()
            alpha = ad @ np.tensordot(np.exp(x[t * S + s]), A, axes=(0, 0))
        total += np.log((alpha + leaky * init * alpha.sum()).sum())
    return total


def test_denominator_backward_is_the_gradient_of_an_independent_forward():
    """The oracle's Backward (beta recursion with the leaky-HMM terms, per-frame rescaling, posterior accumulation) against
    the element-wise gradient of an independent float64 forward pass, on random tiny graphs: every (frame, sequence, pdf)
    posterior, not just a directional derivative."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    @settings(max_examples=25, deadline=None)
    @given(st.integers(2, 6), st.integers(2, 4), st.integers(1, 3), st.integers(1, 4), st.sampled_from([0.1, 0.5, 1e-5]),
           st.booleans(), st.integers(0, 10 ** 6))
    def check(N, P, S, T, leaky, self_loops, seed):
        graph = synth.make_den_graph(N, P, 2.0, seed=seed, self_loops=self_loops)
        g = np.random.default_rng(seed)
        x = g.standard_normal((T * S, P)).astype(np.float32)
        lp, d, ok = O.den_forward_backward(graph, x, S, T, leaky, deriv_weight=-1.0)
        assert ok
        x64 = x.astype(np.float64)
        assert lp == pytest.approx(_den_logprob_f64(graph, x64, S, T, leaky), rel=2e-5, abs=2e-5)
        grad = np.zeros_like(x64)
        eps = 1e-5
        for i in range(x64.shape[0]):
            for j in range(P):
                xp, xm = x64.copy(), x64.copy()
                xp[i, j] += eps
                xm[i, j] -= eps
                grad[i, j] = (_den_logprob_f64(graph, xp, S, T, leaky) - _den_logprob_f64(graph, xm, S, T, leaky)) / (2 * eps)
        np.testing.assert_allclose(-d, grad, rtol=2e-4, atol=2e-5)

    check()
