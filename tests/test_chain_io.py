"""Host-only tests of the graph readers (tdnnf_den_graph_parse_fst_text / tdnnf_num_graph_parse_fst_texts): FSM text ->
the arrays the denominator / numerator kernels take, against the synthetic generators (round trip) and against a direct
Python restatement of DenominatorGraph::SetInitialProbs with final probabilities (kaldi chain-den-graph.cc)."""
import numpy as np
import pytest

from tdnnf_nas_b200 import capi, synth


def test_den_fst_text_round_trip():
    g = synth.make_den_graph(120, 37, 5.0, seed=9)
    back = capi.parse_den_fst_text(synth.den_graph_to_fst_text(g), 37)
    assert back["num_states"] == 120 and back["num_pdfs"] == 37 and back["num_arcs"] == g["num_arcs"]
    for k in ("fwd_ranges", "bwd_ranges", "pdf", "state"):
        np.testing.assert_array_equal(back[k], g[k])
    np.testing.assert_allclose(back["prob"], g["prob"], rtol=1e-6)
    np.testing.assert_allclose(back["init"], g["init"], rtol=1e-5, atol=1e-9)  # rows are stochastic, no final mass


def test_den_initial_probs_with_final_weights_and_scattered_arcs():
    # 3 states, start state 1 (the source of the first line), arcs of a state not contiguous in the text, final weights
    text = "1 2 3 3 0.5\n0 1 1 1 1.0\n1 0 2 2 0.25\n2 0 4 4 0.0\n0 0 1 1 0.7\n2 0.3\n0\n"
    got = capi.parse_den_fst_text(text, 5)
    arcs = [(1, 2, 2, 0.5), (0, 1, 0, 1.0), (1, 0, 1, 0.25), (2, 0, 3, 0.0), (0, 0, 0, 0.7)]  # (src, dst, pdf, weight)
    fwd = sorted(range(5), key=lambda i: arcs[i][0])
    np.testing.assert_array_equal(got["state"][:5], [arcs[i][1] for i in fwd])
    np.testing.assert_array_equal(got["pdf"][:5], [arcs[i][2] for i in fwd])
    np.testing.assert_allclose(got["prob"][:5], [np.exp(-arcs[i][3]) for i in fwd], rtol=1e-6)
    bwd = sorted(fwd, key=lambda i: arcs[i][1])  # stable: grouped by destination, source-state order kept
    np.testing.assert_array_equal(got["state"][5:], [arcs[i][0] for i in bwd])
    np.testing.assert_array_equal(got["fwd_ranges"], [[0, 2], [2, 4], [4, 5]])
    np.testing.assert_array_equal(got["bwd_ranges"], [[5, 8], [8, 9], [9, 10]])
    final = {2: 0.3, 0: 0.0}
    norm = np.array([sum(np.exp(-w) for s, _, _, w in arcs if s == h) + (np.exp(-final[h]) if h in final else 0.0) for h in range(3)])
    cur, avg = np.array([0.0, 1.0, 0.0]), np.zeros(3)
    for _ in range(100):
        avg += cur / 100
        nxt = np.zeros(3)
        for s, d, _, w in arcs:
            nxt[d] += cur[s] / norm[s] * np.exp(-w)
        cur = nxt / nxt.sum()
    np.testing.assert_allclose(got["init"], avg, rtol=1e-5)


def test_den_fst_text_errors():
    with pytest.raises(capi.TdnnfError, match="ilabel"):
        capi.parse_den_fst_text("0 1 9 9 0.1\n1\n", 5)       # pdf-id + 1 beyond num_pdfs
    with pytest.raises(capi.TdnnfError, match="line 2"):
        capi.parse_den_fst_text("0 1 1 1 0.1\n0 1 x\n", 5)
    with pytest.raises(capi.TdnnfError, match="empty"):
        capi.parse_den_fst_text("\n", 5)


def test_numerator_fst_texts_round_trip():
    den = synth.make_den_graph(50, 23, 4.0, seed=2)
    g = synth.make_num_graphs(5, 23, 12, seed=4, den_graph=den)
    back = capi.parse_num_fst_texts(synth.num_graphs_to_fst_texts(g), 23)
    assert back["num_seqs"] == 5 and back["num_arcs"] == g["num_arcs"]
    for k in ("state_offsets", "fwd_ranges", "bwd_ranges", "arc_pdf", "arc_state"):
        np.testing.assert_array_equal(back[k], g[k])
    np.testing.assert_allclose(back["arc_logprob"], g["arc_logprob"], rtol=1e-6)
    np.testing.assert_array_equal(back["final_logprob"] > -1e29, g["final_logprob"] > -1e29)
    # a start state that is not 0 is moved to the front
    alt = capi.parse_num_fst_texts(["2 0 4 4 0.5\n0 1 5 5\n1 0.25\n"], 23)
    np.testing.assert_array_equal(alt["arc_state"][:2], [2, 1])     # local 2 -> 0 (start), local 0 -> 2: arcs 0->2 (pdf 3), 2->1 (pdf 4)
    np.testing.assert_array_equal(alt["arc_pdf"][:2], [3, 4])
    assert alt["final_logprob"][1] == pytest.approx(-0.25) and alt["final_logprob"][0] < -1e29
