"""Synthetic workloads of SURVEY.md section 8d: den graphs, fbank-shaped chunks, supernet shapes.

Everything is generated from numpy's counter-based Philox generator with fixed seeds so that the
GPU path, the oracle and every data-parallel rank see identical inputs.
"""
from __future__ import annotations

import numpy as np

BASE_SEED = 20221


def rng(config_index: int, stream: int = 0) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=BASE_SEED + config_index, counter=[stream, 0, 0, 0]))


def make_den_graph(num_states: int, num_pdfs: int, mean_out_degree: float = 16.0, seed: int = 5,
                   self_loops: bool = True) -> dict:
    """Phone-LM-like synthetic DenominatorGraph (SURVEY 8d cfg 5): out-degree ~ Poisson(mean),
    a self-loop on every state, arc pdf-ids uniform over P, row-normalised Uniform(0.05,1) transition
    probabilities, initial-probs by the 100-step averaging of kaldi chain-den-graph.cc SetInitialProbs."""
    g = rng(seed)
    N = num_states
    deg = np.maximum(g.poisson(mean_out_degree - (1 if self_loops else 0), size=N), 1).astype(np.int64)
    src_list, dst_list = [], []
    for h in range(N):
        d = g.integers(0, N, size=deg[h])
        if self_loops:
            d = np.concatenate([[h], d])
        src_list.append(np.full(len(d), h, dtype=np.int64))
        dst_list.append(d)
    src = np.concatenate(src_list)
    dst = np.concatenate(dst_list)
    A = len(src)
    pdf = g.integers(0, num_pdfs, size=A).astype(np.int32)
    w = g.uniform(0.05, 1.0, size=A)
    rowsum = np.bincount(src, weights=w, minlength=N)
    prob = (w / rowsum[src]).astype(np.float32)
    # forward list: grouped by source state (already sorted); backward list: grouped by destination
    fwd_begin = np.concatenate([[0], np.cumsum(np.bincount(src, minlength=N))])
    order_b = np.argsort(dst, kind="stable")
    bwd_begin = np.concatenate([[0], np.cumsum(np.bincount(dst, minlength=N))])
    t_prob = np.concatenate([prob, prob[order_b]]).astype(np.float32)
    t_pdf = np.concatenate([pdf, pdf[order_b]]).astype(np.int32)
    t_state = np.concatenate([dst, src[order_b]]).astype(np.int32)  # fwd: next state; bwd: previous state
    fwd_ranges = np.stack([fwd_begin[:-1], fwd_begin[1:]], axis=1).astype(np.int32)
    bwd_ranges = (np.stack([bwd_begin[:-1], bwd_begin[1:]], axis=1) + A).astype(np.int32)
    # initial probs: average of the state distribution over 100 steps from state 0, normalised
    cur = np.zeros(N)
    cur[0] = 1.0
    avg = np.zeros(N)
    for _ in range(100):
        avg += cur / 100.0
        cur = np.bincount(dst, weights=cur[src] * prob, minlength=N)
    init = (avg / avg.sum()).astype(np.float32)
    return dict(num_states=N, num_pdfs=num_pdfs, fwd_ranges=fwd_ranges, bwd_ranges=bwd_ranges, prob=t_prob,
                pdf=t_pdf, state=t_state, init=init, num_arcs=A)


def regular_row_offsets(time_offsets, start_t_in: int, start_t_out: int, num_images: int, t_step_in: int = 1,
                        t_step_out: int = 1):
    """row_stride / row_offsets of a regular (t-major, n fastest) grid; tdnn.cc:878-903."""
    n = t_step_out // t_step_in
    offs = []
    for off in time_offsets:
        input_t = (start_t_out + off - start_t_in) // t_step_in
        offs.append(n * (input_t // n) * num_images + input_t % n)
    return n, offs


def make_num_graphs(num_seqs: int, num_pdfs: int, frames: int, seed: int = 6, min_phones: int = 3, max_phones: int = 12,
                    den_graph: dict = None) -> dict:
    """Synthetic unconstrained numerator supervision: per sequence a left-to-right FST over a random "phone" string
    with the chain topology (one forward pdf consumed on entering a phone, then a self-loop pdf), plus an optional
    alternative pronunciation branch so that the FST is not a single path.  Arc lists are given twice (grouped by
    source = forward list, then by destination = backward list) as tdnnf_num_graph_create expects.

    den_graph (from make_den_graph, with self-loops): the phone string is a random WALK in the denominator graph
    (forward pdf = the pdf of the arc taken, self-loop pdf = the pdf of the destination's self-loop), so every
    numerator path is a denominator path as in real chain supervision.  With unrelated random pdfs the LF-MMI
    objective is unbounded above and a long training run on a fixed minibatch drives the outputs to infinity."""
    g = rng(seed)
    state_offsets = [0]
    src, dst, pdf, lp, final = [], [], [], [], []
    if den_graph is not None:
        d_rng, d_state, d_pdf = den_graph["fwd_ranges"], den_graph["state"], den_graph["pdf"]
    for _ in range(num_seqs):
        k = int(g.integers(min_phones, min(max_phones, frames) + 1))
        base = state_offsets[-1]
        ns = k + 1
        f = np.full(ns, -1.0e30, dtype=np.float32)
        f[k] = 0.0
        h = int(g.integers(0, den_graph["num_states"])) if den_graph is not None else 0
        for j in range(k):
            if den_graph is not None:
                a = int(g.integers(d_rng[h][0], d_rng[h][1]))          # an arc leaving den state h
                fwd_pdf, h = int(d_pdf[a]), int(d_state[a])
                loops = [b for b in range(d_rng[h][0], d_rng[h][1]) if d_state[b] == h]
                loop_pdf = int(d_pdf[loops[0]]) if loops else fwd_pdf
            else:
                fwd_pdf, loop_pdf = int(g.integers(0, num_pdfs)), int(g.integers(0, num_pdfs))
            src += [base + j, base + j + 1]
            dst += [base + j + 1, base + j + 1]
            pdf += [fwd_pdf, loop_pdf]
            lp += [float(np.log(0.5)), float(np.log(0.5))]
            if den_graph is None and g.uniform() < 0.3:  # alternative phone on the same transition
                src.append(base + j)
                dst.append(base + j + 1)
                pdf.append(int(g.integers(0, num_pdfs)))
                lp.append(float(np.log(0.25)))
        final.append(f)
        state_offsets.append(base + ns)
    src, dst = np.array(src), np.array(dst)
    pdf, lp = np.array(pdf, dtype=np.int32), np.array(lp, dtype=np.float32)
    N, A = state_offsets[-1], len(src)
    of = np.argsort(src, kind="stable")
    ob = np.argsort(dst, kind="stable")
    fb = np.concatenate([[0], np.cumsum(np.bincount(src, minlength=N))])
    bb = np.concatenate([[0], np.cumsum(np.bincount(dst, minlength=N))]) + A
    return dict(num_seqs=num_seqs, state_offsets=np.array(state_offsets, dtype=np.int32), num_arcs=A,
                fwd_ranges=np.stack([fb[:-1], fb[1:]], 1).astype(np.int32), bwd_ranges=np.stack([bb[:-1], bb[1:]], 1).astype(np.int32),
                arc_logprob=np.concatenate([lp[of], lp[ob]]), arc_pdf=np.concatenate([pdf[of], pdf[ob]]),
                arc_state=np.concatenate([dst[of], src[ob]]).astype(np.int32), final_logprob=np.concatenate(final))


def den_graph_to_fst_text(graph: dict, start_first: bool = True) -> str:
    """A DenominatorGraph as `fstprint den.fst` would show its FST: "src dst ilabel olabel weight" with ilabel = olabel =
    pdf-id + 1 and weight = -log transition probability, every state final with weight 0 (chain den FSTs are)."""
    A = graph["num_arcs"]
    lines = []
    for h in range(graph["num_states"]):
        b, e = graph["fwd_ranges"][h]
        for a in range(b, e):
            lines.append(f"{h} {int(graph['state'][a])} {int(graph['pdf'][a]) + 1} {int(graph['pdf'][a]) + 1} {-np.log(float(graph['prob'][a])):.9g}")
    assert len(lines) == A
    return "\n".join(lines) + "\n"


def num_graphs_to_fst_texts(graph: dict):
    """The per-sequence numerator FSTs of make_num_graphs as FSM texts (local state numbers, start state 0)."""
    texts = []
    offs, A = graph["state_offsets"], graph["num_arcs"]
    for s in range(graph["num_seqs"]):
        lines = []
        for st in range(offs[s], offs[s + 1]):
            b, e = graph["fwd_ranges"][st]
            for a in range(b, e):
                lines.append(f"{st - offs[s]} {int(graph['arc_state'][a]) - offs[s]} {int(graph['arc_pdf'][a]) + 1} "
                             f"{int(graph['arc_pdf'][a]) + 1} {-float(graph['arc_logprob'][a]):.9g}")
        for st in range(offs[s], offs[s + 1]):
            if graph["final_logprob"][st] > -1.0e29:
                lines.append(f"{st - offs[s]} {-float(graph['final_logprob'][st]):.9g}")
        texts.append("\n".join(lines) + "\n")
    return texts
