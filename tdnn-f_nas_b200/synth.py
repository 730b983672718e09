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
