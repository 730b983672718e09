"""GPU test of the training-example reader feeding the chain objective: a binary archive of
single-sequence unconstrained examples and a binary den.fst go through tdnnf_chain_egs_* / tdnnf_den_graph_parse_fst_binary
and the objective computed from them equals the oracle's on the generators' own arrays."""
import numpy as np
import pytest

from tests.util import rel_err

pytestmark = pytest.mark.gpu


def test_objective_from_an_archive_and_a_binary_den_fst(ctx):
    import torch

    from oracle import oracle as O
    from tdnnf_nas_b200 import capi, chain, synth
    from tests import egs_ref as W
    from tests.test_egs_io import synth_supervision_examples

    S, P, T, N = 8, 40, 10, 150
    rng = np.random.default_rng(5)
    ngraph, exs = synth_supervision_examples(rng, S, P, T, seed=3)
    dgraph = synth.make_den_graph(N, P, 5.0, seed=4)
    fst = dict(start=None, num_states=N, arcs=[], finals={})
    for line in synth.den_graph_to_fst_text(dgraph).splitlines():
        f = line.split()
        if len(f) >= 4:
            fst["arcs"].append((int(f[0]), int(f[1]), int(f[2]), float(f[4]) if len(f) > 4 else 0.0))
            if fst["start"] is None:
                fst["start"] = int(f[0])
        elif f:
            fst["finals"][int(f[0])] = float(f[1]) if len(f) > 1 else 0.0
    egs = capi.ChainEgs(W.ark(exs, True))
    m = egs.merge_supervision(0, S, "output", P)
    feats, t0 = egs.merge_input(0, S, "input")
    assert m["num_seqs"] == S and m["frames_per_seq"] == T and feats.shape[1] == S and t0 == -3
    dg = capi.DenGraph(ctx, capi.parse_den_fst_binary(W.fst_vector(fst), P))
    ng = capi.NumeratorGraph(ctx, m["num_graph"])
    x = rng.standard_normal((T * S, P)).astype(np.float32)
    den_lp, den_d, _ = O.den_forward_backward(dgraph, x, S, T, 0.1, deriv_weight=-1.0)
    num_lp, num_d, _ = O.num_forward_backward(ngraph, x, T, deriv_weight=1.0)
    obj = chain.ChainObjective(ctx, dg, ng, S, T, chain.ChainTrainingOptions())
    xd = torch.from_numpy(x).cuda()
    deriv = torch.zeros_like(xd)
    objf, _, weight = obj.compute(xd, deriv)
    assert weight == S * T
    assert objf == pytest.approx(num_lp - den_lp, rel=1e-4)
    assert rel_err(deriv.cpu().numpy(), num_d + den_d) < 1e-3
    obj.close(); ng.close(); dg.close()
