"""End-to-end parity of ONE WHOLE TRAINING STEP of the search-stage supernet: the B200 path (tdnnf_nas_b200.supernet, every
kernel ours) against the CPU reference composed from the oracle's restatements of the reference methods
(oracle/supernet_ref.py), on the same parameters, the same synthetic egs, the same den / numerator graphs and the same
Gumbel draws.  Bars (BASELINE.json north_star): nnet output and LF-MMI objective 1e-4, derivatives and parameter /
architecture-weight deltas 1e-3."""
import numpy as np
import pytest

from tests.util import rel_err

pytestmark = pytest.mark.gpu


def _bn_scale_offset(comp):
    """scale_ / offset_ of a BatchNormTestComponent from its text form (norm.cc:680-713, 956-982)."""
    toks = comp.write(False).decode().split()
    get = lambda name: float(toks[toks.index(name) + 1])
    vec = lambda name: np.array(toks[toks.index(name) + 2: toks.index("]", toks.index(name))], dtype=np.float64)
    mean, var = vec("<StatsMean>"), vec("<StatsVar>")
    scale = get("<TargetRms>") * (np.maximum(var, 0.0) + get("<Epsilon>")) ** -0.5
    return scale.astype(np.float32), (-mean * scale).astype(np.float32)


def _params_of(net):
    cfg = net.cfg
    n, D, B = cfg.num_offsets, cfg.dim, cfg.bottleneck
    p = {}
    for k, v in net.stock.items():
        p[f"{k}.W"] = v["W"].cpu().numpy().copy()
        if v["b"] is not None:
            p[f"{k}.b"] = v["b"].cpu().numpy().copy()
    bns = {"tdnn1": net.t1["bn"], "pc1": net.head["bn1"], "pc2": net.head["bn2"]}
    if cfg.xent:
        bns.update(px1=net.head["xbn1"], px2=net.head["xbn2"])
    for b, blk in enumerate(net.blocks):
        bns[f"blk{b}"] = blk["bn"]
        for h, din, dout in (("lin", D, B), ("aff", B, D)):
            v = blk[h].vectorize()
            p[f"blk{b}.{h}.W"] = v[: dout * n * din].reshape(dout, n * din).copy()
            p[f"blk{b}.{h}.bias"] = v[dout * n * din:].copy()
    for nm, comp in bns.items():
        p[f"bn.{nm}.scale"], p[f"bn.{nm}.offset"] = _bn_scale_offset(comp)
    return p


@pytest.mark.parametrize("xent,planes", [(True, True), (False, True), (True, False)])
def test_whole_step_matches_cpu_reference(xent, planes):
    """planes: the fused tails write the operand planes of what they produce and Propagate keeps its input planes for
    Backprop (the default); False: every GEMM call splits its operands itself."""
    import torch

    from oracle import supernet_ref as R
    from tdnnf_nas_b200 import nnet3, synth
    from tdnnf_nas_b200.supernet import Supernet, SupernetConfig

    # the bottleneck keeps its real width: rank-out 80 of OnlineNaturalGradient must stay well below the dimension it
    # preconditions (with 32 columns the rank is clipped to 31 and X_hat is the ill-conditioned remainder of a near-total
    # projection: rounding differences of 1e-6 in X come out as 1e-2 in the delta)
    cfg = SupernetConfig(num_seqs=8, frames_per_eg=30, dim=256, bottleneck=160, num_blocks=3, prefinal_small=64, num_pdfs=200,
                         den_states=300, den_out_degree=6.0, mode="search", learning_rate=2e-3, darts_lr_factor=0.05, xent=xent,
                         tail_planes=planes)
    nnet3.set_keep_planes(planes)
    net = Supernet(cfg)
    S, T, P, L, n = cfg.num_seqs, net.T, cfg.num_pdfs, cfg.num_blocks, cfg.num_offsets
    den_graph = synth.make_den_graph(cfg.den_states, P, cfg.den_out_degree, seed=5)
    num_graph = synth.make_num_graphs(S, P, T, seed=60, den_graph=den_graph)
    rcfg = R.RefConfig(num_seqs=S, frames_per_eg=cfg.frames_per_eg, feat_dim=cfg.feat_dim, dim=cfg.dim, bottleneck=cfg.bottleneck,
                       num_blocks=L, num_offsets=n, prefinal_small=cfg.prefinal_small, num_pdfs=P, xent=xent,
                       learning_rate=cfg.learning_rate, darts_lr_factor=cfg.darts_lr_factor)
    ref = R.CpuSupernet(rcfg, den_graph, num_graph, _params_of(net))
    assert R.frame_plan(rcfg)[2] == net._frames()[2] and R.frame_plan(rcfg)[4] == net._frames()[4]  # same frames, derived twice
    for step in range(2):  # the second step runs on updated parameters and on preconditioners that have seen a minibatch
        x = net.make_input(step)
        c0 = nnet3.get_rand_counter()
        objf_gpu = net.step(x.pin_memory(), apply_update=False)
        c1 = nnet3.get_rand_counter()
        nnet3.set_rand_counter(c0)
        u = [np.array([nnet3.rand_uniform() for _ in range(n)], np.float32) for _ in range(2 * L)]
        assert nnet3.get_rand_counter() == c1  # exactly the draws the 2 L components made
        # The derivative of ReLU jumps at 0.  With ~1e5 pre-activations per layer some lie within 1e-5 rms of zero, i.e.
        # within the forward tolerance: the GPU (whose split-K reductions even reorder sums from run to run) and the CPU
        # may then disagree on ONE mask element, and everything below that layer moves by ~1e-3.  So the reference uses
        # the GPU's masks, after a check that they differ from its own only where its pre-activation is within the forward
        # tolerance of zero (and in a handful of places at most).
        def gpu_masks(r):
            g = {"pc": net.head["pr"], **({"px": net.head["xr"]} if xent else {}), **{b: blk["aff_out"] for b, blk in enumerate(net.blocks)}}
            masks = {}
            for key, own in r.relu_inputs().items():
                masks[key] = (g[key] > 0).cpu().numpy()
                diff = masks[key] != (own > 0)
                assert diff.sum() <= 4, (key, int(diff.sum()))
                if diff.any():
                    assert np.abs(own[diff]).max() < 1e-4 * np.sqrt((own.astype(np.float64) ** 2).mean()), key
            return masks

        objf_ref = ref.step(x.numpy(), u, apply_update=False, relu_masks=gpu_masks)
        assert rel_err(net.head["out"].cpu().numpy(), ref.st["out"]) < 1e-4
        assert abs(objf_gpu - objf_ref) <= 1e-4 * abs(objf_ref), (objf_gpu, objf_ref)
        assert rel_err(net.head["d_out"].cpu().numpy(), ref.st["d_out"]) < 1e-3
        if xent:
            assert abs(net.last_xent_objf - ref.xent_objf) <= 1e-4 * abs(ref.xent_objf)
        errs = {}
        for b, blk in enumerate(net.blocks):
            errs[(b, "d_aff")] = rel_err(blk["d_aff"].cpu().numpy(), ref.st[b]["d_aff"])
            errs[(b, "d_lin")] = rel_err(blk["d_lin"].cpu().numpy(), ref.st[b]["d_lin"])
            for h in ("lin", "aff"):
                dv = blk[h + "_delta"].vectorize()
                dW, db = ref.delta[(b, h)]
                errs[(b, h, "theta")] = rel_err(dv[: dW.size].reshape(dW.shape), dW)
                errs[(b, h, "bias")] = rel_err(dv[dW.size + n:], db[n:])
                errs[(b, h, "alpha")] = float(np.abs(dv[dW.size: dW.size + n] - db[:n]).max() / (np.abs(db[:n]).max() + 1e-30))
        bad = {k: v for k, v in errs.items() if not v < 1e-3}
        assert not bad, (step, bad, errs)
        # the parameter step (max-change) on both sides
        assert net._update_with_max_change() and ref.update()
        np.testing.assert_allclose(net.last_max_change_factors, ref.last_factors, rtol=1e-3)
        for b, blk in enumerate(net.blocks):
            v = blk["lin"].vectorize()
            W = ref.p[f"blk{b}.lin.W"]
            assert rel_err(v[: W.size].reshape(W.shape), W) < 1e-5
    h0, m0 = net.ctx.operand_cache_stats()
    net.close()
    nnet3.set_keep_planes(True)
    assert h0 > 0
