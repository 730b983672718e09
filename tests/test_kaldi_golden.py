"""The oracle against vectors dumped by the REFERENCE ITSELF (tools/kaldi_golden/dump-nas-golden.cc, run inside a Kaldi tree
patched with TDNN-F_NAS): the one route to pinned parity, for whoever has such a tree -- this repository's container has none,
so tests/golden/kaldi/ is empty and test_reference_dumps skips.  The second test keeps the kit itself honest: it writes a dump
directory in the harness's exact layout from the oracle's own outputs and runs the same checker over it (a check of the
loader and the comparison code, not of parity)."""
import os
import re
import sys

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "kaldi")
sys.path.insert(0, os.path.join(ROOT, "tools", "kaldi_golden"))
import make_inputs as MI  # noqa: E402

FWD_TOL, BWD_TOL = 1e-4, 1e-3   # north_star's tolerances: forward, derivatives / deltas


def read_kaldi_text(path) -> np.ndarray:
    """A Kaldi text matrix (' [\\n a b\\n c d ]') or vector (' [ a b c ]')."""
    body = open(path).read()
    inner = body[body.index("[") + 1:body.rindex("]")]
    rows = [r.split() for r in inner.replace(";", "\n").split("\n") if r.strip()]
    a = np.array([[float(v) for v in r] for r in rows], np.float32)
    return a.reshape(-1) if ("\n" not in inner.strip() and a.shape[0] <= 1) else a


def write_kaldi_text(path, a):
    a = np.asarray(a, np.float32)
    with open(path, "w") as f:
        if a.ndim == 1:
            f.write(" [ " + " ".join(repr(float(v)) for v in a) + " ]\n")
        else:
            f.write(" [" + "".join("\n  " + " ".join(repr(float(v)) for v in r) + " " for r in a) + "]\n")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def load_case(d):
    case = {}
    for line in open(os.path.join(d, "case.txt")):
        k, _, v = line.rstrip("\n").partition(" ")
        case[k] = v
    case["dir"] = d
    return case


def _kv(config):
    return dict(t.split("=", 1) for t in config.split())


def _truth(v, default):
    return default if v is None else v.lower() in ("true", "t", "1")


def _mat(d, name, k):
    m = read_kaldi_text(os.path.join(d, f"{name}.{k}.txt"))
    return m.reshape(1, -1) if m.ndim == 1 else m


def check_darts(case):
    d, kv = case["dir"], _kv(case["config"])
    offsets = [int(t) for t in kv["time-offsets"].split(",")]
    n, din, dout = len(offsets), int(kv["input-dim"]), int(kv["output-dim"])
    flags = ((O.USE_GUMBEL if _truth(kv.get("use-gumbel"), True) else 0) | (O.FREE_SELECT if _truth(kv.get("free-select"), True) else 0)
             | (O.UNIFORM_SAMPLE if _truth(kv.get("uniform-sample"), True) else 0) | (O.USE_ENTROPY if _truth(kv.get("use-entropy"), True) else 0)
             | (O.UPDATE_ALPHA if _truth(kv.get("update-alpha"), True) else 0))          # all-true C++ defaults (tdnn.cc:150-158)
    assert not flags & (O.USE_GUMBEL | O.UNIFORM_SAMPLE), "the harness dumps deterministic modes only"
    temp = float(kv.get("Temp-Proportion", 1.0))
    lr = float(kv.get("learning-rate", 0.001))
    params = read_kaldi_text(os.path.join(d, "params.txt"))
    assert params.size == dout * n * din + n + dout
    W, bias_params = params[: dout * n * din].reshape(dout, n * din).copy(), params[dout * n * din:].copy()
    idx = open(os.path.join(d, "indexes.txt")).read()
    row_stride = int(re.search(r"<RowStride>\s+(-?\d+)", idx).group(1))
    row_offsets = [int(t) for t in re.search(r"<RowOffsets>\s*\[([^\]]*)\]", idx).group(1).split()]
    assert len(row_offsets) == n
    ng_in = O.NaturalGradient(int(kv.get("rank-in", 20)), int(kv.get("update-period", 4)), float(kv.get("num-samples-history", 2000.0)),
                              float(kv.get("alpha-in", 4.0)))
    ng_out = O.NaturalGradient(int(kv.get("rank-out", 80)), int(kv.get("update-period", 4)), float(kv.get("num-samples-history", 2000.0)),
                               float(kv.get("alpha-out", 4.0)))
    worst = {}
    for k in range(int(case["steps"])):
        x, out_ref, od = _mat(d, "in", k), _mat(d, "out", k), _mat(d, "out_deriv", k)
        assert x.shape == (int(case["in_rows"]), din) and out_ref.shape == (int(case["out_rows"]), dout)
        out, coef = O.tdnn_propagate(offsets, flags, temp, W, bias_params, x, out_ref.shape[0], row_offsets, row_stride)
        worst["out"] = max(worst.get("out", 0.0), rel(out, out_ref))
        worst["memo"] = max(worst.get("memo", 0.0), rel(coef, read_kaldi_text(os.path.join(d, f"memo.{k}.txt"))))
        in_deriv = _mat(d, "in_deriv_before", k).copy()
        dW, db = np.zeros_like(W), np.zeros(n + dout, np.float32)
        O.tdnn_backprop(offsets, flags, temp, W, x, od, coef, row_offsets, row_stride, lr, in_deriv=in_deriv, dW=dW, dbias=db,
                        ng_in=ng_in, ng_out=ng_out)
        delta = read_kaldi_text(os.path.join(d, f"delta.{k}.txt"))
        worst["in_deriv"] = max(worst.get("in_deriv", 0.0), rel(in_deriv, _mat(d, "in_deriv", k)))
        worst["delta_theta"] = max(worst.get("delta_theta", 0.0), rel(dW, delta[: dout * n * din].reshape(dout, n * din)))
        worst["delta_bias"] = max(worst.get("delta_bias", 0.0), rel(db[n:], delta[dout * n * din + n:]))
        a_ref = delta[dout * n * din: dout * n * din + n]
        if np.abs(a_ref).max() > 0:
            worst["delta_alpha"] = max(worst.get("delta_alpha", 0.0), rel(db[:n], a_ref))
    assert worst["out"] < FWD_TOL and worst["memo"] < 1e-5, worst
    assert all(worst[k] < BWD_TOL for k in ("in_deriv", "delta_theta", "delta_bias")), worst
    assert worst.get("delta_alpha", 0.0) < 5e-3, worst       # sums of n products that nearly cancel: the bar of tests/test_gpu_ng.py
    return worst


def check_softmax_flops(case):
    d, kv = case["dir"], _kv(case["config"])
    gumbel = case["type"].startswith("Gumbel")
    assert not gumbel, "the harness dumps the noise-free component only"
    worst = {}
    for k in range(int(case["steps"])):
        x, od = _mat(d, "in", k), _mat(d, "out_deriv", k)
        out = O.softmax_flops_fwd(x)
        worst["out"] = max(worst.get("out", 0.0), rel(out, _mat(d, "out", k)))
        in_deriv, od_after = O.softmax_flops_bwd(out, od.copy(), float(kv["scale"]), False, 1.0)
        worst["out_deriv_after"] = max(worst.get("out_deriv_after", 0.0), rel(od_after, _mat(d, "out_deriv_after", k)))   # Q9: the penalty lands in out_deriv
        worst["in_deriv"] = max(worst.get("in_deriv", 0.0), rel(in_deriv, _mat(d, "in_deriv", k)))
    assert worst["out"] < FWD_TOL and worst["out_deriv_after"] < BWD_TOL and worst["in_deriv"] < BWD_TOL, worst
    return worst


def check_copyn(case):
    d, kv = case["dir"], _kv(case["config"])
    scale = float(kv.get("scale", 1.0))
    worst = {}
    for k in range(int(case["steps"])):
        out = _mat(d, "out_before", k).copy()
        O.copyn_fwd(_mat(d, "in", k), out, scale)
        worst["out"] = max(worst.get("out", 0.0), rel(out, _mat(d, "out", k)))
        ind = _mat(d, "in_deriv_before", k).copy()
        O.copyn_bwd(_mat(d, "out_deriv", k), ind, scale)
        worst["in_deriv"] = max(worst.get("in_deriv", 0.0), rel(ind, _mat(d, "in_deriv", k)))
    assert worst["out"] < FWD_TOL and worst["in_deriv"] < BWD_TOL, worst
    return worst


def check_denominator(case):
    d = case["dir"]
    name = os.path.basename(d)
    graph = MI.den_graph(name)
    S, T = int(case["num_sequences"]), int(case["frames"])
    assert int(case["num_states"]) == graph["num_states"] and int(case["num_pdfs"]) == graph["num_pdfs"]
    # DenominatorGraph::SetInitialProbs of the reference's Kaldi on the binary den.fst written by tests/egs_ref.py
    assert rel(graph["init"], read_kaldi_text(os.path.join(d, "initial_probs.txt"))) < 1e-4
    x = _mat(d, "nnet_output", 0)
    lp, deriv, ok = O.den_forward_backward(graph, x, S, T, float(case["leaky"]), deriv_weight=-1.0)
    assert ok == bool(int(case["ok"]))
    assert abs(lp - float(case["logprob"])) <= FWD_TOL * abs(float(case["logprob"]))
    w = dict(deriv=rel(deriv, _mat(d, "deriv", 0)))
    assert w["deriv"] < BWD_TOL, w
    return w


CHECKERS = {"TdnnDARTSV3Component": check_darts, "SoftmaxFlopsComponent": check_softmax_flops, "CopyNComponent": check_copyn,
            "Denominator": check_denominator}


def _dump_dirs():
    return sorted(os.path.join(GOLD, n) for n in os.listdir(GOLD) if os.path.exists(os.path.join(GOLD, n, "case.txt"))) if os.path.isdir(GOLD) else []


@pytest.mark.skipif(not _dump_dirs(), reason="no vectors dumped by a Kaldi + TDNN-F_NAS build under tests/golden/kaldi (see tools/kaldi_golden/README.md)")
@pytest.mark.parametrize("d", _dump_dirs() or ["-"])
def test_reference_dumps(d):
    case = load_case(d)
    print(os.path.basename(d), CHECKERS[case["type"]](case))


def _self_dump(root):
    """What dump-nas-golden.cc writes, from the oracle's own outputs (same file names, same text form)."""
    g = np.random.default_rng(0)
    # -- TdnnDARTSV3Component, 3 steps
    offsets, din, dout, S = [0, 1, 2], 12, 10, 4
    n = len(offsets)
    t_out = list(range(0, 8))
    t_in = list(range(0, 8 + 2))
    d = os.path.join(root, "darts_self")
    os.makedirs(d)
    cfg = (f"input-dim={din} output-dim={dout} time-offsets={','.join(map(str, offsets))} use-bias=true learning-rate=0.02 rank-in=5 rank-out=4 "
           "update-alpha=true update-theta=true use-gumbel=false use-entropy=false free-select=false uniform-sample=false")
    flags = O.UPDATE_ALPHA
    W = (g.standard_normal((dout, n * din)) * 0.2).astype(np.float32)
    bp = np.concatenate([g.standard_normal(n), g.standard_normal(dout)]).astype(np.float32)
    write_kaldi_text(os.path.join(d, "params.txt"), np.concatenate([W.reshape(-1), bp]))
    row_stride, row_offsets = 1, [o * S for o in offsets]
    open(os.path.join(d, "indexes.txt"), "w").write(f"<TdnnDARTSV3ComponentPrecomputedIndexes> <RowStride> {row_stride} <RowOffsets> [ "
                                                    + " ".join(map(str, row_offsets)) + " ]\n</TdnnDARTSV3ComponentPrecomputedIndexes> ")
    steps = 3
    open(os.path.join(d, "case.txt"), "w").write(f"type TdnnDARTSV3Component\nconfig {cfg}\nnum_sequences {S}\nin_rows {len(t_in) * S}\n"
                                                 f"out_rows {len(t_out) * S}\nsteps {steps}\nproperties 0\ninfo self\n")
    ng_in, ng_out = O.NaturalGradient(5, 4, 2000.0, 4.0), O.NaturalGradient(4, 4, 2000.0, 4.0)
    for k in range(steps):
        x = g.standard_normal((len(t_in) * S, din)).astype(np.float32)
        od = (g.standard_normal((len(t_out) * S, dout)) / (len(t_out) * S)).astype(np.float32)
        out, coef = O.tdnn_propagate(offsets, flags, 1.0, W, bp, x, len(t_out) * S, row_offsets, row_stride)
        ind0 = (g.standard_normal(x.shape) * 0.1).astype(np.float32)
        ind, dW, db = ind0.copy(), np.zeros_like(W), np.zeros(n + dout, np.float32)
        O.tdnn_backprop(offsets, flags, 1.0, W, x, od, coef, row_offsets, row_stride, 0.02, in_deriv=ind, dW=dW, dbias=db, ng_in=ng_in, ng_out=ng_out)
        for name, a in (("in", x), ("out", out), ("out_before", np.zeros_like(out)), ("out_deriv", od), ("out_deriv_after", od), ("in_deriv_before", ind0),
                        ("in_deriv", ind)):
            write_kaldi_text(os.path.join(d, f"{name}.{k}.txt"), a)
        write_kaldi_text(os.path.join(d, f"memo.{k}.txt"), coef)
        write_kaldi_text(os.path.join(d, f"delta.{k}.txt"), np.concatenate([dW.reshape(-1), db]))
    # -- SoftmaxFlopsComponent and CopyNComponent, one step each
    R = 24
    d = os.path.join(root, "softmax_self")
    os.makedirs(d)
    open(os.path.join(d, "case.txt"), "w").write(f"type SoftmaxFlopsComponent\nconfig dim=8 scale=0.1\nnum_sequences 4\nin_rows {R}\nout_rows {R}\nsteps 1\n")
    x = g.standard_normal((R, 8)).astype(np.float32)
    od = (g.standard_normal((R, 8)) / R).astype(np.float32)
    out = O.softmax_flops_fwd(x)
    ind, od_after = O.softmax_flops_bwd(out, od.copy(), 0.1, False, 1.0)
    for name, a in (("in", x), ("out", out), ("out_deriv", od), ("out_deriv_after", od_after), ("in_deriv", ind)):
        write_kaldi_text(os.path.join(d, f"{name}.0.txt"), a)
    d = os.path.join(root, "copyn_self")
    os.makedirs(d)
    open(os.path.join(d, "case.txt"), "w").write(f"type CopyNComponent\nconfig input-dim=1 output-dim=5 scale=0.5\nnum_sequences 4\nin_rows {R}\nout_rows {R}\nsteps 1\n")
    x, out0 = g.standard_normal((R, 1)).astype(np.float32), g.standard_normal((R, 5)).astype(np.float32)
    od, ind0 = g.standard_normal((R, 5)).astype(np.float32), g.standard_normal((R, 1)).astype(np.float32)
    out, ind = out0.copy(), ind0.copy()
    O.copyn_fwd(x, out, 0.5)
    O.copyn_bwd(od, ind, 0.5)
    for name, a in (("in", x), ("out_before", out0), ("out", out), ("out_deriv", od), ("in_deriv_before", ind0), ("in_deriv", ind)):
        write_kaldi_text(os.path.join(d, f"{name}.0.txt"), a)
    # -- denominator
    name = "den_small"
    n_states, P, _, _, S, T = MI.DEN_CASES[name]
    graph = MI.den_graph(name)
    d = os.path.join(root, name)
    os.makedirs(d)
    x = np.clip(g.standard_normal((T * S, P)) * 2, -30, 30).astype(np.float32)
    lp, deriv, ok = O.den_forward_backward(graph, x, S, T, 0.1, deriv_weight=-1.0)
    write_kaldi_text(os.path.join(d, "nnet_output.0.txt"), x)
    write_kaldi_text(os.path.join(d, "deriv.0.txt"), deriv)
    write_kaldi_text(os.path.join(d, "initial_probs.txt"), graph["init"])
    open(os.path.join(d, "case.txt"), "w").write(f"type Denominator\nnum_pdfs {P}\nnum_sequences {S}\nframes {T}\nleaky 0.1\nnum_states {n_states}\n"
                                                 f"logprob {lp!r}\nok {int(ok)}\n")


def test_the_kit_checks_what_it_says(tmp_path):
    _self_dump(str(tmp_path))
    seen = set()
    for name in sorted(os.listdir(tmp_path)):
        case = load_case(os.path.join(tmp_path, name))
        worst = CHECKERS[case["type"]](case)
        seen.add(case["type"])
        assert all(v < 1e-5 for v in worst.values()), (name, worst)     # the oracle against itself through the text files
    assert seen == set(CHECKERS)
    # and a wrong vector is caught: a perturbed forward output fails the forward bar
    d = os.path.join(tmp_path, "softmax_self")
    out = read_kaldi_text(os.path.join(d, "out.0.txt"))
    out[3, 2] *= 1.01
    write_kaldi_text(os.path.join(d, "out.0.txt"), out)
    with pytest.raises(AssertionError):
        check_softmax_flops(load_case(d))
