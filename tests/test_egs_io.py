"""Host-only tests of the training-example reader (tdnnf_chain_egs_*, tdnnf_den_graph_parse_fst_binary; csrc/egs_io.cc)
against tests/egs_ref.py, an independent Python writer of the same Kaldi / OpenFst formats.  Reader and writer come from the
same reading of the formats (Kaldi is not available here), so these pin self-consistency and robustness, not Kaldi."""
import struct

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from tdnnf_nas_b200 import capi, synth

from . import egs_ref as W


def _check_example(got, want):
    assert got["key"] == want["key"]
    assert [io["name"] for io in got["inputs"]] == [io["name"] for io in want["inputs"]]
    for g, w in zip(got["inputs"], want["inputs"]):
        np.testing.assert_array_equal(g["indexes"], np.asarray(w["indexes"], np.int32).reshape(-1, 3))
        np.testing.assert_array_equal(g["data"], np.asarray(w["data"], np.float32))
    assert len(got["outputs"]) == len(want["outputs"])
    for g, w in zip(got["outputs"], want["outputs"]):
        for k in ("name", "num_sequences", "frames_per_seq", "label_dim", "e2e"):
            assert g[k] == w[k], k
        assert g["weight"] == pytest.approx(w["weight"])
        np.testing.assert_array_equal(g["indexes"], np.asarray(w["indexes"], np.int32).reshape(-1, 3))
        np.testing.assert_array_equal(g["alignment_pdfs"], np.asarray(w.get("alignment_pdfs", []), np.int32))
        np.testing.assert_allclose(g["deriv_weights"], np.asarray(w["deriv_weights"], np.float32), rtol=0, atol=1e-7)
        assert len(g["fsts"]) == len(w["fsts"])
        for gf, wf in zip(g["fsts"], w["fsts"]):
            assert gf["start"] == wf["start"] and gf["num_states"] == wf["num_states"]
            # both on-disk forms list arcs state by state (the start state's first in text); compare as sorted multisets
            ga = sorted((int(a), int(b), int(l), float(x)) for (a, b, l), x in zip(gf["arcs"], gf["weights"]))
            wa = sorted((a, b, l, float(np.float32(x))) for (a, b, l, x) in wf["arcs"])
            assert ga == wa
            assert dict(zip(gf["final_states"].tolist(), gf["final_weights"].tolist())) == {s: float(np.float32(x)) for s, x in wf["finals"].items()}


@pytest.mark.parametrize("binary", [True, False])
def test_archive_round_trip(binary):
    rng = np.random.default_rng(3)
    exs = [W.random_example(rng, "utt1-0", e2e=True),
           W.random_example(rng, "utt1-1", e2e=False, ivector_dim=0, dw2=True),
           W.random_example(rng, "spk2_utt7-33", e2e=True, num_sequences=3, deriv_weights=False)]
    egs = capi.ChainEgs(W.ark(exs, binary))
    assert len(egs) == 3
    for i, want in enumerate(exs):
        got = egs.example(i)
        assert got["binary"] == binary
        _check_example(got, want)
    assert len(capi.ChainEgs(W.ark(exs, binary), max_examples=2)) == 2
    egs.close()


def test_files_without_the_newer_tokens_are_read():
    # egs written before the unconstrained supervision existed have neither <End2End> nor <AlignmentPdfs>
    rng = np.random.default_rng(5)
    ex = W.random_example(rng, "old", e2e=False)
    ex["outputs"][0]["write_e2e_flag"] = False
    ex["outputs"][0]["write_alignment_pdfs"] = False
    for binary in (True, False):
        _check_example(capi.ChainEgs(W.ark([ex], binary)).example(0), ex)
    ex["outputs"][0]["write_alignment_pdfs"] = True
    ex["outputs"][0]["alignment_pdfs"] = list(range(7))
    for binary in (True, False):
        _check_example(capi.ChainEgs(W.ark([ex], binary)).example(0), ex)


@pytest.mark.parametrize("coding", ["cm1", "cm2", "cm3", "double", "sparse"])
def test_matrix_codings(coding):
    rng = np.random.default_rng(11)
    ex = W.random_example(rng, "k", frames=9, dim=7, coding=coding)
    m = ex["inputs"][0]["data"]
    if coding == "sparse":
        m[rng.random(m.shape) < 0.7] = 0.0
    got = capi.ChainEgs(W.ark([ex], True)).example(0)["inputs"][0]["data"]
    if coding.startswith("cm"):
        blob = W.compressed_matrix(m, int(coding[2:]))
        np.testing.assert_allclose(got, W.decode_compressed(blob), rtol=0, atol=2e-6)
        # and the decoding is the inverse of the coding up to the quantisation step (a wrong byte order would not be)
        span = float(m.max() - m.min())
        assert np.abs(got - m).max() <= span * {"cm1": 0.02, "cm2": 1.0 / 65535, "cm3": 0.6 / 255}[coding] + 1e-6
    else:
        np.testing.assert_array_equal(got, m)
    if coding == "sparse":   # the text form of a sparse matrix
        np.testing.assert_array_equal(capi.ChainEgs(W.ark([ex], False)).example(0)["inputs"][0]["data"], m)


def test_index_vector_coding_edges():
    # one-byte steps up to 124, the 127 escape for larger steps, n changes, a first element that is not (0, small t, 0)
    idx = [(2, -400, 0), (2, -276, 0), (2, -152, 0), (2, -28, 0), (2, 96, 0), (3, 96, 0), (3, -28, 0), (3, -29, 0), (3, 95, 0), (3, 220, 0), (0, 0, 1)]
    rng = np.random.default_rng(0)
    ex = W.random_example(rng, "k", ivector_dim=0)
    ex["inputs"][0]["indexes"] = idx
    ex["inputs"][0]["data"] = rng.standard_normal((len(idx), 3)).astype(np.float32)
    blob = W.ark([ex], True)
    assert W.index_vector(idx[:2], True).count(b"\x7f") == 1       # the 124-step is one byte, the first element escapes
    np.testing.assert_array_equal(capi.ChainEgs(blob).example(0)["inputs"][0]["indexes"], np.asarray(idx, np.int32))
    np.testing.assert_array_equal(capi.ChainEgs(W.ark([ex], False)).example(0)["inputs"][0]["indexes"], np.asarray(idx, np.int32))


@pytest.mark.parametrize("binary", [True, False])
def test_merge_is_what_merge_chain_examples_forms(binary):
    rng = np.random.default_rng(7)
    num_pdfs = 13
    exs = [W.random_example(rng, f"u{i}", num_pdfs=num_pdfs, num_sequences=ns, deriv_weights=(i != 2))
           for i, ns in enumerate([1, 2, 1, 1])]
    # rows of an example in any order: shuffle one input
    perm = rng.permutation(len(exs[1]["inputs"][0]["indexes"]))
    exs[1]["inputs"][0]["indexes"] = [exs[1]["inputs"][0]["indexes"][k] for k in perm]
    exs[1]["inputs"][0]["data"] = exs[1]["inputs"][0]["data"][perm]
    egs = capi.ChainEgs(W.ark(exs, binary))
    first, count = 1, 3
    x, t0 = egs.merge_input(first, count, "input")
    S = 4
    assert t0 == -3 and x.shape == (6 * 3 + 3 + 2, S, 5)
    seq = 0
    for ex in exs[first:first + count]:
        io = ex["inputs"][0]
        for n in range(ex["outputs"][0]["num_sequences"]):
            rows = sorted((t, r) for r, (nn, t, _) in enumerate(io["indexes"]) if nn == n)
            np.testing.assert_array_equal(x[:, seq], io["data"][[r for _, r in rows]])
            seq += 1
    iv, _ = egs.merge_input(first, count, "ivector")
    assert iv.shape == (1, S, 4)
    np.testing.assert_array_equal(iv[0], np.concatenate([ex["inputs"][1]["data"] for ex in exs[first:first + count]]))
    m = egs.merge_supervision(first, count, "output", num_pdfs)
    assert m["num_seqs"] == S and m["frames_per_seq"] == 6 and m["weight"] == pytest.approx(1.0)
    want_dw = np.concatenate([np.asarray(ex["outputs"][0]["deriv_weights"], np.float32).reshape(6, -1) if len(ex["outputs"][0]["deriv_weights"])
                              else np.ones((6, ex["outputs"][0]["num_sequences"]), np.float32) for ex in exs[first:first + count]], axis=1)
    np.testing.assert_allclose(m["deriv_weights"], want_dw, atol=1e-7)
    # the numerator graph equals the one the FSM-text reader builds from the same FSTs, in sequence order
    texts = [W.fst_text_lines(f) for ex in exs[first:first + count] for f in ex["outputs"][0]["fsts"]]
    want = capi.parse_num_fst_texts(texts, num_pdfs)
    for k in ("num_seqs", "num_arcs"):
        assert m["num_graph"][k] == want[k]
    for k in ("state_offsets", "fwd_ranges", "bwd_ranges", "arc_pdf", "arc_state", "arc_logprob", "final_logprob"):
        np.testing.assert_array_equal(m["num_graph"][k], want[k])


def test_merge_refusals():
    rng = np.random.default_rng(8)
    a = W.random_example(rng, "a", frames=6)
    b = W.random_example(rng, "b", frames=5)
    egs = capi.ChainEgs(W.ark([a, b], True))
    with pytest.raises(capi.TdnnfError, match="number of frames"):
        egs.merge_input(0, 2, "input")
    with pytest.raises(capi.TdnnfError, match="frames per sequence"):
        egs.merge_supervision(0, 2, "output", 11)
    with pytest.raises(capi.TdnnfError, match="no input named"):
        egs.merge_input(0, 1, "mfcc")
    with pytest.raises(capi.TdnnfError, match="label dimension"):
        egs.merge_supervision(0, 1, "output", 12)
    with pytest.raises(capi.TdnnfError, match="range"):
        egs.merge_input(1, 2, "input")
    c = W.random_example(rng, "c", e2e=False, num_sequences=2)     # a merged constrained supervision: one FST over two sequences
    egs = capi.ChainEgs(W.ark([c], True))
    assert egs.merge_supervision(0, 1, "output", 11, graph=False)["num_seqs"] == 2
    with pytest.raises(capi.TdnnfError, match="constrained"):
        egs.merge_supervision(0, 1, "output", 11)
    d = W.random_example(rng, "d")
    d["outputs"][0]["fsts"][0]["arcs"][0] = (0, 1, 12, 0.0)        # ilabel beyond num_pdfs
    with pytest.raises(capi.TdnnfError, match="pdf-id"):
        capi.ChainEgs(W.ark([d], True)).merge_supervision(0, 1, "output", 11)


def test_den_fst_binary_equals_text():
    g = synth.make_den_graph(90, 29, 4.0, seed=4)
    text = synth.den_graph_to_fst_text(g)
    want = capi.parse_den_fst_text(text, 29)
    fst = dict(start=None, num_states=90, arcs=[], finals={})
    for line in text.splitlines():
        f = line.split()
        if len(f) >= 4:
            fst["arcs"].append((int(f[0]), int(f[1]), int(f[2]), float(f[4]) if len(f) > 4 else 0.0))
            if fst["start"] is None:
                fst["start"] = int(f[0])
        elif f:
            fst["finals"][int(f[0])] = float(f[1]) if len(f) > 1 else 0.0
    for writer in (W.fst_vector, W.fst_compact_acceptor):
        got = capi.parse_den_fst_binary(writer(fst), 29)
        for k in ("num_states", "num_pdfs", "num_arcs"):
            assert got[k] == want[k]
        for k in ("fwd_ranges", "bwd_ranges", "pdf", "state"):
            np.testing.assert_array_equal(got[k], want[k])
        np.testing.assert_allclose(got["prob"], want["prob"], rtol=1e-6)
        np.testing.assert_allclose(got["init"], want["init"], rtol=1e-5, atol=1e-9)


def test_malformed_archives_are_errors_not_crashes():
    rng = np.random.default_rng(13)
    ex = W.random_example(rng, "key", coding="cm1")
    good = W.ark([ex], True)
    with pytest.raises(capi.TdnnfError, match="example 'key'"):
        capi.ChainEgs(good.replace(b"<NumInputs>", b"<NumInputz>"))
    with pytest.raises(capi.TdnnfError, match="magic"):
        capi.ChainEgs(good.replace(struct.pack("<i", W.FST_MAGIC), struct.pack("<i", 12345)))
    with pytest.raises(capi.TdnnfError, match="arc type"):
        capi.ChainEgs(good.replace(b"standard", b"log\0\0\0\0\0"))
    # a huge count in front of little data is refused before anything is allocated
    k = good.index(b"<I1V> ") + 6
    with pytest.raises(capi.TdnnfError, match="does not fit"):
        capi.ChainEgs(good[:k] + b"\x04" + struct.pack("<i", 2**31 - 1) + good[k + 5:])
    # every proper prefix of a binary and of a text archive is an error (the closing token is missing) and nothing crashes
    for blob in (good, W.ark([W.random_example(rng, "key")], False)):
        end = len(blob.rstrip())
        for cut in list(range(0, end, 7)) + [end - 1]:
            if cut < 4:
                continue    # an empty archive, or white space only, is a valid archive of no examples
            with pytest.raises(capi.TdnnfError):
                capi.ChainEgs(blob[:cut])
    assert len(capi.ChainEgs(b"")) == 0 and len(capi.ChainEgs(b" \n")) == 0


@settings(max_examples=150, deadline=None)
@given(st.data())
def test_corrupted_bytes_never_crash(data):
    rng = np.random.default_rng(data.draw(st.integers(0, 2**31)))
    binary = data.draw(st.booleans())
    blob = bytearray(W.ark([W.random_example(rng, "k", coding=data.draw(st.sampled_from(["full", "cm1", "cm2", "sparse"])),
                                             e2e=data.draw(st.booleans()), num_sequences=data.draw(st.integers(1, 2)))], binary))
    for _ in range(data.draw(st.integers(1, 6))):
        pos = data.draw(st.integers(0, len(blob) - 1))
        blob[pos] = data.draw(st.integers(0, 255))
    try:
        egs = capi.ChainEgs(bytes(blob))
    except capi.TdnnfError:
        return
    for i in range(len(egs)):       # whatever was accepted is internally consistent enough to walk and to merge
        ex = egs.example(i)
        for io in ex["inputs"]:
            assert io["indexes"].shape[0] == io["data"].shape[0]
        try:
            egs.merge_input(i, 1, ex["inputs"][0]["name"])
            if ex["outputs"]:
                egs.merge_supervision(i, 1, ex["outputs"][0]["name"], ex["outputs"][0]["label_dim"])
        except capi.TdnnfError:
            pass


def synth_supervision_examples(rng, S, P, T, seed):
    """One single-sequence unconstrained example per sequence of synth.make_num_graphs (the recipes' egs before merging),
    plus the graph they came from."""
    graph = synth.make_num_graphs(S, P, T, seed=seed)
    exs = []
    for s, text in enumerate(synth.num_graphs_to_fst_texts(graph)):
        fst = dict(start=0, num_states=int(graph["state_offsets"][s + 1] - graph["state_offsets"][s]), arcs=[], finals={})
        for line in text.splitlines():
            f = line.split()
            if len(f) >= 4:
                fst["arcs"].append((int(f[0]), int(f[1]), int(f[2]), float(f[4])))
            else:
                fst["finals"][int(f[0])] = float(f[1])
        ex = W.random_example(rng, f"utt{s}", frames=T, num_pdfs=P, deriv_weights=False)
        ex["outputs"][0]["fsts"] = [fst]
        exs.append(ex)
    return graph, exs


@pytest.mark.parametrize("binary", [True, False])
def test_merged_supervision_through_the_oracle_numerator(binary):
    """archive -> merged numerator graph -> the oracle's GenericNumeratorComputation gives what it gives on the generator's
    own arrays (log-prob and derivative): the graph dict the reader returns is the one the numerator kernels take."""
    from oracle import oracle as O

    S, P, T = 5, 17, 9
    rng = np.random.default_rng(21)
    graph, exs = synth_supervision_examples(rng, S, P, T, seed=6)
    m = capi.ChainEgs(W.ark(exs, binary)).merge_supervision(0, S, "output", P)
    x = (rng.standard_normal((T * S, P)) * 2).astype(np.float32)
    lp_ref, d_ref, ok_ref = O.num_forward_backward(graph, x, T, deriv_weight=1.0)
    lp, d, ok = O.num_forward_backward(m["num_graph"], x, T, deriv_weight=1.0)
    assert ok and ok_ref and lp == pytest.approx(lp_ref, rel=1e-6)
    np.testing.assert_allclose(d, d_ref, rtol=1e-5, atol=1e-7)


def test_archive_drives_a_whole_cpu_training_step():
    """Data path end to end on the host: an archive of single-sequence examples -> the merged minibatch (features in the
    t-major / sequence-fastest row order, numerator graph) -> one whole search-stage training step of the CPU reference
    (oracle/supernet_ref.py) gives the objective the same step gives on the generator's own arrays; with the rows left in the
    archive's sequence-major order it does not (the layout is what is being checked)."""
    from oracle import supernet_ref as R

    cfg = R.RefConfig(num_seqs=3, frames_per_eg=12, feat_dim=10, dim=24, bottleneck=8, num_blocks=2, num_offsets=3, prefinal_small=12,
                      num_pdfs=19, xent=False)
    T, _, _, _, in_t = R.frame_plan(cfg)
    rng = np.random.default_rng(40)
    den = synth.make_den_graph(30, cfg.num_pdfs, 4.0, seed=5)
    graph, exs = synth_supervision_examples(rng, cfg.num_seqs, cfg.num_pdfs, T, seed=9)
    feats = rng.standard_normal((len(in_t), cfg.num_seqs, cfg.feat_dim)).astype(np.float32)     # [t, sequence, dim]
    for s, ex in enumerate(exs):      # the recipe's shape: 'input' rows over the model's whole input context of one chunk
        ex["inputs"] = [dict(name="input", indexes=[(0, t, 0) for t in in_t], data=feats[:, s].copy(), coding="cm2")]
    egs = capi.ChainEgs(W.ark(exs, True))
    x, t0 = egs.merge_input(0, cfg.num_seqs, "input")
    sup = egs.merge_supervision(0, cfg.num_seqs, "output", cfg.num_pdfs)
    assert t0 == in_t[0] and x.shape == feats.shape
    np.testing.assert_allclose(x, feats, atol=float(np.ptp(feats)) / 65535 + 1e-6)     # two-byte compressed on disk
    draws = [np.full(cfg.num_offsets, 0.5, np.float32) for _ in range(2 * cfg.num_blocks)]

    def objf(rows, num_graph):
        return R.CpuSupernet(cfg, den, num_graph).step(np.ascontiguousarray(rows), draws, apply_update=False)

    want = objf(x.reshape(-1, cfg.feat_dim), graph)
    got = objf(x.reshape(-1, cfg.feat_dim), sup["num_graph"])
    assert got == pytest.approx(want, rel=1e-6)
    wrong = objf(x.transpose(1, 0, 2).reshape(-1, cfg.feat_dim), sup["num_graph"])              # sequence-major rows
    assert abs(wrong - want) > 1e-3 * abs(want)
