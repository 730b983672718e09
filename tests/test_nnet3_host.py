"""CPU tests of the host logic: index handling (bit-exact against an independent restatement),
PrecomputedIndexes I/O, the edit directive's error behaviour, the temperature schedule, and that the
shared library loads and exports every symbol include/*.h declares.  No compute calls (no GPU)."""
import os
import re
import random

import pytest

from tests import index_ref as IR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from tdnnf_nas_b200 import capi

    lib = capi.load()
    header = ""
    for h in sorted(os.listdir(os.path.join(ROOT, "include"))):
        header += open(os.path.join(ROOT, "include", h)).read()
    names = sorted(set(re.findall(r"\b(tdnnf_[a-z0-9_]+)\s*\(", header)))
    assert len(names) > 70
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.tdnnf_abi_version() == 1002


def test_no_gpu_means_loud_failure():
    import torch

    from tdnnf_nas_b200 import capi

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.TdnnfError, match="no CPU path"):
        capi.Context(0)


def _grid(S, t_in, t_out, xs=(0,)):
    inp = [(n, t, x) for t in t_in for x in xs for n in range(S)]
    out = [(n, t, x) for t in t_out for x in xs for n in range(S)]
    return inp, out


CASES = [
    ([0, 1, 2, 3, 4, 5, 6], 4, range(0, 20), range(0, 14)),
    ([-6, -5, -4, -3, -2, -1, 0], 3, range(-6, 10), range(0, 10)),
    ([0, 1, 2, 3, 4, 5, 6], 5, range(0, 21), range(0, 15, 3)),      # frame-subsampled output: reorder_t_in = 3
    ([-3, 0, 3], 2, range(-3, 16, 3), range(0, 13, 3)),             # input on a stride-3 grid
    ([0, 2], 1, range(0, 5), [0]),                                  # a single output frame (t_step_out == 0)
    ([0, 1, 2], 2, range(0, 8), range(0, 6, 2)),                    # needs num_t_in rounded up
]


@pytest.mark.parametrize("offsets,S,t_in,t_out", CASES)
def test_precompute_and_reorder_indexes_bit_exact(offsets, S, t_in, t_out):
    from tdnnf_nas_b200 import nnet3

    comp = nnet3.Component.tdnn_darts_for_indexing(offsets)
    inp, out = _grid(S, list(t_in), list(t_out))
    rnd = random.Random(S)
    inp_s, out_s = inp[:], out[:]
    rnd.shuffle(inp_s)
    rnd.shuffle(out_s)
    # ReorderIndexes on shuffled indexes -> the regular order the kernels assume
    ri, ro = comp.reorder_indexes(inp_s, out_s)
    ei, eo = IR.reorder_indexes(inp_s, out_s)
    assert ri == ei and ro == eo
    # PrecomputeIndexes on the regular order
    nnet3.set_rand_seed(1)
    for _ in range(12):  # the 1-in-11 spot check (tdnn.cc:859-870) is exercised too
        pi = comp.precompute_indexes(ri, ro)
        assert pi.row_stride_and_offsets() == tuple(IR.precompute_indexes(offsets, ri, ro))
    # text and binary round trip of the PrecomputedIndexes
    for binary in (False, True):
        data = pi.write(binary)
        head = b"<TdnnDARTSV3ComponentPrecomputedIndexes> <RowStride> "
        assert data.startswith(head)
        back = nnet3.PrecomputedIndexes.read(data, binary)
        assert back.write(binary) == data


def test_reorder_inserts_blanks_for_missing_indexes():
    from tdnnf_nas_b200 import nnet3

    comp = nnet3.Component.tdnn_darts_for_indexing([0, 1])
    inp, out = _grid(2, [0, 1, 2, 4], [0, 1, 3])   # t=3 missing from the input, t=2 from the output
    ri, ro = comp.reorder_indexes(inp, out)
    assert (ri, ro) == IR.reorder_indexes(inp, out)
    assert any(t == nnet3.kNoTime for (_, t, _) in ri) and any(t == nnet3.kNoTime for (_, t, _) in ro)


def test_get_input_indexes_and_is_computable():
    from tdnnf_nas_b200 import nnet3

    comp = nnet3.Component.tdnn_darts_for_indexing([-2, 0, 3])
    assert comp.get_input_indexes(5, 10, 1) == [(5, 8, 1), (5, 10, 1), (5, 13, 1)]
    avail = [(5, 8, 1), (5, 10, 1), (5, 13, 1)]
    assert comp.is_computable(5, 10, 1, avail)
    assert not comp.is_computable(5, 10, 1, avail[:2])
    assert not comp.is_computable(4, 10, 1, avail)
    with pytest.raises(nnet3.Nnet3Error):  # KALDI_ASSERT(output_index.t != kNoTime), tdnn.cc:767
        comp.get_input_indexes(0, nnet3.kNoTime, 0)


def test_edit_directive_errors():
    from tdnnf_nas_b200 import nnet3

    # missing proportion -> KALDI_ERR (utils.cc:1358-1361)
    with pytest.raises(nnet3.Nnet3Error, match="expected proportion"):
        nnet3.apply_edits("set-temperature-proportion name=*", [])
    with pytest.raises(nnet3.Nnet3Error, match="not currently supported"):
        nnet3.apply_edits("frobnicate name=*", [])
    with pytest.raises(nnet3.Nnet3Error, match="Could not interpret"):
        nnet3.apply_edits("set-temperature-proportion name=* proportion=0.5 bogus=1", [])
    nnet3.apply_edits("set-temperature-proportion name=tdnnf* proportion=0.5; set-temperature-proportion proportion=0.1", [])


def test_temperature_schedule_matches_reference_formula():
    from tdnnf_nas_b200 import nnet3

    # temperature_schedule.py:51 : T = (1 - f) * (1 - 0.03) + 0.03
    assert nnet3.temperature_for_iteration(0, 10) == pytest.approx(1.0)
    assert nnet3.temperature_for_iteration(10, 10) == pytest.approx(0.03)
    assert nnet3.temperature_for_iteration(5, 10) == pytest.approx(0.515)
    assert nnet3.temperature_edit_string(5, 10) == "set-temperature-proportion name=* proportion=0.515"


def test_unknown_component_type_fails():
    from tdnnf_nas_b200 import nnet3

    with pytest.raises(nnet3.Nnet3Error, match="Unknown component type"):
        nnet3.Component.new("NoSuchComponent", "dim=3")
    with pytest.raises(nnet3.Nnet3Error):
        nnet3.Component.read(b"<NoSuchComponent> <Dim> 3 </NoSuchComponent> ", False)


def test_batchnorm_component_config_and_text_form():
    """BatchNormComponent (the stock component the search stage `sed`s into BatchNormTestComponent): config keys,
    Properties() and the on-disk token stream of nnet-normalize-component.cc:289-317, 616-642.  Training mode holds no
    device state, so this runs without a GPU."""
    from tdnnf_nas_b200 import nnet3

    bn = nnet3.Component.new("BatchNormComponent", "dim=12 block-dim=4 epsilon=0.01 target-rms=0.5")
    assert bn.type() == "BatchNormComponent" and bn.input_dim() == 12 and bn.output_dim() == 12
    assert bn.properties() == (nnet3.kSimpleComponent | nnet3.kBackpropNeedsOutput | nnet3.kPropagateInPlace |
                               nnet3.kBackpropInPlace | nnet3.kInputContiguous | nnet3.kOutputContiguous |
                               nnet3.kUsesMemo | nnet3.kStoresStats)
    txt = bn.write(False)
    assert txt.startswith(b"<BatchNormComponent> <Dim> 12 <BlockDim> 4 <Epsilon> 0.01 <TargetRms> 0.5 <TestMode> F <Count> 0 <StatsMean>")
    assert txt.rstrip().endswith(b"</BatchNormComponent>")
    full = nnet3.Component.new("BatchNormComponent", "dim=8")
    assert full.properties() & (nnet3.kInputContiguous | nnet3.kOutputContiguous) == 0
    assert "test-mode=false" in full.info() and "block-dim=8" in full.info()
    for bad in ("block-dim=4", "dim=12 block-dim=5", "dim=8 epsilon=0", "dim=8 bogus=1"):
        with pytest.raises(nnet3.Nnet3Error):
            nnet3.Component.new("BatchNormComponent", bad)


@pytest.mark.parametrize("n", [1, 2, 7, 20, 80, 128])
def test_natural_gradient_eigen_solver_matches_numpy(n):
    """The host half of OnlineNaturalGradient's update diagonalises an R x R matrix Z_t (R = 20 / 80 in the recipes):
    Householder + implicit QL in csrc/nnet3/natural_gradient.cc against numpy.linalg.eigh, including repeated and
    widely spread eigenvalues (the update floors small eigenvalues to a common value)."""
    import ctypes as C

    import numpy as np

    from tdnnf_nas_b200 import capi

    lib = capi.load()
    g = np.random.default_rng(n)
    q, _ = np.linalg.qr(g.standard_normal((n, n)))
    w = np.concatenate([np.full(n // 3, 2.5), 10.0 ** g.uniform(-8, 3, n - n // 3)])  # a repeated block + 11 decades
    a = (q * w) @ q.T
    a = (a + a.T) / 2
    vals, vecs = np.zeros(n), np.zeros((n, n))
    dp = C.POINTER(C.c_double)
    rc = lib.tdnnf_nnet3_symmetric_eigen(np.ascontiguousarray(a).ctypes.data_as(dp), n, vals.ctypes.data_as(dp), vecs.ctypes.data_as(dp))
    assert rc == 0
    np.testing.assert_allclose(np.sort(vals), np.sort(np.linalg.eigvalsh(a)), rtol=1e-9, atol=1e-12 * abs(w).max())
    assert np.abs(vecs.T @ vecs - np.eye(n)).max() < 1e-12
    assert np.abs(a @ vecs - vecs * vals).max() < 1e-11 * abs(w).max()


def test_synthetic_numerator_paths_lie_in_the_denominator_graph():
    """synth.make_num_graphs(den_graph=...): every (forward pdf, self-loop pdf) pair of a numerator FST is an arc of the
    denominator graph followed by the self-loop of its destination, so numerator paths are denominator paths."""
    import numpy as np

    from tdnnf_nas_b200 import synth

    den = synth.make_den_graph(400, 90, 6.0, seed=3)
    A = den["num_arcs"]
    src = np.repeat(np.arange(400), den["fwd_ranges"][:, 1] - den["fwd_ranges"][:, 0])
    arcs = {}
    for a in range(A):
        arcs.setdefault((int(src[a]), int(den["pdf"][a])), []).append(int(den["state"][a]))
    loops = {int(src[a]): int(den["pdf"][a]) for a in range(A) if int(den["state"][a]) == int(src[a])}
    num = synth.make_num_graphs(6, 90, 30, seed=9, den_graph=den)
    nA = num["num_arcs"]
    fr = num["fwd_ranges"]
    for s in range(num["num_seqs"]):
        lo, hi = num["state_offsets"][s], num["state_offsets"][s + 1]
        cand = set(range(400))  # den states the walk may be in before the first phone
        for st in range(lo, hi - 1):
            fwd = [a for a in range(fr[st][0], fr[st][1]) if num["arc_state"][a] == st + 1]
            assert len(fwd) == 1
            pdf = int(num["arc_pdf"][fwd[0]])
            nxt = set(d for h in cand for d in arcs.get((h, pdf), []))
            assert nxt, "forward pdf is not on any arc leaving the walk's state"
            loop = [a for a in range(fr[st + 1][0], fr[st + 1][1]) if num["arc_state"][a] == st + 1]
            assert len(loop) == 1
            nxt = {h for h in nxt if loops.get(h) == int(num["arc_pdf"][loop[0]])}
            assert nxt, "self-loop pdf does not match the destination's self-loop"
            cand = nxt
    assert nA == len(num["arc_pdf"]) // 2


def test_rectified_linear_component_config_and_text_form():
    """RectifiedLinearComponent (the ReLU of every TDNN-F block): config keys of NonlinearComponent::InitFromConfig
    (nnet-component-itf.cc:707-718), Properties() (nnet-simple-component.h:351-354) and the token stream of a fresh
    component (nnet-component-itf.cc:630-687).  No statistics yet, so no device state: runs without a GPU."""
    from tdnnf_nas_b200 import nnet3

    r = nnet3.Component.new("RectifiedLinearComponent", "dim=12 block-dim=4 self-repair-scale=1e-05 self-repair-lower-threshold=0.1")
    assert r.type() == "RectifiedLinearComponent" and r.input_dim() == 12 and r.output_dim() == 12
    assert r.properties() == (nnet3.kSimpleComponent | nnet3.kBackpropNeedsOutput | nnet3.kPropagateInPlace |
                              nnet3.kStoresStats | nnet3.kInputContiguous)
    txt = r.write(False)
    assert txt.startswith(b"<RectifiedLinearComponent> <Dim> 12 <BlockDim> 4 <ValueAvg>  [ ]\n<DerivAvg>  [ ]\n<Count> 0 <OderivRms>  [ ]\n"
                          b"<OderivCount> 0 <NumDimsSelfRepaired> 0 <NumDimsProcessed> 0 <SelfRepairLowerThreshold> 0.1 <SelfRepairScale> 1e-05 ")
    assert txt.rstrip().endswith(b"</RectifiedLinearComponent>")
    assert "self-repair-scale=1e-05" in r.info() and "block-dim=4" in r.info()
    plain = nnet3.Component.new("RectifiedLinearComponent", "dim=8")
    assert plain.properties() & nnet3.kInputContiguous == 0
    assert plain.write(False).startswith(b"<RectifiedLinearComponent> <Dim> 8 <ValueAvg>")
    for bad in ("block-dim=4", "dim=12 block-dim=5", "dim=8 bogus=1"):
        with pytest.raises(nnet3.Nnet3Error):
            nnet3.Component.new("RectifiedLinearComponent", bad)


def test_dropout_schedule():
    """--trainer.dropout-schedule of the recipes (run_tdnn_7q_fbk_40_manual.sh:48): piecewise linear in the data fraction."""
    from tdnnf_nas_b200 import nnet3

    sch = "0,0@0.20,0.5@0.50,0"
    assert nnet3.parse_dropout_schedule(sch) == [(0.0, 0.0), (0.2, 0.0), (0.5, 0.5), (1.0, 0.0)]
    for f, want in [(0.0, 0.0), (0.1, 0.0), (0.2, 0.0), (0.35, 0.25), (0.5, 0.5), (0.75, 0.25), (1.0, 0.0)]:
        assert nnet3.dropout_proportion_for_fraction(sch, f) == pytest.approx(want, abs=1e-12)
    assert nnet3.dropout_proportion_for_fraction("0.1,0.3,0.0", 0.25) == pytest.approx(0.2)  # bare middle value = @0.5
    assert nnet3.dropout_edit_string(sch, 0.5) == "set-dropout-proportion name=* proportion=0.5"
    for bad in ("0.5", "0,0.5@0.6,0.2@0.3,0", "0,1.5@0.5,0"):
        with pytest.raises(ValueError):
            nnet3.parse_dropout_schedule(bad)


def _dropout_indexes(pi):
    toks = pi.write(False).decode().split()
    assert toks[0] == "<GeneralDropoutComponentPrecomputedIndexes>" and toks[1] == "<NumMaskRows>"
    lo, hi = toks.index("["), toks.index("]")
    return int(toks[2]), [int(t) for t in toks[lo + 1: hi]]


def test_general_dropout_host_side():
    """GeneralDropoutComponent: config, Properties, token stream, the set-dropout-proportion directive and PrecomputeIndexes
    (one mask row per sequence; per block of time-period frames when set; block-dim reshaping) -- no device needed."""
    from tdnnf_nas_b200 import nnet3

    c = nnet3.Component.new("GeneralDropoutComponent", "dim=12 dropout-proportion=0.0 continuous=true")
    assert c.type() == "GeneralDropoutComponent" and c.dims()[:2] == (12, 12) or c.input_dim() == 12
    assert c.properties() == (nnet3.kRandomComponent | nnet3.kPropagateInPlace | nnet3.kBackpropInPlace | nnet3.kUsesMemo)
    assert c.write(False).split() == b"<GeneralDropoutComponent> <Dim> 12 <BlockDim> 12 <TimePeriod> 0 <DropoutProportion> 0 <Continuous> </GeneralDropoutComponent>".split()
    for binary in (False, True):
        back = nnet3.Component.read(c.write(binary), binary)
        assert back.write(binary) == c.write(binary)
    nnet3.apply_edits("set-dropout-proportion name=tdnnf*.dropout proportion=0.25",
                      [("tdnnf2.dropout", c), ("tdnnf2.relu", nnet3.Component.new("RectifiedLinearComponent", "dim=12"))])
    assert c.dropout_proportion() == pytest.approx(0.25)
    assert b"<DropoutProportion> 0.25 " in c.write(False)
    with pytest.raises(nnet3.Nnet3Error, match="expected proportion"):
        nnet3.apply_edits("set-dropout-proportion name=*", [("x", c)])
    # t-major grid, 3 sequences: mask row = n
    grid = [(n, t, 0) for t in range(-2, 5) for n in range(3)]
    rows, idx = _dropout_indexes(c.precompute_indexes(grid, grid))
    assert rows == 3 and idx == [n for _t in range(7) for n in range(3)]
    # time-period 3: a new mask row per (n, floor(t / 3)), numbered in order of first appearance
    c3 = nnet3.Component.new("GeneralDropoutComponent", "dim=12 time-period=3 dropout-proportion=0.2")
    rows, idx = _dropout_indexes(c3.precompute_indexes(grid, grid))
    blocks = sorted({t // 3 for t in range(-2, 5)})
    assert rows == 3 * len(blocks)
    assert idx == [blocks.index(t // 3) * 3 + n for t in range(-2, 5) for n in range(3)]
    # block-dim 4 of dim 12: every input row is 3 rows of the reshaped view, each with its own mask row
    cb = nnet3.Component.new("GeneralDropoutComponent", "dim=12 block-dim=4 dropout-proportion=0.2")
    assert cb.properties() & (nnet3.kInputContiguous | nnet3.kOutputContiguous)
    rows, idx = _dropout_indexes(cb.precompute_indexes(grid, grid))
    assert rows == 9 and idx[:9] == [0, 1, 2, 3, 4, 5, 6, 7, 8] and len(idx) == 3 * len(grid)
    pi = cb.precompute_indexes(grid, grid)
    for binary in (False, True):
        assert nnet3.PrecomputedIndexes.read(pi.write(binary), binary).write(binary) == pi.write(binary)
    with pytest.raises(nnet3.Nnet3Error):
        nnet3.Component.new("GeneralDropoutComponent", "dim=12 block-dim=5")
    with pytest.raises(nnet3.Nnet3Error, match="SpecAugment"):
        nnet3.Component.new("GeneralDropoutComponent", "dim=12 specaugment-max-proportion=0.5")


def test_stock_tdnn_precomputed_indexes_tokens():
    """The stock TdnnComponent writes its indexes under upstream's own token (itf.cc: TdnnComponentPrecomputedIndexes)."""
    from tdnnf_nas_b200 import capi, nnet3

    # TdnnComponent's InitFromConfig allocates parameters, so go through a model stream read on the host? No device here:
    # the index type itself is reachable through ComponentPrecomputedIndexes::ReadNew.
    data = b"<TdnnComponentPrecomputedIndexes> <RowStride> 3 <RowOffsets> [ 0 7 14 ]\n</TdnnComponentPrecomputedIndexes> "
    pi = nnet3.PrecomputedIndexes.read(data, False)
    assert pi.row_stride_and_offsets() == (3, [0, 7, 14])
    assert pi.write(False).split() == data.split()
    assert nnet3.PrecomputedIndexes.read(pi.write(True), True).write(False).split() == data.split()


def test_frame_plan_matches_the_survey_row_counts():
    """SURVEY 8d, config 3 at 64 chunks x 150 frames: `.linear` outputs 19 840 (tdnnf2) ... 9 856 (tdnnf15) rows, `.affine`
    outputs 19 456 (tdnnf2), 10 240 (tdnnf14), 3 200 (tdnnf15); 4.17 TFLOP of algorithmic GEMM work per step.  Manual
    system: time-strides 1,1,1,0 then 6 (run_tdnn_7q_fbk_40_manual.sh:138-151)."""
    from tdnnf_nas_b200.supernet import SupernetConfig, algorithmic_flops, block_offsets, frame_plan

    cfg = SupernetConfig()
    left, right = block_offsets(cfg)
    assert left[0] == list(range(-6, 1)) and right[0] == list(range(0, 7))
    T, out_t, lin_t, aff_t, in_t = frame_plan(cfg, left, right)
    S = cfg.num_seqs
    assert T == 50 and out_t[:3] == [0, 3, 6]
    assert [len(t) * S for t in lin_t][0] == 19840 and [len(t) * S for t in lin_t][-1] == 9856
    assert [len(t) * S for t in aff_t][0] == 19456 and [len(t) * S for t in aff_t][-2:] == [10240, 3200]
    for b in range(1, cfg.num_blocks):
        assert len(lin_t[b - 1]) - len(lin_t[b]) == 12  # each layer down adds 12 frames of context
    assert algorithmic_flops(cfg) == pytest.approx(4.17e12, rel=2e-3)
    pre = SupernetConfig(mode="pretrain")
    assert algorithmic_flops(pre) == pytest.approx(algorithmic_flops(cfg) * 2 / 7, rel=1e-12)  # shared + sampled offset
    man = SupernetConfig(mode="manual", num_seqs=128)
    left, right = block_offsets(man)
    assert left[:5] == [[-1, 0], [-1, 0], [-1, 0], [0], [-6, 0]] and right[3] == [0] and right[-1] == [0, 6]
    _, _, lin_t, aff_t, in_t = frame_plan(man, left, right)
    assert len(aff_t[-1]) == 50 and aff_t[-1][-1] == 147
    assert lin_t[-1][-1] == 147 + 6 and in_t[0] == -(6 * 10 + 0 + 3) and in_t[-1] == 147 + 6 * 10 + 3


def test_raw_nnet_container_roundtrip():
    """Nnet::Write / Nnet::Read framing (kaldi nnet3/nnet-nnet.cc): '<Nnet3>', config lines, blank line, '<NumComponents> N',
    '<ComponentName> name' + component ..., '</Nnet3>'; text and binary, parameter-free components (no device here)."""
    from tdnnf_nas_b200 import nnet3

    comps = [("tdnnf2.softmax", nnet3.Component.new("GumbelSoftmaxFlopsComponent", "dim=8 scale=0.001 temp-proportion=0.5")),
             ("tdnnf20.copyn", nnet3.Component.new("CopyNComponent", "input-dim=1 output-dim=25")),
             ("tdnnf20.output", nnet3.Component.new("ElementwiseProductComponent", "input-dim=50 output-dim=25")),
             ("tdnnf2.dropout", nnet3.Component.new("GeneralDropoutComponent", "dim=12 dropout-proportion=0.25 continuous=true"))]
    cfg = ["input-node name=input dim=40", "component-node name=tdnnf2.softmax component=tdnnf2.softmax input=tdnnf2.alpha"]
    text = nnet3.write_nnet(cfg, comps, False)
    lines = text.decode().split("\n")
    assert lines[0] == "<Nnet3> " and lines[1:3] == cfg and lines[3] == "" and lines[4] == "<NumComponents> 4 "
    assert lines[5].startswith("<ComponentName> tdnnf2.softmax <GumbelSoftmaxFlopsComponent> <Dim> 8 ")
    assert lines[-1] == "</Nnet3> "
    for binary in (False, True):
        data = nnet3.write_nnet(cfg, comps, binary)
        cfg2, back = nnet3.read_nnet(data, binary)
        assert cfg2 == cfg and [n for n, _ in back] == [n for n, _ in comps]
        for (_, a), (_, b) in zip(comps, back):
            assert a.type() == b.type() and a.write(binary) == b.write(binary)
    # the edits of the training driver apply to the components read back, by name pattern
    _, back = nnet3.read_nnet(text, False)
    nnet3.apply_edits("set-temperature-proportion name=tdnnf*.softmax proportion=0.1; set-dropout-proportion name=* proportion=0.4", back)
    assert back[0][1].temp_proportion() == pytest.approx(0.1) and back[3][1].dropout_proportion() == pytest.approx(0.4)
    with pytest.raises(nnet3.Nnet3Error):
        nnet3.read_nnet(text[:-12], False)  # truncated: no </Nnet3>
    with pytest.raises(ValueError):
        nnet3.write_nnet(["a", ""], comps)


def test_docs_name_only_declared_entry_points():
    """Every tdnnf_* name in the documents exists in include/*.h (a `tdnnf_x_{fwd,bwd}` shorthand must be the prefix of one)."""
    header = "".join(open(os.path.join(ROOT, "include", h)).read() for h in sorted(os.listdir(os.path.join(ROOT, "include"))))
    declared = set(re.findall(r"\b(tdnnf_[a-z0-9_]+)\b", header))
    for doc in ("INTEGRATION.md", "DESIGN.md", "README.md", os.path.join("profiles", "INDEX.md")):
        names = set(re.findall(r"\b(tdnnf_[a-z0-9_]+)\b", open(os.path.join(ROOT, doc)).read()))
        stale = sorted(n for n in names if n not in declared and not (n.endswith("_") and any(d.startswith(n) for d in declared))
                       and n not in ("tdnnf_nas_b200", "tdnnf_nnet3"))
        assert not stale, (doc, stale)


def _top_level_args(text, open_paren):
    """The argument strings of the call whose '(' is at text[open_paren]; None if the parentheses do not balance."""
    depth, args, cur = 0, [], ""
    for i in range(open_paren, len(text)):
        ch = text[i]
        if ch in "([{":
            depth += 1
            if depth == 1:
                continue
        elif ch in ")]}":
            depth -= 1
            if depth == 0:
                if cur.strip():
                    args.append(cur.strip())
                return args
        if depth == 1 and ch == ",":
            args.append(cur.strip())
            cur = ""
        else:
            cur += ch
    return None


def test_integration_snippets_call_the_abi_with_the_declared_number_of_arguments():
    """Every complete tdnnf_* call inside a ```cpp block of INTEGRATION.md passes as many arguments as include/*.h declares
    (calls abbreviated with '...' are skipped): the reference-side stubs shown to a maintainer cannot drift from the ABI."""
    header = "".join(open(os.path.join(ROOT, "include", h)).read() for h in sorted(os.listdir(os.path.join(ROOT, "include"))))
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    header = re.sub(r"//[^\n]*", " ", header)
    arity = {}
    for m in re.finditer(r"\b(tdnnf_[a-z0-9_]+)\s*\(", header):
        args = _top_level_args(header, m.end() - 1)
        if args is not None and header[:m.start()].rstrip().split()[-1:] not in ([], ["return"]):
            arity.setdefault(m.group(1), 0 if args == ["void"] else len(args))
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    checked, bad = 0, []
    for block in re.findall(r"```cpp\n(.*?)```", doc, flags=re.S):
        code = re.sub(r"//[^\n]*", " ", block)
        code = re.sub(r"/\*.*?\*/", " ", code, flags=re.S)
        for m in re.finditer(r"\b(tdnnf_[a-z0-9_]+)\s*\(", code):
            name = m.group(1)
            args = _top_level_args(code, m.end() - 1)
            if name not in arity or args is None or any(a == "..." or a.endswith("...") for a in args):
                continue
            checked += 1
            if len(args) != arity[name]:
                bad.append((name, len(args), arity[name]))
    assert checked >= 12, checked
    assert not bad, bad


def _header_arity():
    header = "".join(open(os.path.join(ROOT, "include", h)).read() for h in sorted(os.listdir(os.path.join(ROOT, "include"))))
    header = re.sub(r"//[^\n]*", " ", re.sub(r"/\*.*?\*/", " ", header, flags=re.S))
    arity = {}
    for m in re.finditer(r"\b(tdnnf_[a-z0-9_]+)\s*\(", header):
        args = _top_level_args(header, m.end() - 1)
        if args is not None and header[:m.start()].rstrip().split()[-1:] not in ([], ["return"]):
            arity.setdefault(m.group(1), 0 if args == ["void"] else len(args))
    return arity


def test_python_bindings_match_the_declared_arity():
    """ctypes checks nothing against the C prototype: the argtypes lists of capi.py and every call site of the handle API in
    nnet3.py (which declares no argtypes) are compared with the parameter counts of include/*.h."""
    arity = _header_arity()
    pkg = os.path.join(ROOT, "tdnn-f_nas_b200")
    bad, sigs, calls = [], 0, 0
    src = open(os.path.join(pkg, "capi.py")).read()
    for m in re.finditer(r"sig\(\s*\"(tdnnf_[a-z0-9_]+)\"\s*,\s*\[", src):
        args = _top_level_args(src, m.end() - 1)
        sigs += 1
        if arity.get(m.group(1)) != len(args):
            bad.append(("capi.py sig", m.group(1), len(args), arity.get(m.group(1))))
    for f in ("nnet3.py", "supernet.py", "chain.py", "parallel.py"):
        src = re.sub(r"#[^\n]*", " ", open(os.path.join(pkg, f)).read())
        for m in re.finditer(r"\.(tdnnf_nnet3_[a-z0-9_]+)\s*\(", src):
            args = _top_level_args(src, m.end() - 1)
            if args is None or m.group(1) not in arity or any(a.startswith("*") for a in args):
                continue
            calls += 1
            if len(args) != arity[m.group(1)]:
                bad.append((f, m.group(1), len(args), arity[m.group(1)]))
    assert sigs > 100 and calls > 60, (sigs, calls)
    assert not bad, bad


def test_capi_argtypes_match_the_declared_parameter_kinds():
    """... and kind for kind: pointer / int / float / double / 64-bit integer (an int declared where the C side takes a float
    goes to the wrong register without any error)."""
    header = "".join(open(os.path.join(ROOT, "include", h)).read() for h in sorted(os.listdir(os.path.join(ROOT, "include"))))
    header = re.sub(r"//[^\n]*", " ", re.sub(r"/\*.*?\*/", " ", header, flags=re.S))

    def c_kind(p):
        if "*" in p or "[" in p:
            return "ptr"
        toks = p.replace("const", "").split()
        base = " ".join(toks[:-1]) if len(toks) > 1 else toks[0]
        return {"float": "float", "double": "double", "int": "int", "int32_t": "int", "uint64_t": "u64", "int64_t": "i64", "size_t": "u64",
                "unsigned long long": "u64"}[base]

    def py_kind(a):
        if a in ("vp", "c_int_p", "c_float_p", "pp_i", "pp_f", "pp_c", "C.c_char_p", "C.c_void_p") or a.startswith("C.POINTER"):
            return "ptr"
        return {"i": "int", "C.c_int": "int", "C.c_int32": "int", "f": "float", "C.c_float": "float", "C.c_double": "double",
                "C.c_uint64": "u64", "C.c_int64": "i64"}[a]

    proto = {}
    for m in re.finditer(r"\b(tdnnf_[a-z0-9_]+)\s*\(", header):
        args = _top_level_args(header, m.end() - 1)
        if args is not None and header[:m.start()].rstrip().split()[-1:] not in ([], ["return"]):
            proto.setdefault(m.group(1), [] if args == ["void"] else [c_kind(a) for a in args])
    src = open(os.path.join(ROOT, "tdnn-f_nas_b200", "capi.py")).read()
    bad, n = [], 0
    for m in re.finditer(r"sig\(\s*\"(tdnnf_[a-z0-9_]+)\"\s*,\s*\[", src):
        kinds = [py_kind(a.strip()) for a in _top_level_args(src, m.end() - 1)]
        n += 1
        if kinds != proto[m.group(1)]:
            bad.append((m.group(1), [(k, a, b) for k, (a, b) in enumerate(zip(kinds, proto[m.group(1)])) if a != b]))
    assert n > 100 and not bad, bad
