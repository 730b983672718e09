"""Property tests (hypothesis) of the index handling: the C++ ReorderIndexes / PrecomputeIndexes against the independent
Python restatement (tests/index_ref.py) on random regular grids -- random offsets, sequence counts, start frames, input
steps, frame-subsampling factors, extra x values, shuffled input order, and randomly dropped frames.  Bit-exact (ints)."""
import random

import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from tests import index_ref as IR


@st.composite
def grids(draw):
    n_off = draw(st.integers(1, 7))
    step_in = draw(st.sampled_from([1, 1, 1, 3]))
    factor = draw(st.sampled_from([1, 1, 2, 3]))              # t_step_out / t_step_in
    offsets = sorted(draw(st.sets(st.integers(-6, 6), min_size=n_off, max_size=n_off)))
    offsets = [o * step_in for o in offsets]
    S = draw(st.integers(1, 5))
    xs = draw(st.sampled_from([(0,), (0,), (0, 1)]))
    start_out = draw(st.integers(-7, 7)) * step_in
    num_out = draw(st.integers(1, 9))
    t_out = [start_out + k * step_in * factor for k in range(num_out)]
    lo, hi = t_out[0] + min(offsets), t_out[-1] + max(offsets)
    t_in = list(range(lo, hi + 1, step_in))
    seed = draw(st.integers(0, 10 ** 6))
    return offsets, S, xs, t_in, t_out, seed


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(grids())
def test_reorder_and_precompute_match_the_python_restatement(case):
    from tdnnf_nas_b200 import nnet3

    offsets, S, xs, t_in, t_out, seed = case
    comp = nnet3.Component.tdnn_darts_for_indexing(offsets)
    inp = [(n, t, x) for t in t_in for x in xs for n in range(S)]
    out = [(n, t, x) for t in t_out for x in xs for n in range(S)]
    rnd = random.Random(seed)
    rnd.shuffle(inp)
    rnd.shuffle(out)
    ri, ro = comp.reorder_indexes(inp, out)
    ei, eo = IR.reorder_indexes(inp, out)
    assert ri == ei and ro == eo
    # the regular order is a fixed point, holds every requested index exactly once, and only adds blanks
    assert comp.reorder_indexes(ri, ro) == (ri, ro)
    assert sorted(i for i in ri if i[1] != nnet3.kNoTime) == sorted(inp)
    assert sorted(i for i in ro if i[1] != nnet3.kNoTime) == sorted(out)
    nnet3.set_rand_seed(seed)
    pi = comp.precompute_indexes(ri, ro)
    row_stride, row_offsets = pi.row_stride_and_offsets()
    assert (row_stride, list(row_offsets)) == tuple(IR.precompute_indexes(offsets, ri, ro))
    # what the offsets MEAN: output row k (t-major) of offset i reads input row row_offsets[i] + k * row_stride,
    # which holds the index (n, t + offset_i, x) of that output row
    num_images = S * len(xs)
    for i, off in enumerate(offsets):
        for k in rnd.sample(range(len(ro)), min(len(ro), 6)):
            n, t, x = ro[k]
            if t == nnet3.kNoTime:
                continue
            r = row_offsets[i] + k * row_stride
            assert 0 <= r < len(ri) and ri[r] == (n, t + off, x), (k, i, r)
    assert len(ro) % num_images == 0 and len(ri) % num_images == 0


@settings(max_examples=40, deadline=None)
@given(grids(), st.integers(0, 10 ** 6))
def test_reorder_with_missing_frames(case, drop_seed):
    """Frames missing from the request come back as blanks (t = kNoTime) in their regular slot."""
    from tdnnf_nas_b200 import nnet3

    offsets, S, xs, t_in, t_out, _ = case
    rnd = random.Random(drop_seed)
    if len(t_in) > 3:
        t_in = [t for j, t in enumerate(t_in) if j in (0, len(t_in) - 1) or rnd.random() > 0.2]
    comp = nnet3.Component.tdnn_darts_for_indexing(offsets)
    inp = [(n, t, x) for t in t_in for x in xs for n in range(S)]
    out = [(n, t, x) for t in t_out for x in xs for n in range(S)]
    try:
        expect = IR.reorder_indexes(inp, out)
    except AssertionError:
        # dropping frames changed the input step so that t_step_out % t_step_in != 0: the reference asserts too
        with pytest.raises(nnet3.Nnet3Error):
            comp.reorder_indexes(inp, out)
        return
    assert comp.reorder_indexes(inp, out) == expect


@settings(max_examples=150, deadline=None)
@given(st.text(alphabet="ab.-_1", min_size=1, max_size=8), st.text(alphabet="ab.-_1*", min_size=1, max_size=7))
def test_name_patterns_of_edit_directives(name, pattern):
    """NameMatchesPattern (kaldi nnet3/nnet-parse.cc): '*' matches any run of characters (also empty), everything else is
    literal -- checked through set-dropout-proportion against a regular expression."""
    import re

    from tdnnf_nas_b200 import nnet3

    comp = nnet3.Component.new("GeneralDropoutComponent", "dim=4 dropout-proportion=0.0")
    nnet3.apply_edits(f"set-dropout-proportion name={pattern} proportion=0.5", [(name, comp)])
    want = re.fullmatch(".*".join(re.escape(part) for part in pattern.split("*")), name) is not None
    assert (comp.dropout_proportion() == 0.5) == want, (name, pattern)


@settings(max_examples=120, deadline=None)
@given(st.integers(1, 4), st.integers(-9, 9), st.integers(1, 9), st.sampled_from([0, 1, 2, 3, 5]), st.sampled_from([1, 2, 3]),
       st.integers(0, 10 ** 6))
def test_general_dropout_precompute_indexes(S, t0, frames, time_period, multiple, seed):
    """GeneralDropoutComponent::PrecomputeIndexes (kaldi nnet-general-component.cc): one mask row per distinct
    (n, x, floor(t / time_period)) -- per (n, x) when time_period = 0 -- numbered in order of first appearance; with
    block-dim < dim every input row becomes `multiple` reshaped rows with consecutive mask rows.  Any index order,
    negative frames included (floor division)."""
    from tdnnf_nas_b200 import nnet3

    block = 4
    comp = nnet3.Component.new("GeneralDropoutComponent",
                               f"dim={block * multiple} block-dim={block} time-period={time_period} dropout-proportion=0.3")
    idx = [(n, t, x) for t in range(t0, t0 + frames) for x in (0, 1) for n in range(S)]
    random.Random(seed).shuffle(idx)
    toks = comp.precompute_indexes(idx, idx).write(False).decode().split()
    rows = int(toks[2])
    got = [int(t) for t in toks[toks.index("[") + 1: toks.index("]")]]
    seen = {}
    want = []
    for (n, t, x) in idx:
        key = (n, x, 0 if time_period == 0 else t // time_period)  # Python's // floors, like DivideRoundingDown
        r = seen.setdefault(key, len(seen))
        want += [r * multiple + j for j in range(multiple)]
    assert rows == len(seen) * multiple and got == want
