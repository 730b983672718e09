"""GPU tests of compatibility with the reference's scripts and model text:
  * every parameterised component line the reference's config generators emit is accepted (fixture: tests/golden/ref_scripts.json);
  * the text model written here parses with the expressions of generate_top_list.py / bottleneckdim_search_top_model_size.py."""
import numpy as np
import pytest

from tests.test_reference_scripts import DEVICE_TYPES, _component_lines, _fixture, _kv

pytestmark = pytest.mark.gpu


@pytest.fixture()
def nn(ctx):
    from tdnnf_nas_b200 import nnet3

    nnet3.set_context(ctx)
    nnet3.set_rand_seed(777)
    return nnet3


def test_generated_lines_of_parameterised_components_are_accepted(ctx):
    """TdnnDARTSV3Component (context-offset supernet), TdnnComponent (bottleneck search), ConstantFunction / Onehot."""
    from tdnnf_nas_b200 import nnet3

    nnet3.set_context(ctx)
    nnet3.set_rand_seed(3)
    fx = _fixture()
    seen = set()
    for key in ("change_config_gumbel", "bottleneck_final_config", "supernet_final_config"):
        for name, typ, rest in _component_lines(fx[key]):
            if typ not in DEVICE_TYPES or (typ, name[-6:]) in seen:
                continue
            seen.add((typ, name[-6:]))
            comp = nnet3.Component.new(typ, rest)
            kv = _kv(rest)
            assert comp.type() == typ
            assert comp.input_dim() == int(kv["input-dim"]) and comp.output_dim() == int(kv["output-dim"])
            if typ in ("TdnnComponent", "TdnnDARTSV3Component"):
                assert "time-offsets=" + kv["time-offsets"] in comp.info()
                assert comp.orthonormal_constraint() == float(kv.get("orthonormal-constraint", 0.0))
            if typ == "TdnnDARTSV3Component":
                # generate_config.py forces use-bias=true and writes all 7 candidate offsets; alpha slots + bias
                assert "use-bias=true" in rest or "use-bias" not in rest
                n = len(kv["time-offsets"].split(","))
                assert n == 7
                assert comp.num_parameters() == int(kv["output-dim"]) * n * int(kv["input-dim"]) + n + int(kv["output-dim"])
                head = comp.write(False)[:400]  # the pretrain-stage flags of run_TDNN_DARTSV3_fbk_stride_pretrain.sh:124
                assert b"<use-gumbel> F" in head and b"<uniform-sample> T" in head
            if typ == "TdnnComponent" and kv.get("use-bias") == "false":
                assert comp.num_parameters() == int(kv["output-dim"]) * len(kv["time-offsets"].split(",")) * int(kv["input-dim"])
    assert {t for t, _ in seen} == DEVICE_TYPES


def test_text_model_is_readable_by_generate_top_list(nn):
    """NAS/scripts/generate_top_list.py:21-27 recovers the architecture weights from the TEXT model with
    `line.split('[')[1].split(' ')[1:offset+1]` on every line that contains '<BiasParams>': the alpha entries must be
    the first `offset` numbers of that one line (Kaldi's vector text form)."""
    n, din, dout = 7, 16, 12
    comp = nn.Component.new("TdnnDARTSV3Component", f"input-dim={din} output-dim={dout} time-offsets=0,1,2,3,4,5,6")
    v = comp.vectorize()
    alpha = np.array([0.25, -1.5, 3.0, 1e-5, -0.125, 2.75, 0.5], dtype=np.float32)
    v[dout * n * din: dout * n * din + n] = alpha
    comp.unvectorize(v)
    text = comp.write(False).decode()
    hits = [line.strip() for line in text.split("\n") if "<BiasParams>" in line]
    assert len(hits) == 1
    prob = hits[0].split("[")[1].split(" ")[1:n + 1]          # the reference's expression, offset = 7
    np.testing.assert_allclose([float(item) for item in prob], alpha, rtol=1e-5)


def test_raw_nnet_text_is_readable_by_the_reference_scripts(nn):
    """A raw nnet written by write_nnet: generate_top_list.py:21-27 ('<BiasParams>' lines) and
    bottleneckdim_search_top_model_size.py:15-18 ('alpha <ConstantFunctionComponent>' lines, first '[' on that line is
    <Output>) recover the architecture weights with their own expressions; read_nnet restores the components."""
    alpha8 = np.array([0.5, -0.25, 1.5, 2.0, -3.0, 0.125, 0.75, -1.0], dtype=np.float32)
    cf = nn.Component.new("ConstantFunctionComponent", "input-dim=220 output-dim=8 is-updatable=true use-natural-gradient=false")
    cf.unvectorize(alpha8)
    n, din, dout = 7, 16, 12
    darts = nn.Component.new("TdnnDARTSV3Component", f"input-dim={din} output-dim={dout} time-offsets=-6,-5,-4,-3,-2,-1,0")
    v = darts.vectorize()
    alpha7 = np.array([0.25, -1.5, 3.0, 1e-5, -0.125, 2.75, 0.5], dtype=np.float32)
    v[dout * n * din: dout * n * din + n] = alpha7
    darts.unvectorize(v)
    comps = [("tdnnf2.alpha", cf), ("tdnnf2.linear", darts)]
    text = nn.write_nnet(["input-node name=input dim=40"], comps, False).decode()
    got8 = [[float(line.split("[")[1].split(" ")[j + 1]) for j in range(8)]
            for line in text.split("\n") if "alpha <ConstantFunctionComponent>" in line]
    assert len(got8) == 1
    np.testing.assert_allclose(got8[0], alpha8, rtol=1e-5)
    got7 = [[float(item) for item in line.strip().split("[")[1].split(" ")[1:n + 1]]
            for line in text.split("\n") if "<BiasParams>" in line]
    assert len(got7) == 1
    np.testing.assert_allclose(got7[0], alpha7, rtol=1e-5)
    for binary in (False, True):
        _, back = nn.read_nnet(nn.write_nnet([], comps, binary), binary)
        assert [nm for nm, _ in back] == ["tdnnf2.alpha", "tdnnf2.linear"]
        np.testing.assert_allclose(back[0][1].vectorize(), alpha8, rtol=1e-5 if not binary else 0)
        np.testing.assert_allclose(back[1][1].vectorize(), v, rtol=2e-5 if not binary else 0, atol=1e-7)
