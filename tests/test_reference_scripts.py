"""Fixtures produced by RUNNING the reference's own Python (tools/make_golden_ref.py -> tests/golden/ref_scripts.json):
the per-iteration temperature directive, and the component config lines its config generators emit.  Every such line must
be accepted by the matching component here, with the dimensions / offsets / flags it asks for -- this is the one place
where outputs of reference code (not of the oracle) pin this library."""
import json
import os
import re

import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_scripts.json")
HOST_TYPES = {"GumbelSoftmaxFlopsComponent", "SoftmaxFlopsComponent", "CopyNComponent", "ElementwiseProductComponent",
              "RectifiedLinearComponent", "BatchNormComponent", "GeneralDropoutComponent"}
DEVICE_TYPES = {"TdnnComponent", "TdnnDARTSV3Component", "ConstantFunctionComponent", "OnehotFunctionComponent"}
GRAPH_ONLY = {"NoOpComponent"}  # descriptor plumbing of the nnet3 graph, not a component of the path


def _fixture():
    return json.load(open(GOLD))


def _component_lines(lines):
    """[(name, type, rest-of-config)] of the `component name=... type=...` lines (what nnet3-init hands to InitFromConfig)."""
    out = []
    for line in lines:
        m = re.match(r"component name=(\S+) type=(\S+)\s*(.*)$", line)
        if m:
            out.append((m.group(1), m.group(2), m.group(3)))
    return out


def _kv(rest):
    return dict(tok.split("=", 1) for tok in rest.split())


def test_temperature_directive_matches_the_reference_schedule():
    from tdnnf_nas_b200 import nnet3

    fx = _fixture()["temperature_edits"]
    assert len(fx) >= 20
    comp = nnet3.Component.new("GumbelSoftmaxFlopsComponent", "dim=8 scale=0.001 temp-proportion=1.0")
    for done, total, ref_string in fx:
        ours = nnet3.temperature_edit_string(done, total)
        assert "nnet3-copy --edits='" + ours + "' - - |" == ref_string  # character for character (float repr included)
        nnet3.apply_edits(ours, [("tdnnf2.softmax", comp)])
        assert comp.temp_proportion() == pytest.approx(float(ours.rsplit("=", 1)[1]), rel=1e-6)
    assert fx[0][2].endswith("proportion=1.0' - - |") and "proportion=0.03' - - |" in fx[-1][2]


def test_every_generated_component_type_is_known():
    fx = _fixture()
    types = set()
    for key in ("change_config_gumbel", "change_config_softmax", "bottleneck_final_config", "supernet_final_config"):
        types |= {t for _, t, _ in _component_lines(fx[key])}
    assert types <= HOST_TYPES | DEVICE_TYPES | GRAPH_ONLY, types - (HOST_TYPES | DEVICE_TYPES | GRAPH_ONLY)
    assert {"TdnnDARTSV3Component", "GumbelSoftmaxFlopsComponent", "SoftmaxFlopsComponent", "CopyNComponent",
            "OnehotFunctionComponent", "ConstantFunctionComponent"} <= types


def test_generated_lines_of_parameter_free_components_are_accepted():
    """No device needed: the mixing components, CopyN, ElementwiseProduct, ReLU, BatchNorm, GeneralDropout."""
    from tdnnf_nas_b200 import nnet3

    fx = _fixture()
    seen = set()
    copyn_dims = {}
    for key in ("change_config_gumbel", "change_config_softmax", "bottleneck_final_config", "supernet_final_config"):
        for name, typ, rest in _component_lines(fx[key]):
            if typ not in HOST_TYPES:
                continue
            comp = nnet3.Component.new(typ, rest)
            kv = _kv(rest)
            assert comp.type() == typ
            seen.add(typ)
            if "dim" in kv:
                assert comp.input_dim() == comp.output_dim() == int(kv["dim"])
            if "input-dim" in kv:
                assert comp.input_dim() == int(kv["input-dim"]) and comp.output_dim() == int(kv["output-dim"])
            if typ == "GumbelSoftmaxFlopsComponent":
                assert comp.temp_proportion() == pytest.approx(float(kv["temp-proportion"]))
                assert ("<Scale> " + kv["scale"]).encode() in comp.write(False)
            if typ == "SoftmaxFlopsComponent":
                assert ("<Scale> " + kv["scale"]).encode() in comp.write(False)
            if typ == "CopyNComponent":
                m = re.match(r"(tdnnf\d)(\d)\.copyn$", name)
                copyn_dims.setdefault(m.group(1), []).append(comp.output_dim())
            if typ == "GeneralDropoutComponent":
                assert comp.dropout_proportion() == 0.0 and b"<Continuous>" in comp.write(False)
    assert seen == HOST_TYPES
    # the shared bottleneck candidates: blocks of 25/25/30/20/20/40/40/40 = 240
    assert copyn_dims and all(v == [25, 25, 30, 20, 20, 40, 40, 40] for v in copyn_dims.values())


def test_flops_vector_is_the_cumulative_bottleneck_width():
    """The hard-coded FLOPs vector of {Gumbel}SoftmaxFlopsComponent (simple.cc:10145-10152: -25 ... -240) is minus the
    cumulative width of the shared bottleneck candidates -- the CopyN blocks the reference's config generator emits
    (25, 25, 30, 20, 20, 40, 40, 40), and the widths bottleneckdim_search_top_model_size.py:62 lists."""
    import numpy as np

    from oracle import oracle as O

    fx = _fixture()
    dims = [int(_kv(rest)["output-dim"]) for name, typ, rest in _component_lines(fx["bottleneck_final_config"])
            if typ == "CopyNComponent" and name.startswith("tdnnf2")]
    widths = np.cumsum(dims).tolist()
    assert widths == [25, 50, 80, 100, 120, 160, 200, 240]
    # read the vector back from the oracle: out_deriv = 0, one row of 8 columns, scale = rows * cols = 8  =>  e = f
    p = np.full((1, 8), 0.125, np.float32)
    _, e = O.softmax_flops_bwd(p, np.zeros((1, 8), np.float32), 8.0, False, 1.0)
    assert (-np.asarray(e).ravel()).tolist() == [float(w) for w in widths]


def test_xconfig_layer_classes_emit_the_lines_the_generators_were_fed():
    """The reference's own xconfig layer classes (composite_layers.py: XconfigTdnnfDARTSV3Layer, XconfigTdnnfLayer), RUN on the
    recipes' layer lines: every component they emit is a known type, the parameter-free ones are accepted as they are, and
    the TdnnDARTSV3Component / TdnnComponent lines carry the keys InitFromConfig reads here (the device-side acceptance of
    such lines is tests/test_zz_gpu_reference_compat.py, on the generators' output, which main() of tools/make_golden_ref.py
    checks line for line against these)."""
    from tdnnf_nas_b200 import nnet3

    fx = _fixture()["xconfig_layers"]
    assert len(fx) == 5
    for layer_line, lines in fx.items():
        comps = _component_lines(lines)
        assert [t for _, t, _ in comps] == [("TdnnDARTSV3Component" if layer_line.startswith("tdnnfdartsv3") else "TdnnComponent")] * 2 + [
            "RectifiedLinearComponent", "BatchNormComponent", "GeneralDropoutComponent", "NoOpComponent"]
        stride = int(re.search(r"time-stride=(\d+)", layer_line).group(1))
        for name, typ, rest in comps:
            kv = _kv(rest)
            if typ in HOST_TYPES:
                comp = nnet3.Component.new(typ, rest)
                assert comp.input_dim() == comp.output_dim() == 1536
            elif typ in DEVICE_TYPES:
                known = {"input-dim", "output-dim", "l2-regularize", "max-change", "use-bias", "time-offsets", "orthonormal-constraint",
                         "use-gumbel", "use-entropy", "free-select", "update-alpha", "update-theta", "uniform-sample", "Temp-Proportion"}
                assert set(kv) <= known, set(kv) - known
                want = ("0" if stride == 0 else (f"-{stride},0" if name.endswith(".linear") else f"0,{stride}"))
                assert kv["time-offsets"] == want
                assert (kv.get("use-bias") == "false") == name.endswith(".linear")
        # the bypass the step fuses into its tail kernels: Sum(Scale(0.66, input), dropout output)
        assert lines[-1].endswith(f"input=Sum(Scale(0.66, tdnn1.dropout), {comps[0][0].split('.')[0]}.dropout)")


def test_dropout_directive_matches_the_reference_schedule_code():
    """The dropout half of the per-iteration edit string (train.py:524-532): fixture from executing the dropout-schedule code
    the reference keeps (as a comment block, un-commented mechanically) in temperature_schedule.py:68-367 -- the recipes'
    schedule '0,0@0.20,0.5@0.50,0' (run_TDNN_DARTSV3_fbk_stride_pretrain.sh:49) and the block's own self-test forms.
    Character for character, float repr included; the directive is then applied to a GeneralDropoutComponent."""
    from tdnnf_nas_b200 import nnet3

    fx = _fixture()["dropout_edits"]
    assert len(fx) >= 60
    comp = nnet3.Component.new("GeneralDropoutComponent", "dim=8 dropout-proportion=0.0 continuous=true")
    for schedule, pattern, fraction, ref_string in fx:
        ours = nnet3.dropout_edit_string(schedule, fraction, pattern)
        assert "nnet3-copy --edits='" + ours + "' - - |" == ref_string
        nnet3.apply_edits(ours, [("tdnnf2.dropout", comp)])
        assert comp.dropout_proportion() == pytest.approx(float(ours.rsplit("=", 1)[1]), rel=1e-6, abs=1e-9)
