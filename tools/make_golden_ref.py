"""Golden fixtures produced by RUNNING the reference's own Python (needs /root/reference; run in the build container,
the output tests/golden/ref_scripts.json is committed and travels to the GPU box):

  * steps/libs/nnet3/train/temperature_schedule.py::get_temperature_edit_string  -- the per-iteration
    `set-temperature-proportion` directive the training driver emits (imported and called);
  * local/chain_NAS/scripts/add_flopsconstraint.py                               -- the ConstantFunction / {Gumbel}SoftmaxFlops
    component lines of the bottleneck search (executed);
  * local/chain_NAS/scripts/generate_bottleneckCB8share_onehottrain_config.py     -- Onehot / CopyN / TdnnComponent /
    ElementwiseProduct / ReLU / BatchNorm / GeneralDropout lines (executed on a final_ori.config);
  * local/chain_NAS/scripts/generate_config.py                                    -- the TdnnDARTSV3Component lines of the
    context-offset supernet (executed on a final.config_temp).

  * steps/libs/nnet3/xconfig/composite_layers.py::XconfigTdnnfDARTSV3Layer / XconfigTdnnfLayer  -- the component lines
    the xconfig layer classes emit for the recipes' `tdnnfdartsv3-layer` / `tdnnf-layer` lines (imported and run; their
    base class XconfigLayerBase lives in upstream's basic_layers.py, which the reference does not ship, so a stub of it
    that parses the key=value pairs and supplies the input descriptor is put on the import path).

The two *.config inputs are what steps/nnet3/xconfig_to_configs.py would write for the recipes' xconfig: darts_layer() /
stock_layer() below lay them out, and main() checks them line for line against what the reference's layer classes emit."""
import importlib.util
import json
import os
import subprocess
import sys
import tempfile

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPTS = os.path.join(REF, "local", "chain_NAS", "scripts")


def _run(script, *args):
    subprocess.run([sys.executable, os.path.join(SCRIPTS, script), *args], check=True, capture_output=True)


def _lines(path):
    return [l.rstrip("\n") for l in open(path) if l.strip()]


def darts_layer(name, inp, dim=1536, bottleneck=160, flags=("false", "false", "false", "false", "true", "true")):
    g, e, f, a, t, u = flags  # run_TDNN_DARTSV3_fbk_stride_pretrain.sh:124
    common = f"use-gumbel={g} use-entropy={e} free-select={f} update-alpha={a} update-theta={t} uniform-sample={u} Temp-Proportion=1.0"
    return [
        f"component name={name}.linear type=TdnnDARTSV3Component input-dim={dim} output-dim={bottleneck} l2-regularize=0.01 "
        f"max-change=0.75 use-bias=false {common} time-offsets=-1,0 orthonormal-constraint=-1.0",
        f"component-node name={name}.linear component={name}.linear input={inp}",
        f"component name={name}.affine type=TdnnDARTSV3Component input-dim={bottleneck} output-dim={dim} l2-regularize=0.01 "
        f"max-change=0.75 {common} time-offsets=0,1",
        f"component-node name={name}.affine component={name}.affine input={name}.linear",
        f"component name={name}.relu type=RectifiedLinearComponent dim={dim} self-repair-scale=1e-05",
        f"component-node name={name}.relu component={name}.relu input={name}.affine",
        f"component name={name}.batchnorm type=BatchNormComponent dim={dim}",
        f"component-node name={name}.batchnorm component={name}.batchnorm input={name}.relu",
        f"component name={name}.dropout type=GeneralDropoutComponent dim={dim} dropout-proportion=0.0 continuous=true",
        f"component-node name={name}.dropout component={name}.dropout input={name}.batchnorm",
        f"component name={name}.noop type=NoOpComponent dim={dim}",
        f"component-node name={name}.noop component={name}.noop input=Sum(Scale(0.66, {inp}), {name}.dropout)",
    ]


def stock_layer(name, inp, stride, dim=1536, bottleneck=160):
    o1, o2 = (f"-{stride},0", f"0,{stride}") if stride else ("0", "0")
    return [
        f"component name={name}.linear type=TdnnComponent input-dim={dim} output-dim={bottleneck} l2-regularize=0.01 "
        f"max-change=0.75 use-bias=false time-offsets={o1} orthonormal-constraint=-1.0",
        f"component-node name={name}.linear component={name}.linear input={inp}",
        f"component name={name}.affine type=TdnnComponent input-dim={bottleneck} output-dim={dim} l2-regularize=0.01 "
        f"max-change=0.75 time-offsets={o2}",
        f"component-node name={name}.affine component={name}.affine input={name}.linear",
        f"component name={name}.relu type=RectifiedLinearComponent dim={dim} self-repair-scale=1e-05",
        f"component-node name={name}.relu component={name}.relu input={name}.affine",
        f"component name={name}.batchnorm type=BatchNormComponent dim={dim}",
        f"component-node name={name}.batchnorm component={name}.batchnorm input={name}.relu",
        f"component name={name}.dropout type=GeneralDropoutComponent dim={dim} dropout-proportion=0.0 continuous=true",
        f"component-node name={name}.dropout component={name}.dropout input={name}.batchnorm",
        f"component name={name}.noop type=NoOpComponent dim={dim}",
        f"component-node name={name}.noop component={name}.noop input=Sum(Scale(0.66, {inp}), {name}.dropout)",
    ]


STUB_BASE = '''
class XconfigLayerBase(object):
    """Stand-in for upstream steps/libs/nnet3/xconfig/basic_layers.py (not in the reference tree): key=value parsing with
    the defaults' types, and the descriptor of the layer's input."""
    def __init__(self, first_token, key_to_value, prev_names=None):
        self.layer_type = first_token
        self.name = key_to_value["name"]
        self.set_default_configs()
        for k, v in key_to_value.items():
            if k == "name":
                continue
            if k not in self.config:
                raise RuntimeError("unknown option " + k)
            d = self.config[k]
            self.config[k] = (v.lower() == "true") if isinstance(d, bool) else int(v) if isinstance(d, int) else float(v) if isinstance(d, float) else v
        self.descriptors = {"input": {"dim": prev_names["dim"], "final-string": prev_names["name"]}}
        self.set_derived_configs()
        self.check_configs()
'''


def xconfig_layers():
    """Run the reference's xconfig layer classes on the recipes' layer lines."""
    with tempfile.TemporaryDirectory() as d:
        pkg = os.path.join(d, "libs", "nnet3", "xconfig")
        os.makedirs(pkg)
        for sub in ("libs", "libs/nnet3", "libs/nnet3/xconfig"):
            open(os.path.join(d, sub, "__init__.py"), "w").close()
        open(os.path.join(pkg, "basic_layers.py"), "w").write(STUB_BASE)
        sys.path.insert(0, d)
        try:
            spec = importlib.util.spec_from_file_location(
                "ref_composite_layers", os.path.join(REF, "steps", "libs", "nnet3", "xconfig", "composite_layers.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
        finally:
            sys.path.remove(d)
            for m in [m for m in sys.modules if m == "libs" or m.startswith("libs.")]:
                del sys.modules[m]
    # run_TDNN_DARTSV3_fbk_stride_pretrain.sh:124,143 and run_tdnn_7q_fbk_40_manual.sh (tdnnf_opts; time-stride 1 / 0 / 6)
    darts_opts = ("l2-regularize=0.01 dropout-proportion=0.0 bypass-scale=0.66 use-gumbel=false use-entropy=false free-select=false "
                  "update-alpha=false update-theta=true uniform-sample=true")
    stock_opts = "l2-regularize=0.01 dropout-proportion=0.0 bypass-scale=0.66"
    out = {}
    for first, cls, line in (
            ("tdnnfdartsv3-layer", mod.XconfigTdnnfDARTSV3Layer, f"name=tdnnf2 {darts_opts} dim=1536 bottleneck-dim=160 time-stride=6"),
            ("tdnnfdartsv3-layer", mod.XconfigTdnnfDARTSV3Layer, f"name=tdnnf3 {darts_opts} dim=1536 bottleneck-dim=160 time-stride=1"),
            ("tdnnf-layer", mod.XconfigTdnnfLayer, f"name=tdnnf2 {stock_opts} dim=1536 bottleneck-dim=160 time-stride=1"),
            ("tdnnf-layer", mod.XconfigTdnnfLayer, f"name=tdnnf5 {stock_opts} dim=1536 bottleneck-dim=160 time-stride=0"),
            ("tdnnf-layer", mod.XconfigTdnnfLayer, f"name=tdnnf6 {stock_opts} dim=1536 bottleneck-dim=160 time-stride=3")):
        kv = dict(t.split("=", 1) for t in line.split())
        layer = cls(first, kv, {"dim": 1536, "name": "tdnn1.dropout"})
        out[first + " " + line] = [l for name, l in layer.get_full_config() if name == "final"]
    return out


def dropout_edits():
    """The reference keeps upstream's dropout-schedule code in temperature_schedule.py:68-367 as a comment block (its live
    train.py:524-532 calls common_train_lib.get_dropout_edit_string, i.e. upstream's dropout_schedule.py with this very text).
    The block is un-commented mechanically ('# ' stripped) and executed: get_dropout_edit_string(schedule, fraction, iter)."""
    import logging

    src = open(os.path.join(REF, "steps", "libs", "nnet3", "train", "temperature_schedule.py")).read().split("\n")
    body = [(l[2:] if l.startswith("# ") else ("" if l.strip() in ("#", "") else l)) for l in src[67:]]
    ns = {"logger": logging.getLogger("ref_dropout"), "__name__": "ref_dropout"}
    exec(compile("\n".join(body), "temperature_schedule.py[68:] un-commented", "exec"), ns)
    out = []
    # run_TDNN_DARTSV3_fbk_stride_pretrain.sh:49 dropout_schedule, then the forms of the block's own _self_test()
    for sch in ("0,0@0.20,0.5@0.50,0", "0.0,0.5,0.0", "0.0,0.3@0.25,0.0", "0.1,0.4@0.3,0.2@0.8,0.05", "0.2,0.2"):
        for frac in (0.0, 0.1, 0.2, 0.25, 0.3, 0.35, 0.5, 0.6, 0.75, 0.8, 0.9, 1.0):
            out.append([sch, "*", frac, ns["get_dropout_edit_string"](sch, frac, 7)])
    for frac in (0.0, 0.4, 1.0):
        out.append(["0,0.5,0", "tdnnf*", frac, ns["get_dropout_edit_string"]("tdnnf*=0,0.5,0", frac, 7)])
    return out


def main():
    out = {}
    out["dropout_edits"] = dropout_edits()
    out["xconfig_layers"] = xl = xconfig_layers()
    # the hand-laid inputs of the config generators below are what the reference's layer classes emit
    by_name = {k.split()[1] + "/" + k.split()[0] + "/" + [t for t in k.split() if t.startswith("time-stride")][0]: v for k, v in xl.items()}
    assert by_name["name=tdnnf3/tdnnfdartsv3-layer/time-stride=1"] == darts_layer("tdnnf3", "tdnn1.dropout"), "darts_layer() drifted from the reference"
    for nm, stride in (("tdnnf2", 1), ("tdnnf5", 0), ("tdnnf6", 3)):
        assert by_name[f"name={nm}/tdnnf-layer/time-stride={stride}"] == stock_layer(nm, "tdnn1.dropout", stride), "stock_layer() drifted from the reference"
    spec = importlib.util.spec_from_file_location(
        "ref_temperature_schedule", os.path.join(REF, "steps", "libs", "nnet3", "train", "temperature_schedule.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    edits = []
    for num_to_process in (7, 120):
        for done in range(0, num_to_process + 1, max(1, num_to_process // 12)):
            edits.append([done, num_to_process, mod.get_temperature_edit_string(float(done) / num_to_process, done)])
    out["temperature_edits"] = edits

    with tempfile.TemporaryDirectory() as d:
        _run("add_flopsconstraint.py", d, "true", "0.001", "tdnn")
        out["change_config_gumbel"] = _lines(os.path.join(d, "change.config"))
        _run("add_flopsconstraint.py", d, "false", "0.1", "tdnn")
        out["change_config_softmax"] = _lines(os.path.join(d, "change.config"))

        lines, prev = [], "tdnn1.dropout"
        for k, stride in ((2, 1), (3, 1), (4, 1), (5, 0), (6, 3)):
            lines += stock_layer(f"tdnnf{k}", prev, stride)
            prev = f"tdnnf{k}.noop"
        open(os.path.join(d, "final_ori.config"), "w").write("\n".join(lines) + "\n")
        _run("generate_bottleneckCB8share_onehottrain_config.py", d)
        out["bottleneck_final_config"] = _lines(os.path.join(d, "final.config"))

        lines, prev = [], "tdnn1.dropout"
        for k in (2, 3):
            lines += darts_layer(f"tdnnf{k}", prev)
            prev = f"tdnnf{k}.noop"
        for name in ("final.config_temp", "ref.config_temp"):
            open(os.path.join(d, name), "w").write("\n".join(lines) + "\n")
        _run("generate_config.py", "7", d + "/")
        out["supernet_final_config"] = _lines(os.path.join(d, "final.config"))
    path = os.path.join(ROOT, "tests", "golden", "ref_scripts.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
