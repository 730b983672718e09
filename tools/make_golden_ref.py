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

The two *.config inputs are what steps/nnet3/xconfig_to_configs.py would write for the recipes' xconfig; upstream's
xconfig package is not in the reference tree, so they are laid out here from the format strings of
steps/libs/nnet3/xconfig/composite_layers.py:733-770 (tdnnfdartsv3-layer) and :156-190 (tdnnf-layer)."""
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


def main():
    out = {}
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
