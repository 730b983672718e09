"""Static check of the rule programmatic dependent launch rests on: a kernel launched through launch_pdl begins
with griddepcontrol.wait -- no thread returns and nothing is read from or written to global memory before it -- and a kernel
without the wait is never launched through launch_pdl.  Source-level (no GPU): the kernels are found by their __global__
definitions, the launch sites by the first argument of launch_pdl / the name before <<<."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "tdnn-f_nas_b200", "csrc")


def _kernels():
    out = {}
    for f in sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh"))):
        text = open(f).read()
        for m in re.finditer(r"__global__[^{;]*?(\w+)\s*\(([^)]*)\)\s*\{", text, re.S):
            depth, i = 1, m.end()
            while depth and i < len(text):
                depth += (text[i] == "{") - (text[i] == "}")
                i += 1
            out[m.group(1)] = (os.path.basename(f), text[m.end():i])
    return out


def test_every_pdl_launched_kernel_waits_first():
    kernels = _kernels()
    assert len(kernels) > 80
    sources = "".join(open(f).read() for f in glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")))
    pdl_launched = set(re.findall(r"launch_pdl(?:_at)?\(\s*(?:__FILE__\s*,\s*)?(\w+)", sources))
    # kernels chosen through a local function pointer / lambda are launched under that local name: every kernel that has
    # the wait counts as PDL-capable, and the ones without must show up as plain <<< >>> or cluster launches only
    waits = {k for k, (_, body) in kernels.items() if "grid_dep_wait" in body or "griddepcontrol.wait" in body}
    no_wait = set(kernels) - waits
    assert not (no_wait & pdl_launched), no_wait & pdl_launched
    for k in no_wait:
        assert re.search(r"\b%s\b(?:<[^;]*?>)?\s*<<<|launch_cluster\(\s*%s\b|cudaLaunchCooperativeKernel|cudaLaunchKernelEx\(&cfg,\s*%s\b" % (k, k, k), sources), \
            f"{k} has no griddepcontrol.wait and no plain launch site was found"
    bad = []
    for k in sorted(waits):
        f, body = kernels[k]
        w = min(x for x in (body.find("grid_dep_wait"), body.find("griddepcontrol.wait")) if x >= 0)
        pre = re.sub(r"//[^\n]*", "", body[:w])
        if re.search(r"\breturn\b", pre):
            bad.append((f, k, "a thread can return before the wait"))
        if re.search(r"\w\[[^\]]*\]\s*(?:[+\-*]?=)(?!=)|=\s*[\w.>-]+\[[^\]]*\]|__ldg|atomic\w+\(|red\.global|ld\.global|st\.global", pre):
            bad.append((f, k, "memory access before the wait: " + pre.strip()[-120:]))
    assert not bad, bad
