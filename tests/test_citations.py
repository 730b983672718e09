"""Every `file:line[-line]` citation of the reference in the sources and documents points into a file that exists under
/root/reference with at least that many lines (skipped where the reference tree is absent, e.g. on the GPU box).
Citations marked `kaldi:` are to upstream Kaldi, which the reference does not ship, and are not checked."""
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
# shorthands of SURVEY.md / DESIGN.md
SHORT = {"tdnn.cc": "src/nnet3/nnet-tdnn-component.cc", "conv.h": "src/nnet3/nnet-convolutional-component.h",
         "simple.cc": "src/nnet3/nnet-simple-component.cc", "simple.h": "src/nnet3/nnet-simple-component.h",
         "norm.cc": "src/nnet3/nnet-normalize-component.cc", "norm.h": "src/nnet3/nnet-normalize-component.h",
         "itf.cc": "src/nnet3/nnet-component-itf.cc", "utils.cc": "src/nnet3/nnet-utils.cc",
         "cvupdate.sh": "local/chain_NAS/run_TDNN_DARTSV3_fbk_stride_cvupdate.sh",
         "pretrain.sh": "local/chain_NAS/run_TDNN_DARTSV3_fbk_stride_pretrain.sh"}


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is not on this machine")
def test_reference_citations_are_in_range():
    by_name = {}
    for d, _, files in os.walk(REF):
        for f in files:
            by_name.setdefault(f, []).append(os.path.join(d, f))
    own = {os.path.basename(p) for p in glob.glob(os.path.join(ROOT, "**", "*"), recursive=True)}
    sources = (glob.glob(os.path.join(ROOT, "include", "*.h")) + glob.glob(os.path.join(ROOT, "tdnn-f_nas_b200", "csrc", "**", "*.c*"), recursive=True)
               + glob.glob(os.path.join(ROOT, "tdnn-f_nas_b200", "csrc", "**", "*.h"), recursive=True) + glob.glob(os.path.join(ROOT, "tdnn-f_nas_b200", "*.py"))
               + glob.glob(os.path.join(ROOT, "oracle", "*.*")) + glob.glob(os.path.join(ROOT, "tests", "*.py")) + glob.glob(os.path.join(ROOT, "tools", "*.py"))
               + [os.path.join(ROOT, f) for f in ("DESIGN.md", "INTEGRATION.md", "README.md", "bench.py")])
    lines_of = {}
    checked, bad = 0, []
    for src in sources:
        text = open(src, errors="replace").read()
        for m in re.finditer(r"(?<![\w/.-])((?:[\w.-]+/)*[\w-]+\.(?:cc|h|py|sh)):(\d+)(?:-(\d+))?", text):
            path, a, b = m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
            before = text[max(0, m.start() - 12):m.start()]
            if "kaldi:" in before or "kaldi " in before:
                continue
            base = os.path.basename(path)
            if path in SHORT:
                cands = [os.path.join(REF, SHORT[path])]
            elif os.path.exists(os.path.join(REF, path)):
                cands = [os.path.join(REF, path)]
            else:
                cands = by_name.get(base, [])
            if not cands:
                if base not in own:
                    bad.append((os.path.relpath(src, ROOT), m.group(0), "no such file in the reference or here"))
                continue
            checked += 1
            for c in cands:
                if c not in lines_of:
                    lines_of[c] = sum(1 for _ in open(c, errors="replace"))
            if a > b or b > max(lines_of[c] for c in cands):
                bad.append((os.path.relpath(src, ROOT), m.group(0), "beyond the end of the file"))
    assert checked > 300, checked
    assert not bad, bad
