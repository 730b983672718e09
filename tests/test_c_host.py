"""The C ABI from a plain C99 host (tests/chost/egs_host.c): both headers compile as C, the program links against the
shared library alone (no Python, torch or C++ runtime on its side), reads an archive and a den.fst and prints what the
ctypes binding sees in the same files.  Without a GPU the program must report the library's refusal to create a context."""
import os
import subprocess

import numpy as np
import pytest

from tdnnf_nas_b200 import capi, synth

from . import egs_ref as W
from .test_egs_io import synth_supervision_examples

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_headers_are_c99():
    for h in ("tdnnf_nas_b200.h", "tdnnf_nnet3.h"):
        r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c", "-"],
                           input=f'#include "{h}"\n', capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_c_host_reads_examples_and_graphs(tmp_path):
    import torch

    capi.load()
    exe = str(tmp_path / "egs_host")
    libdir = os.path.dirname(capi.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "chost", "egs_host.c"),
                        "-o", exe, "-L", libdir, "-ltdnnf_nas_b200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    S, P, T, N = 6, 23, 8, 70
    rng = np.random.default_rng(17)
    _, exs = synth_supervision_examples(rng, S, P, T, seed=2)
    blob = W.ark(exs, True)
    dgraph = synth.make_den_graph(N, P, 4.0, seed=1)
    fst = dict(start=None, num_states=N, arcs=[], finals={})
    for line in synth.den_graph_to_fst_text(dgraph).splitlines():
        f = line.split()
        if len(f) >= 4:
            fst["arcs"].append((int(f[0]), int(f[1]), int(f[2]), float(f[4]) if len(f) > 4 else 0.0))
            if fst["start"] is None:
                fst["start"] = int(f[0])
        elif f:
            fst["finals"][int(f[0])] = float(f[1]) if len(f) > 1 else 0.0
    (tmp_path / "egs.ark").write_bytes(blob)
    (tmp_path / "den.fst").write_bytes(W.fst_vector(fst))
    r = subprocess.run([exe, str(tmp_path / "egs.ark"), str(tmp_path / "den.fst"), str(P)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    out = dict(line.split(" ", 1) for line in r.stdout.strip().splitlines())
    egs = capi.ChainEgs(blob)
    x, t0 = egs.merge_input(0, S, "input")
    m = egs.merge_supervision(0, S, "output", P)
    checksum = float((x.reshape(-1).astype(np.float64) * ((np.arange(x.size) % 7) + 1)).sum())
    assert out["abi"] == "1002" and out["examples"] == str(S)
    f = out["input"].split()
    assert [int(v) for v in f[:3]] == list(x.shape) and int(f[4]) == t0 and float(f[6]) == pytest.approx(checksum, rel=1e-9, abs=1e-5)
    f = out["supervision"].split()
    assert int(f[1]) == S and int(f[3]) == T and float(f[7]) == pytest.approx(float(m["deriv_weights"].sum()), abs=1e-4)
    f = out["numerator"].split()
    assert int(f[1]) == S and int(f[3]) == m["num_graph"]["num_arcs"] and int(f[5]) == int(m["num_graph"]["state_offsets"][-1])
    f = out["denominator"].split()
    assert int(f[1]) == N and int(f[3]) == P and int(f[5]) == 2 * dgraph["num_arcs"]
    if torch.cuda.is_available():
        assert "device" in out and out["device"].startswith("graphs created")
    else:
        assert "no CPU path" in out["no"]
