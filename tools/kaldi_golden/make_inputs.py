"""Inputs of tools/kaldi_golden/dump-nas-golden.cc that cannot be written down as a formula: the denominator graphs, as binary
OpenFst files (what chain-make-den-fst writes).  python tools/kaldi_golden/make_inputs.py <dir>
tests/test_kaldi_golden.py regenerates the same graphs from the same seeds (DEN_CASES) when it checks the dumped vectors."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

# case name -> (num_states, num_pdfs, mean out-degree, seed, num_sequences, frames); the last two are fixed in the .cc file
DEN_CASES = {"den_small": (60, 37, 4.0, 31, 6, 9), "den_medium": (700, 211, 8.0, 32, 16, 17)}


def den_graph(name):
    from tdnnf_nas_b200 import synth

    n, p, deg, seed, _, _ = DEN_CASES[name]
    return synth.make_den_graph(n, p, deg, seed=seed)


def main():
    from tdnnf_nas_b200 import synth
    from tests import egs_ref as W

    out = sys.argv[1]
    os.makedirs(out, exist_ok=True)
    for name in DEN_CASES:
        g = den_graph(name)
        path = os.path.join(out, name + ".den.fst")
        open(path, "wb").write(W.fst_vector(W.den_graph_to_fst(g, synth.den_graph_to_fst_text(g))))
        print(path, g["num_states"], "states", g["num_arcs"], "arcs")


if __name__ == "__main__":
    main()
