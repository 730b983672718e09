"""Sign tie or race?  Runs the small bottleneck-search trajectory several times in one process and, for every pair of runs
whose final alphas differ, reports where the ReLU masks first differ and how close to zero the pre-activations there are."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tdnnf_nas_b200.supernet import Supernet, SupernetConfig  # noqa: E402

seed = int(os.environ.get("TDNNF_TEST_INPUT_SEED", "0"))
runs = []
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    cfg = SupernetConfig(num_seqs=8, frames_per_eg=30, dim=128, bottleneck=24, num_blocks=3, prefinal_small=64, num_pdfs=200,
                         den_states=300, den_out_degree=6.0, mode="bottleneck", learning_rate=2e-3,
                         candidate_widths=(2, 2, 3, 3, 2, 4, 4, 4), flops_coef=0.1, bottleneck_gumbel=True, fuse_mask=True,
                         strides=[1, 0, 3])
    net = Supernet(cfg)
    x = net.make_input(seed).pin_memory()
    rec = []
    for step in range(3):
        net.step(x)
        rec.append(dict(pre=[blk["aff_out"].cpu().numpy().copy() for blk in net.blocks],
                        head=[net.head[k].cpu().numpy().copy() for k in ("pa", "xa") if k in net.head],
                        out=net.head["out"].cpu().numpy().copy(),
                        alpha=np.stack([blk["alpha"].vectorize() for blk in net.blocks])))
    runs.append(rec)
    net.close()
report = []
for i in range(len(runs)):
    for j in range(i + 1, len(runs)):
        a, b = runs[i][-1]["alpha"], runs[j][-1]["alpha"]
        d = float(np.abs(a - b).max() / np.abs(b).max())
        item = dict(pair=[i, j], alpha_rel_diff=d)
        if d > 1e-5:
            for step in range(3):
                for name, ta, tb in ([(f"blk{k}", runs[i][step]["pre"][k], runs[j][step]["pre"][k]) for k in range(3)] +
                                     [(f"head{k}", runs[i][step]["head"][k], runs[j][step]["head"][k]) for k in range(len(runs[i][step]["head"]))]):
                    mism = (ta > 0) != (tb > 0)
                    item.setdefault("steps", []).append(dict(step=step, tensor=name, mask_mismatches=int(mism.sum()),
                                                             max_abs_at_mismatch=float(np.abs(ta[mism]).max()) if mism.any() else 0.0,
                                                             rms=float(np.sqrt((ta.astype(np.float64) ** 2).mean())),
                                                             value_rel_diff=float(np.abs(ta - tb).max() / (np.abs(tb).max() + 1e-30)),
                                                             alpha_rel_diff=float(np.abs(runs[i][step]["alpha"] - runs[j][step]["alpha"]).max() / np.abs(runs[j][step]["alpha"]).max())))
        report.append(item)
print(json.dumps(report, indent=None))
