"""Diagnostic: per-quantity errors of the whole-step parity check (tests/test_gpu_step_parity.py) over a few steps."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from oracle import supernet_ref as R
from tdnnf_nas_b200 import nnet3, synth
from tdnnf_nas_b200.supernet import Supernet, SupernetConfig
from tests.test_gpu_step_parity import _params_of
from tests.util import rel_err

for planes in (True, False):
    for rep in range(2):
        cfg = SupernetConfig(num_seqs=8, frames_per_eg=30, dim=256, bottleneck=160, num_blocks=3, prefinal_small=64, num_pdfs=200,
                             den_states=300, den_out_degree=6.0, mode="search", learning_rate=2e-3, darts_lr_factor=0.05, xent=True,
                             tail_planes=planes)
        nnet3.set_keep_planes(planes)
        net = Supernet(cfg)
        S, T, P, L, n = cfg.num_seqs, net.T, cfg.num_pdfs, cfg.num_blocks, cfg.num_offsets
        den_graph = synth.make_den_graph(cfg.den_states, P, cfg.den_out_degree, seed=5)
        num_graph = synth.make_num_graphs(S, P, T, seed=60, den_graph=den_graph)
        rcfg = R.RefConfig(num_seqs=S, frames_per_eg=cfg.frames_per_eg, feat_dim=cfg.feat_dim, dim=cfg.dim, bottleneck=cfg.bottleneck,
                           num_blocks=L, num_offsets=n, prefinal_small=cfg.prefinal_small, num_pdfs=P, xent=True,
                           learning_rate=cfg.learning_rate, darts_lr_factor=cfg.darts_lr_factor)
        ref = R.CpuSupernet(rcfg, den_graph, num_graph, _params_of(net))
        for step in range(3):
            x = net.make_input(step)
            c0 = nnet3.get_rand_counter()
            objf_gpu = net.step(x.pin_memory(), apply_update=False)
            c1 = nnet3.get_rand_counter()
            nnet3.set_rand_counter(c0)
            u = [np.array([nnet3.rand_uniform() for _ in range(n)], np.float32) for _ in range(2 * L)]
            objf_ref = ref.step(x.numpy(), u, apply_update=False)
            errs = {"objf": abs(objf_gpu - objf_ref) / abs(objf_ref), "out": rel_err(net.head["out"].cpu().numpy(), ref.st["out"])}
            for b, blk in enumerate(net.blocks):
                errs[f"{b}.d_aff"] = rel_err(blk["d_aff"].cpu().numpy(), ref.st[b]["d_aff"])
                for h in ("lin", "aff"):
                    dv = blk[h + "_delta"].vectorize()
                    dW, db = ref.delta[(b, h)]
                    errs[f"{b}.{h}.theta"] = rel_err(dv[: dW.size].reshape(dW.shape), dW)
                    errs[f"{b}.{h}.alpha"] = float(np.abs(dv[dW.size: dW.size + n] - db[:n]).max() / (np.abs(db[:n]).max() + 1e-30))
                    errs[f"{b}.{h}.alpha_vals"] = [float(v) for v in db[:n]]
                    errs[f"{b}.{h}.alpha_gpu"] = [float(v) for v in dv[dW.size: dW.size + n]]
            net._update_with_max_change()
            ref.update()
            print(json.dumps(dict(planes=planes, rep=rep, step=step, **{k: (round(v, 6) if isinstance(v, float) else v) for k, v in errs.items()})), flush=True)
        net.close()
nnet3.set_keep_planes(True)
