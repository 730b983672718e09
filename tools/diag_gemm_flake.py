"""Diagnostic: the parity test's flow (GPU step, then the CPU reference step in the same process) for several steps, with a
GPU-only check of the head's data-gradient GEMMs (float64 products of the GPU's own operands) after every step."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from oracle import supernet_ref as R
from tdnnf_nas_b200 import nnet3, synth
from tdnnf_nas_b200.supernet import Supernet, SupernetConfig
from tests.test_gpu_step_parity import _params_of

mode = sys.argv[1] if len(sys.argv) > 1 else "oracle"
cfg = SupernetConfig(num_seqs=8, frames_per_eg=30, dim=256, bottleneck=160, num_blocks=3, prefinal_small=64, num_pdfs=200,
                     den_states=300, den_out_degree=6.0, mode="search", learning_rate=2e-3, darts_lr_factor=0.05, xent=False)
net = Supernet(cfg)
S, T, P, L, n = cfg.num_seqs, net.T, cfg.num_pdfs, cfg.num_blocks, cfg.num_offsets
den_graph = synth.make_den_graph(cfg.den_states, P, cfg.den_out_degree, seed=5)
num_graph = synth.make_num_graphs(S, P, T, seed=60, den_graph=den_graph)
rcfg = R.RefConfig(num_seqs=S, frames_per_eg=cfg.frames_per_eg, feat_dim=cfg.feat_dim, dim=cfg.dim, bottleneck=cfg.bottleneck,
                   num_blocks=L, num_offsets=n, prefinal_small=cfg.prefinal_small, num_pdfs=P, xent=False,
                   learning_rate=cfg.learning_rate, darts_lr_factor=cfg.darts_lr_factor)
ref = R.CpuSupernet(rcfg, den_graph, num_graph, _params_of(net))
hd, st = net.head, net.stock
out = []
for step in range(5):
    x = net.make_input(step)
    c0 = nnet3.get_rand_counter()
    net.step(x.pin_memory(), apply_update=False)
    nnet3.set_rand_counter(c0)
    u = [np.array([nnet3.rand_uniform() for _ in range(n)], np.float32) for _ in range(2 * L)]
    torch.cuda.synchronize()
    chk = hd["d_pa"].double() @ st["pc_affine"]["W"].double()
    e_gpu = float((hd["d_pl"].double() - chk).norm() / chk.norm())
    e_ref = None
    if mode == "oracle":
        ref.step(x.numpy(), u, apply_update=False)
        e_ref = float(np.linalg.norm(net.blocks[-1]["d_aff"].cpu().numpy().astype(np.float64) - ref.st[L - 1]["d_aff"]) /
                      np.linalg.norm(ref.st[L - 1]["d_aff"]))
        ref.update()
    net._update_with_max_change()
    out.append(dict(step=step, pc_affine_dgrad_vs_f64=round(e_gpu, 8), d_aff_vs_cpu=None if e_ref is None else round(e_ref, 8)))
print(json.dumps(out))
net.close()
