import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import oracle as O
from tests import np_ref
from tests.util import rel_err
from tdnnf_nas_b200 import capi, nnet3 as nn
ctx = capi.Context(0); ctx.use_current_stream(); nn.set_context(ctx)
def _data(g, N, D, k=6):
    basis = g.standard_normal((k, D))
    return ((g.standard_normal((N, k)) * np.linspace(3.0, 1.0, k)) @ basis + 0.3 * g.standard_normal((N, D))).astype(np.float32)
for (N, D, rank, k) in [(256, 8, 40, 2), (256, 8, 4, 2), (384, 41, 30, 3)]:
    g = np.random.default_rng(11)
    ng = nn.NaturalGradient(rank, 1, 2000.0, 4.0); orc = O.NaturalGradient(rank, 1, 2000.0, 4.0); ref = np_ref.NaturalGradientF64(rank, 1, 2000.0, 4.0)
    print("case", N, D, rank)
    for step in range(8):
        X = _data(g, N, D, k)
        Xo = X.copy(); so = orc.precondition(Xo)
        Xr, sr = ref.precondition(X)
        Xd = torch.from_numpy(X).cuda(); sg = ng.precondition(Xd)
        Xg = Xd.cpu().numpy()
        st_g, st_o = ng.state(), orc.state()
        print(step, "gpu-vs-f64 %.2e  orc-vs-f64 %.2e  scale g %.6f o %.6f r %.6f  reorth g %d o %d  rho g %.5g o %.5g r %.5g" % (
            rel_err(Xg, Xr), rel_err(Xo, Xr), sg, so, sr, st_g["num_reorth"], st_o["num_reorth"], st_g["rho"], st_o["rho"], ref.rho))
        print("    d g", np.round(st_g["d"][:8], 4), "\n    d o", np.round(st_o["d"][:8], 4))
