import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.util import rel_err, max_rel_to_scale
from tdnnf_nas_b200 import capi
ctx = capi.Context(0); ctx.use_current_stream()
def _data(g, N, D, k=6):
    basis = g.standard_normal((k, D))
    return ((g.standard_normal((N, k)) * np.linspace(3.0, 1.0, k)) @ basis + 0.3 * g.standard_normal((N, D))).astype(np.float32)
one = torch.ones(1, device="cuda")
for (N, D, R) in [(384, 41, 30), (256, 8, 7), (1000, 1537, 20)]:
    g = np.random.default_rng(1)
    X = _data(g, N, D, 3); W = (g.standard_normal((R, D)) * 0.2).astype(np.float32)
    Xd, Wd = torch.from_numpy(X).cuda(), torch.from_numpy(W).cuda()
    H = torch.zeros((N, R), device="cuda")
    ctx.darts_propagate(Xd, H, Wd, None, 1, one, [0], 1)
    Href = X.astype(np.float64) @ W.astype(np.float64).T
    H32 = (X @ W.T)
    print(N, D, R, "H: gpu %.2e (max %.2e)  numpy-fp32 %.2e" % (rel_err(H.cpu().numpy(), Href), max_rel_to_scale(H.cpu().numpy(), Href), rel_err(H32, Href)))
    Hn = H.cpu().numpy()
    J = torch.zeros((R, D), device="cuda")
    ctx.darts_backprop_params(Xd, H, None, J, None, one, [0], 1, 1.0, None)
    Jref = Hn.astype(np.float64).T @ X.astype(np.float64)
    print("   J: gpu %.2e (max %.2e) numpy-fp32 %.2e" % (rel_err(J.cpu().numpy(), Jref), max_rel_to_scale(J.cpu().numpy(), Jref), rel_err(Hn.T @ X, Jref)))
    L = torch.zeros((R, R), device="cuda")
    ctx.darts_backprop_params(H, H, None, L, None, one, [0], 1, 1.0, None)
    Lref = Hn.astype(np.float64).T @ Hn.astype(np.float64)
    print("   L: gpu %.2e (max %.2e)" % (rel_err(L.cpu().numpy(), Lref), max_rel_to_scale(L.cpu().numpy(), Lref)))
    Jn = J.cpu().numpy()
    K = torch.zeros((R, R), device="cuda")
    ctx.darts_propagate(J, K, J, None, 1, one, [0], 1)
    Kref = Jn.astype(np.float64) @ Jn.astype(np.float64).T
    print("   K: gpu %.2e (max %.2e)" % (rel_err(K.cpu().numpy(), Kref), max_rel_to_scale(K.cpu().numpy(), Kref)))
    # X - H W
    Xh = Xd.clone()
    m1 = -torch.ones(1, device="cuda")
    ctx.darts_backprop_data(H, Xh, Wd, m1, [0], 1)
    Xref = X.astype(np.float64) - Hn.astype(np.float64) @ W.astype(np.float64)
    print("   Xhat: gpu %.2e" % rel_err(Xh.cpu().numpy(), Xref))
