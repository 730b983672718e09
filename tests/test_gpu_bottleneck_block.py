"""BASELINE.json configs[3]: the bottleneck-dimension search block
(generate_bottleneckCB8share_onehottrain_config.py:11-90 + add_flopsconstraint.py:15-30), run through the
component API the way the nnet3 graph wires it, against a float64 numpy restatement of the whole block:

  lda -> alpha (ConstantFunctionComponent 220->8) -> {Gumbel}SoftmaxFlopsComponent(8)
      -> dim-range j, Sum(p_j..p_7) -> CopyNComponent(1 -> b_j)          b = 25,25,30,20,20,40,40,40
  Append(copyn_j, linear_j) -> ElementwiseProductComponent -> Append all 8 (240 wide)

The descriptor-level ops (dim-range, Sum, Append) are Kaldi graph plumbing, done here with tensor slices."""
import numpy as np
import pytest

from tests.util import rel_err

pytestmark = pytest.mark.gpu

BLOCKS = [25, 25, 30, 20, 20, 40, 40, 40]
FLOPS = np.array([-25, -50, -80, -100, -120, -160, -200, -240], dtype=np.float64)


@pytest.mark.parametrize("gumbel,scale", [(True, 0.1), (False, 0.001), (True, 0.0)])
def test_bottleneck_search_block(ctx, gumbel, scale):
    import torch

    from tdnnf_nas_b200 import nnet3 as nn

    nn.set_context(ctx)
    nn.set_rand_seed(99)
    g = np.random.default_rng(3)
    R, T = 384, 0.7
    lr = 0.05
    alpha = nn.Component.new("ConstantFunctionComponent",
                             f"input-dim=220 output-dim=8 is-updatable=true use-natural-gradient=false learning-rate={lr} output-stddev=1.0")
    soft = (nn.Component.new("GumbelSoftmaxFlopsComponent", f"dim=8 scale={scale} temp-proportion={T}") if gumbel
            else nn.Component.new("SoftmaxFlopsComponent", f"dim=8 scale={scale}"))
    copyn = [nn.Component.new("CopyNComponent", f"input-dim=1 output-dim={b}") for b in BLOCKS]
    prod = [nn.Component.new("ElementwiseProductComponent", f"input-dim={2 * b} output-dim={b}") for b in BLOCKS]
    a = alpha.vectorize().astype(np.float64)
    lin = g.standard_normal((R, 240)).astype(np.float32)
    d_out = (g.standard_normal((R, 240)) / R).astype(np.float32)
    dev = "cuda"
    # ---------------- forward
    lda = torch.zeros((R, 220), device=dev)
    a_rows = torch.zeros((R, 8), device=dev)
    alpha.propagate(None, lda, a_rows)
    c0 = nn.get_rand_counter()
    p = torch.zeros((R, 8), device=dev)
    soft.propagate(None, a_rows, p)
    nn.set_rand_counter(c0)
    u = np.array([nn.rand_uniform() for _ in range(8)]) if gumbel else None
    lin_d = torch.from_numpy(lin).to(dev)
    out = torch.zeros((R, 240), device=dev)
    prod_in, off = [], 0
    for j, b in enumerate(BLOCKS):
        mask_in = p[:, j:].sum(dim=1, keepdim=True).contiguous()       # Sum(softmax_j .. softmax_7) descriptor
        cn = torch.zeros((R, b), device=dev)
        copyn[j].propagate(None, mask_in, cn)                          # kPropagateAdds on a zeroed matrix
        pin = torch.cat([cn, lin_d[:, off:off + b]], dim=1).contiguous()  # Append(copyn_j, linear_j)
        o = torch.zeros((R, b), device=dev)
        prod[j].propagate(None, pin, o)
        out[:, off:off + b] = o
        prod_in.append(pin)
        off += b
    # ---------------- numpy restatement (forward)
    tau = T if gumbel else 1.0
    z = (a + (-np.log(-np.log(u)) if gumbel else 0.0)) / tau
    pr = np.exp(z - z.max())
    pr = np.maximum(pr / pr.sum(), 1e-20)
    m = np.array([pr[j:].sum() for j in range(8)])
    mask = np.repeat(m, BLOCKS)
    assert rel_err(out.cpu().numpy(), lin.astype(np.float64) * mask) < 1e-5
    # ---------------- backward
    d_out_d = torch.from_numpy(d_out).to(dev)
    d_p = torch.zeros((R, 8), device=dev)
    d_lin = torch.zeros((R, 240), device=dev)
    off = 0
    for j, b in enumerate(BLOCKS):
        d_pin = torch.zeros((R, 2 * b), device=dev)
        prod[j].backprop(None, prod_in[j], None, d_out_d[:, off:off + b].contiguous(), None, None, d_pin)
        d_lin[:, off:off + b] = d_pin[:, b:]
        d_mask = torch.zeros((R, 1), device=dev)
        copyn[j].backprop(None, None, None, d_pin[:, :b].contiguous(), None, None, d_mask)  # kBackpropAdds
        d_p[:, j:] += d_mask                                            # transpose of the Sum descriptor
        off += b
    d_a = torch.zeros((R, 8), device=dev)
    soft.backprop(None, a_rows, p, d_p, None, None, d_a)
    delta = alpha.copy()
    delta.scale(0.0)
    alpha.backprop(None, None, None, d_a, None, delta, None)
    # ---------------- numpy restatement (backward)
    dmask = (d_out.astype(np.float64) * lin).reshape(R, 240)
    dm = np.stack([dmask[:, sum(BLOCKS[:j]):sum(BLOCKS[:j + 1])].sum(1) for j in range(8)], axis=1)   # R x 8
    dp = np.cumsum(dm, axis=1)                                          # d p_k = sum_{j <= k} d m_j
    e = dp + scale / R / 8 * FLOPS
    da_rows = (pr * e - pr * (e @ pr)[:, None]) / tau
    d_alpha = 5 * lr * da_rows.sum(0)                                   # ConstantFunction: x5 when NG is off (simple.cc:2636)
    assert rel_err(d_lin.cpu().numpy(), d_out.astype(np.float64) * mask) < 1e-5
    got = delta.vectorize().astype(np.float64)
    assert np.abs(got - d_alpha).max() <= 2e-3 * np.abs(d_alpha).max() + 1e-9, (got, d_alpha)


@pytest.mark.parametrize("R,widths,pad", [(384, BLOCKS, 0), (1001, BLOCKS, 16), (77, [3, 5, 2], 1), (50, [240], 0), (19840, BLOCKS, 0)])
def test_shared_mask_fused_kernels(ctx, R, widths, pad):
    """tdnnf_shared_mask_fwd / _bwd (the block's descriptor sub-graph + CopyN + ElementwiseProduct in one pass) against the
    float64 restatement of the component chain above: same mask, same d_linear, same d_p."""
    import torch

    g = np.random.default_rng(R)
    nb, cols = len(widths), sum(widths)
    p = g.uniform(0.01, 1.0, (R, nb))
    p = (p / p.sum(1, keepdims=True)).astype(np.float32)
    lin = g.standard_normal((R, cols)).astype(np.float32)
    d_out = g.standard_normal((R, cols)).astype(np.float32)
    scale = 0.75
    view = lambda a: torch.from_numpy(np.pad(a, ((0, 0), (0, pad)), constant_values=5.0)).cuda()[:, : a.shape[1]]
    pd, ld, dd = view(p), view(lin), view(d_out)
    out = view(np.zeros_like(lin))
    ctx.shared_mask_fwd(pd, ld, out, widths, scale)
    m = scale * np.cumsum(p.astype(np.float64)[:, ::-1], axis=1)[:, ::-1]            # m_j = sum_{k >= j} p_k
    mask = np.repeat(m, widths, axis=1)
    assert rel_err(out.cpu().numpy(), lin * mask) < 1e-6
    d_lin = view(np.zeros_like(lin))
    d_p = view(np.full_like(p, 3.0))
    ctx.shared_mask_bwd(pd, ld, dd, d_lin, d_p, widths, scale)
    ends = np.cumsum(widths)
    dm = np.stack([(d_out.astype(np.float64) * lin)[:, e - w:e].sum(1) for e, w in zip(ends, widths)], axis=1)
    assert rel_err(d_lin.cpu().numpy(), d_out * mask) < 1e-6
    assert rel_err(d_p.cpu().numpy(), scale * np.cumsum(dm, axis=1)) < 1e-5
    d_p2 = view(np.zeros_like(p))
    ctx.shared_mask_bwd(pd, ld, dd, None, d_p2, widths, scale)                      # d_lin not wanted
    assert torch.equal(d_p2, d_p)
