"""GPU parity of the mixing-component kernels against the CPU oracle (fp32, 1e-4 fwd / 1e-3 grads;
pure copies / selections are bit-exact)."""
import numpy as np
import pytest

from tests.util import padded, rel_err, to_cuda_view

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,cols,pad", [(1, 8, 0), (20000, 8, 0), (777, 8, 3), (513, 16, 0), (300, 11, 1), (64, 240, 0)])
@pytest.mark.parametrize("gumbel", [False, True])
def test_softmax_flops_fwd_bwd(ctx, rows, cols, pad, gumbel):
    import torch

    from oracle import oracle as O

    g = np.random.default_rng(rows + cols)
    x = padded(rows, cols, pad, g, scale=2.0)
    u = g.uniform(0.01, 0.99, cols).astype(np.float32) if gumbel else None
    temp = 0.3 if gumbel else 1.0
    ref = O.softmax_flops_fwd(x, u, temp)
    xd = to_cuda_view(x)
    out = torch.zeros((rows, cols), device="cuda")
    ctx.softmax_flops_fwd(xd, out, u, 1.0 / temp)
    got = out.cpu().numpy()
    assert rel_err(got, ref) < 1e-4
    assert got.min() >= 1e-20

    if cols < 8:
        return
    od = (g.standard_normal((rows, cols)) / rows).astype(np.float32)
    scale = 0.1
    ind_ref, od_after_ref = O.softmax_flops_bwd(ref, od, scale, gumbel, temp)
    penalty = np.float32(np.float32(np.float32(scale) / rows) / cols)
    # not in place: out_deriv is mutated like the reference does (quirk Q9)
    od_d = torch.from_numpy(od).cuda()
    ind_d = torch.zeros_like(od_d)
    ctx.softmax_flops_bwd(torch.from_numpy(ref).cuda(), od_d, ind_d, float(penalty), 1.0 / temp, 1)
    assert rel_err(ind_d.cpu().numpy(), ind_ref) < 1e-3
    assert rel_err(od_d.cpu().numpy(), od_after_ref) < 1e-6
    # in place (kBackpropInPlace)
    od_d2 = torch.from_numpy(od).cuda()
    ctx.softmax_flops_bwd(torch.from_numpy(ref).cuda(), od_d2, od_d2, float(penalty), 1.0 / temp, 1)
    assert rel_err(od_d2.cpu().numpy(), ind_ref) < 1e-3


def test_softmax_flops_bwd_rejects_narrow(ctx):
    import torch

    from tdnnf_nas_b200 import capi

    t = torch.zeros((4, 4), device="cuda")
    with pytest.raises(capi.TdnnfError):
        ctx.softmax_flops_bwd(t, t.clone(), t.clone(), 0.0, 1.0, 1)


@pytest.mark.parametrize("rows,in_cols,blocks", [(1000, 1, 25), (333, 1, 40), (64, 3, 7), (0, 1, 4)])
def test_copyn(ctx, rows, in_cols, blocks):
    import torch

    from oracle import oracle as O

    g = np.random.default_rng(rows)
    x = g.standard_normal((rows, in_cols)).astype(np.float32)
    out0 = g.standard_normal((rows, in_cols * blocks)).astype(np.float32)
    ref = O.copyn_fwd(x, out0.copy(), 0.7) if rows else out0.copy()
    out = torch.from_numpy(out0.copy()).cuda()
    if rows:
        ctx.copyn_fwd(torch.from_numpy(x).cuda(), out, 0.7)
        np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-6, atol=1e-7)
        od = g.standard_normal((rows, in_cols * blocks)).astype(np.float32)
        ind0 = g.standard_normal((rows, in_cols)).astype(np.float32)
        ind_ref = O.copyn_bwd(od, ind0.copy(), 0.7)
        ind = torch.from_numpy(ind0.copy()).cuda()
        ctx.copyn_bwd(torch.from_numpy(od).cuda(), ind, 0.7)
        assert rel_err(ind.cpu().numpy(), ind_ref) < 1e-5


@pytest.mark.parametrize("u", [0.0, 0.1249999, 0.125, 0.5, 0.874, 0.99999])
def test_onehot_bit_exact(ctx, u):
    import torch

    from oracle import oracle as O

    out = torch.full((37, 8), -1.0, device="cuda")
    ctx.onehot_fwd(out, u)
    assert np.array_equal(out.cpu().numpy(), O.onehot_fwd(37, 8, u))


@pytest.mark.parametrize("rows,cols,pad", [(5000, 1536, 0), (129, 100, 4), (7, 3, 1)])
def test_bn_test_scale_offset(ctx, rows, cols, pad):
    import torch

    from oracle import oracle as O

    g = np.random.default_rng(cols)
    x = padded(rows, cols, pad, g)
    count = 1000.0
    mean = g.standard_normal(cols)
    var = g.uniform(0.1, 2.0, cols)
    scale, offset = O.bn_test_derived(mean * count, (var + mean * mean) * count, count, 1e-3, 1.0)
    ref = O.scale_offset_rows(x, scale, offset)
    xd = to_cuda_view(x)
    out = torch.zeros((rows, cols), device="cuda")
    sd, od_ = torch.from_numpy(scale).cuda(), torch.from_numpy(offset).cuda()
    ctx.scale_offset_rows(xd, out, sd, od_)
    assert rel_err(out.cpu().numpy(), ref) < 1e-6
    # Backprop: in_deriv = out_deriv .* scale, in place
    ref_b = O.scale_offset_rows(x, scale, None)
    ctx.scale_offset_rows(xd, xd, sd, None)
    assert rel_err(xd.cpu().numpy(), ref_b) < 1e-6


def test_add_row_sum_and_ewprod(ctx):
    import torch

    from oracle import oracle as O

    g = np.random.default_rng(1)
    m = g.standard_normal((4097, 53)).astype(np.float32)
    v0 = g.standard_normal(53).astype(np.float32)
    ref = O.add_row_sum(m, 0.25, v0.copy())
    v = torch.from_numpy(v0.copy()).cuda()
    ctx.add_row_sum(torch.from_numpy(m).cuda(), 0.25, v)
    assert rel_err(v.cpu().numpy(), ref) < 1e-5

    x = g.standard_normal((301, 2 * 240)).astype(np.float32)
    out = torch.zeros((301, 240), device="cuda")
    xd = torch.from_numpy(x).cuda()
    ctx.elementwise_product_fwd(xd, out)
    assert np.array_equal(out.cpu().numpy(), O.ewprod_fwd(x))
    od = g.standard_normal((301, 240)).astype(np.float32)
    ind = torch.zeros_like(xd)
    ctx.elementwise_product_bwd(xd, torch.from_numpy(od).cuda(), ind)
    assert np.array_equal(ind.cpu().numpy(), O.ewprod_bwd(x, od))
