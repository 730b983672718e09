"""GPU tests of the parameter ops and the stock TDNN-F neighbours (ReLU, bypass sum, row gather/scatter,
BatchNorm training mode) against float64 numpy restatements of the Kaldi formulas (norm.cc:392-398)."""
import numpy as np
import pytest

from tests.util import padded, rel_err, to_cuda_view

pytestmark = pytest.mark.gpu


def test_param_ops(ctx):
    import torch

    g = np.random.default_rng(0)
    a = padded(37, 53, 3, g)
    b = padded(37, 53, 5, g)
    ad, bd = to_cuda_view(a), to_cuda_view(b)
    assert ctx.mat_dot(ad, bd) == pytest.approx(float((a.astype(np.float64) * b).sum()), rel=1e-5)
    ctx.mat_axpy(0.5, ad, bd)
    np.testing.assert_allclose(bd.cpu().numpy(), b + 0.5 * a, rtol=1e-6, atol=1e-6)
    ctx.mat_scale(ad, -2.0)
    np.testing.assert_allclose(ad.cpu().numpy(), -2 * a, rtol=1e-6)
    ctx.mat_set(ad, 1.25)
    assert torch.all(ad == 1.25)


def test_relu_and_add_scaled(ctx):
    import torch

    g = np.random.default_rng(1)
    for rows, cols, pad in [(300, 1536, 0), (17, 50, 2)]:
        x = padded(rows, cols, pad, g)
        xd = to_cuda_view(x)
        y = torch.zeros((rows, cols), device="cuda")
        ctx.relu_fwd(xd, y)
        assert np.array_equal(y.cpu().numpy(), np.maximum(x, 0))
        od = g.standard_normal((rows, cols)).astype(np.float32)
        idv = torch.zeros((rows, cols), device="cuda")
        ctx.relu_bwd(y, torch.from_numpy(od).cuda(), idv)
        assert np.array_equal(idv.cpu().numpy(), od * (x > 0))
        out = torch.zeros((rows, cols), device="cuda")
        ctx.add_scaled(xd, 0.66, y, 1.0, out)
        np.testing.assert_allclose(out.cpu().numpy(), 0.66 * x + np.maximum(x, 0), rtol=1e-6, atol=1e-7)


def test_copy_rows_and_add_to_rows(ctx):
    import torch

    g = np.random.default_rng(2)
    src = g.standard_normal((40, 24)).astype(np.float32)
    m = g.permutation(40)[:30].astype(np.int32)
    m[3] = -1
    dst = torch.full((30, 24), 9.0, device="cuda")
    ctx.copy_rows(torch.from_numpy(src).cuda(), dst, torch.from_numpy(m).cuda())
    ref = np.where(m[:, None] >= 0, src[np.maximum(m, 0)], 0)
    assert np.array_equal(dst.cpu().numpy(), ref)
    acc0 = g.standard_normal((40, 24)).astype(np.float32)
    acc = torch.from_numpy(acc0.copy()).cuda()
    upd = g.standard_normal((30, 24)).astype(np.float32)
    ctx.add_to_rows(0.5, torch.from_numpy(upd).cuda(), acc, torch.from_numpy(m).cuda())
    ref2 = acc0.copy()
    for r, t in enumerate(m):
        if t >= 0:
            ref2[t] += 0.5 * upd[r]
    np.testing.assert_allclose(acc.cpu().numpy(), ref2, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("rows,cols", [(2000, 1536), (333, 70)])
def test_batchnorm_train_mode(ctx, rows, cols):
    import torch

    g = np.random.default_rng(3)
    x = (g.standard_normal((rows, cols)) * 2 + 0.5).astype(np.float32)
    eps, rms = 1e-3, 1.0
    x64 = x.astype(np.float64)
    mean = x64.mean(0)
    var = np.maximum((x64 * x64).mean(0) - mean * mean, 0)
    scale = rms * (var + eps) ** -0.5
    z = (x64 - mean) * scale
    xd = torch.from_numpy(x).cuda()
    zd = torch.zeros_like(xd)
    memo = torch.zeros(5 * cols, device="cuda")
    ctx.batchnorm_train_fwd(xd, zd, memo, eps, rms)
    assert rel_err(zd.cpu().numpy(), z) < 1e-4
    mh = memo.cpu().numpy()
    assert rel_err(mh[:cols], mean) < 1e-4 and rel_err(mh[cols:2 * cols], (x64 * x64).mean(0)) < 1e-4  # row 1: uncentred
    # backward (norm.cc:392-398): x' = scale*(z' - mean(z')) + z * var_deriv_mod
    zp = g.standard_normal((rows, cols))
    vdm = -1.0 / (rms * rms) * (zp * z).mean(0) * scale
    xp = scale * (zp - zp.mean(0)) + z * vdm
    idv = torch.zeros_like(xd)
    ctx.batchnorm_train_bwd(zd, torch.from_numpy(zp.astype(np.float32)).cuda(), idv, memo, rms)
    assert rel_err(idv.cpu().numpy(), xp) < 1e-3
    # finite-difference check of the backward formula against the forward definition (float64)
    d = g.standard_normal((rows, cols)) * 1e-4

    def fwd(v):
        m = v.mean(0)
        s = rms * (np.maximum((v * v).mean(0) - m * m, 0) + eps) ** -0.5
        return (v - m) * s

    fd = ((fwd(x64 + d) - fwd(x64 - d)) * zp).sum() / 2
    assert fd == pytest.approx(float((xp * d).sum()), rel=5e-3)


def test_fused_tail_equals_separate_components(ctx):
    """relu -> BatchNormTest scale/offset -> bypass sum in one pass == the three ops run separately."""
    import torch

    g = np.random.default_rng(4)
    rows, cols = 700, 1536
    x = torch.from_numpy(g.standard_normal((rows, cols)).astype(np.float32)).cuda()
    prev = torch.from_numpy(g.standard_normal((rows, cols)).astype(np.float32)).cuda()
    scale = torch.from_numpy(g.uniform(0.5, 2.0, cols).astype(np.float32)).cuda()
    offset = torch.from_numpy(g.standard_normal(cols).astype(np.float32)).cuda()
    relu, bn, ref = torch.zeros_like(x), torch.zeros_like(x), torch.zeros_like(x)
    ctx.relu_fwd(x, relu)
    ctx.scale_offset_rows(relu, bn, scale, offset)
    ctx.add_scaled(prev, 0.66, bn, 1.0, ref)
    out = torch.zeros_like(x)
    ctx.relu_scale_offset_bypass_fwd(x, scale, offset, prev, 0.66, out)
    assert rel_err(out.cpu().numpy(), ref.cpu().numpy()) < 1e-6
    d_out = torch.from_numpy(g.standard_normal((rows, cols)).astype(np.float32)).cuda()
    d_bn, d_relu = torch.zeros_like(x), torch.zeros_like(x)
    ctx.scale_offset_rows(d_out, d_bn, scale, None)
    ctx.relu_bwd(relu, d_bn, d_relu)
    d_x, d_prev = torch.zeros_like(x), torch.full_like(x, 5.0)
    ctx.relu_scale_offset_bypass_bwd(d_out, x, scale, 0.66, d_x, d_prev)
    assert torch.equal(d_x, d_relu)
    np.testing.assert_allclose(d_prev.cpu().numpy(), 0.66 * d_out.cpu().numpy(), rtol=1e-6)
