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


def test_multi_buffer_parameter_step(ctx):
    """tdnnf_multi_sumsq / tdnnf_multi_axpy_zero (the two launches of UpdateNnetWithMaxChange, nnet-utils.cc:2085-2175)
    against torch, with strided views, grouped buffers and a buffer larger than one block's share."""
    import ctypes as C

    import torch

    from tdnnf_nas_b200 import capi

    lib = capi.load()
    g = torch.Generator(device="cuda").manual_seed(0)
    shapes = [(160, 10752, 10752), (1, 167, 167), (1536, 1120, 1124), (7, 5, 8), (300, 256, 256)]
    groups = [0, 0, 1, 2, 2]
    deltas = [torch.randn(r, ld, device="cuda", generator=g)[:, :c] for r, c, ld in shapes]
    models = [torch.randn(r, ld, device="cuda", generator=g)[:, :c] for r, c, ld in shapes]
    n = len(shapes)
    P, I, F = C.c_void_p * n, C.c_int32 * n, C.c_float * n
    dptr, mptr = P(*[d.data_ptr() for d in deltas]), P(*[m.data_ptr() for m in models])
    rows, cols = I(*[s[0] for s in shapes]), I(*[s[1] for s in shapes])
    ld = I(*[s[2] for s in shapes])
    out = torch.zeros(3, dtype=torch.float64, device="cuda")
    assert lib.tdnnf_multi_sumsq(ctx.h, n, dptr, rows, cols, ld, I(*groups), C.c_void_p(out.data_ptr())) == 0
    ref = [0.0, 0.0, 0.0]
    for d, gi in zip(deltas, groups):
        ref[gi] += float((d.double() ** 2).sum())
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-10)
    factors = [0.5, 0.5, -2.0, 1.0, 0.25]
    expect = [m + f * d for m, d, f in zip(models, deltas, factors)]
    pads = [m.clone() for m in models]
    assert lib.tdnnf_multi_axpy_zero(ctx.h, n, mptr, ld, dptr, ld, rows, cols, F(*factors)) == 0
    torch.cuda.synchronize()
    for m, e, d in zip(models, expect, deltas):
        assert torch.allclose(m, e, rtol=1e-6, atol=1e-6) and torch.all(d == 0)
    with pytest.raises(AssertionError):
        assert lib.tdnnf_multi_sumsq(ctx.h, 200, dptr, rows, cols, ld, None, C.c_void_p(out.data_ptr())) == 0  # > TDNNF_MULTI_MAX


@pytest.mark.parametrize("N,r,n_views", [(5000, 80, 1), (3000, 20, 7), (130, 33, 3),
                                        (40000, 80, 1),   # clusters, two passes over a slab that needs > 48 KB of shared memory
                                        (40001, 128, 2),  # the widest tile
                                        (20000, 48, 2)])
def test_ng_gram_scale_and_w_update(ctx, N, r, n_views):
    """tdnnf_ng_gram_scale (L = H^T H, tr(X X^T) from per-row sums of squares, the scale) and tdnnf_ng_w_update
    (W_next = A J + AC W) against float64 numpy."""
    import ctypes as C

    import torch

    from tdnnf_nas_b200 import capi

    lib = capi.load()
    g = np.random.default_rng(N + r)
    H = g.standard_normal((N, r)).astype(np.float32)
    WWt = (np.eye(r) * 0.5 + 0.01 * g.standard_normal((r, r))).astype(np.float32)
    WWt = ((WWt + WWt.T) / 2).astype(np.float32)
    in_rows = N + 8 * (n_views - 1)
    rowsq = g.uniform(0.5, 2.0, in_rows).astype(np.float32)
    offs = [8 * i for i in range(n_views)]
    weff = g.uniform(0.2, 1.0, n_views).astype(np.float32)
    Hd, Wd, rq, wf = (torch.from_numpy(a).cuda() for a in (H, WWt, rowsq, weff))
    L = torch.full((r, r + 3), 7.0, device="cuda")[:, :r]
    out3 = torch.zeros(3, device="cuda")
    sumsq = torch.zeros(16, dtype=torch.float64, device="cuda")
    I = C.c_int32 * n_views
    rc = lib.tdnnf_ng_gram_scale(ctx.h, C.c_void_p(Hd.data_ptr()), N, r, r, C.c_void_p(L.data_ptr()), r + 3,
                                 C.c_void_p(Wd.data_ptr()), r, C.c_void_p(rq.data_ptr()), C.c_void_p(sumsq.data_ptr()), in_rows,
                                 n_views, I(*offs), 1, C.c_void_p(wf.data_ptr()), C.c_float(float(N)), C.c_void_p(out3.data_ptr()))
    assert rc == 0, lib.tdnnf_last_error()
    H64 = H.astype(np.float64)
    L_ref = H64.T @ H64
    assert rel_err(L.cpu().numpy(), L_ref) < 1e-5
    # a second call finds the scratch (accumulation buffers, counters) re-armed by the first
    L2 = torch.zeros((r, r), device="cuda")
    rc = lib.tdnnf_ng_gram_scale(ctx.h, C.c_void_p(Hd.data_ptr()), N, r, r, C.c_void_p(L2.data_ptr()), r,
                                 C.c_void_p(Wd.data_ptr()), r, C.c_void_p(rq.data_ptr()), C.c_void_p(sumsq.data_ptr()), in_rows,
                                 n_views, I(*offs), 1, C.c_void_p(wf.data_ptr()), C.c_float(float(N)), C.c_void_p(out3.data_ptr()))
    assert rc == 0, lib.tdnnf_last_error()
    assert rel_err(L2.cpu().numpy(), L_ref) < 1e-5
    tr_xx = float(N) + sum(float(weff[i]) ** 2 * float(rowsq[offs[i]: offs[i] + N].astype(np.float64).sum()) for i in range(n_views))
    tr_hat = tr_xx - 2 * np.trace(L_ref) + float((L_ref * WWt.astype(np.float64)).sum())
    o = out3.cpu().numpy()
    assert o[0] == pytest.approx(tr_xx, rel=1e-5) and o[1] == pytest.approx(tr_hat, rel=1e-3)
    assert o[2] == pytest.approx(np.sqrt(tr_xx / tr_hat) if tr_hat > 0 else 1.0, rel=1e-3)
    # W update
    D = 517
    A, AC = g.standard_normal((r, r)).astype(np.float32), g.standard_normal((r, r)).astype(np.float32)
    J, W = g.standard_normal((r, D)).astype(np.float32), g.standard_normal((r, D)).astype(np.float32)
    Ad, ACd, Jd, Wd2 = (torch.from_numpy(a).cuda() for a in (A, AC, J, W))
    Wn = torch.zeros((r, D), device="cuda")
    rc = lib.tdnnf_ng_w_update(ctx.h, C.c_void_p(Ad.data_ptr()), r, C.c_void_p(ACd.data_ptr()), r, C.c_void_p(Jd.data_ptr()), D,
                               C.c_void_p(Wd2.data_ptr()), D, r, D, C.c_void_p(Wn.data_ptr()), D)
    assert rc == 0, lib.tdnnf_last_error()
    ref = A.astype(np.float64) @ J.astype(np.float64) + AC.astype(np.float64) @ W.astype(np.float64)
    assert rel_err(Wn.cpu().numpy(), ref) < 1e-5


@pytest.mark.parametrize("rows,cols,ri,ro", [(160, 10753, 20, 80), (1536, 1121, 20, 80), (37, 130, 5, 0), (70, 65, 0, 33),
                                             (129, 257, 128, 128), (8, 40, 1, 3)])
def test_ng_project_gradient(ctx, rows, cols, ri, ro):
    """tdnnf_ng_project_gradient: G <- (I - Wo^T Wo) G (I - Wi^T Wi) (the rank-r projections of
    OnlineNaturalGradient::PreconditionDirections applied to the gradient, tdnn.cc:598-624) against float64 numpy;
    rank 0 = that side is the identity (NULL)."""
    import ctypes as C

    import torch

    from tdnnf_nas_b200 import capi

    lib = capi.load()
    g = np.random.default_rng(rows * 7 + cols)
    G = g.standard_normal((rows, cols)).astype(np.float32)
    Wi = (g.standard_normal((max(ri, 1), cols)) / np.sqrt(cols)).astype(np.float32)
    Wo = (g.standard_normal((max(ro, 1), rows)) / np.sqrt(rows)).astype(np.float32)
    Gd = torch.full((rows, cols + 5), 3.0, device="cuda")
    Gd[:, :cols] = torch.from_numpy(G).cuda()
    Wid, Wod = torch.from_numpy(Wi).cuda(), torch.from_numpy(Wo).cuda()
    rc = lib.tdnnf_ng_project_gradient(ctx.h, C.c_void_p(Gd.data_ptr()), rows, cols, cols + 5,
                                       C.c_void_p(Wid.data_ptr()) if ri else None, ri, cols if ri else 0,
                                       C.c_void_p(Wod.data_ptr()) if ro else None, ro, rows if ro else 0)
    assert rc == 0, lib.tdnnf_last_error()
    ref = G.astype(np.float64)
    if ri:
        W = Wi.astype(np.float64)
        ref = ref - (ref @ W.T) @ W
    if ro:
        W = Wo.astype(np.float64)
        ref = ref - W.T @ (W @ ref)
    out = Gd.cpu().numpy()
    assert rel_err(out[:, :cols], ref) < 2e-6
    assert (out[:, cols:] == 3.0).all()  # the padding of the stride is not touched


@pytest.mark.parametrize("rows,cols,pad", [(320, 6008, 0), (17, 50, 3), (5, 1, 0), (3, 300, 1)])
def test_log_softmax_component_vs_numpy(ctx, rows, cols, pad):
    """LogSoftmaxComponent (nnet-simple-component.cc:3607-3632): ApplyLogSoftMaxPerRow / DiffLogSoftmaxPerRow in float64."""
    import torch

    from tdnnf_nas_b200 import nnet3

    nnet3.set_context(ctx)
    comp = nnet3.Component.new("LogSoftmaxComponent", f"dim={cols}")
    assert comp.type() == "LogSoftmaxComponent"
    assert comp.properties() == (nnet3.kSimpleComponent | nnet3.kBackpropNeedsOutput | nnet3.kStoresStats)
    g = np.random.default_rng(rows + cols)
    x = (g.standard_normal((rows, cols)) * 4.0).astype(np.float32)
    x[0, :] += 60.0  # large offsets must not overflow (max subtraction)
    od = g.standard_normal((rows, cols)).astype(np.float32)
    xb = torch.full((rows, cols + pad), 9.0, device="cuda")
    xb[:, :cols] = torch.from_numpy(x).cuda()
    xin = xb[:, :cols]
    out = torch.empty((rows, cols), device="cuda")
    comp.propagate(None, xin, out)
    x64 = x.astype(np.float64)
    m = x64.max(1, keepdims=True)
    ref = x64 - m - np.log(np.exp(x64 - m).sum(1, keepdims=True))
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=0, atol=2e-5)
    np.testing.assert_allclose(np.exp(out.cpu().numpy().astype(np.float64)).sum(1), 1.0, rtol=1e-5)
    odd = torch.from_numpy(od).cuda()
    ind = torch.empty_like(odd)
    comp.backprop(None, None, out, odd, None, None, ind)
    ref_d = od.astype(np.float64) - np.exp(ref) * od.astype(np.float64).sum(1, keepdims=True)
    assert rel_err(ind.cpu().numpy(), ref_d) < 1e-5
    # in place (the training step feeds the derivative back through the same buffer)
    comp.backprop(None, None, out, odd, None, None, odd)
    assert torch.equal(odd, ind)
    back = nnet3.Component.read(comp.write(True), True)
    assert back.type() == "LogSoftmaxComponent" and back.input_dim() == cols


@pytest.mark.parametrize("rows,cols,pad", [(37, 1536, 0), (300, 160, 0), (5, 64, 4), (129, 200, 8)])
def test_fused_tail_producers_write_the_operand_planes(ctx, rows, cols, pad):
    """tdnnf_relu_scale_offset_bypass_{fwd,bwd}_planes: same fp32 results as the plain fused tails, and the planes they
    hand over give the same GEMM result as splitting the matrix (tdnnf_darts_propagate with the handle attached counts a
    cache hit and launches no split of the activation), the same bias gradient and the same row sums of squares."""
    import ctypes as C

    import torch

    from tdnnf_nas_b200 import capi

    lib = capi.load()
    g = np.random.default_rng(rows + cols)
    mk = lambda r, c: torch.from_numpy(np.pad(g.standard_normal((r, c)).astype(np.float32), ((0, 0), (0, pad)))).cuda()[:, :c]
    x, prev, d_out = mk(rows, cols), mk(rows, cols), mk(rows, cols)
    scale = torch.from_numpy((g.random(cols) + 0.5).astype(np.float32)).cuda()
    offset = torch.from_numpy(g.standard_normal(cols).astype(np.float32)).cuda()
    ref_out, out = mk(rows, cols), mk(rows, cols)
    ctx.relu_scale_offset_bypass_fwd(x, scale, offset, prev, 0.66, ref_out)
    pl = C.c_void_p()
    m = capi._mat
    capi.check(lib.tdnnf_relu_scale_offset_bypass_fwd_planes(ctx.h, m(x)[0], rows, cols, m(x)[3], scale.data_ptr(), offset.data_ptr(),
                                                             m(prev)[0], m(prev)[3], 0.66, m(out)[0], m(out)[3], C.byref(pl)))
    assert rel_err(out.cpu().numpy(), ref_out.cpu().numpy()) < 1e-6
    # a Propagate-shaped GEMM on `out`: with the producer's planes attached vs. splitting the matrix
    W = torch.from_numpy((g.standard_normal((48, cols)) / np.sqrt(cols)).astype(np.float32)).cuda()
    one = torch.ones(1, device="cuda")
    y_split, y_planes = torch.zeros((rows, 48), device="cuda"), torch.zeros((rows, 48), device="cuda")
    ctx.darts_propagate(out, y_split, W, None, 1, one, [0], 1)
    h0, _ = ctx.operand_cache_stats()
    capi.check(lib.tdnnf_ctx_planes_attach(ctx.h, pl))
    ctx.darts_propagate(out, y_planes, W, None, 1, one, [0], 1)
    capi.check(lib.tdnnf_ctx_planes_detach(ctx.h, pl))
    assert ctx.operand_cache_stats()[0] == h0 + 1
    assert rel_err(y_planes.cpu().numpy(), y_split.cpu().numpy()) < 1e-6  # split-K red.adds: the summation order differs run to run
    capi.check(lib.tdnnf_planes_release(pl))
    # backward producer: d_x, d_prev as the plain kernel; planes + column sums feed the parameter gradient
    ref_dx, ref_dp, dx, dp = mk(rows, cols), mk(rows, cols), mk(rows, cols), mk(rows, cols)
    ctx.relu_scale_offset_bypass_bwd(d_out, x, scale, 0.66, ref_dx, ref_dp)
    pl = C.c_void_p()
    capi.check(lib.tdnnf_relu_scale_offset_bypass_bwd_planes(ctx.h, m(d_out)[0], m(d_out)[3], m(x)[0], m(x)[3], scale.data_ptr(), 0.66,
                                                             m(dx)[0], m(dx)[3], m(dp)[0], m(dp)[3], rows, cols, C.byref(pl)))
    assert rel_err(dx.cpu().numpy(), ref_dx.cpu().numpy()) < 1e-6 and rel_err(dp.cpu().numpy(), ref_dp.cpu().numpy()) < 1e-6
    if rows >= 512 or True:
        ctx.set_wgrad_mn_min_rows(1)  # the MN-major parameter gradient is the consumer of planes + column sums
        try:
            inp = mk(rows, 40)
            dW1, db1 = torch.zeros((cols, 40), device="cuda"), torch.zeros(cols, device="cuda")
            dW2, db2 = torch.zeros((cols, 40), device="cuda"), torch.zeros(cols, device="cuda")
            ctx.darts_backprop_params(inp, dx, None, dW1, db1, one, [0], 1, 0.5)
            capi.check(lib.tdnnf_ctx_planes_attach(ctx.h, pl))
            ctx.darts_backprop_params(inp, dx, None, dW2, db2, one, [0], 1, 0.5)
            capi.check(lib.tdnnf_ctx_planes_detach(ctx.h, pl))
            assert rel_err(dW2.cpu().numpy(), dW1.cpu().numpy()) < 1e-6
            assert rel_err(db2.cpu().numpy(), db1.cpu().numpy()) < 1e-5
        finally:
            ctx.set_wgrad_mn_min_rows(512)
    capi.check(lib.tdnnf_planes_release(pl))
