"""GPU parity of the TdnnDARTSV3 kernels (through the C ABI) against the CPU oracle.

Tolerances are BASELINE.json's: forward activations 1e-4 relative, gradients 1e-3 relative.
"""
import numpy as np
import pytest

from tests.util import max_rel_to_scale, padded, rel_err, to_cuda_view

pytestmark = pytest.mark.gpu

FWD_TOL = 1e-4
GRAD_TOL = 1e-3

MODES = {
    "softmax": 0,
    "gumbel": 1,
    "free": 2,
    "uniform": 4,
    "gumbel_uniform": 5,
}


def _setup(n, in_dim, out_dim, S, t_out, offsets, row_stride=1, seed=0, pad=0, alpha_random=True):
    """Regular grid: input frames t = min(offsets) .. (t_out-1)*row_stride + max(offsets), S sequences."""
    from tdnnf_nas_b200 import synth

    g = np.random.default_rng(seed)
    t_in0 = min(offsets)
    n_t_in = (t_out - 1) * row_stride + max(offsets) - t_in0 + 1
    n_t_in = row_stride * ((n_t_in + row_stride - 1) // row_stride)
    rs, row_offsets = synth.regular_row_offsets(offsets, t_in0, 0, S, 1, row_stride)
    assert rs == row_stride
    in_rows, out_rows = n_t_in * S, t_out * S
    x = padded(in_rows, in_dim, pad, g)
    W = padded(out_dim, n * in_dim, pad, g, scale=1.0 / np.sqrt(in_dim * n))
    bias_params = np.concatenate([g.standard_normal(n) if alpha_random else np.zeros(n),
                                  g.standard_normal(out_dim)]).astype(np.float32)
    od = padded(out_rows, out_dim, pad, g, scale=1.0 / out_rows)
    return dict(x=x, W=W, bias_params=bias_params, od=od, row_offsets=row_offsets, row_stride=row_stride,
                in_rows=in_rows, out_rows=out_rows, rng=g)


CASES = [
    # name, n, in_dim, out_dim, S, t_out, offsets, row_stride, pad
    ("cfg1_fwd_offsets", 7, 40, 160, 64, 150, list(range(0, 7)), 1, 0),
    ("cfg1_mirrored", 7, 40, 160, 16, 40, list(range(-6, 1)), 1, 0),
    ("cfg1_wide_out", 7, 40, 1536, 8, 30, list(range(0, 7)), 1, 8),
    ("linear_1536_160", 7, 1536, 160, 8, 24, list(range(-6, 1)), 1, 0),
    ("affine_160_1536", 7, 160, 1536, 8, 24, list(range(0, 7)), 1, 0),
    ("subsample3", 7, 160, 256, 8, 17, list(range(0, 7)), 3, 4),
    ("two_offsets", 2, 96, 72, 5, 33, [0, 3], 1, 4),
    ("ragged_dims", 3, 52, 44, 3, 19, [0, 1, 4], 1, 4),
]


@pytest.mark.parametrize("grad_fast", [False, True], ids=["grad3x", "gradfast"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("mode", list(MODES))
def test_propagate_backprop_parity(ctx, case, mode, grad_fast):
    """grad_fast: the parameter gradient with ONE tensor-core product (fp16 x fp16, both operands scaled by an exact
    power of two) instead of three bf16 ones -- same 1e-3 tolerance (tdnnf_ctx_set_gradient_mode)."""
    import torch

    from oracle import oracle as O
    from tdnnf_nas_b200 import capi

    name, n, in_dim, out_dim, S, t_out, offsets, row_stride, pad = case
    flags = MODES[mode]
    d = _setup(n, in_dim, out_dim, S, t_out, offsets, row_stride, seed=sum(map(ord, name)), pad=pad)
    g = d["rng"]
    temp = 0.5
    u_g = g.uniform(0.05, 0.95, n).astype(np.float32)
    u_u = float(g.uniform(0, 1))
    share = O.share_index(offsets)
    has_bias_fwd = offsets[1] > 0

    # ---------------- oracle
    out_ref, coef_ref = O.tdnn_propagate(offsets, flags, temp, d["W"], d["bias_params"], d["x"], d["out_rows"],
                                         d["row_offsets"], row_stride, u_g, u_u)
    lr = 0.01
    in_deriv0 = (g.standard_normal((d["in_rows"], in_dim)) * 0.01).astype(np.float32)
    in_deriv_ref = in_deriv0.copy()
    dW_ref = np.zeros((out_dim, n * in_dim), np.float32)
    dbias_ref = np.zeros(n + out_dim, np.float32)
    fl_upd = flags | capi.UPDATE_ALPHA
    s_ref = O.tdnn_backprop(offsets, fl_upd, temp, d["W"], d["x"], d["od"], coef_ref, d["row_offsets"], row_stride, lr,
                            in_deriv=in_deriv_ref, dW=dW_ref, dbias=dbias_ref)

    # ---------------- GPU through the C ABI
    x = to_cuda_view(d["x"])
    W = to_cuda_view(d["W"])
    od = to_cuda_view(d["od"])
    bp = torch.from_numpy(d["bias_params"]).cuda()
    out_buf = torch.full((d["out_rows"], out_dim + pad), 7.0, device="cuda")
    out = out_buf[:, :out_dim]
    coef = torch.zeros(n, device="cuda")
    weff = torch.zeros(n, device="cuda")
    ctx.darts_coef(bp[:n], flags, temp, u_g, u_u, share, coef, weff)
    ctx.darts_propagate(x, out, W, bp[n:] if has_bias_fwd else None, 2 if has_bias_fwd else 1, weff, d["row_offsets"],
                        row_stride)
    torch.cuda.synchronize()
    np.testing.assert_allclose(coef.cpu().numpy(), coef_ref, rtol=2e-6, atol=1e-30)
    got = out.cpu().numpy()
    assert rel_err(got, out_ref) < FWD_TOL, f"fwd rel err {rel_err(got, out_ref)}"
    assert max_rel_to_scale(got, out_ref) < 2 * FWD_TOL
    if pad:
        assert torch.all(out_buf[:, out_dim:] == 7.0), "kernel wrote outside the view"

    in_deriv_buf = torch.zeros((d["in_rows"], in_dim + pad), device="cuda")
    in_deriv = in_deriv_buf[:, :in_dim]
    in_deriv.copy_(torch.from_numpy(in_deriv0))
    weff2 = torch.zeros(n, device="cuda")
    ctx.darts_weff_from_coef(coef, flags, share, weff2)
    ctx.set_gradient_mode(grad_fast)
    try:
        ctx.darts_backprop_data(od, in_deriv, W, weff2, d["row_offsets"], row_stride)
        dW = torch.zeros((out_dim, n * in_dim), device="cuda")
        dbias = torch.zeros(n + out_dim, device="cuda")
        s = torch.zeros(n, device="cuda")
        ctx.darts_backprop_params(x, od, W, dW, dbias[n:], weff2, d["row_offsets"], row_stride, lr, s)
    finally:
        ctx.set_gradient_mode(False)
    ctx.darts_alpha_update(s, coef, fl_upd, temp, share, lr, dbias[:n])
    torch.cuda.synchronize()
    assert torch.equal(weff, weff2)
    e = rel_err(in_deriv.cpu().numpy() - in_deriv0, in_deriv_ref - in_deriv0)
    assert e < GRAD_TOL, f"in_deriv rel err {e}"
    e = rel_err(dW.cpu().numpy(), dW_ref)
    assert e < GRAD_TOL, f"dW rel err {e}"
    e = rel_err(dbias[n:].cpu().numpy(), dbias_ref[n:])
    assert e < GRAD_TOL, f"dbias rel err {e}"
    if not (flags & capi.UNIFORM_SAMPLE):
        e = rel_err(s.cpu().numpy(), s_ref)
        assert e < GRAD_TOL, f"s (alpha inner products) rel err {e}"
        # the alpha delta is a difference of O(s) terms: compare against the scale of the terms
        scale = np.abs(dbias_ref[:n]).max() + 1e-30
        assert np.abs(dbias[:n].cpu().numpy() - dbias_ref[:n]).max() / scale < 5 * GRAD_TOL
    else:
        assert np.all(dbias[:n].cpu().numpy() == 0)


@pytest.mark.parametrize("x_scale,od_scale", [(1.0, 1.0), (3.0e4, 1.0e-12), (1.0e-9, 1.0e7)])
def test_fast_gradients_dynamic_range(ctx, x_scale, od_scale):
    """The one-product parameter gradient stores both operands as fp16 after an exact power-of-two scaling taken
    from their max magnitude: activations of 3e4 (near the fp16 maximum unscaled) and derivatives of 1e-12 (far
    below the smallest fp16 subnormal unscaled) must give the same relative accuracy as O(1) data."""
    import torch

    from oracle import oracle as O

    n, in_dim, out_dim, S, t_out, offsets = 3, 96, 80, 8, 40, [0, 2, 5]
    d = _setup(n, in_dim, out_dim, S, t_out, offsets, 1, seed=11)
    x = (d["x"] * x_scale).astype(np.float32)
    od = (d["od"] * od_scale).astype(np.float32)
    coef = np.array([1.0, 0.3, 0.7], np.float32)
    dW_ref = np.zeros((out_dim, n * in_dim), np.float64)
    for i in range(n):
        xi = x[d["row_offsets"][i]: d["row_offsets"][i] + d["out_rows"]].astype(np.float64)
        dW_ref[:, i * in_dim:(i + 1) * in_dim] = coef[i] * (od.astype(np.float64).T @ xi)
    in_ref = np.zeros((d["in_rows"], in_dim), np.float64)
    for i in range(n):
        Wi = d["W"][:, i * in_dim:(i + 1) * in_dim].astype(np.float64)
        in_ref[d["row_offsets"][i]: d["row_offsets"][i] + d["out_rows"]] += coef[i] * (od.astype(np.float64) @ Wi)
    weff = torch.from_numpy(coef).cuda()
    dW = torch.zeros((out_dim, n * in_dim), device="cuda")
    dbias = torch.zeros(out_dim, device="cuda")
    in_deriv = torch.zeros((d["in_rows"], in_dim), device="cuda")
    ctx.set_gradient_mode(True)
    try:
        ctx.darts_backprop_data(torch.from_numpy(od).cuda(), in_deriv, to_cuda_view(d["W"]), weff, d["row_offsets"], 1)
        ctx.darts_backprop_params(torch.from_numpy(x).cuda(), torch.from_numpy(od).cuda(), None, dW, dbias, weff,
                                  d["row_offsets"], 1, 1.0, None)
    finally:
        ctx.set_gradient_mode(False)
    torch.cuda.synchronize()
    assert rel_err(dW.cpu().numpy(), dW_ref) < GRAD_TOL
    assert rel_err(in_deriv.cpu().numpy(), in_ref) < GRAD_TOL
    assert rel_err(dbias.cpu().numpy(), od.astype(np.float64).sum(0)) < 1e-5


@pytest.mark.parametrize("n,in_dim,rank,S,t_out,offsets,row_stride,planes", [
    (7, 1536, 20, 8, 24, list(range(-6, 1)), 1, 2), (7, 160, 20, 8, 24, list(range(0, 7)), 1, 3),
    (7, 96, 12, 5, 17, list(range(0, 7)), 3, 2), (3, 52, 80, 3, 19, [0, 1, 4], 1, 2), (3, 52, 80, 3, 19, [0, 1, 4], 1, 3),
    (7, 64, 40, 4, 20, list(range(0, 7)), 1, 2)])
def test_project_matches_spliced_product(ctx, n, in_dim, rank, S, t_out, offsets, row_stride, planes):
    """tdnnf_darts_project (one un-spliced GEMM + gather over the offsets; the natural-gradient H = X W^T) against
    float64 numpy on the materialised spliced input [w_1 X_1 | ... | w_n X_n | 1]; the last case (n*rank = 280 > 256)
    takes the fall-back through tdnnf_darts_propagate."""
    import torch

    d = _setup(n, in_dim, rank, S, t_out, offsets, row_stride, seed=n * in_dim + rank)
    g = d["rng"]
    W = (g.standard_normal((rank, n * in_dim + 1)) * 0.1).astype(np.float32)
    weff = g.uniform(0.1, 1.0, n).astype(np.float32)
    x64 = d["x"].astype(np.float64)
    ref = np.tile(W[:, -1].astype(np.float64), (d["out_rows"], 1))
    for i in range(n):
        rows = d["row_offsets"][i] + row_stride * np.arange(d["out_rows"])
        ref += weff[i] * (x64[rows] @ W[:, i * in_dim:(i + 1) * in_dim].astype(np.float64).T)
    out = torch.full((d["out_rows"], rank), 3.0, device="cuda")
    ctx.reserve(64 << 20)
    junk = torch.full((16 << 20,), float("nan"), device="cuda")  # whatever the scratch arena hands out must not leak into Y
    del junk
    ctx.set_gemm_planes(planes)
    try:
        ctx.darts_project(to_cuda_view(d["x"]), out, torch.from_numpy(W).cuda(), torch.from_numpy(W[:, -1].copy()).cuda(),
                          torch.from_numpy(weff).cuda(), d["row_offsets"], row_stride)
    finally:
        ctx.set_gemm_planes(2)
    torch.cuda.synchronize()
    e = rel_err(out.cpu().numpy(), ref)
    assert e < (6e-6 if planes == 3 else 2e-5), e  # 3 planes: fp32 accumulation over K is what is left


def test_propagate_adds_mode(ctx):
    """bias_mode 0 (kPropagateAdds, conv.h:130-134): out is accumulated into."""
    import torch

    from oracle import oracle as O

    d = _setup(3, 64, 48, 4, 21, [0, 2, 5], 1, seed=3)
    g = d["rng"]
    out0 = g.standard_normal((d["out_rows"], 48)).astype(np.float32)
    ref, coef = O.tdnn_propagate([0, 2, 5], 0, 1.0, d["W"], d["bias_params"], d["x"], d["out_rows"], d["row_offsets"],
                                 1)
    ref = ref - d["bias_params"][3:][None, :] + out0
    out = torch.from_numpy(out0).cuda()
    weff = torch.tensor([1.0, coef[1], coef[2]], device="cuda")
    ctx.darts_propagate(to_cuda_view(d["x"]), out, to_cuda_view(d["W"]), None, 0, weff, d["row_offsets"], 1)
    assert rel_err(out.cpu().numpy(), ref) < FWD_TOL


def test_bad_arguments_fail_loudly(ctx):
    import torch

    from tdnnf_nas_b200 import capi

    x = torch.zeros((10, 8), device="cuda")
    out = torch.zeros((8, 8), device="cuda")
    W = torch.zeros((8, 16), device="cuda")
    weff = torch.ones(2, device="cuda")
    with pytest.raises(capi.TdnnfError):  # view does not fit (ref: KALDI_ASSERT in GetInputPart, tdnn.cc:811-813)
        ctx.darts_propagate(x, out, W, None, 1, weff, [0, 5], 1)


@pytest.mark.parametrize("case", [c for c in CASES if c[0] in ("cfg1_mirrored", "linear_1536_160", "affine_160_1536", "subsample3",
                                                                "two_offsets", "ragged_dims")], ids=lambda c: c[0])
def test_param_gradient_mn_major_equals_transposed_form(ctx, case):
    """The parameter gradient contracted MN-major over the row planes (default for >= 512 rows) against the
    transposed-plane form, on the same inputs: same products, different operand layout and K blocking."""
    import torch

    name, n, in_dim, out_dim, S, t_out, offsets, row_stride, pad = case
    d = _setup(n, in_dim, out_dim, S, t_out, offsets, row_stride, seed=7, pad=pad)
    x, W, od = to_cuda_view(d["x"]), to_cuda_view(d["W"]), to_cuda_view(d["od"])
    weff = torch.from_numpy(d["rng"].uniform(0.1, 1.0, n).astype(np.float32)).cuda()
    res = []
    try:
        for min_rows in (-1, 1):
            ctx.set_wgrad_mn_min_rows(min_rows)
            dW = torch.zeros((out_dim, n * in_dim), device="cuda")
            db = torch.zeros(out_dim, device="cuda")
            s = torch.zeros(n, device="cuda")
            ctx.darts_backprop_params(x, od, W, dW, db, weff, d["row_offsets"], row_stride, 0.25, s)
            res.append((dW.cpu().numpy(), db.cpu().numpy(), s.cpu().numpy()))
    finally:
        ctx.set_wgrad_mn_min_rows(512)
    (dW0, db0, s0), (dW1, db1, s1) = res
    assert rel_err(dW1, dW0) < 2e-5 and max_rel_to_scale(dW1, dW0) < 5e-5
    assert rel_err(db1, db0) < 2e-5
    assert np.abs(s1 - s0).max() <= 1e-4 * np.abs(s0).max()
