"""GPU tests of GeneralDropoutComponent: masks replayed from the component layer's counter hash."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def nn(ctx):
    from tdnnf_nas_b200 import nnet3

    nnet3.set_context(ctx)
    nnet3.set_rand_seed(991)
    return nnet3


def _replay_mask(nn, c0, rows, cols, p, continuous):
    nn.set_rand_counter(c0)
    u = np.array([nn.rand_uniform() for _ in range(rows * cols)], dtype=np.float32).reshape(rows, cols)
    p32, one, two, four = np.float32(p), np.float32(1.0), np.float32(2.0), np.float32(4.0)
    if continuous:  # GetMemo: Scale(4 p) then Add(1 - 2 p), all in fp32
        return u * (p32 * four) + (one - two * p32)
    return (u - p32 > 0).astype(np.float32) * (one / (one - p32))


@pytest.mark.parametrize("continuous", [True, False])
@pytest.mark.parametrize("dim,block_dim,time_period,pad", [(48, 48, 0, 0), (48, 48, 3, 4), (30, 30, 0, 1), (48, 16, 0, 0)])
def test_general_dropout_vs_numpy(nn, dim, block_dim, time_period, pad, continuous):
    import torch

    p, S, frames = 0.3, 5, 8
    cfg = f"dim={dim} block-dim={block_dim} time-period={time_period} dropout-proportion={p} continuous={'true' if continuous else 'false'}"
    comp = nn.Component.new("GeneralDropoutComponent", cfg)
    grid = [(n, t, 0) for t in range(frames) for n in range(S)]
    idx = comp.precompute_indexes(grid, grid)
    toks = idx.write(False).decode().split()
    num_mask_rows = int(toks[2])
    index = np.array([int(t) for t in toks[toks.index("[") + 1: toks.index("]")]])
    g = np.random.default_rng(dim + time_period)
    x = g.standard_normal((len(grid), dim)).astype(np.float32)
    buf = torch.full((len(grid), dim + pad), 5.0, device="cuda")
    buf[:, :dim] = torch.from_numpy(x).cuda()
    xin = buf[:, :dim] if pad else buf
    out = torch.empty((len(grid), dim), device="cuda")
    c0 = nn.get_rand_counter()
    memo = comp.propagate(idx, xin, out)
    c1 = nn.get_rand_counter()
    assert memo and c1 - c0 == num_mask_rows * block_dim
    mask = _replay_mask(nn, c0, num_mask_rows, block_dim, p, continuous)
    nn.set_rand_counter(c1)
    mult = dim // block_dim
    ref = (x.reshape(len(grid) * mult, block_dim) * mask[index]).reshape(len(grid), dim)
    np.testing.assert_array_equal(out.cpu().numpy(), ref)
    if continuous:
        assert mask.min() >= 1 - 2 * p - 1e-6 and mask.max() <= 1 + 2 * p + 1e-6
    else:
        assert set(np.unique(mask)) <= {np.float32(0.0), np.float32(1.0) / (np.float32(1.0) - np.float32(p))}
    # every frame of a sequence shares its mask row when time-period = 0
    if time_period == 0 and mult == 1:
        assert all(index[t * S + n] == n for t in range(frames) for n in range(S))
    # Backprop, in place (kBackpropInPlace): in_deriv aliases out_deriv
    od = g.standard_normal((len(grid), dim)).astype(np.float32)
    odd = torch.from_numpy(od).cuda()
    comp.backprop(idx, None, None, odd, memo, None, odd)
    np.testing.assert_array_equal(odd.cpu().numpy(), (od.reshape(-1, block_dim) * mask[index]).reshape(len(grid), dim))
    comp.delete_memo(memo)
    # Propagate in place (kPropagateInPlace) draws a NEW mask
    if not pad:
        c2 = nn.get_rand_counter()
        memo2 = comp.propagate(idx, xin, xin)
        mask2 = _replay_mask(nn, c2, num_mask_rows, block_dim, p, continuous)
        nn.set_rand_counter(c2 + num_mask_rows * block_dim)
        np.testing.assert_array_equal(xin.cpu().numpy(), (x.reshape(-1, block_dim) * mask2[index]).reshape(len(grid), dim))
        comp.delete_memo(memo2)
    elif pad:
        assert bool((buf[:, dim:] == 5.0).all())


def test_general_dropout_passthrough(nn):
    import torch

    comp = nn.Component.new("GeneralDropoutComponent", "dim=32 dropout-proportion=0.0 continuous=true")
    grid = [(n, t, 0) for t in range(4) for n in range(3)]
    idx = comp.precompute_indexes(grid, grid)
    x = torch.randn(12, 32, device="cuda")
    out = torch.zeros_like(x)
    c0 = nn.get_rand_counter()
    assert comp.propagate(idx, x, out) is None and nn.get_rand_counter() == c0  # proportion 0: a copy, no draws, no memo
    assert torch.equal(out, x)
    d = torch.zeros_like(x)
    comp.backprop(idx, None, None, x, None, None, d)
    assert torch.equal(d, x)
    # the schedule switches it on; test mode switches it off again
    nn.apply_edits("set-dropout-proportion name=* proportion=0.4", [("tdnnf2.dropout", comp)])
    memo = comp.propagate(idx, x, out)
    assert memo is not None and not torch.equal(out, x)
    comp.delete_memo(memo)
    comp.set_test_mode(True)
    assert comp.propagate(idx, x, out) is None and torch.equal(out, x)
