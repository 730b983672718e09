"""world_size-2 gloo test (CPU) of the data-parallel scheme: per-rank deltas of a sequence shard,
all-reduced with sum, equal the delta of the whole minibatch, for the oracle's arithmetic."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from oracle import oracle as O
    from tdnnf_nas_b200 import parallel, synth

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    offsets, n, din, dout, S, t_out = [0, 1, 2, 3], 4, 12, 10, 6, 9
    g = np.random.default_rng(0)  # identical on every rank: parameters, noise and the GLOBAL minibatch
    W = (g.standard_normal((dout, n * din)) / 7).astype(np.float32)
    bp = g.standard_normal(n + dout).astype(np.float32)
    t_in = t_out + 3
    x = g.standard_normal((t_in * S, din)).astype(np.float32)
    od = (g.standard_normal((t_out * S, dout)) / (t_out * S)).astype(np.float32)
    ug = g.uniform(0.1, 0.9, n).astype(np.float32)
    flags, temp, lr = O.USE_GUMBEL, 0.7, 0.01

    def deltas(xs, ods, s_local):
        _, ro = synth.regular_row_offsets(offsets, 0, 0, s_local, 1, 1)
        _, coef = O.tdnn_propagate(offsets, flags, temp, W, bp, xs, ods.shape[0], ro, 1, ug)
        dW, db = np.zeros_like(W), np.zeros_like(bp)
        O.tdnn_backprop(offsets, flags, temp, W, xs, ods, coef, ro, 1, lr, dW=dW, dbias=db)
        return dW, db

    b, e = parallel.shard_sequences(S, rank, world)
    xs = np.ascontiguousarray(x[parallel.shard_rows(t_in, S, rank, world)])
    ods = np.ascontiguousarray(od[parallel.shard_rows(t_out, S, rank, world)])
    dW, db = deltas(xs, ods, e - b)
    views = [torch.from_numpy(dW), torch.from_numpy(db)]
    parallel.allreduce_deltas(views)
    if rank == 0:
        dW_full, db_full = deltas(x, od, S)
        np.save(os.path.join(tmp, "res.npy"), np.array([
            np.abs(dW - dW_full).max() / np.abs(dW_full).max(),
            np.abs(db[n:] - db_full[n:]).max() / np.abs(db_full[n:]).max(),
            np.abs(db[:n] - db_full[:n]).max() / np.abs(db_full[:n]).max()]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_helpers():
    from tdnnf_nas_b200 import parallel

    assert [parallel.shard_sequences(64, r, 8) for r in (0, 7)] == [(0, 8), (56, 64)]
    assert [parallel.shard_sequences(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert parallel.shard_rows(2, 4, 1, 2) == [2, 3, 6, 7]
    assert parallel.flops_penalty_normaliser(100, 8) == 800


@pytest.mark.timeout(180)
def test_sum_of_shard_deltas_equals_full_minibatch(tmp_path):
    import torch.multiprocessing as mp

    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    err = np.load(tmp_path / "res.npy")
    # theta and bias deltas are sums over rows: exact up to fp32 summation order.  The alpha delta is a
    # sum over rows of s_i pushed through the (shared) softmax Jacobian: linear in s_i, so it adds up too.
    assert err[0] < 1e-5 and err[1] < 1e-5 and err[2] < 1e-4, err
