"""The oracle's OnlineNaturalGradient (oracle/oracle_ng.inc, fp32, upstream code structure) against an independent
float64 restatement from the equations (tests/np_ref.py) and against the invariants of the method."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import np_ref
from tests.util import rel_err


def _data(g, N, D, k=6):
    """Rows with a few dominant directions (so that the low-rank Fisher estimate has something to find)."""
    basis = g.standard_normal((k, D))
    return (g.standard_normal((N, k)) * np.linspace(3.0, 1.0, k)) @ basis + 0.3 * g.standard_normal((N, D))


@pytest.mark.parametrize("N,D,rank,period", [(300, 41, 8, 1), (500, 97, 20, 4), (64, 33, 10, 4), (200, 161, 80, 4)])
def test_matches_float64_restatement(N, D, rank, period):
    g = np.random.default_rng(7 + N)
    orc = O.NaturalGradient(rank, period, 2000.0, 4.0)
    ref = np_ref.NaturalGradientF64(rank, period, 2000.0, 4.0)
    for step in range(16):  # passes the 10 initial updates and three periodic ones
        X = _data(g, N, D).astype(np.float32)
        Xo = X.copy()
        s_o = orc.precondition(Xo)
        Xr, s_r = ref.precondition(X)
        assert rel_err(Xo, Xr) < 2e-4, (step, rel_err(Xo, Xr))
        assert abs(s_o - s_r) / s_r < 2e-4, (step, s_o, s_r)
    st = orc.state()
    assert st["t"] == 16 and st["rank"] == rank and st["D"] == D
    assert abs(st["rho"] - ref.rho) / ref.rho < 5e-3
    assert rel_err(np.sort(st["d"]), np.sort(ref.d)) < 5e-3
    # W_t is defined up to the sign of each row: compare the projector W^T W
    assert rel_err(st["W"].T.astype(np.float64) @ st["W"], ref.W.T @ ref.W) < 5e-3


def test_invariants():
    g = np.random.default_rng(3)
    N, D, rank = 400, 57, 12
    ng = O.NaturalGradient(rank, 1, 2000.0, 4.0)
    for _ in range(6):
        X = _data(g, N, D).astype(np.float32)
        st0 = ng.state()
        Xh = X.copy()
        scale = ng.precondition(Xh)
        if st0["t"] > 0:
            # X_hat = X (I - W_t^T W_t) with the W_t from BEFORE the call
            W = st0["W"].astype(np.float64)
            assert rel_err(Xh, X - (X.astype(np.float64) @ W.T) @ W) < 1e-5
        # scale restores the Frobenius norm
        assert abs(scale * np.linalg.norm(Xh.astype(np.float64)) / np.linalg.norm(X.astype(np.float64)) - 1.0) < 1e-5
    st = ng.state()
    # R_t = E_t^{-1/2} W_t has orthonormal rows
    D_, d, rho = st["D"], st["d"].astype(np.float64), st["rho"]
    beta = rho * (1 + 4.0) + 4.0 * d.sum() / D_
    e = 1.0 / (beta / d + 1.0)
    R = st["W"].astype(np.float64) / np.sqrt(e)[:, None]
    assert np.abs(R @ R.T - np.eye(rank)).max() < 2e-3
    # the dominant data directions are captured: the largest eigenvalue excess dwarfs the isotropic floor
    assert d.max() > 3 * rho


def test_dim_one_and_frozen():
    ng = O.NaturalGradient(5, 1, 2000.0, 4.0)
    X = np.ones((10, 1), dtype=np.float32)
    assert ng.precondition(X) == 1.0 and (X == 1).all() and ng.state()["t"] == 0
    g = np.random.default_rng(0)
    ng = O.NaturalGradient(5, 1, 2000.0, 4.0)
    for _ in range(3):
        ng.precondition(_data(g, 100, 20).astype(np.float32))
    ng.freeze(True)
    W0 = ng.state()["W"].copy()
    ng.precondition(_data(g, 100, 20).astype(np.float32))
    assert (ng.state()["W"] == W0).all()


def test_rank_clipped_to_dim():
    ng = O.NaturalGradient(80, 4, 2000.0, 4.0)
    g = np.random.default_rng(1)
    ng.precondition(g.standard_normal((50, 9)).astype(np.float32))
    assert ng.state()["rank"] == 8
