"""Shared helpers for the parity tests."""
import numpy as np


def rel_err(a, b) -> float:
    """||a-b||_F / ||b||_F in float64 (b is the oracle)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (den if den > 0 else 1.0))


def max_rel_to_scale(a, b) -> float:
    """max|a-b| / max|b|: catches localised errors (a wrong tile) that a Frobenius norm can hide."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    s = np.abs(b).max()
    return float(np.abs(a - b).max() / (s if s > 0 else 1.0))


def padded(rows, cols, pad, rng=None, scale=1.0, dtype=np.float32):
    """A (rows, cols) view into a (rows, cols+pad) buffer: exercises stride > cols like CuMatrix pitch."""
    buf = np.full((rows, cols + pad), np.nan, dtype=dtype)
    if rng is not None:
        buf[:, :cols] = rng.standard_normal((rows, cols)).astype(dtype) * scale
    else:
        buf[:, :cols] = 0
    return buf[:, :cols]


def to_cuda_view(a):
    """Copy a (possibly strided) numpy 2-D view to the GPU keeping the same row stride."""
    import torch

    base_cols = a.strides[0] // a.itemsize
    full = np.lib.stride_tricks.as_strided(a, shape=(a.shape[0], base_cols), strides=a.strides)
    t = torch.from_numpy(np.nan_to_num(np.ascontiguousarray(full), nan=1e30)).cuda()
    return t[:, : a.shape[1]]
