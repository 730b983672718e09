"""An independent float64 numpy restatement of the DARTS component math,
written from the equations -- not from oracle.cc -- to serve as a second opinion on the oracle."""
import numpy as np

USE_GUMBEL, FREE_SELECT, UNIFORM_SAMPLE, USE_ENTROPY, UPDATE_ALPHA = 1, 2, 4, 8, 16


def coef(alpha, flags, T, u_gumbel=None, u_uniform=0.0):
    a = np.asarray(alpha, dtype=np.float64)
    n = len(a)
    if flags & USE_GUMBEL:
        z = (a - np.log(-np.log(np.asarray(u_gumbel, dtype=np.float64)))) / T
        c = np.exp(z - z.max())
        c = np.maximum(c / c.sum(), 1e-20)
    elif flags & FREE_SELECT:
        c = 1.0 / (1.0 + np.exp(-a))
    else:
        c = np.exp(a - a.max())
        c = np.maximum(c / c.sum(), 1e-20)
    if flags & UNIFORM_SAMPLE:
        i = np.arange(n)
        lo = (i.astype(np.float32) / np.float32(n)).astype(np.float64)
        hi = ((i + 1).astype(np.float32) / np.float32(n)).astype(np.float64)
        c = ((u_uniform >= lo) & (u_uniform < hi)).astype(np.float64)
    return c


def weights(c, flags, share):
    n = len(c)
    if flags & UNIFORM_SAMPLE:
        return np.array([1.0 if (i == share or c[i] == 1.0) else 0.0 for i in range(n)])
    if flags & FREE_SELECT:
        return c.copy()
    w = c.copy()
    w[share] = 1.0
    return w


def views(x, out_rows, row_offsets, row_stride):
    return [x[o: o + out_rows * row_stride: row_stride][:out_rows] for o in row_offsets]


def propagate(W, bias_tail, x, out_rows, row_offsets, row_stride, w):
    n = len(row_offsets)
    din = W.shape[1] // n
    out = np.zeros((out_rows, W.shape[0])) if bias_tail is None else np.tile(bias_tail, (out_rows, 1)).astype(np.float64)
    for i, xi in enumerate(views(x.astype(np.float64), out_rows, row_offsets, row_stride)):
        out += w[i] * xi @ W[:, i * din:(i + 1) * din].astype(np.float64).T
    return out


def backprop(W, x, od, row_offsets, row_stride, w, c, flags, T, share, lr):
    """Returns (in_deriv, dW, dbias_tail, s, dalpha) for zero-initialised deltas."""
    n = len(row_offsets)
    din = W.shape[1] // n
    W = W.astype(np.float64)
    od = od.astype(np.float64)
    x = x.astype(np.float64)
    out_rows = od.shape[0]
    ind = np.zeros_like(x)
    dW = np.zeros_like(W)
    s = np.zeros(n)
    for i, o in enumerate(row_offsets):
        Wi = W[:, i * din:(i + 1) * din]
        sl = slice(o, o + out_rows * row_stride, row_stride)
        xi = x[sl][:out_rows]
        ind_rows = np.arange(o, o + out_rows * row_stride, row_stride)[:out_rows]
        ind[ind_rows] += w[i] * od @ Wi
        G = od.T @ xi
        dW[:, i * din:(i + 1) * din] = lr * w[i] * G
        s[i] = np.sum(G * Wi)
    dbias = lr * od.sum(axis=0)
    dalpha = np.zeros(n)
    if not (flags & UNIFORM_SAMPLE):
        tau = T if (flags & USE_GUMBEL) else 1.0
        for i in range(n):
            if flags & FREE_SELECT:
                dalpha[i] += s[i] * c[i] * (1 - c[i])
            elif i != share:
                dalpha -= (s[i] / tau) * c[i] * c
                dalpha[i] += (s[i] / tau) * c[i]
    else:
        s[:] = 0
    mul = 1.0
    if flags & USE_ENTROPY:
        mul *= 5
    mul *= lr if ((flags & USE_GUMBEL) and not (flags & FREE_SELECT)) else 5 * lr
    if flags & UPDATE_ALPHA:
        mul *= 10000
    return ind, dW, dbias, s, dalpha * mul
