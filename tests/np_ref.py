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


class NaturalGradientF64:
    """Independent float64 restatement of Kaldi's OnlineNaturalGradient, written from the equations of the
    method (F_t ~ R_t^T D_t R_t + rho_t I, W_t = E_t^{1/2} R_t, Z_t eigen-update), numpy.linalg.eigh for the
    eigenproblem.  A second opinion on oracle/oracle_ng.inc (which follows the upstream code structure in fp32)."""

    def __init__(self, rank=40, update_period=1, num_samples_history=2000.0, alpha=4.0):
        self.rank, self.update_period, self.nsh, self.alpha = rank, update_period, num_samples_history, alpha
        self.eps, self.delta, self.t, self.frozen = 1e-10, 5e-4, 0, False
        self.W = self.d = self.rho = None

    def _eta(self, N):
        return min(0.9, 1.0 - np.exp(-N / self.nsh))

    def _e(self, d, rho, D):
        beta = rho * (1 + self.alpha) + self.alpha * d.sum() / D
        return 1.0 / (beta / d + 1.0)

    def _init_default(self, D):
        if self.rank >= D:
            self.rank = D - 1
        R = self.rank
        self.rho, self.d = self.eps, np.full(R, self.eps)
        Rm = np.zeros((R, D))
        for r in range(R):
            cols = np.arange(r, D, R)
            Rm[r, cols] = 1.0
            Rm[r, cols[0]] = 1.1
            Rm[r] /= np.linalg.norm(Rm[r])
        self.W = Rm * np.sqrt(1.0 / (2.0 + (D + R) * self.alpha / D))

    def _updating(self):
        return (not self.frozen) and (self.t <= 10 or (self.t - 10) % self.update_period == 0)

    def precondition(self, X):
        """Returns (X_hat, scale)."""
        X = np.asarray(X, dtype=np.float64)
        N, D = X.shape
        if D == 1:
            return X, 1.0
        if self.t == 0:
            self._init_default(D)
            if N > self.rank:
                saved = self.t
                self.t = 1
                fr, self.frozen = self.frozen, False
                for _ in range(3):
                    self._step(X)
                    self.t += 1
                self.t, self.frozen = saved, fr
        Xh = self._step(X)
        self.t += 1
        ip, fp = (X * X).sum(), (Xh * Xh).sum()
        return Xh, (1.0 if ip <= 0 else float(np.sqrt(ip / fp)))

    def _step(self, X):
        N, D = X.shape
        R, W, d, rho, eta = self.rank, self.W, self.d, self.rho, self._eta(X.shape[0])
        H = X @ W.T
        Xh = X - H @ W
        if not self._updating():
            return Xh
        J = H.T @ X
        L, K = H.T @ H, J @ J.T
        e = self._e(d, rho, D)
        ise = 1.0 / np.sqrt(e)
        dr = d + rho
        etaN, eta1 = eta / N, 1.0 - eta
        Lt = ise[:, None] * L * ise[None, :]
        Z = etaN ** 2 * ise[:, None] * K * ise[None, :] + etaN * eta1 * (Lt * dr[None, :] + dr[:, None] * Lt) \
            + np.diag(eta1 ** 2 * dr ** 2)
        c, U = np.linalg.eigh(0.5 * (Z + Z.T))
        order = np.argsort(-np.abs(c))
        c, U = c[order], U[:, order]
        reorth = c[0] > 1e6 * c[-1]
        floor = (rho * (1 - eta)) ** 2
        if (c < floor).any():
            reorth = True
        c = np.maximum(c, floor)
        sc = np.sqrt(c)
        rho1 = (etaN * (X * X).sum() + eta1 * (D * rho + d.sum()) - sc.sum()) / (D - R)
        d1 = sc - rho1
        fl = max(self.eps, self.delta * sc.max())
        rho1, d1 = max(rho1, fl), np.maximum(d1, fl)
        e1 = self._e(d1, rho1, D)
        B = J + (eta1 / etaN) * dr[:, None] * W
        A = etaN * (np.sqrt(e1) / sc)[:, None] * U.T * ise[None, :]
        W1 = A @ B
        if reorth:
            # R_{t+1} = E^{-1/2} W_{t+1} should have orthonormal rows: restore it (Cholesky form)
            Rm = W1 / np.sqrt(e1)[:, None]
            O = Rm @ Rm.T
            if np.abs(O - np.eye(R)).max() > 1e-3:
                Ci = np.linalg.inv(np.linalg.cholesky(O))
                W1 = np.sqrt(e1)[:, None] * (Ci @ Rm)
        self.W, self.d, self.rho = W1, d1, rho1
        return Xh
