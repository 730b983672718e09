"""ctypes loader + numpy-facing wrappers for the CPU oracle (oracle/oracle.cc).

TEST INFRASTRUCTURE ONLY: the product (tdnn-f_nas_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")

USE_GUMBEL, FREE_SELECT, UNIFORM_SAMPLE, USE_ENTROPY, UPDATE_ALPHA = 1, 2, 4, 8, 16

_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(HERE, f) for f in ("oracle.cc", "oracle_ng.inc")]
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(f) for f in srcs):
        subprocess.run(["make", "-C", HERE, "-B" if force else "-s", "_build/liboracle.so"], check=True,
                       capture_output=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.orc_den_forward_backward.restype = C.c_float
        _lib.orc_num_forward_backward.restype = C.c_double
        _lib.orc_ng_create.restype = C.c_void_p
    return _lib


def _f(a):
    assert a.dtype == np.float32
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    assert a.dtype == np.int32
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _d(a):
    assert a.dtype == np.float64
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _stride(a):
    assert a.ndim == 2 and a.strides[1] == a.itemsize
    return a.strides[0] // a.itemsize


def enable_blas(on: bool = True) -> str:
    """AddMatMat through cblas_sgemm (what a Kaldi CPU build calls) instead of the plain OpenMP loops: binds the OpenBLAS
    that numpy bundles (no system BLAS exists in the image).  Returns a description of what was bound ('' if nothing)."""
    import glob

    if not on:
        lib().orc_use_blas(None)
        return ""
    cands = []
    for pkg in ("numpy", "scipy"):
        try:
            mod = __import__(pkg)
            cands += sorted(glob.glob(os.path.join(os.path.dirname(os.path.dirname(mod.__file__)), pkg + ".libs", "libscipy_openblas*.so")))
        except Exception:
            pass
    for path in cands:
        if lib().orc_use_blas(path.encode()) == 1:
            return "OpenBLAS cblas_sgemm (" + os.path.basename(path) + ", bundled with numpy/scipy)"
    return ""


def num_threads() -> int:
    return lib().orc_num_threads()


def use_all_cores() -> int:
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline is meant to use every host core."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().orc_set_num_threads(n)
    return num_threads()


def share_index(time_offsets) -> int:
    to = np.asarray(time_offsets, dtype=np.int32)
    return lib().orc_share_index(_i(to), len(to))


def darts_coef(log_alpha, flags, temp, u_gumbel=None, u_uniform=0.0):
    la = np.ascontiguousarray(log_alpha, dtype=np.float32)
    n = len(la)
    ug = np.ascontiguousarray(u_gumbel if u_gumbel is not None else np.full(n, 0.5), dtype=np.float32)
    coef = np.zeros(n, dtype=np.float32)
    lib().orc_darts_coef(_f(la), n, flags, C.c_float(temp), _f(ug), C.c_float(u_uniform), _f(coef))
    return coef


def tdnn_propagate(time_offsets, flags, temp, W, bias_params, x, out_rows, row_offsets, row_stride,
                   u_gumbel=None, u_uniform=0.0, out=None):
    """Returns (out, coef_memo).  `out` (if given) is the pre-existing output (kPropagateAdds case)."""
    to = np.asarray(time_offsets, dtype=np.int32)
    n = len(to)
    out_dim = W.shape[0]
    in_dim = W.shape[1] // n
    if out is None:
        out = np.zeros((out_rows, out_dim), dtype=np.float32)
    ro = np.asarray(row_offsets, dtype=np.int32)
    ug = np.ascontiguousarray(u_gumbel if u_gumbel is not None else np.full(n, 0.5), dtype=np.float32)
    coef = np.zeros(n, dtype=np.float32)
    bp = None if bias_params is None else _f(bias_params)
    rc = lib().orc_tdnn_propagate(_i(to), n, flags, C.c_float(temp), _f(W), _stride(W), bp, _f(x), x.shape[0], in_dim,
                                  _stride(x), _f(out), out_rows, out_dim, _stride(out), _i(ro), row_stride, _f(ug),
                                  C.c_float(u_uniform), _f(coef))
    if rc != 0:
        raise RuntimeError(f"oracle: reference behaviour undefined here (rc={rc})")
    return out, coef


class NaturalGradient:
    """OnlineNaturalGradient state (oracle_ng.inc): precondition(X) overwrites X and returns the scale."""

    def __init__(self, rank=40, update_period=1, num_samples_history=2000.0, alpha=4.0):
        self.h = C.c_void_p(lib().orc_ng_create(rank, update_period, C.c_float(num_samples_history), C.c_float(alpha)))

    def precondition(self, X):
        scale = C.c_float(0.0)
        lib().orc_ng_precondition(self.h, _f(X), X.shape[0], X.shape[1], _stride(X), C.byref(scale))
        return float(scale.value)

    def freeze(self, frozen=True):
        lib().orc_ng_freeze(self.h, int(frozen))

    def skip_initial_updates(self):
        """Timing aid: past the 10 initial every-call updates, into the one-in-update_period steady state."""
        lib().orc_ng_skip_initial_updates(self.h)

    def state(self):
        t, rank, D, nre, nfl = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        rho = C.c_float()
        lib().orc_ng_state(self.h, C.byref(t), C.byref(rank), C.byref(D), C.byref(rho), None, None, C.byref(nre), C.byref(nfl))
        d = np.zeros(max(rank.value, 0), dtype=np.float32)
        W = np.zeros((max(rank.value, 0), max(D.value, 0)), dtype=np.float32)
        if t.value > 0 and rank.value > 0:
            lib().orc_ng_state(self.h, None, None, None, None, _f(d), _f(W), None, None)
        return dict(t=t.value, rank=rank.value, D=D.value, rho=float(rho.value), d=d, W=W, num_reorth=nre.value,
                    num_floored=nfl.value)

    def __del__(self):
        try:
            lib().orc_ng_destroy(self.h)
        except Exception:
            pass


def tdnn_backprop(time_offsets, flags, temp, W, x, out_deriv, coef, row_offsets, row_stride, lr,
                  in_deriv=None, dW=None, dbias=None, ng_in=None, ng_out=None, scales=None):
    """in_deriv (added to), dW, dbias (n+out_dim; added to) are modified in place.  Returns raw s (n,).
    ng_in / ng_out: NaturalGradient objects (None = identity); scales: optional float32[2] <- (in_scale, out_scale)."""
    to = np.asarray(time_offsets, dtype=np.int32)
    n = len(to)
    out_dim = W.shape[0]
    in_dim = W.shape[1] // n
    ro = np.asarray(row_offsets, dtype=np.int32)
    s = np.zeros(n, dtype=np.float32)
    coef = np.ascontiguousarray(coef, dtype=np.float32)
    rc = lib().orc_tdnn_backprop_ng(
        _i(to), n, flags, C.c_float(temp), _f(W), _stride(W), _f(x), x.shape[0], in_dim, _stride(x), _f(out_deriv),
        out_deriv.shape[0], out_dim, _stride(out_deriv), _f(coef), _i(ro), row_stride,
        None if in_deriv is None else _f(in_deriv), 0 if in_deriv is None else _stride(in_deriv), C.c_float(lr),
        None if dW is None else _f(dW), 0 if dW is None else _stride(dW), None if dbias is None else _f(dbias), _f(s),
        None if ng_in is None else ng_in.h, None if ng_out is None else ng_out.h,
        None if scales is None else _f(scales))
    if rc != 0:
        raise RuntimeError(f"oracle: reference behaviour undefined here (rc={rc})")
    return s


def plain_tdnn_propagate(W, bias, x, out_rows, row_offsets, row_stride, out=None):
    """Upstream TdnnComponent::Propagate (every w_i = 1; bias (D_out) or None = kPropagateAdds onto `out`)."""
    ro = np.asarray(row_offsets, dtype=np.int32)
    n = len(ro)
    out_dim, in_dim = W.shape[0], W.shape[1] // n
    if out is None:
        out = np.zeros((out_rows, out_dim), dtype=np.float32)
    lib().orc_plain_tdnn_propagate(n, _f(W), _stride(W), None if bias is None else _f(bias), _f(x), x.shape[0], in_dim,
                                   _stride(x), _f(out), out_rows, out_dim, _stride(out), _i(ro), row_stride)
    return out


def plain_tdnn_backprop(W, x, out_deriv, row_offsets, row_stride, lr, in_deriv=None, dW=None, dbias=None,
                        natural_gradient=True, ng_in=None, ng_out=None, scales=None):
    """Upstream TdnnComponent::Backprop (+ UpdateSimple / UpdateNaturalGradient); in_deriv, dW, dbias added to in place."""
    ro = np.asarray(row_offsets, dtype=np.int32)
    n = len(ro)
    out_dim, in_dim = W.shape[0], W.shape[1] // n
    lib().orc_plain_tdnn_backprop(
        n, _f(W), _stride(W), _f(x), x.shape[0], in_dim, _stride(x), _f(out_deriv), out_deriv.shape[0], out_dim,
        _stride(out_deriv), _i(ro), row_stride, None if in_deriv is None else _f(in_deriv),
        0 if in_deriv is None else _stride(in_deriv), C.c_float(lr), None if dW is None else _f(dW),
        0 if dW is None else _stride(dW), None if dbias is None else _f(dbias), int(natural_gradient),
        None if ng_in is None else ng_in.h, None if ng_out is None else ng_out.h, None if scales is None else _f(scales))


def constrain_orthonormal(M, scale):
    """ConstrainOrthonormalInternal (utils.cc:914-1035) on M (rows <= cols) in place; returns info
    (scale used, ratio, update_speed, ||M M^T - scale^2 I||_F)."""
    info = np.zeros(4, dtype=np.float32)
    rc = lib().orc_constrain_orthonormal(C.c_float(scale), _f(M), M.shape[0], M.shape[1], _stride(M), _f(info))
    if rc != 0:
        raise RuntimeError("oracle: the reference asserts here (scale == 0 or ratio <= 0.999)")
    return info


def softmax_flops_fwd(x, u=None, temp=1.0):
    out = np.zeros_like(x)
    uu = None if u is None else np.ascontiguousarray(u, dtype=np.float32)
    lib().orc_softmax_flops_fwd(_f(x), x.shape[0], x.shape[1], _stride(x), _f(out), _stride(out),
                                None if uu is None else _f(uu), C.c_float(temp))
    return out


def softmax_flops_bwd(out_value, out_deriv, scale, is_gumbel, temp=1.0, in_place=False):
    """Returns (in_deriv, out_deriv_after).  out_deriv is copied first; the reference mutates it."""
    od = np.array(out_deriv, dtype=np.float32, copy=True)
    ind = od if in_place else np.zeros_like(od)
    lib().orc_softmax_flops_bwd(_f(out_value), _stride(out_value), _f(od), _stride(od), _f(ind), _stride(ind),
                                od.shape[0], od.shape[1], C.c_float(scale), int(is_gumbel), C.c_float(temp))
    return ind, od


def copyn_fwd(x, out, scale):
    lib().orc_copyn_fwd(_f(x), x.shape[0], x.shape[1], _stride(x), _f(out), out.shape[1], _stride(out), C.c_float(scale))
    return out


def copyn_bwd(od, ind, scale):
    lib().orc_copyn_bwd(_f(od), od.shape[0], od.shape[1], _stride(od), _f(ind), ind.shape[1], _stride(ind),
                        C.c_float(scale))
    return ind


def onehot_fwd(rows, dim, u):
    out = np.zeros((rows, dim), dtype=np.float32)
    lib().orc_onehot_fwd(_f(out), rows, dim, dim, C.c_float(u))
    return out


def add_row_sum(mat, scale, vec):
    lib().orc_add_row_sum(_f(mat), mat.shape[0], mat.shape[1], _stride(mat), C.c_float(scale), _f(vec))
    return vec


def bn_test_derived(stats_sum, stats_sumsq, count, epsilon, target_rms):
    dim = len(stats_sum)
    scale = np.zeros(dim, dtype=np.float32)
    offset = np.zeros(dim, dtype=np.float32)
    lib().orc_bn_test_derived(_d(np.ascontiguousarray(stats_sum, dtype=np.float64)),
                              _d(np.ascontiguousarray(stats_sumsq, dtype=np.float64)), C.c_double(count), dim,
                              C.c_float(epsilon), C.c_float(target_rms), _f(scale), _f(offset))
    return scale, offset


def scale_offset_rows(x, scale, offset=None):
    out = np.zeros_like(x)
    lib().orc_scale_offset_rows(_f(x), x.shape[0], x.shape[1], _stride(x), _f(out), _stride(out), _f(scale),
                                None if offset is None else _f(offset))
    return out


def ewprod_fwd(x):
    D = x.shape[1] // 2
    out = np.zeros((x.shape[0], D), dtype=np.float32)
    lib().orc_ewprod_fwd(_f(x), x.shape[0], D, _stride(x), _f(out), _stride(out))
    return out


def ewprod_bwd(x, od):
    D = x.shape[1] // 2
    ind = np.zeros_like(x)
    lib().orc_ewprod_bwd(_f(x), _stride(x), _f(od), _stride(od), _f(ind), _stride(ind), x.shape[0], D)
    return ind


def den_forward_backward(graph, nnet_output, S, T, leaky, deriv_weight=None):
    """graph: dict(num_states, num_pdfs, fwd_ranges[N,2], bwd_ranges[N,2], prob, pdf, state, init).
    Returns (logprob, deriv or None, ok)."""
    N, P = graph["num_states"], graph["num_pdfs"]
    fr = np.ascontiguousarray(graph["fwd_ranges"], dtype=np.int32)
    br = np.ascontiguousarray(graph["bwd_ranges"], dtype=np.int32)
    pr = np.ascontiguousarray(graph["prob"], dtype=np.float32)
    pd = np.ascontiguousarray(graph["pdf"], dtype=np.int32)
    st = np.ascontiguousarray(graph["state"], dtype=np.int32)
    init = np.ascontiguousarray(graph["init"], dtype=np.float32)
    ok = C.c_int(1)
    deriv = None
    if deriv_weight is not None:
        deriv = np.zeros_like(nnet_output)
    lp = lib().orc_den_forward_backward(
        N, P, _i(fr), _i(br), _f(pr), _i(pd), _i(st), _f(init), S, T, C.c_float(leaky), _f(nnet_output),
        _stride(nnet_output), C.c_float(0.0 if deriv_weight is None else deriv_weight),
        None if deriv is None else _f(deriv), 0 if deriv is None else _stride(deriv), C.byref(ok))
    return float(lp), deriv, bool(ok.value)


def num_forward_backward(graph, nnet_output, T, deriv_weight=None):
    """graph: dict from synth.make_num_graphs.  Returns (logprob, deriv or None, ok)."""
    S = graph["num_seqs"]
    so = np.ascontiguousarray(graph["state_offsets"], dtype=np.int32)
    fr = np.ascontiguousarray(graph["fwd_ranges"], dtype=np.int32)
    lp = np.ascontiguousarray(graph["arc_logprob"], dtype=np.float32)
    pd = np.ascontiguousarray(graph["arc_pdf"], dtype=np.int32)
    st = np.ascontiguousarray(graph["arc_state"], dtype=np.int32)
    fl = np.ascontiguousarray(graph["final_logprob"], dtype=np.float32)
    ok = C.c_int(1)
    deriv = None if deriv_weight is None else np.zeros_like(nnet_output)
    tot = lib().orc_num_forward_backward(S, _i(so), _i(fr), _f(lp), _i(pd), _i(st), _f(fl), _f(nnet_output),
                                         _stride(nnet_output), T, C.c_float(deriv_weight or 0.0),
                                         None if deriv is None else _f(deriv), 0 if deriv is None else _stride(deriv),
                                         C.byref(ok))
    return float(tot), deriv, bool(ok.value)
