"""ctypes binding of the C ABI (include/tdnnf_nas_b200.h) for Python callers.

torch is used only for device memory and streams: every wrapper passes raw device pointers,
sizes and strides to the shared library.  There is no fallback: if the library is missing or a
call fails, a RuntimeError carrying tdnnf_last_error() is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libtdnnf_nas_b200.so")

USE_GUMBEL, FREE_SELECT, UNIFORM_SAMPLE, USE_ENTROPY, UPDATE_ALPHA = 1, 2, 4, 8, 16

_lib = None

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int32)
vp = C.c_void_p


def _declare(lib):
    def sig(name, argtypes, restype=C.c_int):
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype

    i, f = C.c_int, C.c_float
    sig("tdnnf_last_error", [], C.c_char_p)
    sig("tdnnf_abi_version", [])
    sig("tdnnf_ctx_create", [i, C.POINTER(vp)])
    sig("tdnnf_ctx_destroy", [vp])
    sig("tdnnf_ctx_set_stream", [vp, vp])
    sig("tdnnf_ctx_reserve", [vp, C.c_uint64])
    sig("tdnnf_ctx_launch_count", [vp], C.c_uint64)
    sig("tdnnf_ctx_gemm_timing_enable", [vp, i])
    sig("tdnnf_ctx_gemm_timing_read", [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint64)])
    sig("tdnnf_ctx_gemm_timing_read_ex", [vp, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                          C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_uint64)])
    sig("tdnnf_ng_gram_scale", [vp, vp, i, i, i, vp, i, vp, i, vp, vp, i, i, c_int_p, i, vp, f, vp])
    sig("tdnnf_ng_w_update", [vp, vp, i, vp, i, vp, i, vp, i, i, i, vp, i])
    sig("tdnnf_copy_blocks", [vp, i, vp, vp, vp, vp, vp, vp])
    sig("tdnnf_ng_project_gradient", [vp, vp, i, i, i, vp, i, i, vp, i, i])
    sig("tdnnf_multi_sumsq", [vp, i, vp, vp, vp, vp, vp, vp])
    sig("tdnnf_multi_axpy_zero", [vp, i, vp, vp, vp, vp, vp, vp, vp])
    sig("tdnnf_ctx_set_gradient_mode", [vp, i])
    sig("tdnnf_ctx_set_gemm_planes", [vp, i])
    sig("tdnnf_ctx_operand_cache_begin", [vp, C.POINTER(C.c_void_p), i])
    sig("tdnnf_ctx_operand_cache_end", [vp])
    sig("tdnnf_ctx_operand_cache_stats", [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)])
    sig("tdnnf_darts_coef", [vp, vp, i, i, f, c_float_p, f, i, vp, vp])
    sig("tdnnf_darts_weff_from_coef", [vp, vp, i, i, i, vp])
    sig("tdnnf_darts_propagate", [vp, vp, i, i, i, vp, i, i, i, vp, i, vp, i, vp, i, c_int_p, i])
    sig("tdnnf_darts_project", [vp, vp, i, i, i, vp, i, i, i, vp, i, vp, vp, i, c_int_p, i])
    sig("tdnnf_darts_backprop_data", [vp, vp, i, i, i, vp, i, i, i, vp, i, vp, i, c_int_p, i])
    sig("tdnnf_darts_backprop_params", [vp, vp, i, i, i, vp, i, i, i, vp, i, vp, i, vp, vp, i, c_int_p, i, f, vp])
    sig("tdnnf_darts_alpha_update", [vp, vp, vp, i, i, f, i, f, vp])
    sig("tdnnf_softmax_flops_fwd", [vp, vp, i, i, i, vp, i, c_float_p, f])
    sig("tdnnf_softmax_flops_bwd", [vp, vp, i, vp, i, vp, i, i, i, f, f, i])
    sig("tdnnf_copyn_fwd", [vp, vp, i, i, i, vp, i, i, f])
    sig("tdnnf_copyn_bwd", [vp, vp, i, i, i, vp, i, i, f])
    sig("tdnnf_onehot_fwd", [vp, vp, i, i, i, f])
    sig("tdnnf_add_row_sum", [vp, vp, i, i, i, f, vp])
    sig("tdnnf_scale_offset_rows", [vp, vp, i, i, i, vp, i, vp, vp])
    sig("tdnnf_elementwise_product_fwd", [vp, vp, i, i, i, vp, i])
    sig("tdnnf_elementwise_product_bwd", [vp, vp, i, vp, i, vp, i, i, i])
    sig("tdnnf_shared_mask_fwd", [vp, vp, i, i, i, vp, i, i, vp, i, c_int_p, f])
    sig("tdnnf_shared_mask_bwd", [vp, vp, i, vp, i, vp, i, vp, i, vp, i, i, i, i, c_int_p, f])
    sig("tdnnf_mat_set", [vp, vp, i, i, i, f])
    sig("tdnnf_mat_scale", [vp, vp, i, i, i, f])
    sig("tdnnf_mat_axpy", [vp, f, vp, i, vp, i, i, i])
    sig("tdnnf_mat_dot", [vp, vp, i, vp, i, i, i, c_float_p])
    sig("tdnnf_mat_dot_dev", [vp, vp, i, vp, i, i, i, vp])
    sig("tdnnf_copy_rows_from_vec", [vp, vp, vp, i, i, i])
    sig("tdnnf_copy_rows", [vp, vp, i, vp, i, i, i, vp])
    sig("tdnnf_add_to_rows", [vp, f, vp, i, i, i, vp, i, vp])
    sig("tdnnf_relu_fwd", [vp, vp, i, i, i, vp, i])
    sig("tdnnf_relu_bwd", [vp, vp, i, vp, i, vp, i, i, i])
    sig("tdnnf_add_scaled", [vp, vp, i, f, vp, i, f, vp, i, i, i])
    sig("tdnnf_relu_scale_offset_bypass_fwd", [vp, vp, i, i, i, vp, vp, vp, i, f, vp, i])
    sig("tdnnf_relu_scale_offset_bypass_bwd", [vp, vp, i, vp, i, vp, f, vp, i, vp, i, i, i])
    sig("tdnnf_relu_scale_offset_bypass_fwd_planes", [vp, vp, i, i, i, vp, vp, vp, i, f, vp, i, C.POINTER(vp)])
    sig("tdnnf_relu_scale_offset_bypass_bwd_planes", [vp, vp, i, vp, i, vp, f, vp, i, vp, i, i, i, C.POINTER(vp)])
    sig("tdnnf_planes_acquire", [vp, vp, i, i, i, i, C.POINTER(vp)])
    sig("tdnnf_planes_release", [vp])
    sig("tdnnf_ctx_planes_attach", [vp, vp])
    sig("tdnnf_ctx_planes_detach", [vp, vp])
    sig("tdnnf_batchnorm_train_fwd", [vp, vp, i, i, i, vp, i, f, f, vp])
    sig("tdnnf_batchnorm_train_bwd", [vp, vp, i, vp, i, vp, i, i, i, f, vp])
    sig("tdnnf_ctx_set_wgrad_mn_min_rows", [vp, i])
    sig("tdnnf_constrain_orthonormal", [vp, vp, i, i, i, f, vp])
    sig("tdnnf_log_softmax_fwd", [vp, vp, i, i, i, vp, i])
    sig("tdnnf_log_softmax_bwd", [vp, vp, i, vp, i, vp, i, i, i])
    sig("tdnnf_dropout_mask", [vp, C.c_uint64, C.c_uint64, vp, i, i, i, f, i])
    sig("tdnnf_mul_rows_indexed", [vp, vp, i, vp, i, i, i, vp, i, vp])
    sig("tdnnf_num_graph_create", [vp, i, c_int_p, i, c_int_p, c_int_p, c_float_p, c_int_p, c_int_p, c_float_p, C.POINTER(vp)])
    sig("tdnnf_num_graph_destroy", [vp])
    sig("tdnnf_num_graph_update", [vp, i, c_int_p, i, c_int_p, c_int_p, c_float_p, c_int_p, c_int_p, c_float_p])
    sig("tdnnf_num_forward_backward", [vp, vp, vp, i, i, f, vp, i, c_float_p, c_int_p])
    sig("tdnnf_den_graph_create", [vp, i, i, i, c_int_p, c_int_p, c_float_p, c_int_p, c_int_p, c_float_p, C.POINTER(vp)])
    sig("tdnnf_den_graph_destroy", [vp])
    sig("tdnnf_den_create", [vp, vp, i, i, f, C.POINTER(vp)])
    sig("tdnnf_den_destroy", [vp])
    sig("tdnnf_den_describe", [vp, c_int_p, c_int_p, c_int_p, c_int_p])
    sig("tdnnf_den_forward", [vp, vp, i, c_float_p])
    sig("tdnnf_den_backward", [vp, f, vp, i, c_int_p])
    sig("tdnnf_update_with_max_change", [vp, i, vp, vp, vp, vp, vp, vp, vp, i, c_float_p, f, f, f, f, vp, c_float_p, c_int_p,
                                         c_int_p, c_int_p])
    sig("tdnnf_apply_l2_regularization", [vp, i, vp, vp, vp, vp, vp, vp, vp, i, c_float_p, c_float_p, f])
    sig("tdnnf_chain_objf_and_deriv", [vp, vp, vp, vp, i, i, i, i, f, f, f, i, i, vp, i, vp, i, c_float_p, c_float_p, c_float_p])
    sig("tdnnf_penalize_out_of_range", [vp, vp, i, i, i, f, f, i, i, vp, i])
    pp_i, pp_f = C.POINTER(c_int_p), C.POINTER(c_float_p)
    sig("tdnnf_den_graph_parse_fst_text", [C.c_char_p, C.c_uint64, i, C.POINTER(vp)])
    sig("tdnnf_host_graph_dims", [vp, c_int_p, c_int_p, c_int_p])
    sig("tdnnf_host_graph_arrays", [vp, pp_i, pp_i, pp_f, pp_i, pp_i, pp_f])
    sig("tdnnf_host_graph_free", [vp])
    sig("tdnnf_den_graph_create_from_host", [vp, vp, C.POINTER(vp)])
    sig("tdnnf_num_graph_parse_fst_texts", [C.POINTER(C.c_char_p), C.POINTER(C.c_uint64), i, i, C.POINTER(vp)])
    sig("tdnnf_host_num_graph_arrays", [vp, c_int_p, c_int_p, pp_i, pp_i, pp_i, pp_f, pp_i, pp_i, pp_f])
    sig("tdnnf_host_num_graph_free", [vp])
    sig("tdnnf_num_graph_create_from_host", [vp, vp, C.POINTER(vp)])
    sig("tdnnf_den_graph_parse_fst_binary", [C.c_char_p, C.c_uint64, i, C.POINTER(vp)])
    pp_c = C.POINTER(C.c_char_p)
    sig("tdnnf_chain_egs_read_ark", [C.c_char_p, C.c_uint64, i, C.POINTER(vp)])
    sig("tdnnf_chain_egs_free", [vp])
    sig("tdnnf_chain_egs_count", [vp, c_int_p])
    sig("tdnnf_chain_egs_example", [vp, i, pp_c, c_int_p, c_int_p, c_int_p])
    sig("tdnnf_chain_egs_input", [vp, i, i, pp_c, c_int_p, c_int_p, pp_i, pp_f])
    sig("tdnnf_chain_egs_supervision", [vp, i, i, pp_c, c_float_p, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p, pp_i, c_int_p, pp_f,
                                        c_int_p, pp_i])
    sig("tdnnf_chain_egs_fst", [vp, i, i, i, c_int_p, c_int_p, c_int_p, pp_i, pp_f, c_int_p, pp_i, pp_f])
    sig("tdnnf_chain_egs_merge_input", [vp, i, i, C.c_char_p, vp, C.c_int64, c_int_p, c_int_p, c_int_p, c_int_p])
    sig("tdnnf_chain_egs_merge_supervision", [vp, i, i, C.c_char_p, i, vp, i, c_int_p, c_int_p, c_float_p, C.POINTER(vp)])
    sig("tdnnf_dp_unique_id", [C.c_char_p, i])
    sig("tdnnf_dp_comm_create", [vp, i, i, C.c_char_p, C.POINTER(vp)])
    sig("tdnnf_dp_comm_adopt", [vp, vp, i, i, C.POINTER(vp)])
    sig("tdnnf_dp_comm_destroy", [vp])
    sig("tdnnf_dp_allreduce_deltas", [vp, i, vp, C.POINTER(C.c_int64)])
    sig("tdnnf_dp_allreduce_bucket_async", [vp, vp, C.c_int64])
    sig("tdnnf_dp_allreduce_wait", [vp])
    sig("tdnnf_dp_nccl_version", [c_int_p])


def load():
    """Load the shared library (raises if it has not been built: there is no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python tdnn-f_nas_b200/build.py` "
                "(the product has no CPU or PyTorch fallback)")
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


class TdnnfError(RuntimeError):
    pass


def check(rc: int):
    if rc != 0:
        raise TdnnfError(f"tdnnf error {rc}: {load().tdnnf_last_error().decode(errors='replace')}")


def _ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def _mat(t):
    """(ptr, rows, cols, stride) of a 2-D fp32 CUDA tensor whose last dim is contiguous."""
    import torch

    assert t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1, "need a row-major fp32 CUDA matrix"
    return t.data_ptr(), t.shape[0], t.shape[1], t.stride(0)


def _fhost(a: Optional[Sequence[float]]):
    if a is None:
        return None
    arr = (C.c_float * len(a))(*[float(x) for x in a])
    return arr


def _ihost(a: Sequence[int]):
    return (C.c_int32 * len(a))(*[int(x) for x in a])


class Context:
    """RAII wrapper of tdnnf_ctx bound to a CUDA device and (optionally) a torch stream."""

    def __init__(self, device: int = 0, stream=None):
        lib = load()
        h = vp()
        check(lib.tdnnf_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device
        if stream is not None:
            self.set_stream(stream)

    def set_stream(self, stream):
        check(load().tdnnf_ctx_set_stream(self.h, vp(stream.cuda_stream if hasattr(stream, "cuda_stream") else stream)))

    def use_current_stream(self):
        import torch

        self.set_stream(torch.cuda.current_stream(self.device))

    def reserve(self, nbytes: int):
        check(load().tdnnf_ctx_reserve(self.h, nbytes))

    @property
    def launches(self) -> int:
        return int(load().tdnnf_ctx_launch_count(self.h))

    def gemm_timing_enable(self, on: bool):
        check(load().tdnnf_ctx_gemm_timing_enable(self.h, int(on)))

    def gemm_timing_read(self):
        """(total_ms, total_algorithmic_flops, launches) of the tensor-core GEMM launches since enabling."""
        ms, fl, n = C.c_double(), C.c_double(), C.c_uint64()
        check(load().tdnnf_ctx_gemm_timing_read(self.h, C.byref(ms), C.byref(fl), C.byref(n)))
        return ms.value, fl.value, int(n.value)

    def gemm_timing_read_ex(self, min_flops: float):
        """Split by size: dict(ms, flops, pipe_flops, launches) of the launches with >= min_flops algorithmic FLOPs
        (pipe_flops counts every tensor-core product issued) and (other_ms, other_launches) of the rest."""
        ms, fl, raw, oms = C.c_double(), C.c_double(), C.c_double(), C.c_double()
        n, on = C.c_uint64(), C.c_uint64()
        check(load().tdnnf_ctx_gemm_timing_read_ex(self.h, float(min_flops), C.byref(ms), C.byref(fl), C.byref(raw), C.byref(n),
                                                   C.byref(oms), C.byref(on)))
        return dict(ms=ms.value, flops=fl.value, pipe_flops=raw.value, launches=int(n.value), other_ms=oms.value,
                    other_launches=int(on.value))

    def set_gemm_planes(self, planes: int):
        """2 = bf16 hi/lo operand planes (three products), 3 = hi/mid/lo (six products, fp32-level)."""
        check(load().tdnnf_ctx_set_gemm_planes(self.h, int(planes)))

    def set_gradient_mode(self, fast: bool):
        """fast: data gradient with two tensor-core products, parameter gradient with one (see the header)."""
        check(load().tdnnf_ctx_set_gradient_mode(self.h, int(fast)))

    def operand_cache_stats(self):
        h, m = C.c_uint64(), C.c_uint64()
        check(load().tdnnf_ctx_operand_cache_stats(self.h, C.byref(h), C.byref(m)))
        return int(h.value), int(m.value)

    def close(self):
        if getattr(self, "h", None):
            load().tdnnf_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ DARTS
    def darts_coef(self, alpha, flags, temperature, u_gumbel, u_uniform, share_index, coef, weff):
        n = alpha.numel()
        check(load().tdnnf_darts_coef(self.h, _ptr(alpha), n, flags, temperature, _fhost(u_gumbel), u_uniform,
                                      share_index, _ptr(coef), _ptr(weff)))

    def darts_weff_from_coef(self, coef, flags, share_index, weff):
        check(load().tdnnf_darts_weff_from_coef(self.h, _ptr(coef), coef.numel(), flags, share_index, _ptr(weff)))

    def darts_propagate(self, x, out, W, bias, bias_mode, weff, row_offsets, row_stride):
        xp, xr, xc, xs = _mat(x)
        op, orr, oc, os_ = _mat(out)
        wp, wr, wc, ws = _mat(W)
        n = len(row_offsets)
        assert wr == oc and wc == n * xc
        check(load().tdnnf_darts_propagate(self.h, xp, xr, xc, xs, op, orr, oc, os_, wp, ws, _ptr(bias), bias_mode,
                                           _ptr(weff), n, _ihost(row_offsets), row_stride))

    def darts_project(self, x, out, W, bias, weff, row_offsets, row_stride):
        """out = [w_1 X_1 | ... | w_n X_n] W[:, :n*in_dim]^T (+ bias): skinny-W form of darts_propagate."""
        xp, xr, xc, xs = _mat(x)
        op, orr, oc, os_ = _mat(out)
        wp, wr, wc, ws = _mat(W)
        n = len(row_offsets)
        assert wr == oc and wc >= n * xc
        check(load().tdnnf_darts_project(self.h, xp, xr, xc, xs, op, orr, oc, os_, wp, ws, _ptr(bias), _ptr(weff), n,
                                         _ihost(row_offsets), row_stride))

    def darts_backprop_data(self, out_deriv, in_deriv, W, weff, row_offsets, row_stride):
        dp, dr, dc, ds = _mat(out_deriv)
        ip, ir, ic, is_ = _mat(in_deriv)
        wp, wr, wc, ws = _mat(W)
        n = len(row_offsets)
        assert wr == dc and wc == n * ic
        check(load().tdnnf_darts_backprop_data(self.h, dp, dr, dc, ds, ip, ir, ic, is_, wp, ws, _ptr(weff), n,
                                               _ihost(row_offsets), row_stride))

    def darts_backprop_params(self, x, out_deriv, W_model, dW, dbias, weff, row_offsets, row_stride, lr, s=None):
        xp, xr, xc, xs = _mat(x)
        dp, dr, dc, ds = _mat(out_deriv)
        gp, gr, gc, gs = _mat(dW)
        n = len(row_offsets)
        wptr, wstride = (0, 0)
        if W_model is not None:
            wptr, _, _, wstride = _mat(W_model)
        check(load().tdnnf_darts_backprop_params(self.h, xp, xr, xc, xs, dp, dr, dc, ds, wptr, wstride, gp, gs,
                                                 _ptr(dbias), _ptr(weff), n, _ihost(row_offsets), row_stride, lr,
                                                 _ptr(s)))

    def darts_alpha_update(self, s, coef, flags, temperature, share_index, lr, dalpha):
        check(load().tdnnf_darts_alpha_update(self.h, _ptr(s), _ptr(coef), coef.numel(), flags, temperature,
                                              share_index, lr, _ptr(dalpha)))

    # ------------------------------------------------------------------ mixing components
    def softmax_flops_fwd(self, x, out, u=None, inv_temp=1.0):
        xp, r, c, xs = _mat(x)
        op, _, _, os_ = _mat(out)
        check(load().tdnnf_softmax_flops_fwd(self.h, xp, r, c, xs, op, os_, _fhost(u), inv_temp))

    def softmax_flops_bwd(self, out_value, out_deriv, in_deriv, penalty, inv_temp=1.0, write_back_e=1):
        vp_, r, c, vs = _mat(out_value)
        dp, _, _, ds = _mat(out_deriv)
        ip, _, _, is_ = _mat(in_deriv)
        check(load().tdnnf_softmax_flops_bwd(self.h, vp_, vs, dp, ds, ip, is_, r, c, penalty, inv_temp, write_back_e))

    def copyn_fwd(self, x, out, scale):
        xp, r, c, xs = _mat(x)
        op, _, oc, os_ = _mat(out)
        check(load().tdnnf_copyn_fwd(self.h, xp, r, c, xs, op, oc, os_, scale))

    def copyn_bwd(self, out_deriv, in_deriv, scale):
        dp, r, oc, ds = _mat(out_deriv)
        ip, _, ic, is_ = _mat(in_deriv)
        check(load().tdnnf_copyn_bwd(self.h, dp, r, oc, ds, ip, ic, is_, scale))

    def onehot_fwd(self, out, u):
        op, r, c, os_ = _mat(out)
        check(load().tdnnf_onehot_fwd(self.h, op, r, c, os_, u))

    def add_row_sum(self, mat, scale, vec):
        mp, r, c, ms = _mat(mat)
        check(load().tdnnf_add_row_sum(self.h, mp, r, c, ms, scale, _ptr(vec)))

    def scale_offset_rows(self, x, out, scale, offset=None):
        xp, r, c, xs = _mat(x)
        op, _, _, os_ = _mat(out)
        check(load().tdnnf_scale_offset_rows(self.h, xp, r, c, xs, op, os_, _ptr(scale), _ptr(offset)))

    def elementwise_product_fwd(self, x, out):
        xp, r, c, xs = _mat(x)
        op, _, oc, os_ = _mat(out)
        check(load().tdnnf_elementwise_product_fwd(self.h, xp, r, oc, xs, op, os_))

    def elementwise_product_bwd(self, x, out_deriv, in_deriv):
        xp, r, c, xs = _mat(x)
        dp, _, oc, ds = _mat(out_deriv)
        ip, _, _, is_ = _mat(in_deriv)
        check(load().tdnnf_elementwise_product_bwd(self.h, xp, xs, dp, ds, ip, is_, r, oc))


    def shared_mask_fwd(self, p, lin, out, widths, scale=1.0):
        pp, r, nb, ps = _mat(p)
        lp, _, c, ls = _mat(lin)
        op, _, _, os_ = _mat(out)
        check(load().tdnnf_shared_mask_fwd(self.h, pp, r, nb, ps, lp, c, ls, op, os_, _ihost(widths), scale))

    def shared_mask_bwd(self, p, lin, d_out, d_lin, d_p, widths, scale=1.0):
        pp, r, nb, ps = _mat(p)
        lp, _, c, ls = _mat(lin)
        dp, _, _, ds = _mat(d_out)
        gp, gs = (0, 0) if d_lin is None else (_mat(d_lin)[0], _mat(d_lin)[3])
        qp, _, _, qs = _mat(d_p)
        check(load().tdnnf_shared_mask_bwd(self.h, pp, ps, lp, ls, dp, ds, gp, gs, qp, qs, r, c, nb, _ihost(widths), scale))

    # ------------------------------------------------------------------ parameter ops / neighbours
    def mat_set(self, a, value):
        p, r, c, s = _mat(a)
        check(load().tdnnf_mat_set(self.h, p, r, c, s, value))

    def mat_scale(self, a, scale):
        p, r, c, s = _mat(a)
        check(load().tdnnf_mat_scale(self.h, p, r, c, s, scale))

    def mat_axpy(self, alpha, src, dst):
        sp, r, c, ss = _mat(src)
        dp, _, _, ds = _mat(dst)
        check(load().tdnnf_mat_axpy(self.h, alpha, sp, ss, dp, ds, r, c))

    def mat_dot(self, a, b) -> float:
        ap, r, c, as_ = _mat(a)
        bp, _, _, bs = _mat(b)
        out = C.c_float(0)
        check(load().tdnnf_mat_dot(self.h, ap, as_, bp, bs, r, c, C.byref(out)))
        return float(out.value)

    def copy_rows(self, src, dst, row_map):
        sp, _, c, ss = _mat(src)
        dp, r, _, ds = _mat(dst)
        check(load().tdnnf_copy_rows(self.h, sp, ss, dp, ds, r, c, row_map.data_ptr()))

    def add_to_rows(self, alpha, src, dst, row_map):
        sp, r, c, ss = _mat(src)
        dp, _, _, ds = _mat(dst)
        check(load().tdnnf_add_to_rows(self.h, alpha, sp, ss, r, c, dp, ds, row_map.data_ptr()))

    def set_wgrad_mn_min_rows(self, min_rows: int):
        """Parameter gradients of operands with >= min_rows rows use the MN-major form (default 512; 1 = always, -1 = never)."""
        check(load().tdnnf_ctx_set_wgrad_mn_min_rows(self.h, int(min_rows)))

    def penalize_out_of_range(self, nnet_output, deriv, limit: float, scale: float, row_step: int = 1, row_offset: int = 0):
        xp, r, c, xs = _mat(nnet_output)
        dp, _, _, ds = _mat(deriv)
        check(load().tdnnf_penalize_out_of_range(self.h, xp, r, c, xs, limit, scale, row_step, row_offset, dp, ds))

    def constrain_orthonormal(self, m, scale: float, info=None):
        """ConstrainOrthonormalInternal (nnet-utils.cc:914-1035) on the device matrix m, in place; scale < 0 = floating.
        info: optional device float[4] <- (scale used, ratio, update_speed, ||M M^T - scale^2 I||_F)."""
        p, r, c, s = _mat(m)
        check(load().tdnnf_constrain_orthonormal(self.h, p, r, c, s, scale, _ptr(info)))

    def relu_fwd(self, x, out):
        xp, r, c, xs = _mat(x)
        op, _, _, os_ = _mat(out)
        check(load().tdnnf_relu_fwd(self.h, xp, r, c, xs, op, os_))

    def relu_bwd(self, out_value, out_deriv, in_deriv):
        vp_, r, c, vs = _mat(out_value)
        dp, _, _, ds = _mat(out_deriv)
        ip, _, _, is_ = _mat(in_deriv)
        check(load().tdnnf_relu_bwd(self.h, vp_, vs, dp, ds, ip, is_, r, c))

    def add_scaled(self, a, alpha, b, beta, out):
        ap, r, c, as_ = _mat(a)
        bp, _, _, bs = _mat(b)
        op, _, _, os_ = _mat(out)
        check(load().tdnnf_add_scaled(self.h, ap, as_, alpha, bp, bs, beta, op, os_, r, c))

    def relu_scale_offset_bypass_fwd(self, x, scale, offset, prev, bypass_scale, out):
        xp, r, c, xs = _mat(x)
        pp, _, _, ps = _mat(prev)
        op, _, _, os_ = _mat(out)
        check(load().tdnnf_relu_scale_offset_bypass_fwd(self.h, xp, r, c, xs, _ptr(scale), _ptr(offset), pp, ps,
                                                        bypass_scale, op, os_))

    def relu_scale_offset_bypass_bwd(self, d_out, x, scale, bypass_scale, d_x, d_prev):
        dp, r, c, ds = _mat(d_out)
        xp, _, _, xs = _mat(x)
        gp, _, _, gs = _mat(d_x)
        pp, _, _, ps = _mat(d_prev)
        check(load().tdnnf_relu_scale_offset_bypass_bwd(self.h, dp, ds, xp, xs, _ptr(scale), bypass_scale, gp, gs, pp,
                                                        ps, r, c))

    def batchnorm_train_fwd(self, x, out, memo, epsilon=1e-3, target_rms=1.0):
        xp, r, c, xs = _mat(x)
        op, _, _, os_ = _mat(out)
        check(load().tdnnf_batchnorm_train_fwd(self.h, xp, r, c, xs, op, os_, epsilon, target_rms, memo.data_ptr()))

    def batchnorm_train_bwd(self, out_value, out_deriv, in_deriv, memo, target_rms=1.0):
        vp_, r, c, vs = _mat(out_value)
        dp, _, _, ds = _mat(out_deriv)
        ip, _, _, is_ = _mat(in_deriv)
        check(load().tdnnf_batchnorm_train_bwd(self.h, vp_, vs, dp, ds, ip, is_, r, c, target_rms, memo.data_ptr()))


def parse_den_fst_text(text: str, num_pdfs: int) -> dict:
    """den.fst in FSM text form -> the DenominatorGraph arrays (tdnnf_den_graph_parse_fst_text: SetTransitions +
    SetInitialProbs of kaldi chain-den-graph.cc).  Host only: no GPU needed.  Same dict layout as synth.make_den_graph."""
    import numpy as np

    data = text.encode() if isinstance(text, str) else text
    h = vp()
    check(load().tdnnf_den_graph_parse_fst_text(data, len(data), num_pdfs, C.byref(h)))
    try:
        n, p, a = C.c_int32(), C.c_int32(), C.c_int32()
        check(load().tdnnf_host_graph_dims(h, C.byref(n), C.byref(p), C.byref(a)))
        fr, br, pd, st = c_int_p(), c_int_p(), c_int_p(), c_int_p()
        pr, init = c_float_p(), c_float_p()
        check(load().tdnnf_host_graph_arrays(h, C.byref(fr), C.byref(br), C.byref(pr), C.byref(pd), C.byref(st), C.byref(init)))
        arr = lambda ptr, count, dt: np.ctypeslib.as_array(ptr, shape=(count,)).astype(dt).copy()
        N, A2 = n.value, a.value
        return dict(num_states=N, num_pdfs=p.value, fwd_ranges=arr(fr, 2 * N, np.int32).reshape(N, 2),
                    bwd_ranges=arr(br, 2 * N, np.int32).reshape(N, 2), prob=arr(pr, A2, np.float32), pdf=arr(pd, A2, np.int32),
                    state=arr(st, A2, np.int32), init=arr(init, N, np.float32), num_arcs=A2 // 2)
    finally:
        load().tdnnf_host_graph_free(h)


def _host_graph_dict(h) -> dict:
    import numpy as np

    n, p, a = C.c_int32(), C.c_int32(), C.c_int32()
    check(load().tdnnf_host_graph_dims(h, C.byref(n), C.byref(p), C.byref(a)))
    fr, br, pd, st = c_int_p(), c_int_p(), c_int_p(), c_int_p()
    pr, init = c_float_p(), c_float_p()
    check(load().tdnnf_host_graph_arrays(h, C.byref(fr), C.byref(br), C.byref(pr), C.byref(pd), C.byref(st), C.byref(init)))
    arr = lambda ptr, count, dt: np.ctypeslib.as_array(ptr, shape=(count,)).astype(dt).copy()
    N, A2 = n.value, a.value
    return dict(num_states=N, num_pdfs=p.value, fwd_ranges=arr(fr, 2 * N, np.int32).reshape(N, 2),
                bwd_ranges=arr(br, 2 * N, np.int32).reshape(N, 2), prob=arr(pr, A2, np.float32), pdf=arr(pd, A2, np.int32),
                state=arr(st, A2, np.int32), init=arr(init, N, np.float32), num_arcs=A2 // 2)


def parse_den_fst_binary(data: bytes, num_pdfs: int) -> dict:
    """den.fst as chain-make-den-fst writes it (OpenFst binary) -> the same dict as parse_den_fst_text.  Host only."""
    h = vp()
    check(load().tdnnf_den_graph_parse_fst_binary(data, len(data), num_pdfs, C.byref(h)))
    try:
        return _host_graph_dict(h)
    finally:
        load().tdnnf_host_graph_free(h)


def _host_num_graph_dict(h) -> dict:
    import numpy as np

    ns, na = C.c_int32(), C.c_int32()
    so, fr, br, pd, st = c_int_p(), c_int_p(), c_int_p(), c_int_p(), c_int_p()
    lp, fl = c_float_p(), c_float_p()
    check(load().tdnnf_host_num_graph_arrays(h, C.byref(ns), C.byref(na), C.byref(so), C.byref(fr), C.byref(br), C.byref(lp),
                                             C.byref(pd), C.byref(st), C.byref(fl)))
    arr = lambda ptr, count, dt: np.ctypeslib.as_array(ptr, shape=(count,)).astype(dt).copy() if count else np.zeros(0, dt)
    k = ns.value
    offs = arr(so, k + 1, np.int32)
    N, A = int(offs[-1]), na.value
    return dict(num_seqs=k, state_offsets=offs, num_arcs=A, fwd_ranges=arr(fr, 2 * N, np.int32).reshape(N, 2),
                bwd_ranges=arr(br, 2 * N, np.int32).reshape(N, 2), arc_logprob=arr(lp, 2 * A, np.float32),
                arc_pdf=arr(pd, 2 * A, np.int32), arc_state=arr(st, 2 * A, np.int32), final_logprob=arr(fl, N, np.float32))


class ChainEgs:
    """A Kaldi archive of NnetChainExample parsed on the host (tdnnf_chain_egs_*; formats in csrc/egs_io.cc).  `data` is
    the archive's bytes (binary `ark:` or text `ark,t:`).  Examples are dicts of numpy arrays; merge_* give the minibatch
    nnet3-chain-merge-egs would form from examples [first, first + count) in the row order the kernels take."""

    def __init__(self, data: bytes, max_examples: int = 0):
        self.h = vp()
        check(load().tdnnf_chain_egs_read_ark(data, len(data), max_examples, C.byref(self.h)))
        n = C.c_int32()
        check(load().tdnnf_chain_egs_count(self.h, C.byref(n)))
        self.count = n.value

    def close(self):
        if self.h:
            load().tdnnf_chain_egs_free(self.h)
            self.h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return self.count

    def example(self, i: int) -> dict:
        import numpy as np

        lib = load()
        arr = lambda ptr, shape, dt: (np.ctypeslib.as_array(ptr, shape=shape).astype(dt).copy() if int(np.prod(shape)) else np.zeros(shape, dt))
        key, binary, ni, no = C.c_char_p(), C.c_int32(), C.c_int32(), C.c_int32()
        check(lib.tdnnf_chain_egs_example(self.h, i, C.byref(key), C.byref(binary), C.byref(ni), C.byref(no)))
        ex = dict(key=key.value.decode(errors="replace"), binary=bool(binary.value), inputs=[], outputs=[])
        for j in range(ni.value):
            name, r, c_, ix, d = C.c_char_p(), C.c_int32(), C.c_int32(), c_int_p(), c_float_p()
            check(lib.tdnnf_chain_egs_input(self.h, i, j, C.byref(name), C.byref(r), C.byref(c_), C.byref(ix), C.byref(d)))
            ex["inputs"].append(dict(name=name.value.decode(errors="replace"), indexes=arr(ix, (r.value, 3), np.int32),
                                     data=arr(d, (r.value, c_.value), np.float32)))
        for j in range(no.value):
            name, w, ns, fps, ld, e2e, nf = C.c_char_p(), C.c_float(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
            ix, ndw, dw, nap, ap = c_int_p(), C.c_int32(), c_float_p(), C.c_int32(), c_int_p()
            check(lib.tdnnf_chain_egs_supervision(self.h, i, j, C.byref(name), C.byref(w), C.byref(ns), C.byref(fps), C.byref(ld),
                                                  C.byref(e2e), C.byref(nf), C.byref(ix), C.byref(ndw), C.byref(dw), C.byref(nap),
                                                  C.byref(ap)))
            fsts = []
            for k in range(nf.value):
                st, nst, na, nfin = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
                a3, aw, fs, fw = c_int_p(), c_float_p(), c_int_p(), c_float_p()
                check(lib.tdnnf_chain_egs_fst(self.h, i, j, k, C.byref(st), C.byref(nst), C.byref(na), C.byref(a3), C.byref(aw),
                                              C.byref(nfin), C.byref(fs), C.byref(fw)))
                fsts.append(dict(start=st.value, num_states=nst.value, arcs=arr(a3, (na.value, 3), np.int32),
                                 weights=arr(aw, (na.value,), np.float32), final_states=arr(fs, (nfin.value,), np.int32),
                                 final_weights=arr(fw, (nfin.value,), np.float32)))
            ex["outputs"].append(dict(name=name.value.decode(errors="replace"), weight=w.value, num_sequences=ns.value, frames_per_seq=fps.value,
                                      label_dim=ld.value, e2e=bool(e2e.value), fsts=fsts,
                                      indexes=arr(ix, (ns.value * fps.value, 3), np.int32), deriv_weights=arr(dw, (ndw.value,), np.float32),
                                      alignment_pdfs=arr(ap, (nap.value,), np.int32)))
        return ex

    def merge_input(self, first: int, count: int, name: str):
        """-> (array [num_t, num_seqs, dim] (row = t_rank * num_seqs + sequence), first_t)"""
        import numpy as np

        T, S, D, t0 = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        args = (self.h, first, count, name.encode())
        check(load().tdnnf_chain_egs_merge_input(*args, None, 0, C.byref(T), C.byref(S), C.byref(D), C.byref(t0)))
        out = np.empty((T.value, S.value, D.value), np.float32)
        check(load().tdnnf_chain_egs_merge_input(*args, out.ctypes.data, out.size, C.byref(T), C.byref(S), C.byref(D), C.byref(t0)))
        return out, t0.value

    def merge_supervision(self, first: int, count: int, name: str, num_pdfs: int, graph: bool = True) -> dict:
        """-> dict(num_seqs, frames_per_seq, weight, deriv_weights [frames_per_seq, num_seqs], num_graph (as parse_num_fst_texts))"""
        import numpy as np

        S, T, w = C.c_int32(), C.c_int32(), C.c_float()
        args = (self.h, first, count, name.encode(), num_pdfs)
        check(load().tdnnf_chain_egs_merge_supervision(*args, None, 0, C.byref(S), C.byref(T), C.byref(w), None))
        dw = np.empty((T.value, S.value), np.float32)
        g = vp()
        check(load().tdnnf_chain_egs_merge_supervision(*args, dw.ctypes.data, dw.size, C.byref(S), C.byref(T), C.byref(w),
                                                       C.byref(g) if graph else None))
        out = dict(num_seqs=S.value, frames_per_seq=T.value, weight=w.value, deriv_weights=dw)
        if graph:
            try:
                out["num_graph"] = _host_num_graph_dict(g)
            finally:
                load().tdnnf_host_num_graph_free(g)
        return out


def parse_num_fst_texts(texts: Sequence[str], num_pdfs: int) -> dict:
    """Per-sequence numerator FSTs in FSM text form -> the arrays of tdnnf_num_graph_create (host only).  Same dict
    layout as synth.make_num_graphs."""
    import numpy as np

    datas = [t.encode() if isinstance(t, str) else t for t in texts]
    k = len(datas)
    arr_t, arr_l = (C.c_char_p * k)(*datas), (C.c_uint64 * k)(*[len(d) for d in datas])
    h = vp()
    check(load().tdnnf_num_graph_parse_fst_texts(arr_t, arr_l, k, num_pdfs, C.byref(h)))
    try:
        ns, na = C.c_int32(), C.c_int32()
        so, fr, br, pd, st = c_int_p(), c_int_p(), c_int_p(), c_int_p(), c_int_p()
        lp, fl = c_float_p(), c_float_p()
        check(load().tdnnf_host_num_graph_arrays(h, C.byref(ns), C.byref(na), C.byref(so), C.byref(fr), C.byref(br), C.byref(lp),
                                                 C.byref(pd), C.byref(st), C.byref(fl)))
        arr = lambda ptr, count, dt: np.ctypeslib.as_array(ptr, shape=(count,)).astype(dt).copy()
        offs = arr(so, k + 1, np.int32)
        N, A = int(offs[-1]), na.value
        return dict(num_seqs=k, state_offsets=offs, num_arcs=A, fwd_ranges=arr(fr, 2 * N, np.int32).reshape(N, 2),
                    bwd_ranges=arr(br, 2 * N, np.int32).reshape(N, 2), arc_logprob=arr(lp, 2 * A, np.float32),
                    arc_pdf=arr(pd, 2 * A, np.int32), arc_state=arr(st, 2 * A, np.int32), final_logprob=arr(fl, N, np.float32))
    finally:
        load().tdnnf_host_num_graph_free(h)


class ParamTable:
    """The parameter buffers of a network next to their delta buffers, grouped by updatable component: the argument
    block of tdnnf_update_with_max_change / tdnnf_apply_l2_regularization (ref: nnet-utils.cc:2085-2175, 2223-2245).
    bufs: list of (model_ptr, model_stride, delta_ptr, delta_stride, rows, cols, group)."""

    def __init__(self, ctx: "Context", bufs, max_change: Sequence[float]):
        import torch

        n = len(bufs)
        self.ctx, self.n, self.num_groups = ctx, n, len(max_change)
        P, I = vp * n, C.c_int32 * n
        self.model, self.model_ld = P(*[b[0] for b in bufs]), I(*[b[1] for b in bufs])
        self.delta, self.delta_ld = P(*[b[2] for b in bufs]), I(*[b[3] for b in bufs])
        self.rows, self.cols, self.groups = I(*[b[4] for b in bufs]), I(*[b[5] for b in bufs]), I(*[b[6] for b in bufs])
        self.max_change = (C.c_float * self.num_groups)(*[float(m) for m in max_change])
        self.dots = torch.zeros(self.num_groups, dtype=torch.float64, device=torch.device("cuda", ctx.device))
        self.factors = (C.c_float * self.num_groups)()
        self.num_per_component = (C.c_int32 * self.num_groups)()
        self.num_global = C.c_int32(0)

    def update_with_max_change(self, max_param_change: float, max_change_scale: float = 1.0, scale: float = 1.0,
                               momentum: float = 0.0) -> bool:
        """Returns False where the reference returns false ("Infinite parameter change, will not apply.")."""
        applied = C.c_int32(0)
        check(load().tdnnf_update_with_max_change(
            self.ctx.h, self.n, self.model, self.model_ld, self.delta, self.delta_ld, self.rows, self.cols, self.groups,
            self.num_groups, self.max_change, max_param_change, max_change_scale, scale, momentum, vp(self.dots.data_ptr()),
            self.factors, self.num_per_component, C.byref(self.num_global), C.byref(applied)))
        return bool(applied.value)

    def apply_l2_regularization(self, lrate: Sequence[float], l2: Sequence[float], l2_regularize_scale: float):
        check(load().tdnnf_apply_l2_regularization(
            self.ctx.h, self.n, self.model, self.model_ld, self.delta, self.delta_ld, self.rows, self.cols, self.groups,
            self.num_groups, _fhost(lrate), _fhost(l2), l2_regularize_scale))


class DataParallel:
    """tdnnf_dp_*: the NCCL sum of the ranks' deltas through the C ABI.  `exchange` ships rank 0's 128-byte NCCL id to
    the other ranks (e.g. a torch.distributed broadcast): plumbing, not data path."""

    def __init__(self, ctx: "Context", nranks: int, rank: int, exchange):
        self.ctx, self.nranks, self.rank = ctx, nranks, rank
        buf = C.create_string_buffer(128)
        if rank == 0:
            check(load().tdnnf_dp_unique_id(buf, 128))
        ident = exchange(bytes(buf.raw))
        assert len(ident) == 128
        h = vp()
        check(load().tdnnf_dp_comm_create(ctx.h, nranks, rank, ident, C.byref(h)))
        self.h = h

    def allreduce(self, ptrs_and_counts):
        n = len(ptrs_and_counts)
        P, L = vp * n, C.c_int64 * n
        check(load().tdnnf_dp_allreduce_deltas(self.h, n, P(*[p for p, _ in ptrs_and_counts]), L(*[c for _, c in ptrs_and_counts])))

    def allreduce_bucket_async(self, ptr: int, count: int):
        check(load().tdnnf_dp_allreduce_bucket_async(self.h, vp(ptr), count))

    def wait(self):
        check(load().tdnnf_dp_allreduce_wait(self.h))

    @staticmethod
    def nccl_version() -> int:
        v = C.c_int32(0)
        check(load().tdnnf_dp_nccl_version(C.byref(v)))
        return int(v.value)

    def close(self):
        if getattr(self, "h", None):
            load().tdnnf_dp_comm_destroy(self.h)
            self.h = None


class DenGraph:
    """DenominatorGraph on the device.  Arrays are host (numpy) as in the C ABI."""

    def __init__(self, ctx: Context, graph: dict):
        import numpy as np

        self.ctx = ctx
        self.num_states = int(graph["num_states"])
        self.num_pdfs = int(graph["num_pdfs"])
        fr = np.ascontiguousarray(graph["fwd_ranges"], dtype=np.int32)
        br = np.ascontiguousarray(graph["bwd_ranges"], dtype=np.int32)
        pr = np.ascontiguousarray(graph["prob"], dtype=np.float32)
        pd = np.ascontiguousarray(graph["pdf"], dtype=np.int32)
        st = np.ascontiguousarray(graph["state"], dtype=np.int32)
        init = np.ascontiguousarray(graph["init"], dtype=np.float32)
        self.num_transitions = len(pr)
        h = vp()
        check(load().tdnnf_den_graph_create(
            ctx.h, self.num_states, self.num_pdfs, len(pr), fr.ctypes.data_as(c_int_p), br.ctypes.data_as(c_int_p),
            pr.ctypes.data_as(c_float_p), pd.ctypes.data_as(c_int_p), st.ctypes.data_as(c_int_p),
            init.ctypes.data_as(c_float_p), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            load().tdnnf_den_graph_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DenominatorComputation:
    def __init__(self, ctx: Context, graph: DenGraph, num_seqs: int, frames_per_seq: int, leaky: float):
        self.ctx, self.graph = ctx, graph
        h = vp()
        check(load().tdnnf_den_create(ctx.h, graph.h, num_seqs, frames_per_seq, leaky, C.byref(h)))
        self.h = h

    def describe(self) -> dict:
        """Which kernels run: path 'slices' (cluster size, E parts, CTAs), 'frames' or 'resident'."""
        a, b, c, d = C.c_int32(0), C.c_int32(0), C.c_int32(0), C.c_int32(0)
        check(load().tdnnf_den_describe(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(path={0: "frames", 1: "resident", 2: "slices"}[a.value], cluster=b.value, parts=c.value, ctas=d.value)

    def forward(self, nnet_output) -> float:
        p, r, c, s = _mat(nnet_output)
        lp = C.c_float(0)
        check(load().tdnnf_den_forward(self.h, p, s, C.byref(lp)))
        return float(lp.value)

    def backward(self, deriv_weight: float, nnet_output_deriv) -> bool:
        p, r, c, s = _mat(nnet_output_deriv)
        ok = C.c_int32(0)
        check(load().tdnnf_den_backward(self.h, deriv_weight, p, s, C.byref(ok)))
        return bool(ok.value)

    def close(self):
        if getattr(self, "h", None):
            load().tdnnf_den_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class NumeratorGraph:
    """Per-sequence numerator FSTs on the device + GenericNumeratorComputation::ForwardBackward."""

    def __init__(self, ctx: Context, graph: dict):
        import numpy as np

        self.ctx = ctx
        a = lambda k, dt: np.ascontiguousarray(graph[k], dtype=dt)
        so, fr, br = a("state_offsets", np.int32), a("fwd_ranges", np.int32), a("bwd_ranges", np.int32)
        lp, pd, st, fl = a("arc_logprob", np.float32), a("arc_pdf", np.int32), a("arc_state", np.int32), a("final_logprob", np.float32)
        self.num_seqs = int(graph["num_seqs"])
        h = vp()
        check(load().tdnnf_num_graph_create(ctx.h, self.num_seqs, so.ctypes.data_as(c_int_p), int(graph["num_arcs"]),
                                            fr.ctypes.data_as(c_int_p), br.ctypes.data_as(c_int_p), lp.ctypes.data_as(c_float_p),
                                            pd.ctypes.data_as(c_int_p), st.ctypes.data_as(c_int_p), fl.ctypes.data_as(c_float_p),
                                            C.byref(h)))
        self.h = h

    @staticmethod
    def host_arrays(graph: dict):
        """The contiguous host arrays of a supervision dict, prepared once (what a data loader would hand over)."""
        import numpy as np

        a = lambda k, dt: np.ascontiguousarray(graph[k], dtype=dt)
        return dict(num_seqs=int(graph["num_seqs"]), num_arcs=int(graph["num_arcs"]), so=a("state_offsets", np.int32),
                    fr=a("fwd_ranges", np.int32), br=a("bwd_ranges", np.int32), lp=a("arc_logprob", np.float32),
                    pd=a("arc_pdf", np.int32), st=a("arc_state", np.int32), fl=a("final_logprob", np.float32))

    def update(self, h: dict) -> int:
        """The next minibatch's supervision (host_arrays) into this handle: asynchronous copies from a pinned staging
        buffer on the context's stream.  Returns the bytes copied."""
        check(load().tdnnf_num_graph_update(self.h, h["num_seqs"], h["so"].ctypes.data_as(c_int_p), h["num_arcs"],
                                            h["fr"].ctypes.data_as(c_int_p), h["br"].ctypes.data_as(c_int_p),
                                            h["lp"].ctypes.data_as(c_float_p), h["pd"].ctypes.data_as(c_int_p),
                                            h["st"].ctypes.data_as(c_int_p), h["fl"].ctypes.data_as(c_float_p)))
        self.num_seqs = h["num_seqs"]
        return sum(int(h[k].nbytes) for k in ("so", "fr", "br", "lp", "pd", "st", "fl"))

    def forward_backward(self, nnet_output, frames_per_seq: int, deriv_weight: float = 1.0, nnet_output_deriv=None):
        """Returns (logprob, ok); adds deriv_weight * posteriors into nnet_output_deriv if given."""
        p, r, c, s = _mat(nnet_output)
        dp, ds = (0, 0) if nnet_output_deriv is None else (_mat(nnet_output_deriv)[0], _mat(nnet_output_deriv)[3])
        lp, ok = C.c_float(0), C.c_int32(0)
        check(load().tdnnf_num_forward_backward(self.ctx.h, self.h, p, s, frames_per_seq, deriv_weight, dp, ds, C.byref(lp),
                                                C.byref(ok)))
        return float(lp.value), bool(ok.value)

    def close(self):
        if getattr(self, "h", None):
            load().tdnnf_num_graph_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
