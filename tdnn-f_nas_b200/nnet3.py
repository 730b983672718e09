"""Python face of the nnet3 component mirror (csrc/nnet3): the calls NnetComputer would make.

Thin ctypes wrappers over the tdnnf_nnet3_* handle API; matrices are torch CUDA tensors used only
as device memory.  Errors raised inside the C++ components (KALDI_ERR / KALDI_ASSERT) surface as
Nnet3Error with the original message.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

from . import capi

# Component property flags (nnet3/nnet-component-itf.h)
kSimpleComponent, kUpdatableComponent, kPropagateInPlace, kPropagateAdds = 0x001, 0x002, 0x004, 0x008
kReordersIndexes, kBackpropAdds, kBackpropNeedsInput, kBackpropNeedsOutput = 0x010, 0x020, 0x040, 0x080
kBackpropInPlace, kStoresStats, kInputContiguous, kOutputContiguous = 0x100, 0x200, 0x400, 0x800
kUsesMemo, kRandomComponent = 0x1000, 0x2000
kNoTime = -32768

vp = C.c_void_p
_declared = False


class Nnet3Error(RuntimeError):
    pass


def _lib():
    global _declared
    lib = capi.load()
    if not _declared:
        lib.tdnnf_nnet3_last_error.restype = C.c_char_p
        lib.tdnnf_nnet3_get_rand_counter.restype = C.c_uint64
        lib.tdnnf_nnet3_rand_uniform.restype = C.c_float
        lib.tdnnf_nnet3_free.argtypes = [vp]
        lib.tdnnf_nnet3_free.restype = None
        _declared = True
    return lib


def _check(rc):
    if rc != 0:
        raise Nnet3Error(_lib().tdnnf_nnet3_last_error().decode(errors="replace"))  # messages may quote raw stream bytes


def _take_string(p: C.c_char_p, length: Optional[int] = None) -> bytes:
    data = C.string_at(p, length) if length is not None else C.string_at(p)
    _lib().tdnnf_nnet3_free(p)
    return data


def set_context(ctx: capi.Context):
    _check(_lib().tdnnf_nnet3_set_context(ctx.h))


def set_rand_seed(seed: int):
    _check(_lib().tdnnf_nnet3_set_rand_seed(C.c_uint64(seed)))


def set_rand_counter(counter: int):
    _check(_lib().tdnnf_nnet3_set_rand_counter(C.c_uint64(counter)))


def get_rand_counter() -> int:
    return int(_lib().tdnnf_nnet3_get_rand_counter())


def rand_uniform() -> float:
    return float(_lib().tdnnf_nnet3_rand_uniform())


def rand_int(lo: int, hi: int) -> int:
    """The RandInt draw the components make (advances the shared counter)."""
    return int(_lib().tdnnf_nnet3_rand_int(int(lo), int(hi)))


class arena:
    """with nnet3.arena(ptr, nbytes) as a: ... components created / copied inside take their device parameters from
    [ptr, ptr + nbytes) in creation order; a.used = bytes consumed (tdnnf_nnet3_arena_begin / _end)."""

    def __init__(self, ptr: int, nbytes: int):
        self.ptr, self.nbytes, self.used = ptr, nbytes, 0

    def __enter__(self):
        _check(_lib().tdnnf_nnet3_arena_begin(vp(self.ptr), C.c_uint64(self.nbytes)))
        return self

    def __exit__(self, *exc):
        used = C.c_uint64(0)
        _check(_lib().tdnnf_nnet3_arena_end(C.byref(used)))
        self.used = int(used.value)
        return False


def set_dp_world_size(g: int):
    _check(_lib().tdnnf_nnet3_set_dp_world_size(g))


def set_keep_planes(flag: bool):
    """TdnnDARTSV3Component: keep the operand planes of Propagate's input in the memo for Backprop (default on)."""
    _check(_lib().tdnnf_nnet3_set_keep_planes(int(flag)))


def set_print_log_alpha(flag: bool):
    _check(_lib().tdnnf_nnet3_set_print_log_alpha(int(flag)))


def set_fast_gradients(flag: bool):
    """TdnnDARTSV3Component parameter-gradient GEMM with one fp16 product instead of three bf16 ones.  Default False:
    with natural gradient the projection amplifies the 2.9e-4 error beyond the 1e-3 tolerance."""
    _check(_lib().tdnnf_nnet3_set_fast_gradients(int(flag)))


def set_ng_identity(flag: bool):
    """Diagnostic: every PreconditionDirections call becomes the identity with scale 1 (raw gradient)."""
    _check(_lib().tdnnf_nnet3_set_ng_identity(int(flag)))


class NaturalGradient:
    """OnlineNaturalGradient (kaldi natural-gradient-online.h): standalone, or borrowed from a component."""

    def __init__(self, rank=40, update_period=1, num_samples_history=2000.0, alpha=4.0, _borrowed=None):
        self.owned = _borrowed is None
        if self.owned:
            self.h = vp()
            _check(_lib().tdnnf_nnet3_ng_new(rank, update_period, C.c_float(num_samples_history), C.c_float(alpha),
                                             C.byref(self.h)))
        else:
            self.h = _borrowed

    def precondition(self, x) -> float:
        """x: torch CUDA float32 matrix, overwritten by the preconditioned directions; returns the scale."""
        p, r, c, s = capi._mat(x)
        scale = C.c_float(1.0)
        _check(_lib().tdnnf_nnet3_ng_precondition(self.h, vp(p), r, c, s, C.byref(scale)))
        return float(scale.value)

    def freeze(self, frozen=True):
        _check(_lib().tdnnf_nnet3_ng_freeze(self.h, int(frozen)))

    def state(self):
        import numpy as np

        t, rank, dim, nre = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        rho = C.c_float()
        _check(_lib().tdnnf_nnet3_ng_state(self.h, C.byref(t), C.byref(rank), C.byref(dim), C.byref(rho), None, None,
                                           C.byref(nre)))
        d = np.zeros(max(rank.value, 0), dtype=np.float32)
        W = np.zeros((max(rank.value, 0), max(dim.value, 0)), dtype=np.float32)
        if dim.value > 0:
            _check(_lib().tdnnf_nnet3_ng_state(self.h, None, None, None, None, d.ctypes.data_as(C.POINTER(C.c_float)),
                                               W.ctypes.data_as(C.POINTER(C.c_float)), None))
        return dict(t=t.value, rank=rank.value, D=dim.value, rho=float(rho.value), d=d, W=W, num_reorth=nre.value)

    def __del__(self):
        try:
            if self.owned and self.h:
                _lib().tdnnf_nnet3_ng_delete(self.h)
                self.h = None
        except Exception:
            pass


def _idx_array(indexes: Sequence[Tuple[int, int, int]]):
    flat = []
    for n, t, x in indexes:
        flat += [int(n), int(t), int(x)]
    return (C.c_int32 * max(len(flat), 1))(*flat)


def _idx_list(ptr, n) -> List[Tuple[int, int, int]]:
    out = [(ptr[3 * i], ptr[3 * i + 1], ptr[3 * i + 2]) for i in range(n)]
    _lib().tdnnf_nnet3_free(ptr)
    return out


class PrecomputedIndexes:
    def __init__(self, handle):
        self.h = handle

    def write(self, binary: bool = False) -> bytes:
        out, n = C.c_void_p(), C.c_uint64()
        _check(_lib().tdnnf_nnet3_indexes_write(self.h, int(binary), C.byref(out), C.byref(n)))
        return _take_string(out, n.value)

    @staticmethod
    def read(data: bytes, binary: bool = False) -> "PrecomputedIndexes":
        h = vp()
        _check(_lib().tdnnf_nnet3_indexes_read(data, C.c_uint64(len(data)), int(binary), C.byref(h)))
        return PrecomputedIndexes(h)

    def row_stride_and_offsets(self) -> Tuple[int, List[int]]:
        toks = self.write(False).decode().split()
        rs = int(toks[toks.index("<RowStride>") + 1])
        i0 = toks.index("[") + 1
        return rs, [int(t) for t in toks[i0: toks.index("]")]]

    def __del__(self):
        try:
            if self.h:
                _lib().tdnnf_nnet3_indexes_delete(self.h)
                self.h = None
        except Exception:
            pass


class Component:
    """Owner of one C++ Component*."""

    def __init__(self, handle):
        self.h = handle

    # ---- construction
    @staticmethod
    def new(type_name: str, config: str) -> "Component":
        h = vp()
        _check(_lib().tdnnf_nnet3_component_new(type_name.encode(), config.encode(), C.byref(h)))
        return Component(h)

    @staticmethod
    def read(data: bytes, binary: bool = False) -> "Component":
        h = vp()
        _check(_lib().tdnnf_nnet3_component_read(data, C.c_uint64(len(data)), int(binary), C.byref(h)))
        return Component(h)

    @staticmethod
    def read_prefix(data: bytes, binary: bool = False) -> Tuple["Component", int]:
        """Component.read from the START of `data`; also returns how many bytes the component occupied."""
        h = vp()
        used = C.c_uint64()
        _check(_lib().tdnnf_nnet3_component_read_ex(data, C.c_uint64(len(data)), int(binary), C.byref(h), C.byref(used)))
        return Component(h), int(used.value)

    @staticmethod
    def tdnn_darts_for_indexing(time_offsets: Sequence[int]) -> "Component":
        h = vp()
        arr = (C.c_int32 * len(time_offsets))(*time_offsets)
        _check(_lib().tdnnf_nnet3_tdnn_darts_for_indexing(arr, len(time_offsets), C.byref(h)))
        return Component(h)

    def copy(self) -> "Component":
        h = vp()
        _check(_lib().tdnnf_nnet3_component_copy(self.h, C.byref(h)))
        return Component(h)

    def write(self, binary: bool = False) -> bytes:
        out, n = C.c_void_p(), C.c_uint64()
        _check(_lib().tdnnf_nnet3_component_write(self.h, int(binary), C.byref(out), C.byref(n)))
        return _take_string(out, n.value)

    def info(self) -> str:
        out = C.c_void_p()
        _check(_lib().tdnnf_nnet3_component_info(self.h, C.byref(out)))
        return _take_string(out).decode()

    def type(self) -> str:
        out = C.c_void_p()
        _check(_lib().tdnnf_nnet3_component_type(self.h, C.byref(out)))
        return _take_string(out).decode()

    def dims(self) -> Tuple[int, int, int]:
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        _check(_lib().tdnnf_nnet3_component_dims(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def input_dim(self) -> int:
        return self.dims()[0]

    def output_dim(self) -> int:
        return self.dims()[1]

    def properties(self) -> int:
        return self.dims()[2]

    # ---- indexes
    def precompute_indexes(self, input_indexes, output_indexes, need_backprop=True) -> Optional[PrecomputedIndexes]:
        h = vp()
        _check(_lib().tdnnf_nnet3_precompute_indexes(self.h, _idx_array(input_indexes), len(input_indexes),
                                                     _idx_array(output_indexes), len(output_indexes),
                                                     int(need_backprop), C.byref(h)))
        return PrecomputedIndexes(h) if h.value else None

    def reorder_indexes(self, input_indexes, output_indexes):
        pi, po = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)()
        ni, no = C.c_int(), C.c_int()
        _check(_lib().tdnnf_nnet3_reorder_indexes(self.h, _idx_array(input_indexes), len(input_indexes),
                                                  _idx_array(output_indexes), len(output_indexes), C.byref(pi),
                                                  C.byref(ni), C.byref(po), C.byref(no)))
        return _idx_list(pi, ni.value), _idx_list(po, no.value)

    def get_input_indexes(self, n, t, x=0):
        p, cnt = C.POINTER(C.c_int32)(), C.c_int()
        _check(_lib().tdnnf_nnet3_get_input_indexes(self.h, n, t, x, C.byref(p), C.byref(cnt)))
        return _idx_list(p, cnt.value)

    def is_computable(self, n, t, x, available) -> bool:
        r = C.c_int()
        _check(_lib().tdnnf_nnet3_is_computable(self.h, n, t, x, _idx_array(available), len(available), C.byref(r)))
        return bool(r.value)

    # ---- computation
    def propagate(self, indexes: Optional[PrecomputedIndexes], x, out):
        """Returns the memo handle (or None)."""
        xp, xr, xc, xs = capi._mat(x)
        op, orr, oc, os_ = capi._mat(out)
        memo = vp()
        _check(_lib().tdnnf_nnet3_propagate(self.h, indexes.h if indexes else None, vp(xp), xr, xc, xs, vp(op), orr, oc,
                                            os_, C.byref(memo)))
        return memo if memo.value else None

    def backprop(self, indexes, in_value, out_value, out_deriv, memo, to_update: Optional["Component"], in_deriv):
        dp, dr, dc, ds = capi._mat(out_deriv)
        ivp, ir, ic, ivs = (0, 0, 0, 0) if in_value is None else capi._mat(in_value)
        ovp, ovs = (0, 0) if out_value is None else (capi._mat(out_value)[0], capi._mat(out_value)[3])
        idp, ids = (0, 0)
        if in_deriv is not None:
            idp, ir2, ic2, ids = capi._mat(in_deriv)
            if in_value is None:
                ir, ic = ir2, ic2
        _check(_lib().tdnnf_nnet3_backprop(self.h, indexes.h if indexes else None, vp(ivp), ir, ic, ivs, vp(ovp), ovs,
                                           vp(dp), dr, dc, ds, memo, to_update.h if to_update else None, vp(idp), ids))

    def store_stats(self, in_value, out_value, memo):
        """Component::StoreStats (BatchNormComponent in training mode accumulates its minibatch statistics)."""
        ivp, ir, ic, ivs = (0, 0, 0, 0) if in_value is None else capi._mat(in_value)
        op, orr, oc, os_ = capi._mat(out_value)
        _check(_lib().tdnnf_nnet3_store_stats(self.h, vp(ivp), ir, ic, ivs, vp(op), orr, oc, os_, memo))

    def zero_stats(self):
        _check(_lib().tdnnf_nnet3_zero_stats(self.h))

    def bn_count(self) -> float:
        c = C.c_double()
        _check(_lib().tdnnf_nnet3_bn_count(self.h, C.byref(c)))
        return c.value

    def delete_memo(self, memo):
        if memo is not None:
            _check(_lib().tdnnf_nnet3_delete_memo(self.h, memo))

    # ---- updatable
    def scale(self, s: float):
        _check(_lib().tdnnf_nnet3_scale(self.h, C.c_float(s)))

    def add(self, alpha: float, other: "Component"):
        _check(_lib().tdnnf_nnet3_add(self.h, C.c_float(alpha), other.h))

    def dot_product(self, other: "Component") -> float:
        r = C.c_float()
        _check(_lib().tdnnf_nnet3_dot_product(self.h, other.h, C.byref(r)))
        return r.value

    def num_parameters(self) -> int:
        n = C.c_int()
        _check(_lib().tdnnf_nnet3_num_parameters(self.h, C.byref(n)))
        return n.value

    def vectorize(self):
        import numpy as np

        n = self.num_parameters()
        v = np.zeros(n, dtype=np.float32)
        _check(_lib().tdnnf_nnet3_vectorize(self.h, v.ctypes.data_as(C.POINTER(C.c_float)), n))
        return v

    def unvectorize(self, v):
        import numpy as np

        v = np.ascontiguousarray(v, dtype=np.float32)
        _check(_lib().tdnnf_nnet3_unvectorize(self.h, v.ctypes.data_as(C.POINTER(C.c_float)), len(v)))

    def perturb_params(self, stddev: float):
        _check(_lib().tdnnf_nnet3_perturb_params(self.h, C.c_float(stddev)))

    def set_learning_rate(self, lr: float):
        _check(_lib().tdnnf_nnet3_set_learning_rate(self.h, C.c_float(lr)))

    def set_actual_learning_rate(self, lr: float):
        _check(_lib().tdnnf_nnet3_set_actual_learning_rate(self.h, C.c_float(lr)))

    def learning_rate(self) -> float:
        r = C.c_float()
        _check(_lib().tdnnf_nnet3_get_learning_rate(self.h, C.byref(r)))
        return r.value

    def set_test_mode(self, flag: bool):
        _check(_lib().tdnnf_nnet3_set_test_mode(self.h, int(flag)))

    def temp_proportion(self) -> float:
        r = C.c_float()
        _check(_lib().tdnnf_nnet3_temp_proportion(self.h, C.byref(r)))
        return r.value

    def dropout_proportion(self) -> float:
        r = C.c_float()
        _check(_lib().tdnnf_nnet3_dropout_proportion(self.h, C.byref(r)))
        return r.value

    def orthonormal_constraint(self) -> float:
        r = C.c_float()
        _check(_lib().tdnnf_nnet3_orthonormal_constraint(self.h, C.byref(r)))
        return r.value

    def preconditioner(self, which: int = 0) -> "NaturalGradient":
        """preconditioner_in_ (0) / preconditioner_out_ (1) of a TdnnDARTSV3Component, or preconditioner_ (0)."""
        h = vp()
        _check(_lib().tdnnf_nnet3_component_ng(self.h, which, C.byref(h)))
        ng = NaturalGradient(_borrowed=h)
        ng._keepalive = self
        return ng

    def param_buffers(self):
        """[(device_ptr, rows, cols, stride), ...] of the parameter buffers (for the delta all-reduce)."""
        ptrs = (C.c_void_p * 2)()
        rows, cols, strides = (C.c_int * 2)(), (C.c_int * 2)(), (C.c_int * 2)()
        cnt = C.c_int()
        _check(_lib().tdnnf_nnet3_param_buffers(self.h, ptrs, rows, cols, strides, C.byref(cnt)))
        return [(ptrs[i], rows[i], cols[i], strides[i]) for i in range(cnt.value)]

    def bn_test_set_stats(self, dim, block_dim, epsilon, target_rms, count, stats_sum, stats_sumsq):
        import numpy as np

        a = np.ascontiguousarray(stats_sum, dtype=np.float64)
        b = np.ascontiguousarray(stats_sumsq, dtype=np.float64)
        _check(_lib().tdnnf_nnet3_bn_test_set_stats(self.h, dim, block_dim, C.c_float(epsilon), C.c_float(target_rms),
                                                    C.c_double(count), a.ctypes.data_as(C.POINTER(C.c_double)),
                                                    b.ctypes.data_as(C.POINTER(C.c_double))))

    def bn_test_scale_offset(self):
        """(scale_ptr, offset_ptr, dim): device pointers of a BatchNormTestComponent's derived vectors."""
        sp, op, dim = C.c_void_p(), C.c_void_p(), C.c_int()
        _check(_lib().tdnnf_nnet3_bn_test_scale_offset(self.h, C.byref(sp), C.byref(op), C.byref(dim)))
        return sp.value, op.value, dim.value

    def __del__(self):
        try:
            if self.h:
                _lib().tdnnf_nnet3_component_delete(self.h)
                self.h = None
        except Exception:
            pass


def constrain_orthonormal(components: Sequence[Component]) -> int:
    """ConstrainOrthonormal (utils.cc:1037-1077) over the components, in order: every TdnnComponent with a non-zero
    orthonormal-constraint is updated with probability 1/4.  Returns how many were."""
    n = len(components)
    comps = (C.c_void_p * max(n, 1))(*[c.h for c in components])
    k = C.c_int()
    _check(_lib().tdnnf_nnet3_constrain_orthonormal(comps, n, C.byref(k)))
    return k.value


def apply_edits(edits: str, named_components: Sequence[Tuple[str, Component]]):
    """ReadEditConfig (utils.cc:1166-1415 subset) over (name, component) pairs; directives ';' or newline separated."""
    n = len(named_components)
    names = (C.c_char_p * max(n, 1))(*[nm.encode() for nm, _ in named_components])
    comps = (C.c_void_p * max(n, 1))(*[c.h for _, c in named_components])
    _check(_lib().tdnnf_nnet3_apply_edits(edits.encode(), names, comps, n))


def temperature_for_iteration(num_archives_processed: int, num_archives_to_process: int) -> float:
    """The linear anneal of temperature_schedule.py:51: T = (1 - f) * (1 - 0.03) + 0.03."""
    f = float(num_archives_processed) / num_archives_to_process
    return (1.0 - f) * (1 - 0.03) + 0.03


def temperature_edit_string(num_archives_processed: int, num_archives_to_process: int) -> str:
    """get_temperature_edit_string (temperature_schedule.py:34-67): the per-iteration edit directive."""
    t = temperature_for_iteration(num_archives_processed, num_archives_to_process)
    return "set-temperature-proportion name=* proportion={0}".format(t)


def parse_dropout_schedule(schedule: str) -> List[Tuple[float, float]]:
    """--trainer.dropout-schedule for one name pattern, e.g. '0,0@0.20,0.5@0.50,0' -> [(data_fraction, proportion), ...]
    ascending: the first value holds at fraction 0, the last at 1, 'p@f' in between, a bare middle value means @0.5
    (temperature_schedule.py:119-172 keeps upstream dropout_schedule.py's parser as a comment)."""
    parts = schedule.strip().split(",")
    if len(parts) < 2:
        raise ValueError("dropout proportion string must specify at least the start and end dropouts")
    values = [(0.0, float(parts[0]))]
    for part in parts[1:-1]:
        pv = part.split("@")
        if len(pv) not in (1, 2):
            raise ValueError(f"bad dropout-schedule entry {part!r}")
        proportion, fraction = float(pv[0]), (float(pv[1]) if len(pv) == 2 else 0.5)
        if fraction < values[-1][0] or fraction > 1.0:
            raise ValueError("dropout-schedule must be in increasing order of data fractions")
        values.append((fraction, proportion))
    values.append((1.0, float(parts[-1])))
    for fraction, proportion in values:
        if not (0.0 <= fraction <= 1.0 and 0.0 <= proportion <= 1.0):
            raise ValueError("dropout-schedule values must lie in [0, 1]")
    return values


def dropout_proportion_for_fraction(schedule: str, data_fraction: float) -> float:
    """Piecewise-linear interpolation of the schedule at `data_fraction` of the training data (upstream
    _get_component_dropout)."""
    values = parse_dropout_schedule(schedule)
    if data_fraction <= 0.0:
        return values[0][1]
    if data_fraction >= 1.0:
        return values[-1][1]
    for (f0, p0), (f1, p1) in zip(values, values[1:]):
        if f0 <= data_fraction <= f1:
            if f1 == f0:
                return p1
            return p0 + (p1 - p0) * (data_fraction - f0) / (f1 - f0)
    return values[-1][1]


def dropout_edit_string(schedule: str, data_fraction: float, name_pattern: str = "*") -> str:
    """The per-iteration edit directive upstream's get_dropout_edit_string emits (utils.cc:1297-1330 consumes it)."""
    return "set-dropout-proportion name={0} proportion={1}".format(
        name_pattern, dropout_proportion_for_fraction(schedule, data_fraction))


# ---------------------------------------------------------------------------------------------------------------------
# The raw nnet3 model container (kaldi: nnet3/nnet-nnet.cc, Nnet::Write / Nnet::Read): "<Nnet3>", the config lines of the
# graph (kept as opaque text: the graph compiler stays in Kaldi), a blank line, "<NumComponents> N", then
# "<ComponentName> name" + the component for each, "</Nnet3>".  What `nnet3-copy --binary=false` writes, so the
# reference's scripts that grep the text model (generate_top_list.py:21-27 on '<BiasParams>' lines,
# bottleneckdim_search_top_model_size.py:15-18 on 'alpha <ConstantFunctionComponent>' lines) work on models written here.
def _bin_int32(v: int) -> bytes:
    import struct

    return b"\x04" + struct.pack("<i", v)  # WriteBasicType(binary): size byte, then the value


def write_nnet(config_lines: Sequence[str], named_components: Sequence[Tuple[str, Component]], binary: bool = False) -> bytes:
    out = [b"<Nnet3> \n"]
    for line in config_lines:
        if not line.strip():
            raise ValueError("config lines must not be empty (a blank line ends the config section)")
        out.append(line.encode() + b"\n")
    out.append(b"\n<NumComponents> ")
    out.append(_bin_int32(len(named_components)) if binary else str(len(named_components)).encode() + b" \n")
    for name, comp in named_components:
        if not name or any(ch.isspace() for ch in name):
            raise ValueError(f"bad component name {name!r}")
        out.append(b"<ComponentName> " + name.encode() + b" ")
        out.append(comp.write(binary))
        if not binary:
            out.append(b"\n")
    out.append(b"</Nnet3> ")
    return b"".join(out)


def read_nnet(data: bytes, binary: bool = False) -> Tuple[List[str], List[Tuple[str, Component]]]:
    def expect(pos: int, token: bytes) -> int:
        while not binary and data[pos:pos + 1].isspace():
            pos += 1
        if data[pos:pos + len(token)] != token:
            raise Nnet3Error(f"expected {token.decode()} at byte {pos}, got {data[pos:pos + 24]!r}")
        return pos + len(token)

    pos = expect(0, b"<Nnet3>")
    end = data.index(b"\n\n", pos)  # a blank line terminates the config section
    config_lines = [l.decode() for l in data[pos:end].split(b"\n") if l.strip()]
    pos = expect(end + 2, b"<NumComponents> ")
    if binary:
        import struct

        if data[pos:pos + 1] != b"\x04":
            raise Nnet3Error("bad binary <NumComponents>")
        n = struct.unpack("<i", data[pos + 1:pos + 5])[0]
        pos += 5
    else:
        stop = pos
        while not data[stop:stop + 1].isspace():
            stop += 1
        n = int(data[pos:stop])
        pos = stop
    comps = []
    for _ in range(n):
        pos = expect(pos, b"<ComponentName> ")
        stop = data.index(b" ", pos)
        name = data[pos:stop].decode()
        comp, used = Component.read_prefix(data[stop + 1:], binary)
        comps.append((name, comp))
        pos = stop + 1 + used
    expect(pos, b"</Nnet3>")
    return config_lines, comps
