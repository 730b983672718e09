"""ORACLE -- TEST INFRASTRUCTURE ONLY (tests/ and bench.py's CPU legs; the product never imports this).

One whole training step of the context-offset DARTS TDNN-F supernet in the SEARCH stage (BASELINE.json configs[2];
run_TDNN_DARTSV3_fbk_stride_cvupdate.sh) on the CPU, composed from the oracle's restatements of the reference methods in
the order nnet3-chain-train executes them:

  forward   tdnn1 (affine, ReLU, BatchNormTest) -> 14 x {TdnnDARTSV3 linear (-6..0), TdnnDARTSV3 affine (0..6), ReLU,
            BatchNormTest, Sum(Scale(0.66, prev), .)} -> prefinal-l -> {prefinal-chain -> output,
            prefinal-xent -> output-xent -> LogSoftmax}
  objective ComputeChainObjfAndDeriv (denominator forward-backward, generic numerator) + the xent derivative
  backward  data gradients through every layer; TdnnDARTSV3Component::Backprop with UpdateNaturalGradient
            (tdnn.cc:335-431, 457-626: in_value_temp, the extra product per offset for the alpha gradient, both
            OnlineNaturalGradient::PreconditionDirections calls); everything else is frozen (learning-rate-factor 0)
  update    UpdateNnetWithMaxChange (nnet-utils.cc:2085-2175)

Every AddMatMat goes through cblas_sgemm when oracle.enable_blas() found one (numpy's bundled OpenBLAS), as in a Kaldi CPU
build; element-wise work runs on all host threads (torch CPU tensors are used for that: memory and threads only).
Two uses: bench.py --impl reference / cpu_baseline times it at full size; tests/test_gpu_step_parity.py loads the GPU
supernet's parameters into it at a small size and compares objective, derivatives and deltas.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np

from . import oracle as O

FLAGS_SEARCH = O.USE_GUMBEL | O.UPDATE_ALPHA  # use-gumbel=true update-alpha=true, everything else false (cvupdate.sh:129-134)


@dataclass
class RefConfig:
    num_seqs: int = 64
    frames_per_eg: int = 150
    frame_subsampling: int = 3
    feat_dim: int = 220
    dim: int = 1536
    bottleneck: int = 160
    num_blocks: int = 14
    num_offsets: int = 7
    prefinal_small: int = 256
    num_pdfs: int = 6008
    leaky_hmm: float = 0.1
    bypass_scale: float = 0.66
    xent: bool = True
    xent_regularize: float = 0.1
    learning_rate: float = 2.5e-4
    darts_lr_factor: float = 1.0e-4
    temp_proportion: float = 1.0
    max_change: float = 0.75
    max_param_change: float = 2.0
    rank_in: int = 20
    rank_out: int = 80


def frame_plan(cfg: RefConfig):
    """Frames every layer computes for one chunk, derived from the output request backwards (what the nnet3 compiler
    does): output frames 0, 3, ..; the last block's affine reads 0..6 frames to the right, every linear 6 to the left."""
    n = cfg.num_offsets - 1
    T = cfg.frames_per_eg // cfg.frame_subsampling
    out_t = [cfg.frame_subsampling * i for i in range(T)]
    L = cfg.num_blocks
    aff_t: List[List[int]] = [None] * L
    lin_t: List[List[int]] = [None] * L
    aff_t[L - 1] = out_t
    lin_t[L - 1] = list(range(out_t[0], out_t[-1] + n + 1))
    for b in range(L - 2, -1, -1):
        aff_t[b] = list(range(lin_t[b + 1][0] - n, lin_t[b + 1][-1] + 1))
        lin_t[b] = list(range(aff_t[b][0], aff_t[b][-1] + n + 1))
    in_t = list(range(lin_t[0][0] - n, lin_t[0][-1] + 1))
    return T, out_t, lin_t, aff_t, in_t


def _t(a):
    import torch

    return torch.from_numpy(a)


class CpuSupernet:
    """params: dict of float32 numpy arrays
         tdnn1.W (D x feat), tdnn1.b; bn.<name>.scale / .offset for name in tdnn1, blk<b>, pc1, pc2[, px1, px2];
         blk<b>.lin.W (B x 7D), blk<b>.lin.bias (7 + B), blk<b>.aff.W (D x 7B), blk<b>.aff.bias (7 + D);
         prefinal_l.W, pc_affine.W/.b, pc_linear.W, output.W/.b [, px_affine.W/.b, px_linear.W, output_xent.W/.b]"""

    def __init__(self, cfg: RefConfig, den_graph: dict, num_graph: dict, params: Optional[Dict[str, np.ndarray]] = None, seed: int = 1):
        import torch

        self.cfg, self.den_graph, self.num_graph = cfg, den_graph, num_graph
        self.T, self.out_t, self.lin_t, self.aff_t, self.in_t = frame_plan(cfg)
        self.left = list(range(-(cfg.num_offsets - 1), 1))
        self.right = list(range(cfg.num_offsets))
        self.p = params if params is not None else self._random_params(seed)
        n = cfg.num_offsets
        self.ng = {(b, h): (O.NaturalGradient(cfg.rank_in, 4, 2000.0, 4.0), O.NaturalGradient(cfg.rank_out, 4, 2000.0, 4.0))
                   for b in range(cfg.num_blocks) for h in ("lin", "aff")}
        self.delta = {}
        for b in range(cfg.num_blocks):
            for h in ("lin", "aff"):
                self.delta[(b, h)] = (np.zeros_like(self.p[f"blk{b}.{h}.W"]), np.zeros_like(self.p[f"blk{b}.{h}.bias"]))
        torch.set_num_threads(max(1, O.num_threads()))
        assert n == len(self.left)

    # ------------------------------------------------------------------ parameters for the timing arm
    def _random_params(self, seed):
        cfg, g = self.cfg, np.random.default_rng(seed)
        D, B, n, P, Sm = cfg.dim, cfg.bottleneck, cfg.num_offsets, cfg.num_pdfs, cfg.prefinal_small
        rn = lambda r, c, s: (g.standard_normal((r, c)) * s).astype(np.float32)
        p = {"tdnn1.W": rn(D, cfg.feat_dim, cfg.feat_dim ** -0.5), "tdnn1.b": rn(1, D, 0.1)[0]}
        names = ["tdnn1"] + [f"blk{b}" for b in range(cfg.num_blocks)] + ["pc1", "pc2"] + (["px1", "px2"] if cfg.xent else [])
        for nm in names:
            d = Sm if nm in ("pc2", "px2") else D
            p[f"bn.{nm}.scale"] = (0.7 + 0.1 * g.random(d)).astype(np.float32)
            p[f"bn.{nm}.offset"] = (0.1 * g.standard_normal(d)).astype(np.float32)
        for b in range(cfg.num_blocks):
            p[f"blk{b}.lin.W"] = rn(B, n * D, (n * D) ** -0.5)
            p[f"blk{b}.lin.bias"] = np.concatenate([np.zeros(n, np.float32), rn(1, B, 1.0)[0]])
            p[f"blk{b}.aff.W"] = rn(D, n * B, (n * B) ** -0.5)
            p[f"blk{b}.aff.bias"] = np.concatenate([np.zeros(n, np.float32), rn(1, D, 1.0)[0]])
        p["prefinal_l.W"] = rn(Sm, D, D ** -0.5)
        for pre in ("pc",) + (("px",) if cfg.xent else ()):
            p[f"{pre}_affine.W"], p[f"{pre}_affine.b"] = rn(D, Sm, Sm ** -0.5), rn(1, D, 0.1)[0]
            p[f"{pre}_linear.W"] = rn(Sm, D, D ** -0.5)
        p["output.W"], p["output.b"] = rn(P, Sm, Sm ** -0.5), rn(1, P, 0.1)[0]
        if cfg.xent:
            p["output_xent.W"], p["output_xent.b"] = rn(P, Sm, Sm ** -0.5), rn(1, P, 0.1)[0]
        return p

    # ------------------------------------------------------------------ small helpers (torch CPU: threads only)
    @staticmethod
    def _affine(x, W, b=None):
        import torch

        y = torch.mm(_t(x), _t(W).t())
        if b is not None:
            y += _t(b)
        return y.numpy()

    @staticmethod
    def _relu_bn(x, scale, offset):
        import torch

        return (torch.relu(_t(x)) * _t(scale) + _t(offset)).numpy()

    def _blocked(self, x, frames):
        """t-major rows (frame f, sequence n) -> the blocked order ReorderIndexes asks for when the output is frame-
        subsampled by 3: row = (f // 3) * 3S + 3n + f % 3, the frame count rounded up to a multiple of 3."""
        S, r = self.cfg.num_seqs, self.cfg.frame_subsampling
        F, D = len(frames), x.shape[1]
        Fp = (F + r - 1) // r * r
        buf = np.zeros((Fp, S, D), np.float32)
        buf[:F] = x.reshape(F, S, D)
        return np.ascontiguousarray(buf.reshape(Fp // r, r, S, D).transpose(0, 2, 1, 3)).reshape(Fp * S, D), Fp

    def _unblocked(self, xb, F, Fp):
        S, r = self.cfg.num_seqs, self.cfg.frame_subsampling
        D = xb.shape[1]
        return np.ascontiguousarray(xb.reshape(Fp // r, S, r, D).transpose(0, 2, 1, 3)).reshape(Fp, S, D)[:F].reshape(F * S, D)

    # ------------------------------------------------------------------ the step
    def forward(self, x, u_gumbel):
        """x: rows_in x feat_dim (t-major, sequence fastest); u_gumbel: list of 2 * num_blocks arrays of n uniforms in the
        order the components draw them (lin0, aff0, lin1, ...).  Returns nnet_output (T*S x P)."""
        cfg, p, S = self.cfg, self.p, self.cfg.num_seqs
        n = cfg.num_offsets
        st = self.st = {}
        st["t1.aff"] = self._affine(x, p["tdnn1.W"], p["tdnn1.b"])
        prev = self._relu_bn(st["t1.aff"], p["bn.tdnn1.scale"], p["bn.tdnn1.offset"])
        prev_t = self.in_t
        st["in0"] = prev
        for b in range(cfg.num_blocks):
            lin_t, aff_t = self.lin_t[b], self.aff_t[b]
            ro = [(lin_t[0] + o - prev_t[0]) * S for o in self.left]
            lin_out, coef_l = O.tdnn_propagate(self.left, FLAGS_SEARCH, cfg.temp_proportion, p[f"blk{b}.lin.W"], p[f"blk{b}.lin.bias"],
                                               prev, len(lin_t) * S, ro, 1, u_gumbel[2 * b], 0.0)
            last = b == cfg.num_blocks - 1
            if last:  # output frames 0, 3, ..: blocked input order, row_stride 3
                a_in, Fp = self._blocked(lin_out, lin_t)
                r = cfg.frame_subsampling
                ro_a, rs_a = [(o // r) * r * S + o % r for o in self.right], r
            else:
                a_in, Fp, ro_a, rs_a = lin_out, len(lin_t), [o * S for o in self.right], 1
            aff_out, coef_a = O.tdnn_propagate(self.right, FLAGS_SEARCH, cfg.temp_proportion, p[f"blk{b}.aff.W"], p[f"blk{b}.aff.bias"],
                                               a_in, len(aff_t) * S, ro_a, rs_a, u_gumbel[2 * b + 1], 0.0)
            y = self._relu_bn(aff_out, p[f"bn.blk{b}.scale"], p[f"bn.blk{b}.offset"])
            # noop = Sum(Scale(0.66, prev), batchnorm): the rows of prev at this block's output frames
            idx = np.array([t - prev_t[0] for t in aff_t])
            byp = prev.reshape(len(prev_t), S, -1)[idx].reshape(len(aff_t) * S, -1)
            out = (_t(y) + cfg.bypass_scale * _t(byp)).numpy()
            st[b] = dict(prev=prev, prev_t=prev_t, ro=ro, lin_out=lin_out, coef_l=coef_l, a_in=a_in, Fp=Fp, ro_a=ro_a, rs_a=rs_a,
                         aff_out=aff_out, coef_a=coef_a, byp_idx=idx)
            prev, prev_t = out, aff_t
        st["last"] = prev
        st["pl"] = self._affine(prev, p["prefinal_l.W"])

        def branch(pre, bn1, bn2, out_name):
            a = self._affine(st["pl"], p[f"{pre}_affine.W"], p[f"{pre}_affine.b"])
            bnd = self._relu_bn(a, p[f"bn.{bn1}.scale"], p[f"bn.{bn1}.offset"])
            li = self._affine(bnd, p[f"{pre}_linear.W"])
            b2 = (_t(li) * _t(p[f"bn.{bn2}.scale"]) + _t(p[f"bn.{bn2}.offset"])).numpy()
            st[pre] = dict(a=a, bnd=bnd, b2=b2)
            return self._affine(b2, p[f"{out_name}.W"], p[f"{out_name}.b"])

        st["out"] = branch("pc", "pc1", "pc2", "output")
        if cfg.xent:
            import torch

            st["xls"] = torch.log_softmax(_t(branch("px", "px1", "px2", "output_xent")), dim=1).numpy()
        return st["out"]

    def objective(self):
        """ComputeChainObjfAndDeriv + NnetChainTrainer::ProcessOutputs on output-xent.  Returns objf per frame."""
        cfg, st = self.cfg, self.st
        S, T = cfg.num_seqs, self.T
        den_lp, den_d, den_ok = O.den_forward_backward(self.den_graph, st["out"], S, T, cfg.leaky_hmm, deriv_weight=-1.0)
        num_lp, num_d, num_ok = O.num_forward_backward(self.num_graph, st["out"], T, deriv_weight=1.0)
        objf = num_lp - den_lp
        if not (np.isfinite(objf) and den_ok and num_ok):
            st["d_out"] = np.zeros_like(st["out"])
            st["d_xls"] = np.zeros_like(st["out"])
            return -10.0
        st["d_out"] = (_t(den_d) + _t(num_d)).numpy()
        if cfg.xent:
            self.xent_objf = float((st["xls"].astype(np.float64) * num_d).sum()) / (S * T)
            st["d_xls"] = (cfg.xent_regularize * _t(num_d)).numpy()
        return objf / (S * T)

    def relu_inputs(self):
        """The pre-activations whose sign the backward pass depends on: {'pc', 'px', block index} -> array."""
        d = {"pc": self.st["pc"]["a"]}
        if self.cfg.xent:
            d["px"] = self.st["px"]["a"]
        for b in range(self.cfg.num_blocks):
            d[b] = self.st[b]["aff_out"]
        return d

    def backward(self, relu_masks=None):
        """relu_masks (optional, keys as relu_inputs()): boolean arrays to use as the ReLU derivative instead of this
        side's own (x > 0).  The derivative of ReLU is discontinuous at 0: a pre-activation within rounding of zero can
        come out on either side in two correct implementations, and ONE such element moves every derivative below it by
        ~1e-3 relative.  A parity check hands over the other side's masks after checking that they differ from this
        side's only at such elements."""
        import torch

        cfg, p, st, S = self.cfg, self.p, self.st, self.cfg.num_seqs
        lr = cfg.learning_rate * cfg.darts_lr_factor
        own = self.relu_inputs()
        mask = lambda key: _t(np.ascontiguousarray(relu_masks[key])) if relu_masks is not None and key in relu_masks else (_t(own[key]) > 0)

        def branch_bwd(pre, bn1, bn2, out_name, d_top):
            d_b2 = torch.mm(_t(d_top), _t(p[f"{out_name}.W"]))
            d_li = d_b2 * _t(p[f"bn.{bn2}.scale"])
            d_bnd = torch.mm(d_li, _t(p[f"{pre}_linear.W"]))
            d_a = d_bnd * _t(p[f"bn.{bn1}.scale"]) * mask(pre)
            return torch.mm(d_a, _t(p[f"{pre}_affine.W"]))

        d_pl = branch_bwd("pc", "pc1", "pc2", "output", st["d_out"])
        if cfg.xent:
            d = _t(st["d_xls"])
            d_xo = d - torch.exp(_t(st["xls"])) * d.sum(dim=1, keepdim=True)   # DiffLogSoftmaxPerRow
            d_pl = d_pl + branch_bwd("px", "px1", "px2", "output_xent", d_xo.numpy())
        d_out = torch.mm(d_pl, _t(p["prefinal_l.W"])).numpy()
        for b in range(cfg.num_blocks - 1, -1, -1):
            s = st[b]
            d_prev = np.zeros_like(s["prev"])
            # bypass + ReLU / BatchNormTest backward
            d_prev.reshape(len(s["prev_t"]), S, -1)[s["byp_idx"]] = (cfg.bypass_scale * _t(d_out)).numpy().reshape(len(s["byp_idx"]), S, -1)
            d_aff = (_t(d_out) * _t(p[f"bn.blk{b}.scale"]) * mask(b)).numpy()
            dW, db = self.delta[(b, "aff")]
            d_ain = np.zeros_like(s["a_in"])
            ng_in, ng_out = self.ng[(b, "aff")]
            O.tdnn_backprop(self.right, FLAGS_SEARCH, cfg.temp_proportion, p[f"blk{b}.aff.W"], s["a_in"], d_aff, s["coef_a"], s["ro_a"],
                            s["rs_a"], lr, in_deriv=d_ain, dW=dW, dbias=db, ng_in=ng_in, ng_out=ng_out)
            d_lin = self._unblocked(d_ain, len(self.lin_t[b]), s["Fp"]) if s["rs_a"] != 1 else d_ain
            dW, db = self.delta[(b, "lin")]
            ng_in, ng_out = self.ng[(b, "lin")]
            O.tdnn_backprop(self.left, FLAGS_SEARCH, cfg.temp_proportion, p[f"blk{b}.lin.W"], s["prev"], d_lin, s["coef_l"], s["ro"], 1, lr,
                            in_deriv=d_prev if b > 0 else None, dW=dW, dbias=db, ng_in=ng_in, ng_out=ng_out)
            st[b]["d_aff"], st[b]["d_lin"] = d_aff, d_lin
            d_out = d_prev

    def update(self):
        """UpdateNnetWithMaxChange (nnet-utils.cc:2085-2175) + ScaleNnet(0): BaseFloat arithmetic."""
        cfg, f32 = self.cfg, np.float32
        keys = [(b, h) for b in range(cfg.num_blocks) for h in ("lin", "aff")]
        dots = [f32(float((self.delta[k][0].astype(np.float64) ** 2).sum() + (self.delta[k][1].astype(np.float64) ** 2).sum())) for k in keys]
        factors, pds = [], f32(0)
        for d in dots:
            f = f32(1)
            if cfg.max_change != 0 and np.sqrt(d) > f32(cfg.max_change):
                f = f32(cfg.max_change) / np.sqrt(d)
            factors.append(f)
            pds += f * f * d
        pd, scale = np.sqrt(pds), f32(1)
        applied = bool(np.isfinite(pd))
        if applied and cfg.max_param_change != 0 and pd > f32(cfg.max_param_change):
            scale = f32(cfg.max_param_change) / pd
        for k, f in zip(keys, factors):
            dW, db = self.delta[k]
            if applied:
                self.p[f"blk{k[0]}.{k[1]}.W"] += (f * scale) * dW
                self.p[f"blk{k[0]}.{k[1]}.bias"] += (f * scale) * db
            dW[:] = 0
            db[:] = 0
        self.last_factors = [float(f * scale) for f in factors]
        return applied

    def step(self, x, u_gumbel, apply_update=True, relu_masks=None):
        """relu_masks: None, or a callable(self) -> dict evaluated after the forward pass (see backward)."""
        self.forward(x, u_gumbel)
        objf = self.objective()
        self.backward(relu_masks(self) if callable(relu_masks) else relu_masks)
        if apply_update:
            self.update()
        return objf
