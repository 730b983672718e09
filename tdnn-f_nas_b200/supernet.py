"""The context-offset DARTS TDNN-F supernet of run_TDNN_DARTSV3_fbk_stride_{pretrain,cvupdate}.sh
(BASELINE.json configs[2]) as one training step on this library: forward, LF-MMI denominator
forward-backward, backward, data-parallel reduction of the deltas and the parameter step.

What runs where:
  * the 2 x 14 TdnnDARTSV3Component instances (with their OnlineNaturalGradient update), BatchNormTestComponent
    (search mode) / BatchNormComponent (pretrain mode) -> the nnet3 component mirror (csrc/nnet3) exactly as
    NnetComputer would call them (Propagate / StoreStats / Backprop / DeleteMemo);
  * ReLU / bypass sum / the stock affine layers around them (tdnn1, prefinal, output: "N4" neighbours) -> the same
    kernels through the C ABI (a stock TdnnComponent is the DARTS GEMM with one offset and weight 1; their update
    is plain SGD);
  * ComputeChainObjfAndDeriv (chain.py) -> den kernels + the generic (per-sequence FST) numerator kernel;
    UpdateNnetWithMaxChange is applied (host logic, one read-back); L2 / orthonormal constraint are not (N2);
    dropout-proportion is 0.0 as in the recipe; no xent branch.
torch is device memory, the H2D copy, streams and torch.distributed; every kernel launched is ours.

The step is compiled once into flat lists of pre-bound C calls so the per-step Python cost is a loop.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from functools import partial
from typing import List, Optional

import numpy as np

from . import capi, chain, nnet3, parallel, synth


@dataclass
class SupernetConfig:
    num_seqs: int = 64              # chunks per minibatch per GPU (--trainer.num-chunk-per-minibatch 64)
    frames_per_eg: int = 150
    frame_subsampling: int = 3
    feat_dim: int = 220             # Append(-1,0,1) of 40-dim fbank + 100-dim i-vector, after the fixed LDA
    dim: int = 1536
    bottleneck: int = 160
    num_blocks: int = 14            # tdnnf2 .. tdnnf15
    num_offsets: int = 7            # offset_len: left context -6..0, right context 0..6
    prefinal_small: int = 256
    num_pdfs: int = 6008
    den_states: int = 16384
    den_out_degree: float = 16.0
    leaky_hmm: float = 0.1
    bypass_scale: float = 0.66
    mode: str = "search"            # "search": use-gumbel, update-alpha, BatchNormTest; "pretrain": uniform-sample;
                                    # "manual": the TDNN-F 7q system of run_tdnn_7q_fbk_40_manual.sh (BASELINE configs[1])
    strides: Optional[List[int]] = None  # manual mode: time-stride per tdnnf-layer (default 1,1,1,0 then 6: :138-151)
    l2_regularize: float = 0.0      # per-component l2-regularize (0.01 in the manual recipe; ApplyL2Regularization)
    batchnorm_stats_scale: float = 0.8   # ScaleBatchnormStats after every minibatch (train-mode batch-norm only)
    xent: bool = False              # the cross-entropy regularisation branch of the chain recipes (prefinal-xent,
                                    # output-xent + LogSoftmaxComponent; --chain.xent-regularize 0.1)
    dropout: bool = False           # build the GeneralDropoutComponent of tdnn1 and of every block (in the search stage
                                    # this switches the fused ReLU + BatchNormTest + bypass tail off);
                                    # proportion 0 until set_dropout_proportion() (the recipes' schedule starts at 0)
    learning_rate: float = 2.5e-4
    darts_lr_factor: float = 1.0e-4  # <LearningRateFactor> set by the cvupdate recipe's sed (search mode only)
    max_change: float = 0.75         # per-component max-change (xconfig default of the tdnnf layers)
    max_param_change: float = 2.0    # --trainer.max-param-change (global)
    fuse_tail: bool = True           # search mode: ReLU + BatchNormTest + bypass as one pass (else three components)
    # "bottleneck" mode (BASELINE configs[3]; generate_bottleneckCB8share_onehottrain_config.py + add_flopsconstraint.py and
    # run_TDNNf_DARTS_mod_fbk_bottleneckCBshare_cvupdate_flopsconstraint.sh:136-139): every tdnnf-layer is TdnnComponent
    # 1536 -> 240 / 240 -> 1536 (time-strides 1,1,1,0,3 x 10), its bottleneck masked by the shared candidates
    # (25,25,30,20,20,40,40,40) weighted by {Gumbel}SoftmaxFlopsComponent(alpha); all pre-trained components are frozen
    # (learning-rate-factor 0, BatchNormTest), only the 14 alpha vectors (ConstantFunctionComponent) are trained.
    candidate_widths: tuple = (25, 25, 30, 20, 20, 40, 40, 40)
    flops_coef: float = 1.0e-3       # eta: the `scale` of {Gumbel}SoftmaxFlopsComponent (0, 1e-3, 1e-1 in the recipes)
    bottleneck_gumbel: bool = False  # GumbelSoftmaxFlopsComponent (temperature schedule) instead of SoftmaxFlopsComponent
    fuse_mask: bool = True           # the descriptor sub-graph + 8 CopyN + 8 ElementwiseProduct as one pass (else one by one)
    tail_planes: bool = True         # the fused tails also write the bf16 operand planes (+ row / column sums) of what they
                                     # produce: the next GEMM does not split the matrix again
    # learning-rate-factor 0 on everything but the searched parameters, as the search recipes set it
    # (run_TDNN_DARTSV3_fbk_stride_cvupdate.sh:129-134): nnet3 then computes no model derivatives for tdnn1 / prefinal /
    # output.  None = True in the search and bottleneck stages, False otherwise.
    freeze_stock: Optional[bool] = None
    seed: int = 20221 + 3


def block_offsets(cfg: "SupernetConfig"):
    """(left, right): the time offsets of the `linear` and `affine` half of every block.  DARTS stages: all
    `num_offsets` candidates (-6..0 / 0..6); manual: tdnnf-layer time-stride s gives (-s, 0) / (0, s), and the single
    offset 0 for s = 0 (composite_layers.py:145-150; default strides 1,1,1,0 then 6: run_tdnn_7q_fbk_40_manual.sh:138-151)."""
    n = cfg.num_offsets
    if cfg.mode in ("manual", "bottleneck"):
        wide = 3 if cfg.mode == "bottleneck" else 6  # generate_bottleneckCB8share_onehottrain_config.py:43-55: -3,0 / 0,3
        strides = list(cfg.strides) if cfg.strides is not None else ([1, 1, 1, 0] + [wide] * cfg.num_blocks)[:cfg.num_blocks]
        assert len(strides) == cfg.num_blocks
        return [[-s_, 0] if s_ else [0] for s_ in strides], [[0, s_] if s_ else [0] for s_ in strides]
    return [list(range(-(n - 1), 1))] * cfg.num_blocks, [list(range(n))] * cfg.num_blocks


def frame_plan(cfg: "SupernetConfig", left_offsets, right_offsets):
    """Which frames every layer computes for one chunk (what the nnet3 compiler derives from the output request):
    (T, output frames, per block the frames of the linear output, per block the frames of the affine output, the frames
    of tdnn1).  Pure host arithmetic (no device)."""
    T = cfg.frames_per_eg // cfg.frame_subsampling
    sub = cfg.frame_subsampling
    out_t = [sub * i for i in range(T)]
    L = cfg.num_blocks
    lw = [-min(o) for o in left_offsets]   # left context of each block's `linear` half
    rw = [max(o) for o in right_offsets]   # right context of its `affine` half
    aff_t: List[List[int]] = [None] * L
    lin_t: List[List[int]] = [None] * L
    aff_t[L - 1] = out_t
    lin_t[L - 1] = list(range(out_t[0], out_t[-1] + rw[L - 1] + 1))
    for b in range(L - 2, -1, -1):
        lo, hi = lin_t[b + 1][0] - lw[b + 1], lin_t[b + 1][-1]
        aff_t[b] = list(range(lo, hi + 1))
        lin_t[b] = list(range(lo, hi + rw[b] + 1))
    in_t = list(range(lin_t[0][0] - lw[0], lin_t[0][-1] + 1))  # tdnn1 output frames
    return T, out_t, lin_t, aff_t, in_t


def algorithmic_flops(cfg: "SupernetConfig") -> float:
    """fwd + dgrad + wgrad FLOPs of the block GEMMs per step and GPU without building the net: every
    candidate offset in the search stage, the shared + the sampled one in the pretrain stage, the layer's own in manual."""
    left, right = block_offsets(cfg)
    _, _, lin_t, aff_t, _ = frame_plan(cfg, left, right)
    tot = 0.0
    for b in range(cfg.num_blocks):
        if cfg.mode in ("manual", "bottleneck"):
            n_lin, n_aff = len(left[b]), len(right[b])
        else:
            n_lin = n_aff = cfg.num_offsets if cfg.mode == "search" else 2
        passes = 2 if cfg.mode == "bottleneck" else 3  # frozen layers: forward + data gradient, no parameter gradient
        tot += passes * 2.0 * len(lin_t[b]) * cfg.num_seqs * n_lin * cfg.dim * cfg.bottleneck
        tot += passes * 2.0 * len(aff_t[b]) * cfg.num_seqs * n_aff * cfg.bottleneck * cfg.dim
    return tot


class _Plan:
    """A flat list of zero-argument C calls; run() checks every status."""

    def __init__(self):
        self.calls = []

    def add(self, kind: str, fn, *args):
        self.calls.append((kind, partial(fn, *args)))

    def add_py(self, fn):
        """A Python callable (returns None) interleaved with the C calls, e.g. launching an async all-reduce."""
        def call():
            fn()
            return 0
        self.calls.append(("py", call))

    def run(self):
        for kind, call in self.calls:
            rc = call()
            if rc != 0:
                lib = capi.load()
                msg = lib.tdnnf_nnet3_last_error() if kind == "nnet3" else lib.tdnnf_last_error()
                raise RuntimeError(f"supernet step failed in a {kind} call: {msg.decode()}")


def _m(t):
    p, r, c, s = capi._mat(t)
    return C.c_void_p(p), r, c, s


class Supernet:
    def __init__(self, cfg: SupernetConfig, device: int = 0, rank: int = 0, world_size: int = 1,
                 process_group=None, dp_buckets: int = 1, standalone: bool = False):
        """dp_buckets: 1 = one in-place all-reduce of the whole delta arena after the backward pass; k > 1 = the arena is
        cut into k buckets in backward order, each reduced on a side stream as soon as the backward pass has produced it
        (tdnnf_dp_allreduce_bucket_async), the parameter step waits for all of them."""
        import torch

        self.cfg, self.rank, self.world, self.pg = cfg, rank, world_size, process_group
        self.dp_buckets = max(1, int(dp_buckets))
        self._reduce_now = False
        self.standalone = standalone  # a replica of rank `rank` of a `world_size` job built without a communicator (self-check)
        self.dev = torch.device("cuda", device)
        torch.cuda.set_device(self.dev)
        self.ctx = capi.Context(device)
        self.ctx.use_current_stream()
        nnet3.set_context(self.ctx)
        nnet3.set_rand_seed(cfg.seed)          # identical on every rank: same Gumbel noise / one-hot choices
        nnet3.set_dp_world_size(world_size)
        self.lib = capi.load()
        self._build()

    # ------------------------------------------------------------------ construction
    def _frames(self):
        return frame_plan(self.cfg, self.left_offsets, self.right_offsets)

    def _build(self):
        import torch

        cfg, dev, ctx = self.cfg, self.dev, self.ctx
        S, D, B, n = cfg.num_seqs, cfg.dim, cfg.bottleneck, cfg.num_offsets
        manual = cfg.mode in ("manual", "bottleneck")   # stock TdnnComponent layers
        bneck = cfg.mode == "bottleneck"
        if bneck:
            assert sum(cfg.candidate_widths) == cfg.bottleneck, "candidate widths must add up to the bottleneck dimension"
        self.frozen = cfg.freeze_stock if cfg.freeze_stock is not None else cfg.mode in ("search", "bottleneck")
        self.left_offsets, self.right_offsets = block_offsets(cfg)
        T, out_t, lin_t, aff_t, in_t = self._frames()
        self.T, self.in_frames = T, len(in_t)
        g = synth.rng(3, stream=1)
        zeros = lambda r, c: torch.zeros((r, c), device=dev, dtype=torch.float32)
        randn = lambda r, c, s: torch.from_numpy((g.standard_normal((r, c)) * s).astype(np.float32)).to(dev)
        self.keep = []  # tensors / ctypes arrays that must outlive the plans

        search = cfg.mode == "search"
        self.frozen_bn = cfg.mode in ("search", "bottleneck")
        flags_cfg = ("use-gumbel=true use-entropy=false free-select=false update-alpha=true update-theta=false uniform-sample=false"
                     if search else
                     "use-gumbel=false use-entropy=false free-select=false update-alpha=false update-theta=true uniform-sample=true")
        lrf = cfg.darts_lr_factor if search else 1.0

        def grid(ts):
            return [(s, t, 0) for t in ts for s in range(S)]

        # ---------------- stock affine layers (weights as plain device matrices)
        def affine_params(din, dout, bias=True, zero=False):
            W = zeros(dout, din) if zero else randn(dout, din, 1.0 / np.sqrt(din))
            b = (torch.zeros(dout, device=dev) if zero else randn(1, dout, 0.1)[0]) if bias else None
            return dict(W=W, b=b, dW=None, db=None)  # the deltas live in the delta arena (_alloc_deltas)

        self.one = torch.ones(1, device=dev)
        self.zero_off = (C.c_int32 * 1)(0)
        self.stock = dict(tdnn1=affine_params(cfg.feat_dim, D), prefinal_l=affine_params(D, cfg.prefinal_small, False),
                          pc_affine=affine_params(cfg.prefinal_small, D), pc_linear=affine_params(D, cfg.prefinal_small, False),
                          # output-layer: param-stddev=0 at the start of training; a frozen (pre-trained) one is not zero
                          output=affine_params(cfg.prefinal_small, cfg.num_pdfs, zero=not self.frozen))
        if cfg.xent:  # prefinal-layer name=prefinal-xent input=prefinal-l; output-layer name=output-xent (log-softmax)
            self.stock.update(px_affine=affine_params(cfg.prefinal_small, D), px_linear=affine_params(D, cfg.prefinal_small, False),
                              output_xent=affine_params(cfg.prefinal_small, cfg.num_pdfs, zero=not self.frozen))

        # ---------------- frozen batch-norm (BatchNormTestComponent) or train-mode batch-norm
        def make_bn(dim):
            # BatchNormComponent in training mode (the supernet pretrain stage).  In search mode these are replaced,
            # after one calibration forward pass with StoreStats, by BatchNormTestComponents read from the `sed`-ed
            # model text -- "pretrain the supernet, then sed BatchNormComponent -> BatchNormTestComponent".
            return dict(comp=nnet3.Component.new("BatchNormComponent", f"dim={dim}"), memo=C.c_void_p())

        # ---------------- DARTS blocks
        self.blocks = []
        for b in range(cfg.num_blocks):
            left = ",".join(str(i) for i in self.left_offsets[b])
            right = ",".join(str(i) for i in self.right_offsets[b])
            if manual:
                # the config lines XconfigTdnnfLayer emits (composite_layers.py:156-169)
                common = (f"learning-rate={0.0 if bneck else cfg.learning_rate} l2-regularize={cfg.l2_regularize} max-change={cfg.max_change}")
                lin = nnet3.Component.new("TdnnComponent", f"input-dim={D} output-dim={B} {common} use-bias=false "
                                                           f"time-offsets={left} orthonormal-constraint=-1.0")
                aff = nnet3.Component.new("TdnnComponent", f"input-dim={B} output-dim={D} {common} time-offsets={right}")
            else:
                common = f"learning-rate={cfg.learning_rate * lrf} learning-rate-factor={lrf} {flags_cfg} use-bias=true"
                lin = nnet3.Component.new("TdnnDARTSV3Component", f"input-dim={D} output-dim={B} time-offsets={left} {common}")
                aff = nnet3.Component.new("TdnnDARTSV3Component", f"input-dim={B} output-dim={D} time-offsets={right} {common}")
            prev_t = in_t if b == 0 else aff_t[b - 1]
            # linear: input = previous block output (t-major), output = lin_t[b]
            li_in, li_out = lin.reorder_indexes(grid(prev_t), grid(lin_t[b]))
            assert li_in == grid(prev_t) and li_out == grid(lin_t[b])
            lin_idx = lin.precompute_indexes(li_in, li_out)
            # affine: ReorderIndexes may ask for the blocked (reorder_t_in = subsampling) input order
            ai_in, ai_out = aff.reorder_indexes(grid(lin_t[b]), grid(aff_t[b]))
            assert ai_out == grid(aff_t[b])
            aff_idx = aff.precompute_indexes(ai_in, ai_out)
            reorder = None
            if ai_in != grid(lin_t[b]):
                pos = {ix: r for r, ix in enumerate(grid(lin_t[b]))}
                fwd_map = np.array([pos.get(ix, -1) for ix in ai_in], dtype=np.int32)      # blocked row <- t-major row
                inv_map = np.full(len(pos), -1, dtype=np.int32)
                for r, m in enumerate(fwd_map):
                    if m >= 0:
                        inv_map[m] = r
                reorder = dict(fwd=torch.from_numpy(fwd_map).to(dev), inv=torch.from_numpy(inv_map).to(dev),
                               rows=len(ai_in))
            # bypass: rows of the previous block output that line up with this block's output frames
            if b == 0:
                bypass = None  # tdnnf2's bypass input is tdnn1 (dim matches): Sum(Scale(0.66, tdnn1), .)
            prev_pos = {ix: r for r, ix in enumerate(grid(prev_t))}
            by_rows = np.array([prev_pos[ix] for ix in grid(aff_t[b])], dtype=np.int32)
            contiguous = bool(np.all(np.diff(by_rows) == 1))
            bypass = dict(offset=int(by_rows[0]), contiguous=contiguous,
                          map=None if contiguous else torch.from_numpy(by_rows).to(dev))
            rows_lin, rows_aff, rows_prev = len(lin_t[b]) * S, len(aff_t[b]) * S, len(prev_t) * S
            blk = dict(lin=lin, aff=aff, lin_idx=lin_idx, aff_idx=aff_idx, reorder=reorder, bypass=bypass,
                       lin_delta=None, aff_delta=None, bn=make_bn(D),
                       lin_out=zeros(rows_lin, B), aff_in=zeros(reorder["rows"], B) if reorder else None,
                       aff_out=zeros(rows_aff, D), relu=zeros(rows_aff, D), bn_out=zeros(rows_aff, D), out=zeros(rows_aff, D),
                       d_out=zeros(rows_aff, D), d_aff=zeros(rows_aff, D), d_aff_in=zeros(reorder["rows"], B) if reorder else None,
                       d_lin=zeros(rows_lin, B), byp_tmp=None if contiguous else zeros(rows_aff, D),
                       memo_lin=C.c_void_p(), memo_aff=C.c_void_p(), rows_prev=rows_prev)
            if cfg.dropout:
                blk.update(self._make_dropout(D, grid(aff_t[b]), rows_aff, zeros))
            if bneck:
                # tdnnfK.alpha / tdnnfK.softmax (add_flopsconstraint.py:15-30) on the rows of the `linear` output, the CopyN /
                # ElementwiseProduct components of generate_bottleneckCB8share_onehottrain_config.py:22-75
                soft = (f"GumbelSoftmaxFlopsComponent dim=8 scale={cfg.flops_coef} temp-proportion=1.0" if cfg.bottleneck_gumbel
                        else f"SoftmaxFlopsComponent dim=8 scale={cfg.flops_coef}").split(" ", 1)
                nb = len(cfg.candidate_widths)
                blk.update(alpha=nnet3.Component.new("ConstantFunctionComponent",
                                                     f"input-dim={cfg.feat_dim} output-dim={nb} is-updatable=true use-natural-gradient=false "
                                                     f"learning-rate={cfg.learning_rate}"),
                           soft=nnet3.Component.new(soft[0], soft[1].replace("dim=8", f"dim={nb}")),
                           a_rows=zeros(rows_lin, nb), p=zeros(rows_lin, nb), d_p=zeros(rows_lin, nb), d_a=zeros(rows_lin, nb),
                           masked=zeros(rows_lin, B), d_masked=zeros(rows_lin, B), alpha_delta=None,
                           widths=(C.c_int32 * nb)(*cfg.candidate_widths))
                if not cfg.fuse_mask:
                    blk.update(copyn=[nnet3.Component.new("CopyNComponent", f"input-dim=1 output-dim={w}") for w in cfg.candidate_widths],
                               prod=[nnet3.Component.new("ElementwiseProductComponent", f"input-dim={2 * w} output-dim={w}")
                                     for w in cfg.candidate_widths],
                               m_in=[zeros(rows_lin, 1) for _ in cfg.candidate_widths],
                               cn=[zeros(rows_lin, w) for w in cfg.candidate_widths],
                               pin=[zeros(rows_lin, 2 * w) for w in cfg.candidate_widths],
                               d_pin=[zeros(rows_lin, 2 * w) for w in cfg.candidate_widths])
            self.blocks.append(blk)

        rows_in, rows_T = len(in_t) * S, T * S
        self.rows_in = rows_in
        self.x = zeros(rows_in, cfg.feat_dim)
        self.x_host = torch.zeros((rows_in, cfg.feat_dim), dtype=torch.float32).pin_memory()
        self.t1 = dict(aff=zeros(rows_in, D), relu=zeros(rows_in, D), out=zeros(rows_in, D), bn=make_bn(D),
                       d_out=zeros(rows_in, D), d_aff=zeros(rows_in, D))
        if cfg.dropout:
            self.t1.update(self._make_dropout(D, grid(in_t), rows_in, zeros))
        P, Ssm = cfg.num_pdfs, cfg.prefinal_small
        self.head = dict(pl=zeros(rows_T, Ssm), pa=zeros(rows_T, D), pr=zeros(rows_T, D), pb=zeros(rows_T, D),
                         pli=zeros(rows_T, Ssm), pb2=zeros(rows_T, Ssm), out=zeros(rows_T, P), bn1=make_bn(D), bn2=make_bn(Ssm),
                         d_out=zeros(rows_T, P), d_pb2=zeros(rows_T, Ssm), d_pli=zeros(rows_T, Ssm), d_pb=zeros(rows_T, D),
                         d_pa=zeros(rows_T, D), d_pl=zeros(rows_T, Ssm))
        if cfg.xent:
            self.head.update(xa=zeros(rows_T, D), xr=zeros(rows_T, D), xb=zeros(rows_T, D), xli=zeros(rows_T, Ssm),
                             xb2=zeros(rows_T, Ssm), xo=zeros(rows_T, P), xls=zeros(rows_T, P), xbn1=make_bn(D), xbn2=make_bn(Ssm),
                             d_xls=zeros(rows_T, P), d_xb2=zeros(rows_T, Ssm), d_xb=zeros(rows_T, D), d_xa=zeros(rows_T, D),
                             log_softmax=nnet3.Component.new("LogSoftmaxComponent", f"dim={P}"))
        # ---------------- denominator graph + synthetic numerator alignment
        graph = synth.make_den_graph(cfg.den_states, P, cfg.den_out_degree, seed=5)
        self.den_arcs = graph["num_arcs"]
        self.den_graph_host = graph
        self.den_graph = capi.DenGraph(ctx, graph)
        # per-sequence numerator FSTs of this rank's shard (unconstrained supervision, `--constrained false`): phone
        # strings are random walks in the denominator graph, so numerator paths are denominator paths (bounded objective)
        self.num_graph = capi.NumeratorGraph(ctx, synth.make_num_graphs(S, P, T, seed=60 + self.rank, den_graph=graph))
        self.objective = chain.ChainObjective(ctx, self.den_graph, self.num_graph, S, T,
                                              chain.ChainTrainingOptions(leaky_hmm_coefficient=cfg.leaky_hmm))
        self._alloc_deltas()
        self._compile()
        if self.frozen_bn:
            self._freeze_batchnorm()
            self._compile()

    def _backward_order(self):
        """Updatable components in the order the backward pass finishes them (= their order in the delta arena)."""
        head = ["output", "pc_linear", "pc_affine"] + (["output_xent", "px_linear", "px_affine"] if self.cfg.xent else []) + ["prefinal_l"]
        order = [] if self.frozen else [("stock", nm) for nm in head]
        for b in range(len(self.blocks) - 1, -1, -1):
            order += [("comp", (b, "alpha"))] if self.cfg.mode == "bottleneck" else [("comp", (b, "aff")), ("comp", (b, "lin"))]
        return order + ([] if self.frozen else [("stock", "tdnn1")])

    def _alloc_deltas(self):
        """delta_nnet_ as ONE device range (nnet3.arena): the delta copy of every updatable component, in backward order,
        so that the data-parallel reduction is a single in-place ncclAllReduce (or a few contiguous buckets issued while
        the backward pass is still running) instead of a gather / all-reduce / scatter."""
        import torch

        pad = lambda nbytes: (nbytes + 255) // 256 * 256
        pitch = lambda cols: (cols + 63) // 64 * 64
        total = 0
        for kind, key in self._backward_order():
            if kind == "stock":
                p = self.stock[key]
                total += pad(p["W"].numel() * 4) + (pad(p["b"].numel() * 4) if p["b"] is not None else 0)
            else:
                for _, r, c, _ in self.blocks[key[0]][key[1]].param_buffers():
                    total += pad(r * (pitch(c) if r > 1 else c) * 4)
        total += 4096
        self.delta_arena = torch.zeros(total // 4, device=self.dev, dtype=torch.float32)
        base = self.delta_arena.data_ptr()
        assert base % 256 == 0
        off = 0
        self.delta_spans = []  # (kind, key, begin_float, end_float) in backward order
        for kind, key in self._backward_order():
            begin = off
            if kind == "stock":
                p = self.stock[key]
                n = p["W"].numel()
                p["dW"] = self.delta_arena[off // 4: off // 4 + n].view(p["W"].shape)
                off += pad(n * 4)
                if p["b"] is not None:
                    nb = p["b"].numel()
                    p["db"] = self.delta_arena[off // 4: off // 4 + nb]
                    off += pad(nb * 4)
            else:
                blk = self.blocks[key[0]]
                with nnet3.arena(base + off, total - off) as a:
                    delta = blk[key[1]].copy()
                delta.scale(0.0)
                for ptr, _, _, _ in delta.param_buffers():
                    assert base + off <= ptr < base + off + a.used, "delta parameters were not carved from the arena"
                blk[key[1] + "_delta"] = delta
                off += pad(a.used)
            self.delta_spans.append((kind, key, begin // 4, off // 4))
        self.delta_floats = off // 4
        # buckets for the overlapped reduction: contiguous runs of spans of about equal size
        k = min(self.dp_buckets, len(self.delta_spans))
        self.delta_buckets = []   # (last (kind, key) of the bucket, begin_float, end_float)
        target, start = self.delta_floats / k, 0
        for i, (kind, key, b, e) in enumerate(self.delta_spans):
            last = i == len(self.delta_spans) - 1
            if last or (len(self.delta_buckets) < k - 1 and e >= target * (len(self.delta_buckets) + 1)):
                self.delta_buckets.append(((kind, key), start, e))
                start = e
        self.dp = None
        if self.world > 1 and not self.standalone:
            import torch.distributed as dist

            def exchange(ident: bytes) -> bytes:  # rank 0's NCCL id to everybody: plumbing over torch.distributed
                t = torch.tensor(list(ident), dtype=torch.uint8, device=self.dev if dist.get_backend(self.pg) == "nccl" else "cpu")
                dist.broadcast(t, src=0, group=self.pg)
                return bytes(t.cpu().tolist())

            self.dp = capi.DataParallel(self.ctx, self.world, self.rank, exchange)

    def _all_bn(self):
        return ([self.t1["bn"]] + [blk["bn"] for blk in self.blocks] +
                [self.head[k] for k in ("bn1", "bn2", "xbn1", "xbn2") if k in self.head])

    def _make_dropout(self, dim, indexes, rows, zeros):
        """`component name=X.dropout type=GeneralDropoutComponent dim=D dropout-proportion=0.0 continuous=true`
        (composite_layers.py: tdnnf-layer / relu-batchnorm-dropout-layer with dropout-per-dim-continuous=true): one mask
        row per sequence, shared by all its frames.  Out of place: the batch-norm backward needs its own output."""
        comp = nnet3.Component.new("GeneralDropoutComponent", f"dim={dim} dropout-proportion=0.0 continuous=true")
        return dict(drop=comp, drop_idx=comp.precompute_indexes(indexes, indexes), drop_memo=C.c_void_p(),
                    drop_in=zeros(rows, dim))

    def set_dropout_proportion(self, proportion: float):
        """What the per-iteration `set-dropout-proportion name=* proportion=p` edit does (utils.cc:1297-1330)."""
        assert self.cfg.dropout, "built without dropout nodes"
        named = [("tdnn1.dropout", self.t1["drop"])] + [(f"tdnnf{b + 2}.dropout", blk["drop"]) for b, blk in enumerate(self.blocks)]
        nnet3.apply_edits(f"set-dropout-proportion name=* proportion={proportion}", named)

    def _dropout_fwd(self, plan, d, out):
        """d["drop_in"] (the batch-norm output) -> out."""
        ip, r, c, is_ = _m(d["drop_in"])
        op, _, _, os_ = _m(out)
        plan.add("nnet3", self.lib.tdnnf_nnet3_propagate, d["drop"].h, d["drop_idx"].h, ip, r, c, is_, op, r, c, os_,
                 C.byref(d["drop_memo"]))

    def _dropout_bwd(self, plan, d, d_out, d_in):
        dp, r, c, ds = _m(d_out)
        ip, _, _, is_ = _m(d_in)
        plan.add("nnet3", self.lib.tdnnf_nnet3_backprop, d["drop"].h, d["drop_idx"].h, None, r, c, 0, None, 0, dp, r, c, ds,
                 d["drop_memo"], None, ip, is_)
        plan.add("nnet3", self.lib.tdnnf_nnet3_delete_memo, d["drop"].h, d["drop_memo"])

    def _freeze_batchnorm(self):
        """The synthetic stand-in for "pretrain the supernet, then sed BatchNormComponent -> BatchNormTestComponent":
        a few forward passes with train-mode batch-norm (a new Gumbel draw each), StoreStats accumulating, then every
        batch-norm becomes a BatchNormTestComponent read from the sed-ed model text (test mode).

        The frozen variances are doubled.  A 14-block stack out = 0.66 prev + BN_frozen(ReLU(affine(linear(prev)))) is
        linear in the scale of `prev`, and with frozen statistics it is only MARGINALLY stable: the per-block variance
        gain is 0.44 + 0.56 g^2 with g = (gain of this minibatch's Gumbel draw) / (gain at calibration), which is 1 at
        g = 1 and compounds exponentially with depth for g > 1 -- random-initialised weights then produce outputs of
        1e3..1e5 on some draws (observed), and eventually NaNs.  A pretrained supernet is adapted to its draws; the
        margin makes the random one contractive for typical draws (g^2 / 2 < 1)."""
        import torch

        def drop_memos():
            for blk in self.blocks:
                self.lib.tdnnf_nnet3_delete_memo(blk["lin"].h, blk["memo_lin"])
                self.lib.tdnnf_nnet3_delete_memo(blk["aff"].h, blk["memo_aff"])

        bns = self._all_bn()
        passes = 4
        for k in range(passes):
            self.x.copy_(self.make_input(-1 - k).to(self.dev))
            self.fwd_plan.run()
            torch.cuda.synchronize(self.dev)
            drop_memos()
            for bn in bns:
                self.lib.tdnnf_nnet3_delete_memo(bn["comp"].h, bn["memo"])

        def freeze(bn, variance_margin=2.0):
            text = bn["comp"].write(False).replace(b"BatchNormComponent", b"BatchNormTestComponent")
            head, rest = text.split(b"<StatsVar>")
            body, tail = rest.split(b"]", 1)
            var = np.array(body.split(b"[")[1].split(), dtype=np.float64) * variance_margin
            text = head + b"<StatsVar>  [ " + " ".join(repr(float(v)) for v in var).encode() + b" ]" + tail
            test = nnet3.Component.read(text, False)
            test.set_test_mode(True)
            return test

        self.t1["bn"] = freeze(self.t1["bn"])
        for blk in self.blocks:
            blk["bn"] = freeze(blk["bn"])
        for k in ("bn1", "bn2", "xbn1", "xbn2"):
            if k in self.head:
                self.head[k] = freeze(self.head[k])

    # ------------------------------------------------------------------ the step as pre-bound calls
    def _affine_fwd(self, plan, x, p, out):
        lib, h = self.lib, self.ctx.h
        xp, xr, xc, xs = _m(x)
        op, orr, oc, os_ = _m(out)
        wp, _, _, ws = _m(p["W"])
        b = p["b"]
        plan.add("abi", lib.tdnnf_darts_propagate, h, xp, xr, xc, xs, op, orr, oc, os_, wp, ws,
                 C.c_void_p(b.data_ptr()) if b is not None else None, 2 if b is not None else 1,
                 C.c_void_p(self.one.data_ptr()), 1, self.zero_off, 1)

    def _affine_bwd(self, plan, x, p, d_out, d_in, lr, zero=True):
        """zero=False: d_in already holds another branch's derivative (kBackpropAdds)."""
        lib, h = self.lib, self.ctx.h
        xp, xr, xc, xs = _m(x)
        dp, dr, dc, ds = _m(d_out)
        wp, _, _, ws = _m(p["W"])
        gp, gs = (_m(p["dW"])[0], _m(p["dW"])[3]) if p["dW"] is not None else (None, 0)
        one = C.c_void_p(self.one.data_ptr())
        if d_in is not None:
            ip, ir, ic, is_ = _m(d_in)
            if zero:
                plan.add("abi", lib.tdnnf_mat_set, h, ip, ir, ic, is_, 0.0)
            plan.add("abi", lib.tdnnf_darts_backprop_data, h, dp, dr, dc, ds, ip, ir, ic, is_, wp, ws, one, 1, self.zero_off, 1)
        if not self.frozen:  # learning-rate-factor 0: no model derivative (nnet3 leaves it out of the computation)
            plan.add("abi", lib.tdnnf_darts_backprop_params, h, xp, xr, xc, xs, dp, dr, dc, ds, None, 0, gp, gs,
                     C.c_void_p(p["db"].data_ptr()) if p["db"] is not None else None, one, 1, self.zero_off, 1, lr, None)

    def _mask_fwd(self, plan, blk):
        """tdnnfK.alpha -> tdnnfK.softmax -> Sum(p_j..) -> CopyN_j -> Append -> ElementwiseProduct_j -> Append: lin_out -> masked."""
        cfg, lib, h = self.cfg, self.lib, self.ctx.h
        R = blk["lin_out"].shape[0]
        lda = self.x[:R]  # stands for the `lda` rows the graph names as input; ConstantFunctionComponent ignores its input
        self.keep.append(lda)
        xp, xr, xc, xs = _m(lda)
        ap, ar, ac, as_ = _m(blk["a_rows"])
        plan.add("nnet3", lib.tdnnf_nnet3_propagate, blk["alpha"].h, None, xp, xr, xc, xs, ap, ar, ac, as_, None)
        pp, _, _, ps = _m(blk["p"])
        plan.add("nnet3", lib.tdnnf_nnet3_propagate, blk["soft"].h, None, ap, ar, ac, as_, pp, ar, ac, ps, None)
        lp, _, lc, ls = _m(blk["lin_out"])
        mp, _, _, ms = _m(blk["masked"])
        if cfg.fuse_mask:
            plan.add("abi", lib.tdnnf_shared_mask_fwd, h, pp, ar, ac, ps, lp, lc, ls, mp, ms, blk["widths"], 1.0)
            return
        off = 0
        for j, w in enumerate(cfg.candidate_widths):
            # Sum(softmax_j, .., softmax_7): matrix-add commands on 1-column sub-matrices
            m_in = blk["m_in"][j]
            plan.add("abi", lib.tdnnf_mat_set, h, *_m(m_in), 0.0)
            for k in range(j, ac):
                col = blk["p"][:, k:k + 1]
                self.keep.append(col)
                plan.add("abi", lib.tdnnf_mat_axpy, h, 1.0, _m(col)[0], _m(col)[3], _m(m_in)[0], _m(m_in)[3], ar, 1)
            cn = blk["cn"][j]
            plan.add("abi", lib.tdnnf_mat_set, h, *_m(cn), 0.0)                      # kPropagateAdds
            plan.add("nnet3", lib.tdnnf_nnet3_propagate, blk["copyn"][j].h, None, *_m(m_in), *_m(cn), None)
            pin = blk["pin"][j]                                                      # Append(copyn_j, linear_j)
            left, right, lin_j = pin[:, :w], pin[:, w:], blk["lin_out"][:, off:off + w]
            self.keep += [left, right, lin_j]
            plan.add("abi", lib.tdnnf_add_scaled, h, _m(cn)[0], _m(cn)[3], 1.0, _m(cn)[0], _m(cn)[3], 0.0, _m(left)[0], _m(left)[3], ar, w)
            plan.add("abi", lib.tdnnf_add_scaled, h, _m(lin_j)[0], _m(lin_j)[3], 1.0, _m(lin_j)[0], _m(lin_j)[3], 0.0, _m(right)[0],
                     _m(right)[3], ar, w)
            out_j = blk["masked"][:, off:off + w]
            self.keep.append(out_j)
            plan.add("nnet3", lib.tdnnf_nnet3_propagate, blk["prod"][j].h, None, *_m(pin), *_m(out_j), None)
            off += w

    def _mask_bwd(self, plan, blk):
        """d_masked -> d_lin (for the `linear` layer) and the alpha update (softmax backprop with the FLOPs penalty, then
        ConstantFunctionComponent::Backprop into the delta component)."""
        cfg, lib, h = self.cfg, self.lib, self.ctx.h
        ar, ac = blk["p"].shape
        pp, _, _, ps = _m(blk["p"])
        lp, _, lc, ls = _m(blk["lin_out"])
        dm, _, _, dms = _m(blk["d_masked"])
        dl, _, _, dls = _m(blk["d_lin"])
        dp, _, _, dps = _m(blk["d_p"])
        if cfg.fuse_mask:
            plan.add("abi", lib.tdnnf_shared_mask_bwd, h, pp, ps, lp, ls, dm, dms, dl, dls, dp, dps, ar, lc, ac, blk["widths"], 1.0)
        else:
            plan.add("abi", lib.tdnnf_mat_set, h, dp, ar, ac, dps, 0.0)
            off = 0
            for j, w in enumerate(cfg.candidate_widths):
                d_out_j, d_pin = blk["d_masked"][:, off:off + w], blk["d_pin"][j]
                self.keep.append(d_out_j)
                pin = blk["pin"][j]
                plan.add("nnet3", lib.tdnnf_nnet3_backprop, blk["prod"][j].h, None, _m(pin)[0], ar, 2 * w, _m(pin)[3], None, 0,
                         _m(d_out_j)[0], ar, w, _m(d_out_j)[3], None, None, _m(d_pin)[0], _m(d_pin)[3])
                left, right, d_lin_j = d_pin[:, :w], d_pin[:, w:], blk["d_lin"][:, off:off + w]
                self.keep += [left, right, d_lin_j]
                plan.add("abi", lib.tdnnf_add_scaled, h, _m(right)[0], _m(right)[3], 1.0, _m(right)[0], _m(right)[3], 0.0,
                         _m(d_lin_j)[0], _m(d_lin_j)[3], ar, w)
                d_m = blk["m_in"][j]                                                 # reused as the derivative of Sum(..)
                plan.add("abi", lib.tdnnf_mat_set, h, *_m(d_m), 0.0)                 # kBackpropAdds
                plan.add("nnet3", lib.tdnnf_nnet3_backprop, blk["copyn"][j].h, None, None, ar, 1, 0, None, 0, _m(left)[0], ar, w,
                         _m(left)[3], None, None, _m(d_m)[0], _m(d_m)[3])
                for k in range(j, ac):                                               # transpose of the Sum descriptor
                    col = blk["d_p"][:, k:k + 1]
                    self.keep.append(col)
                    plan.add("abi", lib.tdnnf_mat_axpy, h, 1.0, _m(d_m)[0], _m(d_m)[3], _m(col)[0], _m(col)[3], ar, 1)
                off += w
        ap, _, _, as_ = _m(blk["a_rows"])
        da, _, _, das = _m(blk["d_a"])
        plan.add("nnet3", lib.tdnnf_nnet3_backprop, blk["soft"].h, None, ap, ar, ac, as_, pp, ps, dp, ar, ac, dps, None, None, da, das)
        plan.add("nnet3", lib.tdnnf_nnet3_backprop, blk["alpha"].h, None, None, ar, ac, 0, None, 0, da, ar, ac, das, None,
                 blk["alpha_delta"].h, None, 0)

    def _with_planes(self, plan, holder, add_consumer):
        """The consumer call(s) added by add_consumer() run with the planes handle in `holder` (a ctypes pointer filled at
        run time by a producer kernel) attached to the context; afterwards the producer's reference is dropped (a
        TdnnDARTSV3Component keeps its own in the memo until Backprop)."""
        if holder is None:
            add_consumer()
            return
        lib, h = self.lib, self.ctx.h
        plan.add("abi", lib.tdnnf_ctx_planes_attach, h, holder)
        add_consumer()
        plan.add("abi", lib.tdnnf_ctx_planes_detach, h, holder)
        plan.add("abi", lib.tdnnf_planes_release, holder)

    def _bucket_hook(self, plan, kind, key):
        """Inside the backward plan: once the backward pass has finished the last component of a delta bucket, start its
        all-reduce on the side stream (only with dp_buckets > 1 and more than one rank)."""
        if self.dp is None or self.dp_buckets <= 1:
            return
        for (last, begin, end) in self.delta_buckets:
            if last == (kind, key):
                ptr, count = self.delta_arena.data_ptr() + 4 * begin, end - begin
                plan.add_py(partial(self._bucket_reduce, ptr, count))

    def _bucket_reduce(self, ptr, count):
        if self._reduce_now:
            self.dp.allreduce_bucket_async(ptr, count)

    def _bn_fwd(self, plan, bn, x, out):
        lib = self.lib
        xp, xr, xc, xs = _m(x)
        op, _, _, os_ = _m(out)
        if isinstance(bn, nnet3.Component):  # BatchNormTestComponent (search stage)
            plan.add("nnet3", lib.tdnnf_nnet3_propagate, bn.h, None, xp, xr, xc, xs, op, xr, xc, os_, None)
        else:  # BatchNormComponent in training mode: Propagate returns the memo, StoreStats accumulates it
            plan.add("nnet3", lib.tdnnf_nnet3_propagate, bn["comp"].h, None, xp, xr, xc, xs, op, xr, xc, os_, C.byref(bn["memo"]))
            plan.add("nnet3", lib.tdnnf_nnet3_store_stats, bn["comp"].h, None, 0, 0, 0, op, xr, xc, os_, bn["memo"])

    def _bn_bwd(self, plan, bn, out_value, d_out, d_in):
        lib = self.lib
        vp_, r, c, vs = _m(out_value)
        dp, _, _, ds = _m(d_out)
        ip, _, _, is_ = _m(d_in)
        if isinstance(bn, nnet3.Component):
            plan.add("nnet3", lib.tdnnf_nnet3_backprop, bn.h, None, None, r, c, 0, vp_, vs, dp, r, c, ds, None, None, ip, is_)
        else:
            plan.add("nnet3", lib.tdnnf_nnet3_backprop, bn["comp"].h, None, None, r, c, 0, vp_, vs, dp, r, c, ds, bn["memo"], None, ip, is_)
            plan.add("nnet3", lib.tdnnf_nnet3_delete_memo, bn["comp"].h, bn["memo"])

    def _compile(self):
        cfg, lib, h = self.cfg, self.lib, self.ctx.h
        lr = cfg.learning_rate
        fwd, bwd = _Plan(), _Plan()
        st, t1, hd = self.stock, self.t1, self.head
        # ---- forward
        self._affine_fwd(fwd, self.x, st["tdnn1"], t1["aff"])
        fwd.add("abi", lib.tdnnf_relu_fwd, h, *_m(t1["aff"]), _m(t1["relu"])[0], _m(t1["relu"])[3])
        if cfg.dropout:
            self._bn_fwd(fwd, t1["bn"], t1["relu"], t1["drop_in"])
            self._dropout_fwd(fwd, t1, t1["out"])
        else:
            self._bn_fwd(fwd, t1["bn"], t1["relu"], t1["out"])
        prev = t1["out"]
        prev_planes = None  # the planes handle of `prev` when its producer wrote them (fused tail with tail_planes)
        for blk in self.blocks:
            pp, pr, pc, ps = _m(prev)
            lp, lr_, lc, ls = _m(blk["lin_out"])
            if cfg.mode in ("manual", "bottleneck"):  # use-bias=false => kPropagateAdds: the computer hands over a zeroed matrix
                fwd.add("abi", lib.tdnnf_mat_set, h, lp, lr_, lc, ls, 0.0)
            self._with_planes(fwd, prev_planes, lambda: fwd.add(
                "nnet3", lib.tdnnf_nnet3_propagate, blk["lin"].h, blk["lin_idx"].h, pp, pr, pc, ps, lp, lr_, lc, ls,
                C.byref(blk["memo_lin"])))
            a_in = blk["lin_out"]
            if cfg.mode == "bottleneck":
                self._mask_fwd(fwd, blk)
                a_in = blk["masked"]
                lp, lr_, lc, ls = _m(a_in)
            if blk["reorder"]:
                ip, ir, ic, is_ = _m(blk["aff_in"])
                fwd.add("abi", lib.tdnnf_copy_rows, h, lp, ls, ip, is_, ir, ic, C.c_void_p(blk["reorder"]["fwd"].data_ptr()))
                a_in = blk["aff_in"]
            ap, ar, ac, as_ = _m(a_in)
            op, orr, oc, os_ = _m(blk["aff_out"])
            fwd.add("nnet3", lib.tdnnf_nnet3_propagate, blk["aff"].h, blk["aff_idx"].h, ap, ar, ac, as_, op, orr, oc, os_,
                    C.byref(blk["memo_aff"]))
            fused = cfg.fuse_tail and isinstance(blk["bn"], nnet3.Component) and not cfg.dropout  # the fused tail has no dropout node
            # noop = Sum(Scale(0.66, prev[rows]), batchnorm(relu(affine)))
            byp = blk["bypass"]
            if byp["contiguous"]:
                src = prev[byp["offset"]: byp["offset"] + orr]
            else:
                fwd.add("abi", lib.tdnnf_copy_rows, h, pp, ps, _m(blk["byp_tmp"])[0], _m(blk["byp_tmp"])[3], orr, oc,
                        C.c_void_p(byp["map"].data_ptr()))
                src = blk["byp_tmp"]
            self.keep.append(src)
            prev_planes = None
            if fused and cfg.tail_planes and oc <= 4096:
                sp, ofp, _ = blk["bn"].bn_test_scale_offset()
                blk["out_pl"] = C.c_void_p()
                fwd.add("abi", lib.tdnnf_relu_scale_offset_bypass_fwd_planes, h, op, orr, oc, os_, C.c_void_p(sp), C.c_void_p(ofp),
                        _m(src)[0], _m(src)[3], cfg.bypass_scale, _m(blk["out"])[0], _m(blk["out"])[3], C.byref(blk["out_pl"]))
                prev_planes = blk["out_pl"]
            elif fused:
                sp, ofp, _ = blk["bn"].bn_test_scale_offset()
                fwd.add("abi", lib.tdnnf_relu_scale_offset_bypass_fwd, h, op, orr, oc, os_, C.c_void_p(sp), C.c_void_p(ofp),
                        _m(src)[0], _m(src)[3], cfg.bypass_scale, _m(blk["out"])[0], _m(blk["out"])[3])
            else:
                fwd.add("abi", lib.tdnnf_relu_fwd, h, op, orr, oc, os_, _m(blk["relu"])[0], _m(blk["relu"])[3])
                if cfg.dropout:  # batchnorm -> dropout -> noop = Sum(Scale(bypass, input), dropout)
                    self._bn_fwd(fwd, blk["bn"], blk["relu"], blk["drop_in"])
                    self._dropout_fwd(fwd, blk, blk["bn_out"])
                else:
                    self._bn_fwd(fwd, blk["bn"], blk["relu"], blk["bn_out"])
                fwd.add("abi", lib.tdnnf_add_scaled, h, _m(src)[0], _m(src)[3], cfg.bypass_scale, _m(blk["bn_out"])[0],
                        _m(blk["bn_out"])[3], 1.0, _m(blk["out"])[0], _m(blk["out"])[3], orr, oc)
            prev = blk["out"]
        self._with_planes(fwd, prev_planes, lambda: self._affine_fwd(fwd, prev, st["prefinal_l"], hd["pl"]))
        self._affine_fwd(fwd, hd["pl"], st["pc_affine"], hd["pa"])
        fwd.add("abi", lib.tdnnf_relu_fwd, h, *_m(hd["pa"]), _m(hd["pr"])[0], _m(hd["pr"])[3])
        self._bn_fwd(fwd, hd["bn1"], hd["pr"], hd["pb"])
        self._affine_fwd(fwd, hd["pb"], st["pc_linear"], hd["pli"])
        self._bn_fwd(fwd, hd["bn2"], hd["pli"], hd["pb2"])
        self._affine_fwd(fwd, hd["pb2"], st["output"], hd["out"])
        if cfg.xent:
            self._affine_fwd(fwd, hd["pl"], st["px_affine"], hd["xa"])
            fwd.add("abi", lib.tdnnf_relu_fwd, h, *_m(hd["xa"]), _m(hd["xr"])[0], _m(hd["xr"])[3])
            self._bn_fwd(fwd, hd["xbn1"], hd["xr"], hd["xb"])
            self._affine_fwd(fwd, hd["xb"], st["px_linear"], hd["xli"])
            self._bn_fwd(fwd, hd["xbn2"], hd["xli"], hd["xb2"])
            self._affine_fwd(fwd, hd["xb2"], st["output_xent"], hd["xo"])
            op, orr, oc, os_ = _m(hd["xo"])
            fwd.add("nnet3", lib.tdnnf_nnet3_propagate, hd["log_softmax"].h, None, op, orr, oc, os_, _m(hd["xls"])[0], orr, oc,
                    _m(hd["xls"])[3], None)

        # ---- backward (d_out of the output layer is filled by the objective)
        self._affine_bwd(bwd, hd["pb2"], st["output"], hd["d_out"], hd["d_pb2"], lr)
        self._bucket_hook(bwd, "stock", "output")
        self._bn_bwd(bwd, hd["bn2"], hd["pb2"], hd["d_pb2"], hd["d_pb2"])
        self._affine_bwd(bwd, hd["pb"], st["pc_linear"], hd["d_pb2"], hd["d_pb"], lr)
        self._bucket_hook(bwd, "stock", "pc_linear")
        self._bn_bwd(bwd, hd["bn1"], hd["pb"], hd["d_pb"], hd["d_pb"])
        bwd.add("abi", lib.tdnnf_relu_bwd, h, _m(hd["pr"])[0], _m(hd["pr"])[3], _m(hd["d_pb"])[0], _m(hd["d_pb"])[3],
                _m(hd["d_pa"])[0], _m(hd["d_pa"])[3], hd["pr"].shape[0], hd["pr"].shape[1])
        self._affine_bwd(bwd, hd["pl"], st["pc_affine"], hd["d_pa"], hd["d_pl"], lr)
        self._bucket_hook(bwd, "stock", "pc_affine")
        if cfg.xent:
            # output-xent: d_xls = xent_regularize * numerator posteriors (filled by the objective); its affine runs at
            # learning-rate-factor 0.5 / xent_regularize (the recipes' `learning_rate_factor`)
            xp_, xr_, xc_, xs_ = _m(hd["xls"])
            bwd.add("nnet3", lib.tdnnf_nnet3_backprop, hd["log_softmax"].h, None, None, xr_, xc_, 0, xp_, xs_, _m(hd["d_xls"])[0], xr_, xc_,
                    _m(hd["d_xls"])[3], None, None, _m(hd["d_xls"])[0], _m(hd["d_xls"])[3])
            self._affine_bwd(bwd, hd["xb2"], st["output_xent"], hd["d_xls"], hd["d_xb2"], lr * 0.5 / self.objective.opts.xent_regularize)
            self._bucket_hook(bwd, "stock", "output_xent")
            self._bn_bwd(bwd, hd["xbn2"], hd["xb2"], hd["d_xb2"], hd["d_xb2"])
            self._affine_bwd(bwd, hd["xb"], st["px_linear"], hd["d_xb2"], hd["d_xb"], lr)
            self._bucket_hook(bwd, "stock", "px_linear")
            self._bn_bwd(bwd, hd["xbn1"], hd["xb"], hd["d_xb"], hd["d_xb"])
            bwd.add("abi", lib.tdnnf_relu_bwd, h, _m(hd["xr"])[0], _m(hd["xr"])[3], _m(hd["d_xb"])[0], _m(hd["d_xb"])[3],
                    _m(hd["d_xa"])[0], _m(hd["d_xa"])[3], hd["xr"].shape[0], hd["xr"].shape[1])
            self._affine_bwd(bwd, hd["pl"], st["px_affine"], hd["d_xa"], hd["d_pl"], lr, zero=False)
            self._bucket_hook(bwd, "stock", "px_affine")
        last = self.blocks[-1]
        self._affine_bwd(bwd, last["out"], st["prefinal_l"], hd["d_pl"], last["d_out"], lr)
        self._bucket_hook(bwd, "stock", "prefinal_l")
        for bi in range(len(self.blocks) - 1, -1, -1):
            blk = self.blocks[bi]
            prev_out = self.blocks[bi - 1]["out"] if bi > 0 else t1["out"]
            d_prev = self.blocks[bi - 1]["d_out"] if bi > 0 else t1["d_out"]
            dp_, dr, dc, ds = _m(blk["d_out"])
            qp, qr, qc, qs = _m(d_prev)
            fused = cfg.fuse_tail and isinstance(blk["bn"], nnet3.Component) and not cfg.dropout  # the fused tail has no dropout node
            byp = blk["bypass"]
            daff_planes = None  # planes of d_aff when the fused tail wrote them
            if fused and byp["contiguous"]:
                # zero only the halo rows of d_prev; the fused kernel overwrites the matching rows with the bypass term
                off = byp["offset"]
                for lo_, hi_ in ((0, off), (off + dr, qr)):
                    if hi_ > lo_:
                        halo = d_prev[lo_:hi_]
                        self.keep.append(halo)
                        bwd.add("abi", lib.tdnnf_mat_set, h, _m(halo)[0], hi_ - lo_, qc, _m(halo)[3], 0.0)
                dst = d_prev[off: off + dr]
                self.keep.append(dst)
                sp, _, _ = blk["bn"].bn_test_scale_offset()
                if cfg.tail_planes and dc <= 4096:
                    daff_planes = blk["daff_pl"] = C.c_void_p()
                    bwd.add("abi", lib.tdnnf_relu_scale_offset_bypass_bwd_planes, h, dp_, ds, _m(blk["aff_out"])[0], _m(blk["aff_out"])[3],
                            C.c_void_p(sp), cfg.bypass_scale, _m(blk["d_aff"])[0], _m(blk["d_aff"])[3], _m(dst)[0], _m(dst)[3], dr, dc,
                            C.byref(daff_planes))
                else:
                    bwd.add("abi", lib.tdnnf_relu_scale_offset_bypass_bwd, h, dp_, ds, _m(blk["aff_out"])[0], _m(blk["aff_out"])[3],
                            C.c_void_p(sp), cfg.bypass_scale, _m(blk["d_aff"])[0], _m(blk["d_aff"])[3], _m(dst)[0], _m(dst)[3], dr, dc)
            else:
                # d_prev = 0 everywhere, then the bypass term on the matching rows
                bwd.add("abi", lib.tdnnf_mat_set, h, qp, qr, qc, qs, 0.0)
                if byp["contiguous"]:
                    dst = d_prev[byp["offset"]: byp["offset"] + dr]
                    self.keep.append(dst)
                    bwd.add("abi", lib.tdnnf_mat_axpy, h, cfg.bypass_scale, dp_, ds, _m(dst)[0], _m(dst)[3], dr, dc)
                else:
                    bwd.add("abi", lib.tdnnf_add_to_rows, h, cfg.bypass_scale, dp_, ds, dr, dc, qp, qs, C.c_void_p(byp["map"].data_ptr()))
                if fused:
                    sp, _, _ = blk["bn"].bn_test_scale_offset()
                    scratch = blk["bn_out"]  # d_prev of the non-contiguous (last) block is handled above; discard the kernel's copy
                    if cfg.tail_planes and dc <= 4096:
                        daff_planes = blk["daff_pl"] = C.c_void_p()
                        bwd.add("abi", lib.tdnnf_relu_scale_offset_bypass_bwd_planes, h, dp_, ds, _m(blk["aff_out"])[0],
                                _m(blk["aff_out"])[3], C.c_void_p(sp), 0.0, _m(blk["d_aff"])[0], _m(blk["d_aff"])[3], _m(scratch)[0],
                                _m(scratch)[3], dr, dc, C.byref(daff_planes))
                    else:
                        bwd.add("abi", lib.tdnnf_relu_scale_offset_bypass_bwd, h, dp_, ds, _m(blk["aff_out"])[0], _m(blk["aff_out"])[3],
                                C.c_void_p(sp), 0.0, _m(blk["d_aff"])[0], _m(blk["d_aff"])[3], _m(scratch)[0], _m(scratch)[3], dr, dc)
                else:
                    # (dropout,) batchnorm, relu
                    if cfg.dropout:
                        self._dropout_bwd(bwd, blk, blk["d_out"], blk["d_aff"])
                        self._bn_bwd(bwd, blk["bn"], blk["drop_in"], blk["d_aff"], blk["d_aff"])
                    else:
                        self._bn_bwd(bwd, blk["bn"], blk["bn_out"], blk["d_out"], blk["d_aff"])
                    bwd.add("abi", lib.tdnnf_relu_bwd, h, _m(blk["relu"])[0], _m(blk["relu"])[3], _m(blk["d_aff"])[0], _m(blk["d_aff"])[3],
                            _m(blk["d_aff"])[0], _m(blk["d_aff"])[3], dr, dc)
            # affine DARTS: in_deriv (kBackpropAdds) must start from zero
            bneck = cfg.mode == "bottleneck"
            a_src = blk["masked"] if bneck else blk["lin_out"]
            d_src = blk["d_masked"] if bneck else blk["d_lin"]
            a_in = blk["aff_in"] if blk["reorder"] else a_src
            d_ain = blk["d_aff_in"] if blk["reorder"] else d_src
            ip, ir, ic, is_ = _m(a_in)
            gp, gr, gc, gs = _m(d_ain)
            delta_h = lambda d: d.h if d is not None else None   # frozen component: no to_update, data gradient only
            bwd.add("abi", lib.tdnnf_mat_set, h, gp, gr, gc, gs, 0.0)
            self._with_planes(bwd, daff_planes, lambda: bwd.add(
                "nnet3", lib.tdnnf_nnet3_backprop, blk["aff"].h, blk["aff_idx"].h, ip, ir, ic, is_, None, 0,
                _m(blk["d_aff"])[0], dr, dc, _m(blk["d_aff"])[3], blk["memo_aff"], delta_h(blk["aff_delta"]), gp, gs))
            bwd.add("nnet3", lib.tdnnf_nnet3_delete_memo, blk["aff"].h, blk["memo_aff"])
            self._bucket_hook(bwd, "comp", (bi, "aff"))
            if blk["reorder"]:
                lp, lr_, lc, ls = _m(d_src)
                bwd.add("abi", lib.tdnnf_copy_rows, h, gp, gs, lp, ls, lr_, lc, C.c_void_p(blk["reorder"]["inv"].data_ptr()))
            if bneck:
                self._mask_bwd(bwd, blk)
                self._bucket_hook(bwd, "comp", (bi, "alpha"))
            # linear DARTS: accumulates into d_prev on top of the bypass term.  With everything below frozen (search stages)
            # the first block needs no input derivative.
            pp, pr, pc, ps = _m(prev_out)
            lp, lr_, lc, ls = _m(blk["d_lin"])
            want_in = not (self.frozen and bi == 0)
            if want_in or blk["lin_delta"] is not None:
                bwd.add("nnet3", lib.tdnnf_nnet3_backprop, blk["lin"].h, blk["lin_idx"].h, pp, pr, pc, ps, None, 0, lp, lr_, lc, ls,
                        blk["memo_lin"], delta_h(blk["lin_delta"]), qp if want_in else None, qs)
            bwd.add("nnet3", lib.tdnnf_nnet3_delete_memo, blk["lin"].h, blk["memo_lin"])
            self._bucket_hook(bwd, "comp", (bi, "lin"))
        if self.frozen:  # tdnn1 and everything below it: no derivative is needed
            self.fwd_plan, self.bwd_plan = fwd, bwd
            self._compile_update()
            return
        if cfg.dropout:
            self._dropout_bwd(bwd, t1, t1["d_out"], t1["d_out"])
            self._bn_bwd(bwd, t1["bn"], t1["drop_in"], t1["d_out"], t1["d_out"])
        else:
            self._bn_bwd(bwd, t1["bn"], t1["out"], t1["d_out"], t1["d_out"])
        bwd.add("abi", lib.tdnnf_relu_bwd, h, _m(t1["relu"])[0], _m(t1["relu"])[3], _m(t1["d_out"])[0], _m(t1["d_out"])[3],
                _m(t1["d_aff"])[0], _m(t1["d_aff"])[3], t1["relu"].shape[0], t1["relu"].shape[1])
        self._affine_bwd(bwd, self.x, st["tdnn1"], t1["d_aff"], None, lr)
        self._bucket_hook(bwd, "stock", "tdnn1")

        self.fwd_plan, self.bwd_plan = fwd, bwd
        self._compile_update()

    def _compile_update(self):
        """The parameter step (UpdateNnetWithMaxChange, utils.cc:2085-2175) as a table of (model, delta) buffers."""
        cfg, st = self.cfg, self.stock
        self.updatables = []  # (kind, model, delta, max_change)
        for blk in self.blocks:
            if cfg.mode == "bottleneck":
                self.updatables.append(("comp", blk["alpha"], blk["alpha_delta"], 0.0))  # UpdatableComponent default max-change
            else:
                self.updatables.append(("comp", blk["lin"], blk["lin_delta"], cfg.max_change))
                self.updatables.append(("comp", blk["aff"], blk["aff_delta"], cfg.max_change))
        stock_names = [] if self.frozen else list(st)
        for name in stock_names:
            self.updatables.append(("stock", st[name], None, 1.5 if name.startswith("output") else cfg.max_change))
        # every parameter buffer of the model with the matching delta buffer: (model ptr, stride, delta ptr, stride, rows, cols, group)
        bufs = []
        for i, (kind, m, d, _) in enumerate(self.updatables):
            if kind == "comp":
                for (mp, r, c, ms), (dp, r2, c2, ds) in zip(m.param_buffers(), d.param_buffers()):
                    assert (r, c) == (r2, c2)
                    bufs.append((mp, ms, dp, ds, r, c, i))
            else:
                wp, wr, wc, ws = capi._mat(m["W"])
                gp, _, _, gs = capi._mat(m["dW"])
                bufs.append((wp, ws, gp, gs, wr, wc, i))
                if m["b"] is not None:
                    n = m["b"].numel()
                    bufs.append((m["b"].data_ptr(), n, m["db"].data_ptr(), n, 1, n, i))
        self.param_table = capi.ParamTable(self.ctx, bufs, [u[3] for u in self.updatables])
        # per updatable component: learning rate and l2-regularize of ApplyL2Regularization (utils.cc:2223-2245)
        self.l2_lrate, self.l2_value = [], []
        for kind, m, d, _ in self.updatables:
            if kind == "comp":
                self.l2_lrate.append(d.learning_rate())
                self.l2_value.append(cfg.l2_regularize)
        for name in stock_names:
            self.l2_lrate.append(cfg.learning_rate * (0.5 / self.objective.opts.xent_regularize if name == "output_xent" else 1.0))
            self.l2_value.append(0.002 if name.startswith("output") else cfg.l2_regularize)  # output_opts (run_tdnn_7q_fbk_40_manual.sh:123)

    def _allreduce_deltas(self):
        """Sum of the ranks' deltas (tdnnf_dp_*: NCCL over NVLink through the C ABI), in place on the delta arena."""
        if self.dp_buckets > 1:
            self.dp.wait()  # the buckets were issued from inside the backward plan
        else:
            self.dp.allreduce([(self.delta_arena.data_ptr(), self.delta_floats)])

    # ------------------------------------------------------------------ public API
    @property
    def frames_per_step(self) -> int:
        """Input frames consumed per step per GPU (chunks x frames_per_eg): the unit of the headline metric."""
        return self.cfg.num_seqs * self.cfg.frames_per_eg

    def algorithmic_flops(self) -> float:
        """fwd + dgrad + wgrad FLOPs of the block GEMMs per step; see the module-level function."""
        return algorithmic_flops(self.cfg)

    def make_input(self, step: int = 0):
        """Synthetic egs for this rank: N(0,1) features, different per rank (its shard of the minibatch)."""
        import torch

        g = synth.rng(3, stream=100 + self.rank * 1000 + step)
        return torch.from_numpy(g.standard_normal((self.rows_in, self.cfg.feat_dim)).astype(np.float32))

    def make_supervision(self, step: int = 0) -> dict:
        """Host arrays of a synthetic numerator supervision for this rank's chunks (what the egs reader would deliver with
        every minibatch); pass to step(supervision=...)."""
        g = synth.make_num_graphs(self.cfg.num_seqs, self.cfg.num_pdfs, self.T, seed=60 + self.rank + 1000 * (step + 1),
                                  den_graph=self.den_graph_host)
        return capi.NumeratorGraph.host_arrays(g)

    def step(self, x_host=None, apply_update: bool = True, reduce: bool = True, supervision: Optional[dict] = None) -> float:
        """One training step.  x_host: pinned host tensor (rows_in x feat_dim) or None to reuse device input.
        reduce=False leaves this rank's own deltas in the delta arena (no all-reduce): bench.py's data-parallel self-check.
        (Prefetching the next input on a copy stream was measured: it made the step 3 ms SLOWER than this in-stream
        copy, which costs 0.4 ms.)
        Returns the LF-MMI objective per output frame (numerator - denominator) of this rank."""
        import torch

        cfg = self.cfg
        self._reduce_now = reduce and self.dp is not None
        if x_host is not None:
            self.x.copy_(x_host, non_blocking=True)
        if supervision is not None:  # this minibatch's numerator FSTs: asynchronous upload into the resident handle
            self.last_supervision_bytes = self.num_graph.update(supervision)
        self.fwd_plan.run()
        # ComputeChainObjfAndDeriv: denominator fwd-bwd, numerator fwd-bwd, objf = num - den (host scalars, like Kaldi)
        if cfg.xent:
            objf, _, weight = self.objective.compute(self.head["out"], self.head["d_out"], self.head["d_xls"])
            # NnetChainTrainer::ProcessOutputs on output-xent: objective <log-softmax, numerator posteriors>, derivative x 0.1
            self.last_xent_objf = self.objective.xent_objf_and_deriv(self.head["xls"], self.head["d_xls"]) / weight
        else:
            objf, _, weight = self.objective.compute(self.head["out"], self.head["d_out"])
        if cfg.l2_regularize != 0.0:
            # additive and independent of the derivatives: applied before the backward pass so that the overlapped
            # bucket reductions (dp_buckets > 1) see complete deltas
            self._apply_l2_regularization()
        self.bwd_plan.run()
        if self._reduce_now:
            self._allreduce_deltas()
        if apply_update:
            self._update_with_max_change()
            self._after_update()
        return objf / weight

    def _apply_l2_regularization(self):
        """ApplyL2Regularization (utils.cc:2223-2245): delta += -2 * l2_regularize_scale * lrate * l2 * model for every
        updatable component, l2_regularize_scale = the number of sequences of the minibatch (NnetChainTrainer passes
        GetNumNvalues() * l2_regularize_factor).  Each rank adds its share (its own sequences), the all-reduce sums."""
        self.param_table.apply_l2_regularization(self.l2_lrate, self.l2_value, float(self.cfg.num_seqs))

    def _after_update(self):
        """What NnetChainTrainer::TrainInternal does after the parameter step: ConstrainOrthonormal(nnet_)
        (utils.cc:1037-1077; touches TdnnComponents / LinearComponents with a constraint: the `linear` halves, prefinal-l and
        the prefinal layers' linear part, in network order, each with probability 1/4) and ScaleBatchnormStats
        (utils.cc:527-539; train-mode batch-norm only).  Identical on every rank (same RNG counter, deterministic kernels)."""
        cfg = self.cfg
        if cfg.mode == "manual":
            nnet3.constrain_orthonormal([blk["lin"] for blk in self.blocks])
            # linear-component prefinal-l and the `linear` part of each prefinal-layer: orthonormal-constraint=-1.0
            for name in ("prefinal_l", "pc_linear", "px_linear"):
                if name in self.stock and nnet3.rand_int(0, 3) == 0:
                    self.ctx.constrain_orthonormal(self.stock[name]["W"], -1.0)
        if cfg.mode != "search" and cfg.batchnorm_stats_scale != 1.0:
            for bn in self._all_bn():
                if isinstance(bn, dict):
                    bn["comp"].scale(cfg.batchnorm_stats_scale)

    def _update_with_max_change(self, scale: float = 1.0, max_change_scale: float = 1.0):
        """UpdateNnetWithMaxChange + ScaleNnet(momentum=0) (utils.cc:2085-2175, common.py:877-878): one C call
        (tdnnf_update_with_max_change).  Returns False where the reference refuses an infinite parameter change (the
        model is then untouched and the deltas are zeroed)."""
        t = self.param_table
        self.last_update_applied = t.update_with_max_change(self.cfg.max_param_change, max_change_scale, scale, momentum=0.0)
        self.last_max_change_factors = np.array(list(t.factors), dtype=np.float64)
        return self.last_update_applied

    def _raise_nnet3(self):
        raise RuntimeError(self.lib.tdnnf_nnet3_last_error().decode())

    def close(self):
        import torch

        torch.cuda.synchronize(self.dev)
        if getattr(self, "dp", None) is not None:
            self.dp.close()
            self.dp = None
        self.objective.close()
        self.num_graph.close()
        self.den_graph.close()
