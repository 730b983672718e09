"""ComputeChainObjfAndDeriv (kaldi: chain/chain-training.cc) on the den / num kernels.

    weight = sup.weight * num_sequences * frames_per_sequence
    deriv  = 0
    den    = sup.weight * Denominator.Forward();  Denominator.Backward(-sup.weight, &deriv)      (denominator first)
    num    = sup.weight * Numerator.Forward();    Numerator.Backward(&deriv)  (scaled by sup.weight)
    objf   = num - den;   non-finite objf or a failed check  =>  deriv = 0, objf = -10 * weight
    l2     = -0.5 * sup.weight * l2_regularize * ||nnet_output||^2,  deriv += -sup.weight * l2_regularize * nnet_output
xent branch (xent_output_deriv != None): the numerator posteriors (x sup.weight) are returned on their own, as the
targets of the cross-entropy output; NnetChainTrainer then takes objf_xent = <xent_output, xent_deriv> and feeds
xent_regularize * xent_deriv back as that output's derivative (xent_objf_and_deriv below).
out-of-range penalty: PenalizeOutOfRange(limit 30, scale 2 * out_of_range_regularize) on every oor_row_step-th row
starting at a RandInt(0, step - 1) offset drawn by the caller, the scale multiplied by the step.  Upstream's exact
sub-sampling is not part of the reference tree (unpinned); the penalty is zero whenever |nnet_output| <= 30.

The arithmetic lives in the C entry tdnnf_chain_objf_and_deriv (csrc/chain_step.cu), so that a C++ host needs no
Python; this class only holds the handles.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import ctypes as C

from . import capi


@dataclass
class ChainTrainingOptions:
    l2_regularize: float = 0.0            # the recipes pass --chain.l2-regularize 0.0
    leaky_hmm_coefficient: float = 0.1    # --chain.leaky-hmm-coefficient 0.1
    xent_regularize: float = 0.1          # --chain.xent-regularize 0.1: see xent_objf_and_deriv
    out_of_range_regularize: float = 0.01
    oor_row_step: int = 4                 # rows penalised: every 4th (upstream sub-samples; unpinned)


class ChainObjective:
    """Holds the denominator computation and the per-minibatch numerator graphs for a fixed (S, T)."""

    def __init__(self, ctx: capi.Context, den_graph: capi.DenGraph, num_graph: capi.NumeratorGraph, num_seqs: int,
                 frames_per_seq: int, opts: ChainTrainingOptions = ChainTrainingOptions(), supervision_weight: float = 1.0):
        self.ctx, self.opts, self.S, self.T, self.sup_weight = ctx, opts, num_seqs, frames_per_seq, supervision_weight
        self.den = capi.DenominatorComputation(ctx, den_graph, num_seqs, frames_per_seq, opts.leaky_hmm_coefficient)
        self.num = num_graph

    def compute(self, nnet_output, nnet_output_deriv, xent_output_deriv=None, oor_row_offset: int = 0):
        """Returns (objf, l2_term, weight); fills nnet_output_deriv (overwritten) and, if given, xent_output_deriv with
        the numerator posteriors (overwritten).  oor_row_offset: the caller's RandInt(0, oor_row_step - 1) draw."""
        xp, rows, cols, xs = capi._mat(nnet_output)
        assert rows == self.S * self.T
        dp, ds = (0, 0) if nnet_output_deriv is None else (capi._mat(nnet_output_deriv)[0], capi._mat(nnet_output_deriv)[3])
        qp, qs = (0, 0) if xent_output_deriv is None else (capi._mat(xent_output_deriv)[0], capi._mat(xent_output_deriv)[3])
        objf, l2_term, weight = C.c_float(0), C.c_float(0), C.c_float(0)
        capi.check(capi.load().tdnnf_chain_objf_and_deriv(
            self.ctx.h, self.den.h, self.num.h, xp, xs, self.S, self.T, cols, self.sup_weight, self.opts.l2_regularize,
            self.opts.out_of_range_regularize, self.opts.oor_row_step, oor_row_offset, dp, ds, qp, qs, C.byref(objf),
            C.byref(l2_term), C.byref(weight)))
        return float(objf.value), float(l2_term.value), float(weight.value)

    def xent_objf_and_deriv(self, xent_output, xent_output_deriv):
        """NnetChainTrainer::ProcessOutputs for the 'output-xent' node (kaldi: nnet3/nnet-chain-training.cc): the
        cross-entropy objective <log-softmax output, numerator posteriors>, then the derivative scaled by
        xent_regularize in place.  Returns the objective (un-scaled, as logged upstream)."""
        objf = self.ctx.mat_dot(xent_output, xent_output_deriv)
        self.ctx.mat_scale(xent_output_deriv, self.opts.xent_regularize)
        return objf

    def close(self):
        self.den.close()
