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
The out-of-range penalty (|x| > 30, sub-sampled rows upstream) is NOT applied: its exact sub-sampling is not
recoverable from the reference and it is inert for the bounded synthetic outputs used here.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

from . import capi


@dataclass
class ChainTrainingOptions:
    l2_regularize: float = 0.0            # the recipes pass --chain.l2-regularize 0.0
    leaky_hmm_coefficient: float = 0.1    # --chain.leaky-hmm-coefficient 0.1
    xent_regularize: float = 0.1          # --chain.xent-regularize 0.1: see xent_objf_and_deriv
    out_of_range_regularize: float = 0.01


class ChainObjective:
    """Holds the denominator computation and the per-minibatch numerator graphs for a fixed (S, T)."""

    def __init__(self, ctx: capi.Context, den_graph: capi.DenGraph, num_graph: capi.NumeratorGraph, num_seqs: int,
                 frames_per_seq: int, opts: ChainTrainingOptions = ChainTrainingOptions(), supervision_weight: float = 1.0):
        self.ctx, self.opts, self.S, self.T, self.sup_weight = ctx, opts, num_seqs, frames_per_seq, supervision_weight
        self.den = capi.DenominatorComputation(ctx, den_graph, num_seqs, frames_per_seq, opts.leaky_hmm_coefficient)
        self.num = num_graph

    def compute(self, nnet_output, nnet_output_deriv, xent_output_deriv=None):
        """Returns (objf, l2_term, weight); fills nnet_output_deriv (overwritten) and, if given, xent_output_deriv with
        the numerator posteriors (overwritten)."""
        w = self.sup_weight
        weight = w * self.S * self.T
        self.ctx.mat_set(nnet_output_deriv, 0.0)
        den = w * self.den.forward(nnet_output)
        den_ok = self.den.backward(-w, nnet_output_deriv)
        if xent_output_deriv is not None:
            self.ctx.mat_set(xent_output_deriv, 0.0)
            num, num_ok = self.num.forward_backward(nnet_output, self.T, w, xent_output_deriv)
            if num_ok:
                self.ctx.mat_axpy(1.0, xent_output_deriv, nnet_output_deriv)
        else:
            num, num_ok = self.num.forward_backward(nnet_output, self.T, w, nnet_output_deriv)
        num *= w
        objf = num - den
        if not math.isfinite(objf) or not den_ok or not num_ok:
            self.ctx.mat_set(nnet_output_deriv, 0.0)
            if xent_output_deriv is not None:
                self.ctx.mat_set(xent_output_deriv, 0.0)
            objf = -10.0 * weight
        l2_term = 0.0
        if self.opts.l2_regularize != 0.0 and num_ok:
            scale = w * self.opts.l2_regularize
            l2_term = -0.5 * scale * self.ctx.mat_dot(nnet_output, nnet_output)
            self.ctx.mat_axpy(-scale, nnet_output, nnet_output_deriv)
        return objf, l2_term, weight

    def xent_objf_and_deriv(self, xent_output, xent_output_deriv):
        """NnetChainTrainer::ProcessOutputs for the 'output-xent' node (kaldi: nnet3/nnet-chain-training.cc): the
        cross-entropy objective <log-softmax output, numerator posteriors>, then the derivative scaled by
        xent_regularize in place.  Returns the objective (un-scaled, as logged upstream)."""
        objf = self.ctx.mat_dot(xent_output, xent_output_deriv)
        self.ctx.mat_scale(xent_output_deriv, self.opts.xent_regularize)
        return objf

    def close(self):
        self.den.close()
