#!/usr/bin/env python
"""bench.py -- the driver-facing benchmark (see the task contract).

  python bench.py --gpus N --steps K --warmup W            our arm: DARTS TDNN-F supernet training step
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU arithmetic (oracle) on host cores

Metric (BASELINE.json): supernet training frames/sec (input frames = chunks x frames_per_eg consumed per
second; forward + LF-MMI numerator/denominator forward-backward + backward with the natural-gradient update of
every TdnnDARTSV3Component + delta reduction + max-change parameter step), on the context-offset search supernet
of configs[2] with 64 chunks x 150 frames PER GPU (weak scaling).  One JSON line on stdout from rank 0.

OnlineNaturalGradient refreshes its Fisher estimate on each of its first 10 calls and then on every 4th
(update_period 4): NG_SETTLE_STEPS untimed steps run before the W warm-up steps so that the K timed steps sample
the steady state (K = 10 holds 2-3 refresh steps, as a long run does on average).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "TDNN-F DARTS supernet train frames/sec"
# dram__bytes_read.sum + dram__bytes_write.sum per splice_gemm_kernel launch: read at run time from the committed summary of
# the `ncu --set full` capture (profiles/gemm_traffic.json, written by tools/summarize_profile.py from the .ncu-rep).
# A tensor-bound kernel: this is context, not the roofline numerator.


def load_gemm_traffic():
    path = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if not os.path.exists(path):
        return None, "profiles/gemm_traffic.json missing"
    d = json.load(open(path))
    return float(d["mean_dram_bytes_per_launch"]), f"profiles/gemm_traffic.json ({d['source']}; {d['launches']} launches)"

UNIT = "frames/s"
NG_SETTLE_STEPS = 12
# launches with fewer algorithmic FLOPs than this are the skinny natural-gradient products (H = X W^T with 20-80
# columns: <= 8.5 GFLOP; rank-r corrections), not the TdnnDARTSV3 Propagate / Backprop GEMMs (>= 11 GFLOP at this config)
MAIN_GEMM_MIN_FLOPS = 1.0e10


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"], bf16_tflops_sustained=p["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms DURING the timed regions."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(pw) if pw else None, samples=len(sm), reasons=sorted(reasons))


# ------------------------------------------------------------------ CPU baseline (the oracle, timed)
CPU_REF_CHUNKS = 16  # chunks per CPU "step": a bounded sample (a quarter of the 64-chunk minibatch, every layer at full width)


def cpu_reference_steps(cfg, steps: int, warmup: int, chunks: int = CPU_REF_CHUNKS):
    """The search-stage training step on the host cores, WHOLE (oracle/supernet_ref.py: every layer, LF-MMI numerator and
    denominator, TdnnDARTSV3 Backprop with both OnlineNaturalGradient preconditioners, max-change update), every AddMatMat
    through cblas_sgemm (numpy's bundled OpenBLAS) as in a Kaldi CPU build.  Each step is a minibatch of `chunks` chunks of
    150 frames (the GPU step has 64): nothing is extrapolated, frames/s = chunks * 150 * steps / wall time."""
    import numpy as np

    from oracle import oracle as O
    from oracle import supernet_ref as R
    from tdnnf_nas_b200 import synth

    O.use_all_cores()
    blas = O.enable_blas()
    rcfg = R.RefConfig(num_seqs=chunks, frames_per_eg=cfg.frames_per_eg, frame_subsampling=cfg.frame_subsampling, feat_dim=cfg.feat_dim,
                       dim=cfg.dim, bottleneck=cfg.bottleneck, num_blocks=cfg.num_blocks, num_offsets=cfg.num_offsets,
                       prefinal_small=cfg.prefinal_small, num_pdfs=cfg.num_pdfs, leaky_hmm=cfg.leaky_hmm, xent=cfg.xent,
                       learning_rate=cfg.learning_rate, darts_lr_factor=cfg.darts_lr_factor)
    T = cfg.frames_per_eg // cfg.frame_subsampling
    den = synth.make_den_graph(cfg.den_states, cfg.num_pdfs, cfg.den_out_degree, seed=5)
    num = synth.make_num_graphs(chunks, cfg.num_pdfs, T, seed=60, den_graph=den)
    net = R.CpuSupernet(rcfg, den, num)
    g = np.random.default_rng(7)
    x = g.standard_normal((len(net.in_t) * chunks, cfg.feat_dim)).astype(np.float32)
    draws = lambda: [g.uniform(0.05, 0.95, cfg.num_offsets).astype(np.float32) for _ in range(2 * cfg.num_blocks)]
    for _ in range(warmup):
        net.step(x, draws())
    if warmup > 0:  # like the NG_SETTLE_STEPS of the GPU arm: time the steady state (one Fisher refresh per 4 minibatches)
        for pair in net.ng.values():
            for ng in pair:
                ng.skip_initial_updates()
    t0 = time.perf_counter()
    objf = None
    for _ in range(steps):
        objf = net.step(x, draws())
    wall = time.perf_counter() - t0
    frames = chunks * cfg.frames_per_eg
    return dict(fps=frames * steps / wall, step_s=wall / steps, wall=wall, cores=O.num_threads(), chunks=chunks, objf=objf,
                sample=(f"whole search-stage training step on the host (oracle/supernet_ref.py: 14 blocks at full width, LF-MMI numerator + "
                        f"denominator on the {cfg.den_states}-state graph, TdnnDARTSV3 Backprop with both OnlineNaturalGradient "
                        f"preconditioners, max-change update) at {chunks} chunks x {cfg.frames_per_eg} frames per step (the GPU step: "
                        f"{cfg.num_seqs}), {steps} steps timed after {warmup} warm-up; GEMMs: {blas or 'plain OpenMP loops (no BLAS found)'}; "
                        "measured, not extrapolated"))


def cpu_baseline_sample(cfg, repeats: int = 1):
    """Times the oracle on a bounded sample of the workload and extrapolates to frames/sec.
    Sample: one TDNN-F block (TdnnDARTSV3 1536->160 and 160->1536, 7 offsets; Propagate + Backprop incl. the
    parameter/alpha update) at 64 sequences x 96 output frames, plus the denominator forward-backward on the
    bench graph with 4 sequences x T frames.  Extrapolation: GEMM time scales with algorithmic FLOPs, the
    denominator with the number of sequences."""
    import numpy as np

    from oracle import oracle as O
    from tdnnf_nas_b200 import synth

    O.use_all_cores()
    g = np.random.default_rng(1)
    n, D, B, S = cfg.num_offsets, cfg.dim, cfg.bottleneck, 64
    t_out = 96
    flags = O.USE_GUMBEL | O.UPDATE_ALPHA if cfg.mode == "search" else O.UNIFORM_SAMPLE
    n_eff = n if cfg.mode == "search" else 2

    def comp(din, dout, offsets):
        rs, ro = synth.regular_row_offsets(offsets, min(offsets), 0, S, 1, 1)
        in_rows, out_rows = (t_out + n - 1) * S, t_out * S
        x = g.standard_normal((in_rows, din)).astype(np.float32)
        W = (g.standard_normal((dout, n * din)) / np.sqrt(n * din)).astype(np.float32)
        bp = g.standard_normal(n + dout).astype(np.float32)
        od = (g.standard_normal((out_rows, dout)) / out_rows).astype(np.float32)
        return dict(offsets=offsets, x=x, W=W, bp=bp, od=od, ro=ro, out_rows=out_rows,
                    flops=3 * 2.0 * out_rows * n_eff * din * dout)

    comps = [comp(D, B, list(range(-(n - 1), 1))), comp(B, D, list(range(n)))]
    ug = g.uniform(0.1, 0.9, n).astype(np.float32)
    graph = synth.make_den_graph(cfg.den_states, cfg.num_pdfs, cfg.den_out_degree, seed=5)
    T, S_den = cfg.frames_per_eg // cfg.frame_subsampling, 8
    xo = g.standard_normal((T * S_den, cfg.num_pdfs)).astype(np.float32)
    t_gemm = t_den = 0.0
    for _ in range(repeats):
        t0 = time.perf_counter()
        for c in comps:
            out, coef = O.tdnn_propagate(c["offsets"], flags, 0.5, c["W"], c["bp"], c["x"], c["out_rows"], c["ro"], 1, ug, 0.3)
            O.tdnn_backprop(c["offsets"], flags, 0.5, c["W"], c["x"], c["od"], coef, c["ro"], 1, 1e-3,
                            in_deriv=np.zeros_like(c["x"]), dW=np.zeros_like(c["W"]), dbias=np.zeros_like(c["bp"]))
        t1 = time.perf_counter()
        O.den_forward_backward(graph, xo, S_den, T, cfg.leaky_hmm, deriv_weight=-1.0)
        t2 = time.perf_counter()
        t_gemm += t1 - t0
        t_den += t2 - t1
    t_gemm /= repeats
    t_den /= repeats
    return dict(t_gemm=t_gemm, t_den=t_den, sample_flops=sum(c["flops"] for c in comps), den_seqs=S_den,
                cores=O.num_threads(),
                sample=("oracle (CPU restatement, OpenMP): 1 TDNN-F block (TdnnDARTSV3 1536->160 + 160->1536, 7 offsets) "
                        "Propagate+Backprop+update at 64 seq x 96 frames (plain OpenMP/AVX2 loops: no BLAS in the image; "
                        "natural-gradient preconditioning NOT included, which favours this baseline), + denominator fwd-bwd on the bench graph at "
                        f"{S_den} seq x {T} frames; extrapolated to the full step by algorithmic GEMM FLOPs and by sequences"))


def cpu_frames_per_sec(cfg, sample, total_flops, frames_per_step):
    step_s = sample["t_gemm"] * total_flops / sample["sample_flops"] + sample["t_den"] * cfg.num_seqs / sample["den_seqs"]
    return frames_per_step / step_s, step_s


def supernet_flops(cfg):
    """Algorithmic block-GEMM FLOPs per step without building the net (Supernet.algorithmic_flops)."""
    from tdnnf_nas_b200.supernet import algorithmic_flops

    return algorithmic_flops(cfg)


def den_report(cfg, arcs, den_ms, peaks):
    """Denominator fwd-bwd roofline figures: algorithmic HBM bytes and the L2 gather bytes that
    actually bound the recursion (two 4-byte row gathers per arc, sequence and frame, both directions)."""
    S, T, N, P = cfg.num_seqs, cfg.frames_per_eg // cfg.frame_subsampling, cfg.den_states, cfg.num_pdfs
    alg = 4.0 * S * (2 * (T + 1) * (N + 1) + 3 * P * T) + 12.0 * arcs * 2
    gather = 8.0 * arcs * S * T * 2
    gbs = alg / den_ms / 1e6
    return dict(ms=den_ms, states=N, arcs=arcs, seqs=S, frames=T, algorithmic_bytes=alg, achieved_GBps=gbs,
                hbm_peak_GBps=peaks["hbm_gbs"], frac_of_hbm=gbs / peaks["hbm_gbs"], l2_gather_bytes=gather,
                l2_gather_GBps=gather / den_ms / 1e6,
                note="bound by L2 row gathers (8 B per arc, sequence and frame), ~25x the algorithmic HBM bytes")


def mixing_kernel_report(net, peaks, reps: int = 20):
    """Achieved HBM GB/s of the bandwidth-bound mixing kernels (algorithmic bytes = 4 x (elements read once +
    written once)) at R = the rows of the widest layer, each launch timed alone with CUDA events on the launch stream and
    the 126 MB L2 flushed (a 256 MB buffer is rewritten) before every launch."""
    import torch

    ctx, dev = net.ctx, net.dev
    R, D, Bn = max(blk["lin_out"].shape[0] for blk in net.blocks), net.cfg.dim, 240
    widths = [25, 25, 30, 20, 20, 40, 40, 40]
    g = torch.Generator(device=dev).manual_seed(1)
    rnd = lambda r, c: torch.randn((r, c), device=dev, generator=g)
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    a8, p8, d8 = rnd(R, 8), torch.zeros((R, 8), device=dev), rnd(R, 8)
    ctx.softmax_flops_fwd(a8, p8, None, 1.0)
    x240, y240, d240 = rnd(R, Bn), torch.zeros((R, Bn), device=dev), rnd(R, Bn)
    xD, yD, dD, prev = rnd(R, D), torch.zeros((R, D), device=dev), rnd(R, D), rnd(R, D)
    sc, of = torch.rand(D, device=dev) + 0.5, torch.randn(D, device=dev, generator=g)
    one, c40 = rnd(R, 1), torch.zeros((R, 40), device=dev)
    u8 = [0.3, 0.6, 0.2, 0.9, 0.5, 0.4, 0.7, 0.1]
    cases = [
        ("softmax_flops_fwd (SoftmaxFlops R x 8)", 2 * R * 8 * 4, lambda: ctx.softmax_flops_fwd(a8, p8, None, 1.0)),
        ("softmax_flops_fwd (Gumbel R x 8)", 2 * R * 8 * 4, lambda: ctx.softmax_flops_fwd(a8, p8, u8, 2.0)),
        ("softmax_flops_bwd (R x 8)", 3 * R * 8 * 4, lambda: ctx.softmax_flops_bwd(p8, d8, d8, 1e-3 / (R * 8), 1.0, 0)),
        ("copyn_fwd (R x 1 -> R x 40)", (R + 2 * R * 40) * 4, lambda: ctx.copyn_fwd(one, c40, 1.0)),
        ("copyn_bwd (R x 40 -> R x 1)", (R * 40 + 2 * R) * 4, lambda: ctx.copyn_bwd(c40, one, 1.0)),
        ("shared_mask_fwd (R x 240; Sum + 8 CopyN + 8 ElementwiseProduct fused)", (R * 8 + 2 * R * Bn) * 4,
         lambda: ctx.shared_mask_fwd(p8, x240, y240, widths, 1.0)),
        ("shared_mask_bwd (R x 240)", (2 * R * 8 + 3 * R * Bn) * 4, lambda: ctx.shared_mask_bwd(p8, x240, d240, y240, d8, widths, 1.0)),
        ("scale_offset_rows (BatchNormTest fwd, R x 1536)", 2 * R * D * 4, lambda: ctx.scale_offset_rows(xD, yD, sc, of)),
        ("scale_offset_rows (BatchNormTest bwd, R x 1536)", 2 * R * D * 4, lambda: ctx.scale_offset_rows(dD, yD, sc, None)),
        ("tail_fwd (ReLU + BatchNormTest + bypass fused, R x 1536)", 3 * R * D * 4,
         lambda: ctx.relu_scale_offset_bypass_fwd(xD, sc, of, prev, 0.66, yD)),
        ("tail_bwd (R x 1536)", 4 * R * D * 4, lambda: ctx.relu_scale_offset_bypass_bwd(dD, xD, sc, 0.66, yD, prev)),
    ]
    out = []
    for name, nbytes, fn in cases:
        fn()
        torch.cuda.synchronize(dev)
        total = 0.0
        for _ in range(reps):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize(dev)
            total += e0.elapsed_time(e1)
        us = total / reps * 1e3
        gbs = nbytes / (us * 1e-6) / 1e9
        out.append(dict(kernel=name, rows=R, algorithmic_bytes=nbytes, us=round(us, 2), achieved_GBps=round(gbs, 1),
                        frac_of_hbm=round(gbs / peaks["hbm_gbs"], 4)))
    return out


def dp_self_check(net, cfg, world, rank, local_rank, x_host):
    """Data-parallel correctness on the hardware (world > 1), one un-timed step on fresh state: (1) the deltas that come out
    of the NCCL path (tdnnf_dp_allreduce_deltas) equal the sum of the ranks' own deltas (collected with an all-gather and
    summed in rank order); (2) rank 0 ALSO runs other ranks' shards itself (a second Supernet built as that rank: same seed,
    that rank's input and numerator supervision) and compares with what that rank computed; its own shard run a second time
    gives the run-to-run spread to read that against (split-K sums are red.global.add in arrival order, and at full size a
    few of the 3e8 ReLU pre-activations per step lie within that rounding noise of zero: each flipped sign moves the
    derivatives below it by ~1e-4).  The model is not updated."""
    import torch
    import torch.distributed as dist

    from tdnnf_nas_b200 import nnet3
    from tdnnf_nas_b200.supernet import Supernet

    n = net.delta_floats
    net.step(x_host, apply_update=False, reduce=False)
    own = net.delta_arena[:n].clone()
    gathered = [torch.empty_like(own) for _ in range(world)]
    dist.all_gather(gathered, own)
    net._reduce_now = True
    if net.dp_buckets > 1:  # the one-shot form of the same NCCL path
        net.dp.allreduce([(net.delta_arena.data_ptr(), n)])
    else:
        net._allreduce_deltas()
    reduced = net.delta_arena[:n].clone()
    net.delta_arena.zero_()
    total = torch.zeros_like(own, dtype=torch.float64)
    for t in gathered:
        total += t.double()
    err_sum = float((reduced.double() - total).norm() / total.norm())
    out = None
    if rank == 0:
        counter = nnet3.get_rand_counter()
        others = sorted(set(range(1, world)) if world <= 4 else {1, world // 2, world - 1})
        errs = {}
        for r in [0] + others:
            rep = Supernet(cfg, device=local_rank, rank=r, world_size=world, process_group=None, dp_buckets=1, standalone=True)
            rep.step(rep.make_input(0).pin_memory(), apply_update=False, reduce=False)
            mine = rep.delta_arena[:n]
            errs[str(r)] = float((mine.double() - gathered[r].double()).norm() / gathered[r].double().norm())
            rep.close()
            del rep
        nnet3.set_context(net.ctx)
        nnet3.set_rand_counter(counter)  # the replicas re-seeded the shared RNG: put rank 0 back in step with the other ranks
        spread = errs.pop("0")
        out = dict(ok=bool(err_sum <= 1e-5 and all(e <= 2e-3 for e in errs.values())), allreduce_vs_sum_of_shards=err_sum,
                   shards_recomputed_on_rank0=errs, run_to_run_spread_rank0=spread, delta_floats=int(n),
                   note=("allreduce_vs_sum_of_shards: ||NCCL result - sum_r delta_r|| / ||sum||, bar 1e-5; shards_recomputed_on_rank0: "
                         "relative error between rank r's delta and the same shard run on rank 0, bar 2e-3 (two runs, each within the 1e-3 "
                         "derivative tolerance of the exact result); "
                         "run_to_run_spread_rank0: rank 0's own shard run twice -- the floor of that comparison (split-K red.global.add "
                         "order + ReLU sign ties)"))
    dist.barrier()
    return out


# ------------------------------------------------------------------ arms
def run_reference(args, cfg, rank, world):
    if rank != 0:
        return
    if cfg.mode == "search":
        r = cpu_reference_steps(cfg, args.steps, args.warmup)
        line = dict(impl="reference", metric=METRIC, value=r["fps"], unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                    ms_per_step=r["step_s"] * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                    data="synthetic", config=workload_config(cfg, args.gpus),
                    cpu_baseline=dict(value=r["fps"], unit=UNIT, cores=r["cores"], kind="port", sample=r["sample"]),
                    e2e=dict(value=r["fps"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                    objf_per_frame=r["objf"],
                    note=("the reference (a patch set on upstream Kaldi) cannot be built here; this arm runs the in-repo CPU restatement "
                          f"of its methods on one host ({r['cores']} threads, {r['chunks']} chunks per step); single-host number, does "
                          "not scale with --gpus"))
        print(json.dumps(line), flush=True)
        return
    frames = cfg.num_seqs * cfg.frames_per_eg
    flops = supernet_flops(cfg)
    for _ in range(args.warmup):
        cpu_baseline_sample(cfg)
    t0 = time.perf_counter()
    acc = dict(t_gemm=0.0, t_den=0.0)
    s = None
    for _ in range(args.steps):
        s = cpu_baseline_sample(cfg)
        acc["t_gemm"] += s["t_gemm"]
        acc["t_den"] += s["t_den"]
    wall = time.perf_counter() - t0
    s["t_gemm"], s["t_den"] = acc["t_gemm"] / args.steps, acc["t_den"] / args.steps
    fps, step_s = cpu_frames_per_sec(cfg, s, flops, frames)
    line = dict(impl="reference", metric=METRIC, value=fps, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=step_s * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", config=workload_config(cfg, args.gpus),
                cpu_baseline=dict(value=fps, unit=UNIT, cores=s["cores"], kind="port", sample=s["sample"]),
                e2e=dict(value=fps, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                note=("the reference (a patch set on upstream Kaldi) cannot be built here; this arm times the in-repo CPU "
                      f"oracle on one host ({s['cores']} threads); measured sample wall {wall / args.steps:.2f} s/step; "
                      "it is a single-host number and does not scale with --gpus"))
    print(json.dumps(line), flush=True)


def workload_config(cfg, gpus):
    if cfg.mode == "bottleneck":
        soft = "GumbelSoftmaxFlopsComponent" if cfg.bottleneck_gumbel else "SoftmaxFlopsComponent"
        return dict(workload=("bottleneck-dimension search (BASELINE.json configs[3]; generate_bottleneckCB8share_onehottrain_config.py + "
                              "add_flopsconstraint.py + run_TDNNf_DARTS_mod_fbk_bottleneckCBshare_cvupdate_flopsconstraint.sh): tdnn1 220->1536, "
                              "14 tdnnf-layers {TdnnComponent 1536->240, ConstantFunctionComponent alpha(8) -> " + soft +
                              f"(scale {cfg.flops_coef}) -> Sum(p_j..p_7) -> 8 x CopyNComponent (25,25,30,20,20,40,40,40) -> 8 x "
                              "ElementwiseProductComponent on the 240-wide bottleneck, TdnnComponent 240->1536, ReLU, BatchNormTest, bypass "
                              "0.66} with time-strides 1,1,1,0,3x10, prefinal 256/1536, output 6008; every pre-trained component frozen "
                              "(learning-rate-factor 0), the 14 alpha vectors trained; LF-MMI on a synthetic "
                              f"{cfg.den_states}-state den graph"),
                    mode=cfg.mode, chunks_per_gpu=cfg.num_seqs, frames_per_eg=cfg.frames_per_eg, global_chunks=cfg.num_seqs * gpus,
                    num_pdfs=cfg.num_pdfs, den_states=cfg.den_states, parallelism=f"dp{gpus}", flops_coef=cfg.flops_coef,
                    fused_mask=cfg.fuse_mask,
                    cache="per-step working set exceeds the 126 MB L2: no explicit flush needed",
                    included=("forward, LF-MMI numerator and denominator (+ xent branch), data gradients through every layer, the "
                              "mask / softmax (FLOPs penalty) / alpha backward, all-reduce of the alpha deltas, UpdateNnetWithMaxChange"),
                    not_included="parameter gradients of the frozen layers (none are computed by the reference either)")
    if cfg.mode == "manual":
        # secondary workload (python bench.py --mode manual --chunks 128): NOT the headline line, see profiles/
        return dict(workload=("manual TDNN-F 7q fbk-40 (BASELINE.json configs[1], run_tdnn_7q_fbk_40_manual.sh): tdnn1 220->1536, 14 "
                              "tdnnf-layers {TdnnComponent 1536->160 offsets (-s,0) orthonormal, TdnnComponent 160->1536 offsets (0,s), "
                              "ReLU, BatchNorm (train mode), bypass 0.66} with time-strides 1,1,1,0,6x10, prefinal 256/1536, output 6008; "
                              f"LF-MMI on a synthetic {cfg.den_states}-state den graph"),
                    mode=cfg.mode, chunks_per_gpu=cfg.num_seqs, frames_per_eg=cfg.frames_per_eg, global_chunks=cfg.num_seqs * gpus,
                    num_pdfs=cfg.num_pdfs, den_states=cfg.den_states, parallelism=f"dp{gpus}",
                    cache="per-step working set exceeds the 126 MB L2: no explicit flush needed",
                    included=("natural-gradient update of all 28 TdnnComponents, LF-MMI numerator and denominator, l2-regularize 0.01 "
                              "(ApplyL2Regularization), UpdateNnetWithMaxChange, ConstrainOrthonormal, ScaleBatchnormStats"),
                    ng_settle_steps=NG_SETTLE_STEPS,
                    not_included="natural gradient of the stock affine layers, dropout" + ("" if cfg.xent else ", xent output branch"))
    return dict(workload=("context-offset DARTS TDNN-F supernet, search stage (BASELINE.json configs[2]): 14 x "
                          "{TdnnDARTSV3 1536->160 offsets -6..0, TdnnDARTSV3 160->1536 offsets 0..6, ReLU, BatchNormTest, "
                          "bypass 0.66}, tdnn1 220->1536, prefinal 256/1536, output 6008" +
                          (", prefinal-xent / output-xent (log-softmax) with xent-regularize 0.1" if cfg.xent else "") +
                          f"; LF-MMI denominator fwd-bwd on a synthetic {cfg.den_states}-state den graph"),
                mode=cfg.mode, chunks_per_gpu=cfg.num_seqs, frames_per_eg=cfg.frames_per_eg, global_chunks=cfg.num_seqs * gpus,
                num_pdfs=cfg.num_pdfs, den_states=cfg.den_states, parallelism=f"dp{gpus}",
                cache="per-step working set (~14 GB of activations) exceeds the 126 MB L2: no explicit flush needed",
                included=("natural-gradient update (OnlineNaturalGradient rank 20/80, update period 4) of all 28 TdnnDARTSV3 "
                          "components, LF-MMI numerator (per-sequence FST) and denominator, UpdateNnetWithMaxChange" +
                          (", the cross-entropy regularisation branch (numerator posteriors -> output-xent)" if cfg.xent else "") +
                          ("; tdnn1 / prefinal / output are also trained (plain SGD; --train-stock), which the search recipe does not do"
                           if cfg.freeze_stock is False else
                           "; everything but the TdnnDARTSV3 components is frozen (learning-rate-factor 0, "
                           "run_TDNN_DARTSV3_fbk_stride_cvupdate.sh:129-134): no model derivatives for tdnn1 / prefinal / output, as in nnet3")),
                ng_settle_steps=NG_SETTLE_STEPS,
                not_included=("L2 regularisation (the recipe's l2 applies to frozen layers only), dropout nodes (the recipes' schedule "
                              "0,0@0.20,0.5@0.50,0 starts and ends at proportion 0; SupernetConfig.dropout builds them); the orthonormal "
                              "constraint does not cover TdnnDARTSV3 (utils.cc:1047-1061)"))


def run_ours(args, cfg, rank, world, local_rank):
    import torch

    from tdnnf_nas_b200.supernet import Supernet

    pg = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        pg = dist.group.WORLD
    net = Supernet(cfg, device=local_rank, rank=rank, world_size=world, process_group=pg, dp_buckets=args.dp_buckets)
    dev = net.dev
    host_inputs = [net.make_input(i).pin_memory() for i in range(2)]
    host_sups = [net.make_supervision(i) for i in range(2)]  # every minibatch brings its own numerator supervision
    dp_check = dp_self_check(net, cfg, world, rank, local_rank, host_inputs[0]) if world > 1 and not args.no_dp_check else None

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    def timed(n_steps, with_copy):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = net.ctx.launches
        e0.record()
        last = None
        for i in range(n_steps):
            last = net.step(host_inputs[i % 2], supervision=host_sups[i % 2]) if with_copy else net.step(None)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms, net.ctx.launches - launches0, last

    net.x.copy_(host_inputs[0])
    for i in range(NG_SETTLE_STEPS + args.warmup):
        net.step(host_inputs[i % 2])
    # ---- device-resident number (`value`)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, launches, objf = timed(args.steps, with_copy=False)
    # ---- end-to-end number (`e2e`): pinned host input copied in and the objective read back every step
    ms_e2e, _, _ = timed(args.steps, with_copy=True)
    clocks = sampler.stop() if rank == 0 else None
    # ---- roofline of the dominant kernel: per-launch CUDA events around every tensor-core GEMM of 2 more steps
    net.ctx.gemm_timing_enable(True)
    ROOF_STEPS = 4  # one natural-gradient period
    for _ in range(ROOF_STEPS):
        net.step(None)
    gt = net.ctx.gemm_timing_read_ex(MAIN_GEMM_MIN_FLOPS)
    net.ctx.gemm_timing_enable(False)
    # ---- the denominator forward-backward on its own (second half of BASELINE.json's metric): CUDA events around
    # DenominatorComputation::Forward + Backward on the step's (T*S) x P output; algorithmic bytes per SURVEY 8d
    den_reps = 5
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    net.objective.den.forward(net.head["out"])
    torch.cuda.synchronize(dev)
    d0.record()
    for _ in range(den_reps):
        net.objective.den.forward(net.head["out"])
        net.objective.den.backward(-1.0, net.head["d_out"])
    d1.record()
    torch.cuda.synchronize(dev)
    den_ms = d0.elapsed_time(d1) / den_reps
    # ---- the collective on its own: the in-place all-reduce of the delta arena, CUDA events, max over ranks
    allreduce_ms = None
    if world > 1:
        net.delta_arena.zero_()
        for _ in range(2):
            net.dp.allreduce([(net.delta_arena.data_ptr(), net.delta_floats)])
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(5):
            net.dp.allreduce([(net.delta_arena.data_ptr(), net.delta_floats)])
        a1.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([a0.elapsed_time(a1) / 5], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        allreduce_ms = float(t.item())
    mixing = mixing_kernel_report(net, load_peaks()) if (rank == 0 and (cfg.mode == "bottleneck" or args.mixing)) else None
    barrier()
    if rank != 0:
        net.close()
        torch.distributed.destroy_process_group()
        return
    peaks = load_peaks()
    gemm_traffic, gemm_traffic_source = load_gemm_traffic()
    frames_all = net.frames_per_step * world
    value = frames_all * args.steps / (ms_dev / 1e3)
    e2e = frames_all * args.steps / (ms_e2e / 1e3)
    achieved = gt["flops"] / (gt["ms"] / 1e3) / 1e12 if gt["ms"] > 0 else 0.0
    achieved_pipe = gt["pipe_flops"] / (gt["ms"] / 1e3) / 1e12 if gt["ms"] > 0 else 0.0
    peak = peaks["bf16_tflops_sustained"]
    step_ms = ms_dev / args.steps
    cpu_line = None
    if world == 1 and not args.no_cpu_baseline:  # the CPU baseline is reported at N = 1 only
        if cfg.mode == "search":
            r = cpu_reference_steps(cfg, steps=2, warmup=1)
            cpu_line = dict(value=r["fps"], unit=UNIT, cores=r["cores"], kind="port", sample=r["sample"])
        else:
            cpu = cpu_baseline_sample(cfg)
            cpu_fps, _ = cpu_frames_per_sec(cfg, cpu, net.algorithmic_flops(), net.frames_per_step)
            cpu_line = dict(value=cpu_fps, unit=UNIT, cores=cpu["cores"], kind="port", sample=cpu["sample"])
    line = dict(
        metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=step_ms,
        higher_is_better=True, scaling="weak", vs_baseline=None,
        dtype="f32 (bf16 hi/lo split operands, 3 tensor-core products, fp32 accumulate)", data="synthetic",
        config=workload_config(cfg, world), clocks=clocks,
        e2e=dict(value=e2e, unit=UNIT, h2d_bytes_per_step=int(net.x.numel() * 4 + getattr(net, "last_supervision_bytes", 0)),
                 d2h_bytes_per_step=int(8 * net.param_table.num_groups + 16 + (4 if cfg.xent else 0)), ms_per_step=ms_e2e / args.steps,
                 per_step=("pinned-host features copied in, this minibatch's numerator FSTs uploaded (tdnnf_num_graph_update), the "
                           "LF-MMI objective (2 scalars + 2 check flags), the xent objective and the max-change norms read back")),
        gpu_launches=int(launches),
        roofline=dict(bound="tensor", kernel="splice_gemm_kernel (tcgen05, the TdnnDARTSV3 Propagate / data-gradient / parameter-gradient GEMMs)",
                      achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak, traffic=gemm_traffic,
                      traffic_source=gemm_traffic_source,
                      peak_source=peaks["source"] + ", bf16 sustained",
                      achieved_tensor_pipe=achieved_pipe, frac_tensor_pipe=achieved_pipe / peak,
                      launches_timed=gt["launches"], steps_timed=ROOF_STEPS, gemm_ms_per_step=gt["ms"] / ROOF_STEPS,
                      gemm_share_of_step=(gt["ms"] / ROOF_STEPS) / step_ms,
                      skinny_ng_gemm_launches=gt["other_launches"], skinny_ng_gemm_ms_per_step=gt["other_ms"] / ROOF_STEPS,
                      note=("achieved = algorithmic fp32-equivalent FLOPs (2MNK over un-padded operands, one pass) / CUDA-event "
                            "time, over every launch with >= 10 GFLOP (the TdnnDARTSV3 GEMMs); fp32-level accuracy comes from "
                            "bf16 hi/lo operand planes and 3 tensor-core products per K step, so the tensor pipe itself runs at "
                            "achieved_tensor_pipe (counts each product issued); launches below 10 GFLOP are the skinny "
                            "natural-gradient products, timed separately")),
        den=den_report(cfg, net.den_arcs, den_ms, peaks),
        cpu_baseline=cpu_line, objf_per_frame=objf, den_arcs=net.den_arcs)
    if world > 1:
        line["dp_check"] = dp_check
        line["allreduce_ms"] = allreduce_ms
        line["allreduce"] = dict(bytes=int(net.delta_floats * 4), ms=allreduce_ms, buckets=net.dp_buckets,
                                 path="tdnnf_dp_allreduce_deltas (ncclAllReduce in place on the delta arena, NVLink)",
                                 algbw_GBps=net.delta_floats * 4 / (allreduce_ms * 1e-3) / 1e9 if allreduce_ms else None)
    if mixing is not None:
        line["mixing_kernels"] = mixing
    print(json.dumps(line), flush=True)
    net.close()
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="search", choices=["search", "pretrain", "manual", "bottleneck"])
    ap.add_argument("--flops-coef", type=float, default=1e-3, help="bottleneck mode: eta of the FLOPs penalty (0, 1e-3, 1e-1)")
    ap.add_argument("--gumbel", action="store_true", help="bottleneck mode: GumbelSoftmaxFlopsComponent instead of SoftmaxFlopsComponent")
    ap.add_argument("--unfused-mask", action="store_true", help="bottleneck mode: CopyN / ElementwiseProduct one by one")
    ap.add_argument("--mixing", action="store_true", help="add the per-kernel GB/s table of the mixing kernels to the line")
    ap.add_argument("--dp-buckets", type=int, default=1, help="N > 1: delta buckets reduced while the backward pass runs")
    ap.add_argument("--no-dp-check", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU baseline leg (profiling runs)")
    ap.add_argument("--train-stock", action="store_true", help="search mode: also train tdnn1 / prefinal / output (round-1 behaviour)")
    ap.add_argument("--den-states", type=int, default=16384)
    ap.add_argument("--blocks", type=int, default=14)
    ap.add_argument("--chunks", type=int, default=64)
    ap.add_argument("--no-xent", action="store_true", help="leave out the xent-regularisation branch (prefinal-xent / output-xent)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from tdnnf_nas_b200.supernet import SupernetConfig

    cfg = SupernetConfig(mode=args.mode, den_states=args.den_states, num_blocks=args.blocks, num_seqs=args.chunks,
                         l2_regularize=0.01 if args.mode == "manual" else 0.0, xent=not args.no_xent,
                         bottleneck=240 if args.mode == "bottleneck" else 160, flops_coef=args.flops_coef,
                         bottleneck_gumbel=args.gumbel, fuse_mask=not args.unfused_mask,
                         freeze_stock=False if args.train_stock else None)
    if args.impl == "reference":
        if cfg.mode != "search" and args.steps > 3:
            args.steps = 3  # the sampled legs of the secondary workloads: ~10 s of CPU work each
        args.warmup = min(args.warmup, 1) if cfg.mode != "search" else args.warmup
        run_reference(args, cfg, rank, world)
        return
    args.warmup = max(args.warmup, 3)
    run_ours(args, cfg, rank, world, local_rank)


if __name__ == "__main__":
    main()
