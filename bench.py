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
# dram__bytes_read.sum + dram__bytes_write.sum per splice_gemm_kernel launch, from the committed `ncu --set full`
# capture (profiles/r01_summary.md).  A tensor-bound kernel: this is context, not the roofline numerator.
NCU_GEMM_TRAFFIC_BYTES = 131.2e6
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


# ------------------------------------------------------------------ arms
def run_reference(args, cfg, rank, world):
    if rank != 0:
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
                          "; the stock affine layers (tdnn1, prefinal, output) are also trained (plain SGD) although the search recipe "
                          "freezes them with learning-rate-factor 0 (run_TDNN_DARTSV3_fbk_stride_cvupdate.sh:129): extra work in the step"),
                ng_settle_steps=NG_SETTLE_STEPS,
                not_included=("natural gradient of the 5 stock affine layers around the blocks (tdnn1, prefinal, output: plain SGD "
                              "there; all 28 TdnnDARTSV3 components are preconditioned), L2 regularisation, dropout "
                              "(GeneralDropoutComponent is upstream Kaldi and not built; the recipes' schedule 0,0@0.20,0.5@0.50,0 starts "
                              "and ends at proportion 0); the orthonormal constraint does not cover TdnnDARTSV3 (utils.cc:1047-1061)"))


def run_ours(args, cfg, rank, world, local_rank):
    import torch

    from tdnnf_nas_b200.supernet import Supernet

    pg = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        pg = dist.group.WORLD
    net = Supernet(cfg, device=local_rank, rank=rank, world_size=world, process_group=pg)
    dev = net.dev
    host_inputs = [net.make_input(i).pin_memory() for i in range(2)]

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
            last = net.step(host_inputs[i % 2] if with_copy else None)
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
    barrier()
    if rank != 0:
        net.close()
        torch.distributed.destroy_process_group()
        return
    peaks = load_peaks()
    frames_all = net.frames_per_step * world
    value = frames_all * args.steps / (ms_dev / 1e3)
    e2e = frames_all * args.steps / (ms_e2e / 1e3)
    achieved = gt["flops"] / (gt["ms"] / 1e3) / 1e12 if gt["ms"] > 0 else 0.0
    achieved_pipe = gt["pipe_flops"] / (gt["ms"] / 1e3) / 1e12 if gt["ms"] > 0 else 0.0
    peak = peaks["bf16_tflops_sustained"]
    step_ms = ms_dev / args.steps
    cpu_line = None
    if world == 1:  # the CPU baseline is reported at N = 1 only
        cpu = cpu_baseline_sample(cfg)
        cpu_fps, _ = cpu_frames_per_sec(cfg, cpu, net.algorithmic_flops(), net.frames_per_step)
        cpu_line = dict(value=cpu_fps, unit=UNIT, cores=cpu["cores"], kind="port", sample=cpu["sample"])
    line = dict(
        metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=step_ms,
        higher_is_better=True, scaling="weak", vs_baseline=None,
        dtype="f32 (bf16 hi/lo split operands, 3 tensor-core products, fp32 accumulate)", data="synthetic",
        config=workload_config(cfg, world), clocks=clocks,
        e2e=dict(value=e2e, unit=UNIT, h2d_bytes_per_step=int(net.x.numel() * 4), d2h_bytes_per_step=12,
                 ms_per_step=ms_e2e / args.steps),
        gpu_launches=int(launches),
        roofline=dict(bound="tensor", kernel="splice_gemm_kernel (tcgen05, the TdnnDARTSV3 Propagate / data-gradient / parameter-gradient GEMMs)",
                      achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak, traffic=NCU_GEMM_TRAFFIC_BYTES,
                      traffic_source="profiles/r01_summary.md: mean dram__bytes_read+write over the 10 launches of the ncu --set full capture",
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
    ap.add_argument("--mode", default="search", choices=["search", "pretrain", "manual"])
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
                         l2_regularize=0.01 if args.mode == "manual" else 0.0, xent=not args.no_xent)
    if args.impl == "reference":
        if args.steps > 3:
            args.steps = 3  # each step is ~10 s of CPU work: keep the whole run within a few minutes
        args.warmup = min(args.warmup, 1)
        run_reference(args, cfg, rank, world)
        return
    args.warmup = max(args.warmup, 3)
    run_ours(args, cfg, rank, world, local_rank)


if __name__ == "__main__":
    main()
