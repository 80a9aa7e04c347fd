#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric on B200: sampled images/s of the 500-step
(499-evaluation) 32x32 DDPM reverse loop, plus training images/s as a secondary figure.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
  python bench.py --impl reference ...                     (CPU restatement of the Flux path)

A "step" is one full pass of the hot path over one batch: `--images` images per GPU taken
through all T-1 = 499 U-Net evaluations + reverse updates (5200 = 4 graph chunks of 1300, the chunk size that fills
the persistent kernels' tile rounds exactly; 12.6 such steps == BASELINE config 4).  The end-to-end leg (`e2e`) runs
config 4 LITERALLY: one `ddpm_sample(N=65536)` call through the C ABI per rank, pinned host x_T in, images out.
Weak scaling: every rank samples its own images, global image indices are disjoint, no data-path collective.

Prints ONE JSON line on rank 0 (see the contract in the task description / DESIGN.md).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_STEPS = 500
FLOP_PER_EVAL_FOLDED = 735.31e6     # SURVEY.md Appendix A with the embedding fold (what the kernels execute)
FLOP_PER_EVAL_REFERENCE = 886.31e6  # the reference's 129-channel formulation
FLOP_PER_TRAIN_IMG = 2204.76e6      # fwd + dgrad + wgrad with the fold


def ncu_traffic(kernel: str, images: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of `kernel` at `images` images per launch, read from
    the committed summary of the `ncu --set full` captures (profiles/ncu_kernels.json, written by
    profiles/summarize_ncu.py from the .ncu-rep files); None when no capture at that size exists."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_kernels.json")))
        e = d.get(kernel, {}).get(str(images))
        return None if e is None else float(e["dram_bytes_read"]) + float(e["dram_bytes_write"])
    except Exception:
        return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    td = None
    if world > 1:
        import torch
        import torch.distributed as td_

        torch.cuda.set_device(local)
        td_.init_process_group("nccl", device_id=torch.device("cuda", local))
        td = td_
    return rank, world, local, td


def max_over_ranks(td, value, local):
    if td is None:
        return value
    import torch

    t = torch.tensor([value], dtype=torch.float64, device=torch.device("cuda", local))
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t.item())


def barrier(td, local):
    import torch

    if td is not None:
        td.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize(local)


# ------------------------------------------------------------------------------------ reference arm
def cpu_reference_sampling(n_images, n_steps_sample, threads=None):
    """Times the CPU restatement of the reference's Flux path (oracle/ddpm_oracle.py, torch-CPU
    fp32 conv = im2col+GEMM like NNlib) on a bounded sample: n_images images through
    n_steps_sample of the 499 reverse steps; returns (images/s extrapolated to 499 steps, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch

    import ddpm_oracle as O
    import igdm_b200  # noqa: F401
    from igdm_b200 import bson_io

    if threads:
        torch.set_num_threads(threads)
    arrs, _ = bson_io.load_checkpoint(os.path.join(ROOT, "fixtures", "trained_model.bson"))
    net = O.Net([a.flat for a in arrs])
    _, _, acum = O.schedule(T_STEPS)
    pe = O.embedding_table(T_STEPS)
    rng = np.random.default_rng(0)
    xT = rng.standard_normal((n_images, 1, 32, 32)).astype(np.float32)
    z = rng.standard_normal((n_steps_sample, n_images, 1, 32, 32)).astype(np.float32)
    O.generate_image(net, xT[:1], z[:2, :1], acum, pe, t_start=3)  # warm-up
    t0 = time.perf_counter()
    O.generate_image(net, xT, z, acum, pe, t_start=n_steps_sample + 1)
    dt = time.perf_counter() - t0
    per_eval = dt / (n_images * n_steps_sample)
    return 1.0 / (per_eval * (T_STEPS - 1)), dt, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    # torchrun exports OMP_NUM_THREADS=1 to every rank: the CPU arm must ask for the host's cores itself
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    n_img, n_st = args.ref_images, args.ref_steps
    # bound the whole `--steps K --warmup W` run to ~ref_budget seconds: calibrate on a small sample, then shrink the
    # per-step sample (reverse steps first, then images) -- the metric is extrapolated per evaluation anyway
    _, dt_cal, _ = cpu_reference_sampling(16, 4, threads)
    per_eval = dt_cal / (16 * 4)
    budget = args.ref_budget / max(1.0, args.steps + 0.25 * args.warmup)
    while n_img * n_st * per_eval > budget and n_st > 4:
        n_st = max(4, n_st // 2)
    while n_img * n_st * per_eval > budget and n_img > 8:
        n_img = max(8, n_img // 2)
    vals, tot = [], 0.0
    for _ in range(args.warmup):
        cpu_reference_sampling(max(1, n_img // 4), max(2, n_st // 4), threads)
    for _ in range(args.steps):
        v, dt, threads = cpu_reference_sampling(n_img, n_st, threads)
        vals.append(v)
        tot += dt
    value = float(np.mean(vals))
    sample = f"{n_img} images x {n_st} of 499 reverse steps per step, extrapolated linearly to 499"
    line = {
        "impl": "reference", "metric": "sampled img/s (500-step DDPM, 32x32)", "value": value, "unit": "img/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "generate_image: 499-evaluation reverse loop, 32x32, trained_model.bson weights",
                   "images_per_step_per_gpu": n_img, "T": T_STEPS,
                   "note": "Flux-semantics CPU restatement (Julia unavailable in this image); rank 0 only, all host threads"},
        "cpu_baseline": {"value": value, "unit": "img/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import igdm_b200  # noqa: F401
    from igdm_b200 import api, capi, tables

    rank, world, local, td = dist_setup(args.gpus)
    if capi.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: libddpm has no CPU fallback")
    peaks = load_peaks()
    prec = {"fp32": capi.PREC_FP32, "fp16": capi.PREC_FP16, "bf16": capi.PREC_BF16, "tf32": capi.PREC_TF32}[args.precision]
    h = capi.Handle(T=T_STEPS, precision=prec, device=local)
    beta, _, acum = tables.beta_schedule(T_STEPS)
    h.set_tables(beta, acum, tables.embedding_table(T_STEPS))
    model = api.SimpleUNet.load()
    h.set_weights(model.arrays)
    h.set_option("sample_chunk", args.chunk)
    if args.streams > 0:
        h.set_option("sample_streams", args.streams)
    for kv in filter(None, os.environ.get("DDPM_OPTS", "").split(",")):   # e.g. DDPM_OPTS=tc_pair=0,fuse_final=0
        k, v = kv.split("=")
        h.set_option(k, int(v))
    t_start = args.t_start
    evals = t_start - 1

    def first_index(step_no):
        return (step_no * world + rank) * args.images

    if args.train_images <= 0:
        args.train_images = max(2, 4096 // world)
    common = (h, td, rank, world, local, peaks)
    if args.workload in ("sample", "both"):
        line = run_workload(args, "sample", *common, args.images, t_start, evals, first_index, api, capi, tables)
    else:
        line = run_workload(args, "train", *common, args.train_images, t_start, evals, first_index, api, capi, tables)
    if args.workload == "both":
        # secondary figure of BASELINE.json's metric: training images/s on the same GPUs, same JSON line
        tl = run_workload(args, "train", *common, args.train_images, t_start, evals, first_index, api, capi, tables)
        if rank == 0:
            line["train"] = {k: tl[k] for k in ("metric", "value", "unit", "ms_per_step", "e2e", "gpu_launches", "config")}
            line["train"]["scaling"] = "strong (global batch fixed at %d)" % (args.train_images * world)
            line["train"]["tflops"] = tl["roofline"]["whole_step_tflops"]
            line["train"]["frac_of_sustained_peak"] = tl["roofline"]["whole_step_frac_of_sustained"]
            # strong-scaling efficiency against the committed 1-GPU figure of the same global batch (the driver
            # recomputes it from its own per-N runs; this is the builder-side number for the record)
            line["train"]["per_gpu_value"] = tl["value"] / world
            try:
                one = json.load(open(os.path.join(ROOT, "profiles", "r2", "train_1gpu.json")))
                line["train"]["efficiency_vs_1gpu"] = 1.0 if world == 1 else tl["value"] / (world * one["value"])
                line["train"]["one_gpu_value"] = tl["value"] if world == 1 else one["value"]
            except Exception:
                line["train"]["efficiency_vs_1gpu"] = 1.0 if world == 1 else None
    if rank == 0:
        emit(line)
    if td is not None:
        td.barrier()
        td.destroy_process_group()


def run_workload(args, workload, h, td, rank, world, local, peaks, N, t_start, evals, first_index, api, capi, tables):
    if workload == "sample":
        # ---- device-resident timing (value)
        for w in range(args.warmup):
            h.sample_device(N, seed=args.seed, first_index=first_index(w), t_start=t_start)
        barrier(td, local)
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
        l0 = h.counter("launches")
        h.timer_start()
        for k in range(args.steps):
            h.sample_device(N, seed=args.seed, first_index=first_index(args.warmup + k), t_start=t_start)
        ms = h.timer_stop()
        barrier(td, local)
        launches = h.counter("launches") - l0
        clk = clocks.stop() if rank == 0 else None
        ms = max_over_ranks(td, ms, local)
        value = args.steps * N * world / (ms * 1e-3)
        # ---- end to end through the public C-ABI call with pinned host buffers: BASELINE config 4 literally,
        #      ONE ddpm_sample(N = 65536) call per rank (x_T from pinned host memory, images back to pinned host memory)
        import torch

        NE = args.e2e_images if args.e2e_images > 0 else N
        x_host = torch.empty((NE, 1024), dtype=torch.float32, pin_memory=True)
        o_host = torch.empty((NE, 1024), dtype=torch.float32, pin_memory=True)
        x_host.normal_(generator=torch.Generator().manual_seed(1 + rank))
        xin, oout = x_host.numpy(), o_host.numpy()
        # warm-up: the chunk sizes of the big call (full chunks + the remainder chunk, see sample_impl in csrc/libddpm.cu)
        rem = NE % args.chunk
        nw = min(NE, args.chunk + rem) if (NE > args.chunk and rem * 4 >= args.chunk) else min(NE, 2 * args.chunk)
        h.sample(nw, x_T=xin[:nw], seed=args.seed, first_index=0, t_start=t_start, out=oout[:nw])
        barrier(td, local)
        t0 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        for k in range(e2e_steps):
            h.sample(NE, x_T=xin, seed=args.seed, first_index=(k * world + rank) * NE, t_start=t_start, out=oout)
        barrier(td, local)
        e2e_s = max_over_ranks(td, time.perf_counter() - t0, local)
        e2e = {"value": e2e_steps * NE * world / e2e_s, "unit": "img/s", "h2d_bytes_per_step": int(xin.nbytes),
               "d2h_bytes_per_step": int(oout.nbytes), "steps": e2e_steps, "images_per_call_per_gpu": NE,
               "note": "one ddpm_sample call of NE images per step per rank (NE = 65536: BASELINE config 4 as quoted)"}
        metric = "sampled img/s (500-step DDPM, 32x32)"
        workload_desc = (f"generate_image: {evals}-evaluation reverse loop (t={t_start}..2), 32x32, trained_model.bson weights, "
                    f"device Philox noise; value: {N} device-resident images per step and GPU (BASELINE config 4 == {65536 / N:.1f} such steps); "
                    f"e2e: config 4 as quoted, one ddpm_sample(N={args.e2e_images if args.e2e_images > 0 else N}) call with host buffers")
        flop_per_unit = FLOP_PER_EVAL_FOLDED * evals
    else:
        # ---- training throughput: data parallel, global batch = images * world
        data = (api.load_dataset() * np.float32(2) - np.float32(1)).astype(np.float32)
        rng = np.random.default_rng(4)
        synth = rng.uniform(-1, 1, (max(N, 512), 1, 32, 32)).astype(np.float32)
        synth[:500] = data
        h.upload_dataset(synth)
        h.set_adam(1e-4, 0.9, 0.999, 1e-8)
        if world > 1:
            from igdm_b200 import dist

            dist.init_data_parallel(h, sync_bn=bool(args.sync_bn))
        idx = (np.arange(N) % synth.shape[0]).astype(np.int32)
        for w in range(args.warmup):
            h.train_step_device(N, args.seed, w, idx=idx, want_loss=False)
        barrier(td, local)
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
        l0 = h.counter("launches")
        h.timer_start()
        for k in range(args.steps):
            h.train_step_device(N, args.seed, args.warmup + k, idx=idx, want_loss=False)
        ms = h.timer_stop()
        barrier(td, local)
        launches = h.counter("launches") - l0
        clk = clocks.stop() if rank == 0 else None
        ms = max_over_ranks(td, ms, local)
        value = args.steps * N * world / (ms * 1e-3)
        # e2e: host batches in, loss out, through ddpm_train_step
        import torch

        xb = torch.empty((N, 1024), dtype=torch.float32, pin_memory=True).uniform_(-1, 1)
        eb = torch.empty((N, 1024), dtype=torch.float32, pin_memory=True).normal_()
        tsb = np.random.default_rng(rank).integers(1, T_STEPS + 1, N).astype(np.int32)
        h.train_step(xb.numpy(), tsb, eb.numpy())
        barrier(td, local)
        t0 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        for k in range(e2e_steps):
            h.train_step(xb.numpy(), tsb, eb.numpy())
        barrier(td, local)
        e2e_s = max_over_ranks(td, time.perf_counter() - t0, local)
        e2e = {"value": e2e_steps * N * world / e2e_s, "unit": "img/s",
               "h2d_bytes_per_step": int(2 * N * 4096 + 4 * N), "d2h_bytes_per_step": 4, "steps": e2e_steps}
        metric = "train img/s (U-Net fwd+bwd+Adam, 32x32)"
        workload_desc = (f"train_step: q_sample + U-Net fwd/bwd + MSE + Adam, 32x32 synthetic U(-1,1) data, per-GPU batch {N}, "
                    f"global batch {N * world}, sync_bn={args.sync_bn}")
        flop_per_unit = FLOP_PER_TRAIN_IMG

    if rank != 0:
        return None

    # ---- roofline of the dominant kernel: the 64->64 3x3 convolution at 32x32 (5 of the 10 convs,
    #      51% of the U-Net FLOPs), timed alone with CUDA events on the engine's stream
    chunk = min(args.chunk, N)
    kern = {}
    for name in ("conv_l2", "conv_l4", "conv_l9", "conv_l3", "conv1", "pool", "up2", "reverse_update", "qsample", "mse", "adam",
                 "probe_fill", "probe_read"):
        if workload == "train" and args.workload == "both":
            break   # already measured in the sampling pass of this run
        try:
            kms, by, fl = h.time_kernel(name, chunk, 20)
            kern[name] = {"ms": kms, "tflops": fl / kms / 1e9 if fl else None, "gbs": by / kms / 1e6 if by else None}
        except Exception as ex:  # pragma: no cover
            kern[name] = {"error": str(ex)}
    dom = kern.get("conv_l2", {})
    uses_tc = h.counter("uses_tc") == 1
    peak_tf = peaks["bf16_tflops"]  # burst figure: kernel timed alone
    if not uses_tc and args.precision == "fp32":
        peak_note = "FP32 CUDA-core mode measured against the bf16 tensor peak"
    else:
        peak_note = "kind::f16 tensor peak (cuBLAS bf16 burst)"
    roofline = {"bound": "tensor", "kernel": "conv3x3 64->64 @32x32 (layer 2)", "achieved": dom.get("tflops"),
                "peak": peak_tf, "unit": "TFLOP/s", "frac": (dom.get("tflops") or 0.0) / peak_tf, "traffic": ncu_traffic("conv_l2", chunk),
                "peak_source": peaks["source"], "note": peak_note,
                "whole_step_tflops": value * flop_per_unit / 1e12 / world,
                "whole_step_frac_of_sustained": value * flop_per_unit / 1e12 / world / peaks["bf16_tflops_sustained"]}
    # largest HBM-side kernel of the default sampler path: the first conv (writes one 32x32x64 tensor per launch).  Peak =
    # the copy figure of MEASURED_PEAKS; the pure write / read stream rates of this box and run are listed beside it
    # (probe_fill / probe_read in the kernel table: a write-only kernel cannot exceed the former)
    c1 = kern.get("conv1", {})
    roofline_hbm = {"bound": "hbm", "kernel": "first conv of the sampler (tcgen05, writes 64 ch per pixel)", "achieved": c1.get("gbs"),
                    "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": (c1.get("gbs") or 0.0) / peaks["hbm_gbs"],
                    "stream_write_gbs": kern.get("probe_fill", {}).get("gbs"), "stream_read_gbs": kern.get("probe_read", {}).get("gbs"),
                    "frac_of_stream_write": ((c1.get("gbs") or 0.0) / kern["probe_fill"]["gbs"]) if kern.get("probe_fill", {}).get("gbs") else None}

    # ---- CPU baseline on the box's host cores, bounded sample
    cpu = None
    if world == 1 and not args.no_cpu and workload == "sample":
        v, dt, threads = cpu_reference_sampling(args.ref_images, args.ref_steps, os.cpu_count())
        cpu = {"value": v, "unit": "img/s", "cores": threads, "kind": "port",
               "sample": f"{args.ref_images} images x {args.ref_steps} of 499 reverse steps ({dt:.1f} s), extrapolated to 499; "
                         "Flux-semantics CPU restatement (oracle/ddpm_oracle.py, torch-CPU fp32)"}

    line = {
        "metric": metric, "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "fp16": "f16", "bf16": "bf16", "tf32": "tf32"}[args.precision], "data": "synthetic",
        "config": {"workload": workload_desc, "images_per_step_per_gpu": N, "T": T_STEPS, "chunk": args.chunk,
                   "precision": args.precision, "tensor_cores": bool(uses_tc),
                   "l2_policy": "per-step activation working set exceeds the 126 MB L2" if N * 0.4 > 126 else
                                "activations of one chunk are L2-resident by design (chunked sampler); inputs regenerated per step"},
        "roofline": roofline, "roofline_hbm": roofline_hbm, "kernels": kern,
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
    }
    return line


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, warnings) was diverted to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    # libraries (NCCL prints its version banner on stdout) must not pollute the one-line contract
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="both", choices=["sample", "train", "both"])
    ap.add_argument("--images", type=int, default=5200, help="sampled images per step per GPU (4 chunks of 1300)")
    ap.add_argument("--train-images", type=int, default=0,
                    help="training batch per step per GPU (default: BASELINE config 5, global batch 4096 => 4096/n_gpus)")
    ap.add_argument("--chunk", type=int, default=1300, help="images per captured reverse-loop graph")
    ap.add_argument("--streams", type=int, default=0, help="concurrent chunk streams of the sampler (0: library default)")
    ap.add_argument("--precision", default="fp16", choices=["fp32", "fp16", "bf16", "tf32"])
    ap.add_argument("--t-start", type=int, default=T_STEPS)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--sync-bn", type=int, default=1)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-images", type=int, default=65536,
                    help="images of the ONE ddpm_sample call the end-to-end leg times per step (BASELINE config 4; 0: --images)")
    ap.add_argument("--ref-budget", type=float, default=200.0, help="seconds the --impl reference run may take in total")
    ap.add_argument("--ref-images", type=int, default=128)
    ap.add_argument("--ref-steps", type=int, default=100)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
