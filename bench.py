#!/usr/bin/env python
"""Benchmark of the DDIM denoising hot path (BASELINE.json metric: motion-seconds generated per second,
50-step DDIM).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2|C3|C1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one complete 50-step DDIM sampling loop over one batch of synthetic clips (C2: 64 clips x
6 s = 180 frames x 26 keypoint coordinates, conditioned on 64-d music features per frame).  Every rank
generates its own batch (weak scaling, no data-path collective); for N > 1 the generated motion is
all-gathered over NCCL inside the timed region.  Prints ONE JSON line on rank 0.
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOP_PER_TOKEN_STEP = 8_500_224           # BASELINE.md §3 (algorithmic, step-invariant work excluded)
LAYER_MAC_PER_TOKEN = 200_704 + 165_888 + 163_840
WORKLOADS = {  # name: (B, T, S, sampler)
    "C1": (1, 180, 25, "ddim"),
    "C2": (64, 180, 50, "ddim"),
    "C3": (32, 1800, 50, "ddim"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path, timed on the host cores (bounded sample)
# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(B, T, S, n_steps_sample, batch_sample=None):
    """motion-s/s of the reference algorithm on this host: times `n_steps_sample` denoise steps (after one
    warm-up step) of a `batch_sample`-clip batch with all host threads and extrapolates to the S-step loop."""
    import torch

    from diffusion_conductor_b200.synth import synth_features, synth_inputs, synth_state_dict
    from oracle import motion_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs = batch_sample or B
    sd = synth_state_dict(0, num_layers=8)
    xf_proj, xf_out = synth_features(Bs, T, seed=1)
    _, noise = synth_inputs(Bs, T, seed=1)
    tb = O.Tables(O.linear_betas(S))
    O.sample_loop(sd, tb, noise, [T] * Bs, xf_proj, xf_out, max_steps=1)
    t0 = time.perf_counter()
    O.sample_loop(sd, tb, noise, [T] * Bs, xf_proj, xf_out, max_steps=n_steps_sample)
    dt = (time.perf_counter() - t0) / n_steps_sample
    rate = (Bs * T / 30.0) / (dt * S)
    return rate, cores, dt, f"{n_steps_sample} denoise steps of a {Bs}x{T}-frame batch after 1 warm-up step, x{S} extrapolated"


def eager_gpu_rate(B, T, S, dev, n_calls=3):
    """the reference algorithm in PyTorch eager fp32 ON THE SAME GPU (the oracle restatement of
    MotionTransformer.forward with every tensor on the device; the sampler update is negligible next to it) -- the
    "existing GPU kernels" bar.  Times n_calls forward passes after one warm-up and extrapolates to the S-step loop."""
    import torch

    from diffusion_conductor_b200.synth import synth_features, synth_inputs, synth_state_dict
    from oracle import motion_oracle as O

    sd = {k: v.to(dev) for k, v in synth_state_dict(0, num_layers=8).items()}
    xf_proj, xf_out = (t.to(dev) for t in synth_features(B, T, seed=1))
    x = synth_inputs(B, T, seed=1)[1].to(dev)
    t = torch.full((B,), S - 1, dtype=torch.long, device=dev)
    with torch.no_grad():
        O.motion_transformer_forward(sd, x, t, [T] * B, xf_proj, xf_out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_calls):
            O.motion_transformer_forward(sd, x, t, [T] * B, xf_proj, xf_out)
        e1.record()
        torch.cuda.synchronize(dev)
    dt = e0.elapsed_time(e1) / 1e3 / n_calls
    return (B * T / 30.0) / (dt * S), f"{n_calls} eager forward passes of a {B}x{T}-frame batch after 1 warm-up, x{S} extrapolated"


def conditioning_block(torch, model, diff, eng, synth_inputs, B, T, rank, dev, noise_d, hout, time):
    """the once-per-clip conditioning -- music encoder (mel -> features) and the step-invariant precompute, timed
    with CUDA events -- and the whole generate_music_motion-equivalent call from a pinned host mel to a host motion array."""
    mel, _ = synth_inputs(B, T, seed=100 + rank)
    hmel = mel.pin_memory()
    mel_d = mel.to(dev)
    cond = {}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for i in range(3):
        ev[0].record()
        fp, fo = model.encode_music(mel_d, dev)
        ev[1].record()
        eng.prepare(fp, fo, [T] * B, B, T)
        ev[2].record()
        torch.cuda.synchronize(dev)
        cond = {"encode_music_ms": round(ev[0].elapsed_time(ev[1]), 3), "prepare_cond_ms": round(ev[1].elapsed_time(ev[2]), 3)}
    t0 = 0.0
    for it in range(6):
        if it == 1:                                                                   # iteration 0 is the warm-up
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
        fp, fo = model.encode_music(hmel.to(dev, non_blocking=True), dev)            # what generate_music_motion does per rank
        out = diff.ddim_sample_loop(model, (B, T, 26), noise=noise_d, clip_denoised=False,
                                    model_kwargs=dict(xf_proj=fp, xf_out=fo, length=[T] * B))
        hout.copy_(out, non_blocking=True)
        torch.cuda.synchronize(dev)
    cond["e2e_from_mel_motion_s_per_s"] = round((B * T / 30.0) / ((time.perf_counter() - t0) / 5), 2)
    cond["mel_h2d_bytes"] = hmel.numel() * 4
    return cond


def run_reference(args, rank):
    if rank != 0:
        return
    B, T, S, _ = WORKLOADS[args.workload]
    vals = []
    for _ in range(max(1, args.warmup) - 1):
        cpu_reference_rate(B, T, S, 1)
    for _ in range(args.steps):
        rate, cores, dt, sample = cpu_reference_rate(B, T, S, 2)
        vals.append((rate, dt))
    rate = sum(v[0] for v in vals) / len(vals)
    dt = sum(v[1] for v in vals) / len(vals)
    line = {
        "impl": "reference", "metric": "motion-seconds generated/sec (50-step DDIM)", "value": round(rate, 3),
        "unit": "motion-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt * S * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {S}-step DDIM, batch {B} x {T} frames (26 coords), 8-layer MotionTransformer D=128",
                   "where": "host CPU, torch fp32, oracle port of the reference path (the reference tree itself does not travel to the GPU box)"},
        "cpu_baseline": {"value": round(rate, 3), "unit": "motion-s/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(rate, 3), "unit": "motion-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--operand", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-conditioning", action="store_true", help="skip the once-per-clip conditioning block (music encoder timing); "
                    "used for the ncu launch list so that it covers the sampling loop only")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    from diffusion_conductor_b200 import GaussianDiffusion, MotionTransformer, _lib
    from diffusion_conductor_b200.gaussian_diffusion import LossType, ModelMeanType, ModelVarType, get_named_beta_schedule
    from diffusion_conductor_b200.synth import synth_features, synth_inputs, synth_state_dict

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(3, args.warmup)
    B, T, S, _ = WORKLOADS[args.workload]

    model = MotionTransformer(26, num_frames=1800, num_layers=8, latent_dim=128, device=dev, music_model_path=None,
                              operand_dtype=args.operand)
    model.load_state_dict(synth_state_dict(0, num_layers=8), strict=True)
    model = model.to(dev).eval()
    diff = GaussianDiffusion(betas=get_named_beta_schedule("linear", S), model_mean_type=ModelMeanType.START_X,
                             model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE)
    xf_proj, xf_out = synth_features(B, T, seed=100 + rank)
    _, noise = synth_inputs(B, T, seed=100 + rank)
    xf_proj_d, xf_out_d, noise_d = xf_proj.to(dev), xf_out.to(dev), noise.to(dev)
    kw = dict(xf_proj=xf_proj_d, xf_out=xf_out_d, length=[T] * B)
    eng = model.engine(dev)                      # base handle: C-ABI host-buffer call and the per-kernel profile
    plan = model.engine_for(dev, B, T)           # what ddim_sample_loop uses (chunks of clips when the batch exceeds the SMs)
    gathered = [torch.empty(B, T, 26, device=dev) for _ in range(world)] if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def one_loop():
        out = diff.ddim_sample_loop(model, (B, T, 26), noise=noise_d, clip_denoised=False, model_kwargs=kw)
        if world > 1:
            dist.all_gather(gathered, out)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(warmup):
        one_loop()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = plan.kernel_launches()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for e0, e1 in evs:
        flush.fill_(1)
        e0.record()
        one_loop()
        e1.record()
    barrier()
    launches = plan.kernel_launches() - launches0
    total_ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    motion_seconds = world * B * T / 30.0
    value = motion_seconds / (ms_per_step / 1e3)

    # ---- end to end through the C ABI with HOST buffers (pinned): H2D features + noise, conditioning
    # precompute, the loop, D2H of the generated motion -- every step
    hp, ho, hn = xf_proj.pin_memory(), xf_out.pin_memory(), noise.pin_memory()
    hout = torch.empty(B, T, 26).pin_memory()
    flags = _lib.DC_SAMPLER_DDIM

    def one_e2e():
        eng.generate_host(flags, hp, ho, [T] * B, hn, hout, B, T)
        if world > 1:
            dist.all_gather(gathered, hout.to(dev, non_blocking=True))
            torch.cuda.synchronize(dev)

    for _ in range(warmup):
        one_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = motion_seconds / (float(t.item()) / args.steps)
    clk = clocks.stop() if rank == 0 else None

    # ---- the once-per-clip conditioning: music encoder (mel -> features) and the step-invariant precompute, timed
    # with CUDA events; and the whole generate_music_motion-equivalent call from a pinned host mel to a host motion array
    cond = None
    if not args.no_conditioning:
        cond = conditioning_block(torch, model, diff, eng, synth_inputs, B, T, rank, dev, noise_d, hout, time)

    # ---- roofline of the dominant kernel, timed live with CUDA events
    pk = peaks()
    chunked = hasattr(plan, "bounds")
    Bp = (plan.bounds[0][1] - plan.bounds[0][0]) if chunked else B      # the profile runs on one chunk of clips
    x = noise_d[:Bp].clone()
    eng.prepare(xf_proj_d[:Bp], xf_out_d[:Bp], [T] * Bp, Bp, T)
    agg, cnt = {}, {}
    reps = 5
    for i in range(reps + 2):
        ms, c = eng.profile_step(_lib.DC_SAMPLER_DDIM, x, S - 1 - (i % S))
        if i >= 2:
            for k in ms:
                agg[k] = agg.get(k, 0.0) + ms[k] / reps
                cnt[k] = c[k]
    step_ms = sum(agg.values())
    persistent = cnt["layer"] == 1
    tp = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    tj = json.load(open(tp)) if os.path.exists(tp) else {}
    if persistent:
        # the whole sampling loop is ONE launch of the cluster-per-clip kernel: time that launch with CUDA events on the
        # launching (= torch current) stream; algorithmic FLOPs per launch = B * T * S * 8,500,224
        ms = []
        for _ in range(5):
            xl = noise_d[:Bp].clone()
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.sample_loop(_lib.DC_SAMPLER_DDIM, xl)
            e1.record()
            torch.cuda.synchronize(dev)
            ms.append(e0.elapsed_time(e1))
        launch_ms = sum(ms) / len(ms)
        flop_per_launch = float(Bp) * T * S * FLOP_PER_TOKEN_STEP
        kname = "dc::clip_kernel (persistent; one thread-block cluster per clip; the whole sampling loop is one launch)"
        launches_per_loop = 1
        share = min(1.0, launch_ms / ms_per_step) if Bp == B else None
    else:
        launch_ms = agg["layer"] / max(cnt["layer"], 1)
        flop_per_launch = 2.0 * LAYER_MAC_PER_TOKEN * 8 * Bp * T / max(cnt["layer"], 1)
        kname = "dc::layer_kernel"
        launches_per_loop = cnt["layer"] * S
        share = round(agg["layer"] / step_ms, 3)
    achieved = flop_per_launch / (launch_ms * 1e-3) / 1e12
    traffic = tj.get(kname.split(" ")[0], {}).get(args.workload, {}).get("dram_bytes_per_launch")
    roofline = {"bound": "tensor", "kernel": kname, "achieved": round(achieved, 2), "peak": pk["bf16_tflops"],
                "unit": "TFLOP/s", "frac": round(achieved / pk["bf16_tflops"], 4), "traffic": traffic,
                "peak_source": pk["source"], "frac_of_sustained_peak": round(achieved / pk["bf16_tflops_sustained"], 4) if pk.get("bf16_tflops_sustained") else None,
                "launch_ms": round(launch_ms, 4), "flop_per_launch": flop_per_launch,
                "launches_per_loop": launches_per_loop, "share_of_loop": share if share is None else round(share, 3),
                "single_denoise_step_ms_by_kernel": {k: round(v, 4) for k, v in agg.items()}, "profiled_clips": Bp,
                "clip_chunks": len(plan.bounds) if chunked else 1,
                "whole_loop_frac_of_peak": round(B * T * S * FLOP_PER_TOKEN_STEP / (ms_per_step / 1e3) / 1e12 / pk["bf16_tflops"], 4)}

    if rank == 0:
        cpu, eager = None, None
        if not args.no_cpu_baseline and world == 1:
            rate, cores, dt, sample = cpu_reference_rate(B, T, S, 3)
            cpu = {"value": round(rate, 3), "unit": "motion-s/s", "cores": cores, "kind": "port", "sample": sample}
            try:
                grate, gsample = eager_gpu_rate(B, T, S, dev)
                eager = {"value": round(grate, 2), "unit": "motion-s/s", "kind": "port, PyTorch eager fp32 on the same B200", "sample": gsample}
            except Exception as exc:                      # a baseline leg must never take the product line down
                eager = {"unavailable": repr(exc)[:200]}
        n_in = (hp.numel() + ho.numel() + hn.numel()) * 4
        line = {
            "metric": "motion-seconds generated/sec (50-step DDIM)", "value": round(value, 2), "unit": "motion-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.operand, "data": "synthetic",
            "config": {"workload": f"{args.workload}: {S}-step DDIM, batch {B} x {T} frames (26 coords) per GPU, 8-layer "
                                   f"MotionTransformer D=128, random-init weights, eta=0",
                       "token_steps_per_step": B * T * S, "l2": "256 MiB buffer written between timed iterations",
                       "collective": "NCCL all_gather of the generated motion" if world > 1 else "none"},
            "e2e": {"value": round(e2e_value, 2), "unit": "motion-s/s", "h2d_bytes_per_step": n_in, "d2h_bytes_per_step": hout.numel() * 4},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "torch_eager_gpu_baseline": eager, "clocks": clk, "conditioning": cond,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
