#!/usr/bin/env python
"""Benchmark of the DDIM denoising hot path (BASELINE.json metric: motion-seconds generated per second,
50-step DDIM).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C1|C2|C3|C4|C5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one complete sampling loop over one batch of synthetic clips (C2: 50-step DDIM, 64 clips x 6 s =
180 frames x 26 keypoint coordinates per GPU, conditioned on 64-d music features per frame).  The batch is the
GLOBAL batch of the workload (C1-C3, C5: per-GPU batch x N ranks = weak scaling; C4: 512 clips in total = strong
scaling); it goes through the product's multi-GPU path, generate.sharded_sample: contiguous shards by
generate.shard_range, no data-path collective, one NCCL all_gather_into_tensor of the generated motion inside the
timed region.  Prints ONE JSON line on rank 0.
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
WORKLOADS = {  # name: (clips, T, S, sampler, scaling); clips are per GPU for "weak", in total for "strong"  (BASELINE.json configs)
    "C1": (1, 180, 25, "ddim", "weak"),
    "C2": (64, 180, 50, "ddim", "weak"),
    "C3": (32, 1800, 50, "ddim", "weak"),
    "C4": (512, 180, 50, "ddim", "strong"),
    "C5": (32, 1800, 1000, "ddpm", "weak"),
}
GOLDEN = {"C1": "c1.npz", "C2": "c2_pair.npz", "C3": "c3_clip.npz", "C4": "c2_pair.npz", "C5": "ddpm1000.npz"}


def workload_text(name, world):
    """The SAME string in both arms (the driver compares them)."""
    B, T, S, sampler, scaling = WORKLOADS[name]
    what = f"{S}-step {'DDIM' if sampler == 'ddim' else 'DDPM'}"
    per = (f"batch {B} x {T} frames (26 coords) per GPU" if scaling == "weak" else
           f"batch {B} x {T} frames (26 coords) in total, split over the GPUs")
    return f"{name}: {what}, {per}, 8-layer MotionTransformer D=128, random-init weights" + (", eta=0" if sampler == "ddim" else "")


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
# CPU arm: the oracle port of the reference path, timed on the host cores (bounded sample: WHOLE sampling loops)
# ------------------------------------------------------------------------------------------------
CPU_SAMPLE = {  # name: (clips of the batch that the CPU arm runs, denoise steps per loop it runs) -- whole loops wherever they fit
    "C1": (1, 25), "C2": (64, 50), "C3": (4, 50), "C4": (64, 50), "C5": (1, 100),
}


class CpuReference:
    """The reference algorithm (oracle port, torch fp32, all host threads) on a bounded sample of the workload."""

    def __init__(self, name):
        import torch

        from diffusion_conductor_b200.synth import synth_features, synth_inputs, synth_state_dict
        from oracle import motion_oracle as O

        self.torch, self.O, self.name = torch, O, name
        B, self.T, self.S, self.sampler, _ = WORKLOADS[name]
        self.clips, self.run_steps = CPU_SAMPLE[name]
        self.clips = min(self.clips, B)
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.sd = synth_state_dict(0, num_layers=8)
        self.xf_proj, self.xf_out = synth_features(self.clips, self.T, seed=1)
        _, self.noise = synth_inputs(self.clips, self.T, seed=1)
        self.tb = O.Tables(O.linear_betas(self.S))
        part = "" if self.clips == B else f" on {self.clips} of the {B} clips"
        steps = (f"whole {self.S}-step loops" if self.run_steps == self.S else
                 f"the first {self.run_steps} of {self.S} steps per loop, x{self.S // self.run_steps} extrapolated")
        self.sample = f"{steps}{part} ({self.clips} x {self.T} frames)"

    def loop(self, max_steps=None):
        """One sampling loop over the sample; returns seconds scaled to the full S steps."""
        torch = self.torch
        n = max_steps or self.run_steps
        nz = None
        if self.sampler == "ddpm":
            nz = torch.randn((n,) + tuple(self.noise.shape))
        t0 = time.perf_counter()
        self.O.sample_loop(self.sd, self.tb, self.noise, [self.T] * self.clips, self.xf_proj, self.xf_out, kind=self.sampler,
                           step_noise=nz, max_steps=n)
        return (time.perf_counter() - t0) * (self.S / n)

    def rate(self, seconds_per_loop):
        return (self.clips * self.T / 30.0) / seconds_per_loop


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
    calls = []
    for it in range(10):                                                              # iteration 0 is the warm-up; median of the other 9 calls
        torch.cuda.synchronize(dev)                                                   # (each call ends with a synchronize: per-call times, robust to one-off stalls of the host link)
        t0 = time.perf_counter()
        fp, fo = model.encode_music(hmel.to(dev, non_blocking=True), dev)            # what generate_music_motion does per rank
        out = diff.ddim_sample_loop(model, (B, T, 26), noise=noise_d, clip_denoised=False,
                                    model_kwargs=dict(xf_proj=fp, xf_out=fo, length=[T] * B))
        hout.copy_(out, non_blocking=True)
        torch.cuda.synchronize(dev)
        if it:
            calls.append(time.perf_counter() - t0)
    calls.sort()
    cond["e2e_from_mel_motion_s_per_s"] = round((B * T / 30.0) / calls[len(calls) // 2], 2)
    cond["e2e_from_mel_calls"] = len(calls)
    cond["mel_h2d_bytes"] = hmel.numel() * 4
    return cond


def parity_block(torch, np, model, name, dev):
    """Measured error of THIS build against the reference-generated fixture closest to the workload (tests/golden/, written by
    oracle/make_golden.py from the unmodified reference): reported next to `dtype` so that the precision of the timed path is
    on the same line as its speed."""
    from diffusion_conductor_b200 import GaussianDiffusion, _lib
    from diffusion_conductor_b200.gaussian_diffusion import LossType, ModelMeanType, ModelVarType, get_named_beta_schedule
    from diffusion_conductor_b200.generate import generate_music_motion
    from diffusion_conductor_b200.synth import synth_features, synth_inputs

    fx = GOLDEN[name]
    g = np.load(os.path.join(ROOT, "tests", "golden", fx))

    def diffusion(S):
        return GaussianDiffusion(betas=get_named_beta_schedule("linear", S), model_mean_type=ModelMeanType.START_X,
                                 model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE)

    if fx == "c1.npz":
        mel, noise = synth_inputs(1, 180, seed=0)
        got, ref, what = generate_music_motion(model, diffusion(25), mel, 26, noise=noise.to(dev)), g["final"], "25-step DDIM, 1 x 180 frames, music encoder included"
    elif fx == "c3_clip.npz":
        mel, noise = synth_inputs(1, 1800, seed=3)
        got, ref, what = generate_music_motion(model, diffusion(50), mel, 26, noise=noise.to(dev)), g["final"], "50-step DDIM, one 1800-frame clip, music encoder included"
    elif fx == "c2_pair.npz":
        xf_proj, xf_out = synth_features(2, 180, seed=21)
        _, noise = synth_inputs(2, 180, seed=21)
        kw = dict(xf_proj=xf_proj.to(dev), xf_out=xf_out.to(dev), length=[180, 180])
        got = diffusion(50).ddim_sample_loop(model, noise.shape, noise=noise.to(dev), clip_denoised=False, model_kwargs=kw)
        ref, what = g["final"], "50-step DDIM, 2 x 180 frames"
    else:
        xf_proj, xf_out = synth_features(1, 180, seed=22)
        _, noise = synth_inputs(1, 180, seed=22)
        torch.manual_seed(int(g["seed"]))          # the reference's noise stream: one CPU randn_like per step
        nz = torch.stack([torch.randn(1, 180, 26) for _ in range(1000)]).to(dev)
        d = diffusion(1000)
        eng = d._bind(model, noise.to(dev), dict(xf_proj=xf_proj.to(dev), xf_out=xf_out.to(dev), length=[180]))
        got = noise.to(dev).clone()
        eng.sample_loop(_lib.DC_SAMPLER_DDPM, got, step_noise=nz)
        ref, what = g["ddpm_sample"][list(g["steps"]).index(0)], "1000-step DDPM on the reference's noise stream, 1 x 180 frames"
    a, b = got.detach().float().cpu().numpy(), np.asarray(ref)
    rms = float(np.sqrt((b ** 2).mean()))
    return {"rel_rms": float(f"{np.sqrt(((a - b) ** 2).mean()) / rms:.3e}"), "max_abs": float(f"{np.abs(a - b).max():.3e}"),
            "ref_rms": round(rms, 3), "vs": f"tests/golden/{fx}: final keypoints of the unmodified reference (CPU fp32), {what}"}


def run_reference(args, rank):
    if rank != 0:
        return
    name = args.workload
    B, T, S, sampler, scaling = WORKLOADS[name]
    ref = CpuReference(name)
    ref.loop(max_steps=1)                                   # page-in / thread-pool warm-up
    for _ in range(max(0, min(args.warmup, 1))):            # one whole warm-up loop is enough on the CPU; the budget goes to timed loops
        ref.loop()
    secs = [ref.loop() for _ in range(args.steps)]
    dt = sum(secs) / len(secs)
    rate = ref.rate(dt)
    line = {
        "impl": "reference", "metric": "motion-seconds generated/sec (50-step DDIM)", "value": round(rate, 3),
        "unit": "motion-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt * 1e3, 2), "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(name, args.gpus),
                   "where": "host CPU, torch fp32, all host threads: oracle port of the reference path (the reference tree itself does "
                            "not travel to the GPU box); one step = one sampling loop over the sample below, rate = its motion-seconds / its time",
                   "sample": ref.sample},
        "cpu_baseline": {"value": round(rate, 3), "unit": "motion-s/s", "cores": ref.cores, "kind": "port", "sample": ref.sample},
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
    ap.add_argument("--no-conditioning", action="store_true", help="skip the once-per-clip conditioning block (music encoder timing) and the "
                    "parity block; used for the ncu launch list so that it covers the sampling loop only")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    from diffusion_conductor_b200 import GaussianDiffusion, MotionTransformer, _lib
    from diffusion_conductor_b200.gaussian_diffusion import LossType, ModelMeanType, ModelVarType, get_named_beta_schedule
    from diffusion_conductor_b200.generate import gather_shards, shard_range, sharded_sample
    from diffusion_conductor_b200.synth import synth_features, synth_inputs, synth_state_dict

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(3, args.warmup)
    name = args.workload
    Bw, T, S, sampler, scaling = WORKLOADS[name]
    Bg = Bw * world if scaling == "weak" else Bw               # global batch of the job
    lo, hi = shard_range(Bg, rank, world)
    B = hi - lo                                                # this rank's clips
    ddpm = sampler == "ddpm"
    flags = _lib.DC_SAMPLER_DDPM if ddpm else _lib.DC_SAMPLER_DDIM

    model = MotionTransformer(26, num_frames=1800, num_layers=8, latent_dim=128, device=dev, music_model_path=None,
                              operand_dtype=args.operand)
    model.load_state_dict(synth_state_dict(0, num_layers=8), strict=True)
    model = model.to(dev).eval()
    diff = GaussianDiffusion(betas=get_named_beta_schedule("linear", S), model_mean_type=ModelMeanType.START_X,
                             model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE)
    # the GLOBAL batch (same on every rank: synthetic, seeded), resident in HBM; each rank samples its shard_range slice
    xf_proj_g, xf_out_g = (t.to(dev) for t in synth_features(Bg, T, seed=100))
    noise_g = synth_inputs(Bg, T, seed=100)[1].to(dev)
    length_g = [T] * Bg
    xf_proj_d, xf_out_d, noise_d = xf_proj_g[lo:hi], xf_out_g[lo:hi], noise_g[lo:hi]
    eng = model.engine(dev)                      # base handle: C-ABI host-buffer call and the per-kernel profile
    plan = model.engine_for(dev, B, T)           # what the sampling loops use
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def sample_fn(fp, fo, nz, ln):               # one rank's shard through the public sampling-loop API
        kw = dict(xf_proj=fp, xf_out=fo, length=ln)
        if ddpm:
            return diff.p_sample_loop(model, tuple(nz.shape), noise=nz, clip_denoised=False, model_kwargs=kw)
        return diff.ddim_sample_loop(model, tuple(nz.shape), noise=nz, clip_denoised=False, model_kwargs=kw)

    def one_loop():                              # the product's multi-GPU path: shard, sample, ONE all_gather_into_tensor
        return sharded_sample(sample_fn, xf_proj_g, xf_out_g, noise_g, length_g)

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
    per_rank_ms = [round(total_ms / args.steps, 4)]
    if world > 1:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank_ms = [round(float(v.item()) / args.steps, 4) for v in allt]       # every rank's own loop time (the line reports the max)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    motion_seconds = Bg * T / 30.0
    value = motion_seconds / (ms_per_step / 1e3)

    # ---- end to end with HOST buffers (pinned): H2D features + noise, conditioning precompute, the loop, D2H of the generated
    # motion -- every step.  DDIM: one C-ABI call (dc_generate_host); DDPM: the public p_sample_loop fed from pinned host tensors
    hp, ho, hn = xf_proj_d.cpu().pin_memory(), xf_out_d.cpu().pin_memory(), noise_d.cpu().pin_memory()
    hout = torch.empty(B, T, 26).pin_memory()

    def one_e2e():
        if ddpm:
            out = sample_fn(hp.to(dev, non_blocking=True), ho.to(dev, non_blocking=True), hn.to(dev, non_blocking=True), [T] * B)
            hout.copy_(out, non_blocking=True)
            torch.cuda.synchronize(dev)
        else:
            eng.generate_host(flags, hp, ho, [T] * B, hn, hout, B, T)
        if world > 1:
            gather_shards(hout.to(dev, non_blocking=True), Bg)
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
    cond, parity = None, None
    if not args.no_conditioning:
        if not ddpm:
            cond = conditioning_block(torch, model, diff, eng, synth_inputs, B, T, rank, dev, noise_d, hout, time)
        if rank == 0:
            parity = parity_block(torch, np, model, name, dev)

    # ---- roofline of the dominant kernel, timed live with CUDA events on the launching (= torch current) stream
    pk = peaks()
    diff._bind(model, noise_d, dict(xf_proj=xf_proj_d, xf_out=xf_out_d, length=[T] * B))      # the timed loops' schedule / conditioning
    x = noise_d.clone()
    agg, cnt = {}, {}
    reps = 5
    for i in range(reps + 2):
        ms, c = plan.profile_step(flags, x, S - 1 - (i % S))
        if i >= 2:
            for k in ms:
                agg[k] = agg.get(k, 0.0) + ms[k] / reps
                cnt[k] = c[k]
    step_ms = sum(agg.values())
    persistent = cnt["layer"] == 1
    tp = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    tj = json.load(open(tp)) if os.path.exists(tp) else {}
    if persistent:
        # One launch of the cluster-per-clip kernel runs a whole block of steps (DDIM: all S; DDPM: the steps whose pre-drawn
        # noise fits the bounded buffer).  Algorithmic FLOPs per launch = clips * T * steps * 8,500,224.
        n_launch_steps = S if not ddpm else max(1, min(S, diff.NOISE_BLOCK_BYTES // max(1, x.numel() * 4)))
        nz = torch.randn((n_launch_steps,) + tuple(x.shape), device=dev) if ddpm else None
        ms = []
        for _ in range(5):
            xl = noise_d.clone()
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.sample_range(flags, xl, S - 1, n_launch_steps, step_noise=nz)
            e1.record()
            torch.cuda.synchronize(dev)
            ms.append(e0.elapsed_time(e1))
        launch_ms = sum(ms) / len(ms)
        flop_per_launch = float(B) * T * n_launch_steps * FLOP_PER_TOKEN_STEP
        kname = "dc::clip_kernel (persistent; one thread-block cluster per clip; one launch = a whole block of denoise steps)"
        launches_per_loop = -(-S // n_launch_steps)
        share = round(min(1.0, launch_ms * (S / n_launch_steps) / ms_per_step), 3)
    else:
        launch_ms = agg["layer"] / max(cnt["layer"], 1)
        flop_per_launch = 2.0 * LAYER_MAC_PER_TOKEN * 8 * B * T / max(cnt["layer"], 1)
        kname = "dc::layer_kernel"
        launches_per_loop = cnt["layer"] * S
        share = round(agg["layer"] / step_ms, 3)
    achieved = flop_per_launch / (launch_ms * 1e-3) / 1e12
    traffic = tj.get(kname.split(" ")[0], {}).get(name, {}).get("dram_bytes_per_launch")
    roofline = {"bound": "tensor", "kernel": kname, "achieved": round(achieved, 2), "peak": pk["bf16_tflops"],
                "unit": "TFLOP/s", "frac": round(achieved / pk["bf16_tflops"], 4), "traffic": traffic,
                "peak_source": pk["source"], "frac_of_sustained_peak": round(achieved / pk["bf16_tflops_sustained"], 4) if pk.get("bf16_tflops_sustained") else None,
                "launch_ms": round(launch_ms, 4), "flop_per_launch": flop_per_launch,
                "launches_per_loop": launches_per_loop, "share_of_loop": share,
                "single_denoise_step_ms_by_kernel": {k: round(v, 4) for k, v in agg.items()}, "profiled_clips": B,
                "whole_loop_frac_of_peak": round(B * T * S * FLOP_PER_TOKEN_STEP / (ms_per_step / 1e3) / 1e12 / pk["bf16_tflops"], 4)}

    if rank == 0:
        cpu, eager = None, None
        if not args.no_cpu_baseline and world == 1:
            ref = CpuReference(name)
            ref.loop(max_steps=1)
            n_loops = 2 if name in ("C1", "C2", "C4") else 1
            dt = sum(ref.loop() for _ in range(n_loops)) / n_loops
            cpu = {"value": round(ref.rate(dt), 3), "unit": "motion-s/s", "cores": ref.cores, "kind": "port",
                   "sample": f"{n_loops} x {ref.sample} after a 1-step warm-up"}
            try:
                grate, gsample = eager_gpu_rate(min(B, 64), T, S, dev)
                eager = {"value": round(grate, 2), "unit": "motion-s/s", "kind": "port, PyTorch eager fp32 on the same B200", "sample": gsample}
            except Exception as exc:                      # a baseline leg must never take the product line down
                eager = {"unavailable": repr(exc)[:200]}
        n_in = (hp.numel() + ho.numel() + hn.numel()) * 4
        line = {
            "metric": "motion-seconds generated/sec (50-step DDIM)", "value": round(value, 2), "unit": "motion-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": round(ms_per_step, 4), "ms_per_step_by_rank": per_rank_ms,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": args.operand, "parity": parity, "data": "synthetic",
            "config": {"workload": workload_text(name, world), "global_batch": Bg, "clips_on_rank0": B,
                       "token_steps_per_step": Bg * T * S, "l2": "256 MiB buffer written between timed iterations",
                       "path": "generate.sharded_sample -> " + ("p_sample_loop" if ddpm else "ddim_sample_loop") + " on this rank's shard_range slice",
                       "collective": "NCCL all_gather_into_tensor of the generated motion (generate.gather_shards)" if world > 1 else "none"},
            "e2e": {"value": round(e2e_value, 2), "unit": "motion-s/s", "h2d_bytes_per_step": n_in, "d2h_bytes_per_step": hout.numel() * 4},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "torch_eager_gpu_baseline": eager, "clocks": clk, "conditioning": cond,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
