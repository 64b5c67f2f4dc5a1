"""BASELINE.json configs[4] per-GPU shard: 1000-step DDPM, 32 clips x 1800 frames (256 clips over 8 GPUs), noise stream
from a seeded generator.  Times p_sample_loop through the public API with CUDA events (run under gpurun)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusion_conductor_b200 import GaussianDiffusion, MotionTransformer  # noqa: E402
from diffusion_conductor_b200.gaussian_diffusion import LossType, ModelMeanType, ModelVarType, get_named_beta_schedule  # noqa: E402
from diffusion_conductor_b200.synth import synth_features, synth_inputs, synth_state_dict  # noqa: E402

B, T, S = 32, 1800, 1000
dev = torch.device("cuda", 0)
m = MotionTransformer(26, num_frames=1800, num_layers=8, latent_dim=128, device=dev, music_model_path=None)
m.load_state_dict(synth_state_dict(0, num_layers=8), strict=True)
m = m.to(dev).eval()
d = GaussianDiffusion(betas=get_named_beta_schedule("linear", S), model_mean_type=ModelMeanType.START_X,
                      model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE)
xf_proj, xf_out = synth_features(B, T, seed=7)
_, noise = synth_inputs(B, T, seed=7)
kw = dict(xf_proj=xf_proj.to(dev), xf_out=xf_out.to(dev), length=[T] * B)
times = []
for it in range(3):
    torch.manual_seed(1234)                      # the per-step randn_like stream of the reference loop
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = d.p_sample_loop(m, (B, T, 26), noise=noise.to(dev), clip_denoised=False, model_kwargs=kw)
    e1.record()
    torch.cuda.synchronize(dev)
    times.append(e0.elapsed_time(e1))
assert torch.isfinite(out).all()
ms = min(times[1:])
print(json.dumps({"workload": "C5 shard: 1000-step DDPM, 32 clips x 1800 frames on one B200 (256 clips over 8 GPUs)",
                  "loop_ms": round(ms, 1), "motion_s_per_s": round(B * T / 30.0 / (ms / 1e3), 1),
                  "token_steps_per_s": round(B * T * S / (ms / 1e3) / 1e6, 2), "unit2": "M token-steps/s",
                  "frac_of_bf16_peak": round(B * T * S * 8500224 / (ms / 1e3) / 1e12 / 1689.9, 4),
                  "noise_stream_bytes": S * B * T * 26 * 4}))
