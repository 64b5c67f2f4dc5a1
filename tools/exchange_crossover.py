"""Clusters (distributed shared memory) vs independent CTAs (L2) for the per-clip exchange, by tiles per clip (run under gpurun).
usage: DC_GX=0|1 python tools/exchange_crossover.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusion_conductor_b200 import GaussianDiffusion, MotionTransformer  # noqa: E402
from diffusion_conductor_b200.gaussian_diffusion import LossType, ModelMeanType, ModelVarType, get_named_beta_schedule  # noqa: E402
from diffusion_conductor_b200.synth import synth_features, synth_inputs, synth_state_dict  # noqa: E402

dev = torch.device("cuda", 0)
m = MotionTransformer(26, num_frames=2048, num_layers=8, latent_dim=128, device=dev, music_model_path=None)
m.load_state_dict(synth_state_dict(0, num_layers=8, num_frames=2048), strict=True)
m = m.to(dev).eval()
d = GaussianDiffusion(betas=get_named_beta_schedule("linear", 50), model_mean_type=ModelMeanType.START_X,
                      model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE)
for (B, T) in [(296, 256), (150, 384), (120, 512), (120, 640), (100, 768), (100, 896), (80, 1024), (64, 1280), (48, 1536), (40, 2048)]:
    xf_proj, xf_out = synth_features(B, T, seed=B)
    _, noise = synth_inputs(B, T, seed=B)
    kw = dict(xf_proj=xf_proj.to(dev), xf_out=xf_out.to(dev), length=[T] * B)
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d.ddim_sample_loop(m, (B, T, 26), noise=noise.to(dev), clip_denoised=False, model_kwargs=kw)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(f"DC_GX={os.environ.get('DC_GX', 'default')} B={B} T={T} tiles/clip={-(-T // 128)}: {best * 1e3:.1f} ms, {B * T / 30 / best:.0f} motion-s/s", flush=True)
