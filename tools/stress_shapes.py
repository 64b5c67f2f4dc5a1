"""Odd shapes and oversubscription of the cluster-per-clip kernel: every case must finish and stay finite (run under gpurun)."""
import sys, time, torch
sys.path.insert(0, '.')
from diffusion_conductor_b200 import GaussianDiffusion, MotionTransformer
from diffusion_conductor_b200.gaussian_diffusion import LossType, ModelMeanType, ModelVarType, get_named_beta_schedule
from diffusion_conductor_b200.synth import synth_features, synth_inputs, synth_state_dict
dev = torch.device("cuda", 0)
m = MotionTransformer(26, num_frames=1800, num_layers=8, latent_dim=128, device=dev, music_model_path=None)
m.load_state_dict(synth_state_dict(0, num_layers=8), strict=True); m = m.to(dev).eval()
d = GaussianDiffusion(betas=get_named_beta_schedule("linear", 50), model_mean_type=ModelMeanType.START_X, model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE)
for (B, T) in [(1000, 180), (40, 1800), (333, 300), (777, 64), (100, 700), (5, 1283), (148, 128), (149, 129)]:
    xf_proj, xf_out = synth_features(B, T, seed=B); _, noise = synth_inputs(B, T, seed=B)
    length = [T - (i % 7) for i in range(B)]
    kw = dict(xf_proj=xf_proj.to(dev), xf_out=xf_out.to(dev), length=length)
    for rep in range(2):                      # the first call allocates the workspace for this shape
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = d.ddim_sample_loop(m, (B, T, 26), noise=noise.to(dev), clip_denoised=False, model_kwargs=kw)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"B={B} T={T} tiles/clip={-(-T//128)}: {dt*1e3:.1f} ms, finite={bool(torch.isfinite(out).all())}, {B*T/30/dt:.0f} motion-s/s", flush=True)
