"""mel (pinned host) -> motion (pinned host) without phase syncs, as bench.py's conditioning block times it (run under gpurun)."""
import sys, time, torch
sys.path.insert(0, '.')
from diffusion_conductor_b200 import GaussianDiffusion, MotionTransformer
from diffusion_conductor_b200.gaussian_diffusion import LossType, ModelMeanType, ModelVarType, get_named_beta_schedule
from diffusion_conductor_b200.synth import synth_inputs, synth_state_dict
dev = torch.device("cuda", 0)
B, T, S = 64, 180, 50
model = MotionTransformer(26, num_frames=1800, num_layers=8, latent_dim=128, device=dev, music_model_path=None)
model.load_state_dict(synth_state_dict(0, num_layers=8), strict=True)
model = model.to(dev).eval()
diff = GaussianDiffusion(betas=get_named_beta_schedule("linear", S), model_mean_type=ModelMeanType.START_X,
                         model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE)
mel, noise = synth_inputs(B, T, seed=1)
hmel = mel.pin_memory(); noise_d = noise.to(dev); hout = torch.empty(B, T, 26).pin_memory()
for rep in range(3):
    for it in range(8):
        if it == 2:
            torch.cuda.synchronize(dev); t0 = time.perf_counter()
        fp, fo = model.encode_music(hmel.to(dev, non_blocking=True), dev)
        out = diff.ddim_sample_loop(model, (B, T, 26), noise=noise_d, clip_denoised=False, model_kwargs=dict(xf_proj=fp, xf_out=fo, length=[T] * B))
        hout.copy_(out, non_blocking=True)
        torch.cuda.synchronize(dev)
    ms = 1e3 * (time.perf_counter() - t0) / 6
    print(f"mel -> motion {ms:.3f} ms per call = {B * T / 30 / ms * 1e3:.0f} motion-s/s")
