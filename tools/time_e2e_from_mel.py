import time, torch, sys
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
def sync(): torch.cuda.synchronize(dev)
for it in range(4):
    sync(); t0 = time.perf_counter()
    md = hmel.to(dev, non_blocking=True); sync(); t1 = time.perf_counter()
    fp, fo = model.encode_music(md, dev); sync(); t2 = time.perf_counter()
    out = diff.ddim_sample_loop(model, (B, T, 26), noise=noise_d, clip_denoised=False, model_kwargs=dict(xf_proj=fp, xf_out=fo, length=[T] * B)); sync(); t3 = time.perf_counter()
    hout.copy_(out, non_blocking=True); sync(); t4 = time.perf_counter()
    print(f"iter {it}: h2d {1e3*(t1-t0):.2f} encode {1e3*(t2-t1):.2f} loop {1e3*(t3-t2):.2f} d2h {1e3*(t4-t3):.2f} ms")
