"""2+ GPUs under torchrun: event-time the sampling loop alone, the gather alone, and both (where does the N > 1 overhead go?)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusion_conductor_b200 import GaussianDiffusion, MotionTransformer  # noqa: E402
from diffusion_conductor_b200.gaussian_diffusion import LossType, ModelMeanType, ModelVarType, get_named_beta_schedule  # noqa: E402
from diffusion_conductor_b200.generate import gather_shards  # noqa: E402
from diffusion_conductor_b200.synth import synth_features, synth_inputs, synth_state_dict  # noqa: E402

lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
world, rank = dist.get_world_size(), dist.get_rank()
B, T, S = 64, 180, 50
m = MotionTransformer(26, num_frames=1800, num_layers=8, latent_dim=128, device=dev, music_model_path=None)
m.load_state_dict(synth_state_dict(0), strict=True)
m = m.to(dev).eval()
d = GaussianDiffusion(betas=get_named_beta_schedule("linear", S), model_mean_type=ModelMeanType.START_X,
                      model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE)
xp, xo = (t.to(dev) for t in synth_features(B, T, seed=rank))
noise = synth_inputs(B, T, seed=rank)[1].to(dev)
kw = dict(xf_proj=xp, xf_out=xo, length=[T] * B)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / n


out = d.ddim_sample_loop(m, (B, T, 26), noise=noise, clip_denoised=False, model_kwargs=kw)
t_loop = timed(lambda: d.ddim_sample_loop(m, (B, T, 26), noise=noise, clip_denoised=False, model_kwargs=kw))
t_gather = timed(lambda: gather_shards(out, B * world))
t_both = timed(lambda: gather_shards(d.ddim_sample_loop(m, (B, T, 26), noise=noise, clip_denoised=False, model_kwargs=kw), B * world))
print(f"rank {rank}: loop {t_loop:.3f} ms, gather {t_gather:.3f} ms, loop+gather {t_both:.3f} ms", flush=True)
dist.destroy_process_group()
