"""Print the measured parity numbers quoted in DESIGN.md (run under gpurun): C1 golden trajectory, bf16 and fp16 operands."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusion_conductor_b200 import GaussianDiffusion, MotionTransformer  # noqa: E402
from diffusion_conductor_b200.gaussian_diffusion import LossType, ModelMeanType, ModelVarType, get_named_beta_schedule  # noqa: E402
from diffusion_conductor_b200.synth import synth_inputs, synth_state_dict  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "c1.npz"))
mel, noise = synth_inputs(1, 180, seed=0)
for operand in ("bf16", "fp16"):
    m = MotionTransformer(26, num_frames=1800, num_layers=8, latent_dim=128, device="cuda", music_model_path=None, operand_dtype=operand)
    m.load_state_dict(synth_state_dict(0, num_layers=8), strict=True)
    m = m.cuda().eval()
    d = GaussianDiffusion(betas=get_named_beta_schedule("linear", 25), model_mean_type=ModelMeanType.START_X,
                          model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE)
    xp, xo = m.encode_music(mel.cuda(), "cuda")
    print(operand, "encode_music max abs vs reference golden:", float(np.abs(xo.cpu().numpy() - g["xf_out"]).max()))
    kw = dict(xf_proj=torch.from_numpy(g["xf_proj"]).cuda(), xf_out=torch.from_numpy(g["xf_out"]).cuda(), length=[180])
    rels, mxs = [], []
    for n, o in enumerate(d.ddim_sample_loop_progressive(m, noise.shape, noise=noise.cuda(), clip_denoised=False, model_kwargs=kw)):
        a, b = o["pred_xstart"].cpu().numpy(), g["ddim_x0"][n]
        rels.append(float(np.sqrt(((a - b) ** 2).mean()) / np.sqrt((b ** 2).mean())))
        mxs.append(float(np.abs(a - b).max()))
    fin = d.ddim_sample_loop(m, noise.shape, noise=noise.cuda(), clip_denoised=False, model_kwargs=kw).cpu().numpy()
    rf = float(np.sqrt(((fin - g["final"]) ** 2).mean()) / np.sqrt((g["final"] ** 2).mean()))
    print(f"{operand}: per-step pred_xstart rel-RMS max {max(rels):.2e} mean {np.mean(rels):.2e}, max-abs {max(mxs):.2e}; final keypoints rel-RMS {rf:.2e}")
