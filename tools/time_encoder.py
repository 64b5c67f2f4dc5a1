"""Run the music encoder on a C2-sized mel batch a few times (for an ncu launch list / event timing). Run under gpurun."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusion_conductor_b200 import MotionTransformer  # noqa: E402
from diffusion_conductor_b200.synth import synth_inputs, synth_state_dict  # noqa: E402

B, T = (int(v) for v in (sys.argv[1:3] if len(sys.argv) > 2 else (64, 180)))
m = MotionTransformer(26, num_frames=1800, num_layers=8, latent_dim=128, device="cuda", music_model_path=None)
m.load_state_dict(synth_state_dict(0), strict=True)
m = m.cuda().eval()
mel, _ = synth_inputs(B, T, seed=1)
mel = mel.cuda()
for _ in range(3):
    m.encode_music(mel, "cuda")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    m.encode_music(mel, "cuda")
e1.record()
torch.cuda.synchronize()
print(f"encode_music B={B} T={T}: {e0.elapsed_time(e1) / 5:.3f} ms")
