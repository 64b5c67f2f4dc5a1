"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference, CPU, fp32) on synthetic weights/inputs.  Run in the build container only:

    python oracle/make_golden.py            # everything
    python oracle/make_golden.py pinned     # only c3_clip / c2_pair / ddpm1000 (round 2)
    python oracle/make_golden.py eval       # only eval_features / stgcn_state_dict_layout (round 2)

The reference cannot travel to the GPU box, so the fixtures are committed together with
this script.  TEST INFRASTRUCTURE ONLY (see oracle/motion_oracle.py header).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/Diffusion_Stage")

from models import GaussianDiffusion, MotionTransformer  # noqa: E402  (the reference)
from models.gaussian_diffusion import LossType, ModelMeanType, ModelVarType, get_named_beta_schedule  # noqa: E402
from models.transformer import timestep_embedding  # noqa: E402

from diffusion_conductor_b200.synth import synth_features, synth_inputs, synth_state_dict  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

TABLE_NAMES = ["betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod",
               "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
               "posterior_mean_coef1", "posterior_mean_coef2"]


def build(num_layers, seed):
    m = MotionTransformer(26, num_frames=1800, num_layers=num_layers, latent_dim=128, device="cpu",
                          music_model_path=None)
    sd = synth_state_dict(seed, num_layers=num_layers)
    missing, unexpected = m.load_state_dict(sd, strict=True)
    m.eval()
    return m


def diffusion(S):
    return GaussianDiffusion(betas=get_named_beta_schedule("linear", S), model_mean_type=ModelMeanType.START_X,
                             model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE)


def run_loop(m, d, noise, kw, kind, eta=0.0):
    x0s, samples = [], []
    gen = (d.ddim_sample_loop_progressive(m, noise.shape, noise=noise, clip_denoised=False, model_kwargs=kw, eta=eta)
           if kind == "ddim" else
           d.p_sample_loop_progressive(m, noise.shape, noise=noise, clip_denoised=False, model_kwargs=kw))
    for out in gen:
        x0s.append(out["pred_xstart"].numpy().copy())
        samples.append(out["sample"].numpy().copy())
    return np.stack(x0s), np.stack(samples)


def write_state_dict_layout():
    """tests/golden/state_dict_layout.json: key -> [shape, dtype] of a fresh reference model (checkpoint contract)."""
    import json
    m = MotionTransformer(26, num_frames=1800, num_layers=8, latent_dim=128, device="cpu", music_model_path=None)
    d = {k: [list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()}
    json.dump(d, open(os.path.join(OUT, "state_dict_layout.json"), "w"), indent=0)



PIN_STEPS = (49, 40, 25, 10, 0)        # timestep indices whose pred_xstart is stored by the pinned configs


def pinned_configs():
    """Reference-generated fixtures of the BASELINE.json configs the round-1 fixtures did not pin
    (gaussian_diffusion.py:871-965, 667-781; transformer.py:447-497):

      c3_clip.npz   north_star target: ONE full 60 s clip (T = 1800, mel 5400 x 128) through the music encoder and
                    50-step DDIM; pred_xstart at timesteps PIN_STEPS, the final keypoints, every 8th feature row.
      c2_pair.npz   the C2 schedule (50-step DDIM, 8 layers, T = 180) on a 2-clip batch, synthetic features.
      ddpm1000.npz  the C5 sampler: 1000-step DDPM, B = 1, T = 180; the noise stream is NOT stored -- the reference
                    draws it from torch's global CPU generator after manual_seed(DDPM_SEED), the test regenerates it."""
    m8 = build(8, seed=0)
    d50 = diffusion(50)

    def pick(x0s, S):
        return np.stack([x0s[S - 1 - t] for t in PIN_STEPS])           # x0s[n] belongs to timestep S - 1 - n

    mel, noise = synth_inputs(1, 1800, seed=3)
    with torch.no_grad():
        xp, xo = m8.encode_music(mel, "cpu")
        x0s, smp = run_loop(m8, d50, noise, dict(xf_proj=xp, xf_out=xo, length=[1800]), "ddim")
    np.savez_compressed(os.path.join(OUT, "c3_clip.npz"), steps=np.array(PIN_STEPS), ddim_x0=pick(x0s, 50), final=smp[-1],
                        xf_out_rows8=xo.numpy()[:, ::8], xf_proj_rows8=xp.numpy()[:, ::8])

    xf_proj, xf_out = synth_features(2, 180, seed=21)
    _, noise = synth_inputs(2, 180, seed=21)
    with torch.no_grad():
        x0s, smp = run_loop(m8, d50, noise, dict(xf_proj=xf_proj, xf_out=xf_out, length=[180, 180]), "ddim")
    np.savez_compressed(os.path.join(OUT, "c2_pair.npz"), steps=np.array(PIN_STEPS), ddim_x0=pick(x0s, 50), final=smp[-1])

    d1000 = diffusion(1000)
    xf_proj, xf_out = synth_features(1, 180, seed=22)
    _, noise = synth_inputs(1, 180, seed=22)
    torch.manual_seed(DDPM_SEED)
    with torch.no_grad():
        x0s, smp = run_loop(m8, d1000, noise, dict(xf_proj=xf_proj, xf_out=xf_out, length=[180]), "ddpm")
    keep = (999, 500, 100, 0)
    np.savez_compressed(os.path.join(OUT, "ddpm1000.npz"), steps=np.array(keep), seed=np.array(DDPM_SEED),
                        ddpm_sample=np.stack([smp[999 - t] for t in keep]), ddpm_x0=np.stack([x0s[999 - t] for t in keep]))


DDPM_SEED = 2024


def eval_features():
    """tests/golden/eval_features.npz: the UNMODIFIED reference ST_GCN module (models/ST_GCN/ST_GCN.py) wrapped
    exactly as MotionEncoder_STGCN does (tools/eval_new_metrics.py:38-74 -- that file itself imports mmcv / librosa at module
    level, which are absent, so the 10-line wrapper class is restated here), and the reference's OWN source of
    motion_peak_onehot / alignment_score / calculate_frechet_distance, extracted from the file with ast and executed."""
    import ast
    import json
    import textwrap

    import scipy.signal as scisignal
    from scipy import linalg
    from torch import nn

    from models.ST_GCN.ST_GCN import ST_GCN  # noqa: E402  (the reference)

    from diffusion_conductor_b200.synth import synth_motion, synth_stgcn_state_dict

    class MotionEncoder_STGCN(nn.Module):
        def __init__(self):
            super().__init__()
            self.st_gcn = ST_GCN(in_channels=2, out_channels=32, graph_args={}, edge_importance_weighting=True, mode="M2S")
            self.fc = nn.Sequential(nn.Conv1d(32 * 13, 64, kernel_size=1), nn.BatchNorm1d(64))

        def features(self, input):
            input = input.transpose(1, 2).transpose(1, 3).unsqueeze(4)
            output = self.st_gcn(input).transpose(1, 2)
            output = torch.flatten(output, start_dim=2)
            output = self.fc(output.transpose(1, 2)).transpose(1, 2)
            features = self.st_gcn.extract_feature(input)
            features.append(output.transpose(1, 2))
            return features

    enc = MotionEncoder_STGCN()
    json.dump({k: [list(v.shape), str(v.dtype)] for k, v in enc.state_dict().items()},
              open(os.path.join(OUT, "stgcn_state_dict_layout.json"), "w"), indent=0)
    graph_A = enc.st_gcn.A.numpy().copy()
    sd = synth_stgcn_state_dict(5, A=graph_A)
    enc.load_state_dict(sd, strict=True)
    enc.eval()
    motion = synth_motion(3, 200, seed=9)
    with torch.no_grad():
        feats = enc.features(motion)
    latent = feats[-1].transpose(1, 2).numpy() if feats[-1].shape[1] == 64 else feats[-1].numpy()     # (N, T, 64)
    layer3 = feats[4].numpy()                                                                              # after st_gcn layer 3: (N, T, 416)

    src = open("/root/reference/Diffusion_Stage/tools/eval_new_metrics.py").read()
    tree = ast.parse(src)
    want = {"alignment_score", "normalize", "motion_peak_onehot", "calculate_frechet_distance"}
    class _Linalg:                     # this scipy dropped sqrtm's `disp` argument (disp=False returned (sqrtm, error estimate))
        @staticmethod
        def sqrtm(a, disp=True):
            r = linalg.sqrtm(a)
            return r if disp else (r, 0.0)

    ns = {"np": np, "scisignal": scisignal, "linalg": _Linalg}
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == "Evaluator":
            body = [n for n in node.body if isinstance(n, ast.FunctionDef) and n.name in want]
            cls = ast.ClassDef(name="RefEvaluator", bases=[], keywords=[], body=body, decorator_list=[])
            mod = ast.Module(body=[cls], type_ignores=[])
            ast.fix_missing_locations(mod)
            exec(compile(mod, "eval_new_metrics.py (extract)", "exec"), ns)
    ev = ns["RefEvaluator"]()
    beats = np.stack([ev.motion_peak_onehot(motion[i].numpy(), "generated") for i in range(motion.shape[0])])
    rng = np.random.RandomState(3)
    music = np.zeros((motion.shape[0], 600), dtype=np.float32)
    for i in range(motion.shape[0]):
        music[i, np.sort(rng.choice(600, 40, replace=False))] = 1.0
    scores = np.array([ev.alignment_score(music[i], beats[i], "generated", sigma=3) for i in range(motion.shape[0])], dtype=np.float64)
    a, b = latent[:2].reshape(-1, 64), latent[1:].reshape(-1, 64)
    mu_a, cov_a, mu_b, cov_b = np.mean(a, axis=0), np.cov(a, rowvar=False), np.mean(b, axis=0), np.cov(b, rowvar=False)
    fgd = ev.calculate_frechet_distance(mu_a, cov_a, mu_b, cov_b)
    l1 = np.mean(np.sum(np.absolute(a - b), axis=-1))
    np.savez_compressed(os.path.join(OUT, "eval_features.npz"), graph_A=graph_A, latent=latent, layer3=layer3, beats=beats, music_beats=music,
                        beat_scores=scores, mu_a=mu_a, cov_a=cov_a, fgd=np.array(fgd), l1=np.array(l1))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "eval":        # only the N4 fixtures
        eval_features()
        for f in sorted(os.listdir(OUT)):
            print(f, os.path.getsize(os.path.join(OUT, f)))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "pinned":      # only the round-2 fixtures (the round-1 ones stay byte-identical)
        pinned_configs()
        for f in sorted(os.listdir(OUT)):
            print(f, os.path.getsize(os.path.join(OUT, f)))
        return

    # ---- (1) schedule tables, bit-exact fp64, and the fp32 values the samplers actually use
    tabs = {}
    for S in (25, 50, 1000):
        d = diffusion(S)
        for n in TABLE_NAMES:
            tabs[f"S{S}_{n}"] = getattr(d, n)
    np.savez_compressed(os.path.join(OUT, "tables.npz"), **tabs)

    # ---- (2) timestep embedding + time MLP for a few integer timesteps
    m2 = build(2, seed=7)
    t = torch.tensor([0, 1, 7, 24, 49, 500, 999])
    with torch.no_grad():
        temb = timestep_embedding(t, 128)
        te = m2.time_embed(temb)
    np.savez_compressed(os.path.join(OUT, "time_embed.npz"), t=t.numpy(), sinusoid=temb.numpy(), te=te.numpy())

    # ---- (3) small masked config: 2 layers, B=3, T=40, ragged lengths, arbitrary per-sample t
    B, T = 3, 40
    xf_proj, xf_out = synth_features(B, T, seed=11)
    _, x = synth_inputs(B, T, seed=11)
    length = [40, 33, 1]
    tt = torch.tensor([24, 3, 0])
    with torch.no_grad():
        y = m2(x, tt, length=length, xf_proj=xf_proj, xf_out=xf_out)
    d25 = diffusion(25)
    kw = dict(xf_proj=xf_proj, xf_out=xf_out, length=length)
    with torch.no_grad():
        x0s, smp = run_loop(m2, d25, x, kw, "ddim")
    torch.manual_seed(123)
    with torch.no_grad():
        x0s_p, smp_p = run_loop(m2, d25, x, kw, "ddpm")
    # the DDPM noise stream the reference drew (global CPU generator, one randn_like per step)
    torch.manual_seed(123)
    ddpm_noise = np.stack([torch.randn_like(x).numpy() for _ in range(25)])
    np.savez_compressed(os.path.join(OUT, "small_masked.npz"), length=np.array(length), t=tt.numpy(),
                        forward=y.numpy(), ddim_x0=x0s, ddim_sample=smp, ddpm_x0=x0s_p, ddpm_sample=smp_p,
                        ddpm_noise=ddpm_noise)

    # ---- (4) C1: 8 layers, B=1, T=180, S=25, through the music encoder (BASELINE.json configs[0])
    m8 = build(8, seed=0)
    mel, noise = synth_inputs(1, 180, seed=0)
    with torch.no_grad():
        xp, xo = m8.encode_music(mel, "cpu")
        kw = dict(xf_proj=xp, xf_out=xo, length=[180])
        x0s, smp = run_loop(m8, d25, noise, kw, "ddim")
    np.savez_compressed(os.path.join(OUT, "c1.npz"), xf_proj=xp.numpy(), xf_out=xo.numpy(), ddim_x0=x0s,
                        final=smp[-1])
    write_state_dict_layout()
    pinned_configs()
    eval_features()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
