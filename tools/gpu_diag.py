"""Exploratory GPU diagnostics (run under gpurun): prints error numbers for every building block so a
failure can be localised from one run.  Not part of the test-suite; see tests/ for the asserted versions."""
import ctypes as C
import json
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusion_conductor_b200 import GaussianDiffusion, MotionTransformer, _lib  # noqa: E402
from diffusion_conductor_b200.gaussian_diffusion import LossType, ModelMeanType, ModelVarType, get_named_beta_schedule  # noqa: E402
from diffusion_conductor_b200.synth import synth_features, synth_inputs, synth_state_dict  # noqa: E402
from oracle import motion_oracle as O  # noqa: E402

OUT = {}
ONLY = set(sys.argv[1:])


def section(name):
    def deco(fn):
        if ONLY and not any(name.startswith(o) for o in ONLY):
            return fn
        t0 = time.time()
        try:
            OUT[name] = fn()
        except Exception as e:  # noqa: BLE001
            OUT[name] = {"error": repr(e), "trace": traceback.format_exc()[-1500:]}
        OUT[name + "_sec"] = round(time.time() - t0, 2)
        print(name, json.dumps(OUT[name])[:1500], flush=True)
        return fn
    return deco


def round16(a, bf16):
    t = torch.from_numpy(a)
    return t.to(torch.bfloat16 if bf16 else torch.float16).float().numpy()


@section("gemm")
def _gemm():
    lib = _lib.load()
    res = {}
    rng = np.random.default_rng(0)
    for (M, N, K) in [(128, 128, 64), (128, 256, 512), (300, 64, 128), (128, 128, 128), (257, 16, 64)]:
        for op, bf in ((0, True), (1, False)):
            A = rng.standard_normal((M, K), dtype=np.float32)
            W = rng.standard_normal((N, K), dtype=np.float32) * 0.1
            b = rng.standard_normal(N, dtype=np.float32)
            out = np.zeros((M, N), dtype=np.float32)
            rc = lib.dc_selftest_gemm(0, op, M, N, K, A.ctypes.data, W.ctypes.data, b.ctypes.data, out.ctypes.data)
            ref = round16(A, bf).astype(np.float64) @ round16(W, bf).astype(np.float64).T + b
            res[f"{M}x{N}x{K}_{'bf16' if bf else 'fp16'}"] = [int(rc), float(np.abs(out - ref).max()), float(np.abs(ref).max())]
    return res


def make_model(num_layers, seed, operand="bf16"):
    m = MotionTransformer(26, num_frames=1800, num_layers=num_layers, latent_dim=128, device="cuda", music_model_path=None,
                          operand_dtype=operand)
    sd = synth_state_dict(seed, num_layers=num_layers)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), sd


def diffusion(S):
    return GaussianDiffusion(betas=get_named_beta_schedule("linear", S), model_mean_type=ModelMeanType.START_X,
                             model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE)


def err(a, b):
    a = a.detach().float().cpu().numpy() if isinstance(a, torch.Tensor) else a
    b = b.detach().float().cpu().numpy() if isinstance(b, torch.Tensor) else b
    return {"max_abs": float(np.abs(a - b).max()), "rel_rms": float(np.sqrt(((a - b) ** 2).mean() / ((b ** 2).mean() + 1e-30))),
            "ref_rms": float(np.sqrt((b ** 2).mean())), "nan": bool(np.isnan(a).any())}


for OPERAND in ("bf16", "fp16"):
    @section(f"forward_small_{OPERAND}")
    def _fwd():
        g = np.load(os.path.join(ROOT, "tests/golden/small_masked.npz"))
        m, sd = make_model(2, 7, OPERAND)
        B, T = 3, 40
        xf_proj, xf_out = synth_features(B, T, seed=11)
        _, x = synth_inputs(B, T, seed=11)
        length = [int(v) for v in g["length"]]
        t = torch.from_numpy(g["t"])
        res = {}
        # per-layer residual-stream check against the oracle (uses an un-masked and a masked run)
        for tag, ln in (("full", [T] * B), ("masked", length)):
            y = m(x.cuda(), t.cuda(), length=ln, xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda())
            torch.cuda.synchronize()
            with torch.no_grad():
                yo = O.motion_transformer_forward(sd, x, t, ln, xf_proj, xf_out)
            res[tag] = err(y, yo)
        res["vs_golden_masked"] = err(y, g["forward"])
        return res

    @section(f"loop_small_{OPERAND}")
    def _loop():
        g = np.load(os.path.join(ROOT, "tests/golden/small_masked.npz"))
        m, sd = make_model(2, 7, OPERAND)
        B, T = 3, 40
        xf_proj, xf_out = synth_features(B, T, seed=11)
        _, x = synth_inputs(B, T, seed=11)
        length = [int(v) for v in g["length"]]
        d = diffusion(25)
        kw = dict(xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda(), length=length)
        res = {}
        x0s = []
        for out in d.ddim_sample_loop_progressive(m, x.shape, noise=x.cuda(), clip_denoised=False, model_kwargs=kw):
            x0s.append(out["pred_xstart"].cpu())
            last = out
        x0s = torch.stack(x0s)
        res["progressive_x0"] = err(x0s, g["ddim_x0"])
        res["progressive_x0_step0"] = err(x0s[0], g["ddim_x0"][0])
        res["progressive_final"] = err(last["sample"], g["ddim_sample"][-1])
        res["final_equals_x0"] = bool(torch.equal(last["sample"].cpu(), last["pred_xstart"].cpu()))
        fin = d.ddim_sample_loop(m, x.shape, noise=x.cuda(), clip_denoised=False, model_kwargs=kw)
        torch.cuda.synchronize()
        res["graph_final"] = err(fin, g["ddim_sample"][-1])
        res["graph_vs_progressive_equal"] = bool(torch.equal(fin.cpu(), last["sample"].cpu()))
        tr = d.ddim_sample_loop(m, x.shape, noise=x.cuda(), clip_denoised=False, model_kwargs=kw, idxs=[0, 5, 24])
        res["idxs_keys"] = sorted(int(k) for k in tr.keys())
        res["idxs_5"] = err(tr[5], g["ddim_sample"][5])
        # DDPM with the reference's own noise stream
        eng = m.engine(torch.device("cuda", 0))
        xs = x.cuda().clone()
        nz = torch.from_numpy(g["ddpm_noise"]).cuda()
        eng.sample_loop(_lib.DC_SAMPLER_DDPM, xs, step_noise=nz)
        torch.cuda.synchronize()
        res["ddpm_final"] = err(xs, g["ddpm_sample"][-1])
        # update rule alone, bit-exact given the reference's x0
        img = x.cuda().clone()
        exact = True
        for n, i in enumerate(range(24, -1, -1)):
            eng.sampler_update(_lib.DC_SAMPLER_DDIM, img, torch.from_numpy(g["ddim_x0"][n]).cuda(), i)
            exact &= bool(np.array_equal(img.cpu().numpy(), g["ddim_sample"][n]))
            img = torch.from_numpy(g["ddim_sample"][n]).cuda()
        res["ddim_update_bit_exact"] = exact
        img = x.cuda().clone()
        exact = True
        for n, i in enumerate(range(24, -1, -1)):
            eng.sampler_update(_lib.DC_SAMPLER_DDPM, img, torch.from_numpy(g["ddpm_x0"][n]).cuda(), i, nz[n])
            exact &= bool(np.array_equal(img.cpu().numpy(), g["ddpm_sample"][n]))
            img = torch.from_numpy(g["ddpm_sample"][n]).cuda()
        res["ddpm_update_bit_exact"] = exact
        return res


@section("time_embed")
def _te():
    g = np.load(os.path.join(ROOT, "tests/golden/time_embed.npz"))
    m, sd = make_model(2, 7)
    # te through forward is not observable; check engine.forward sensitivity instead via the oracle in forward_small.
    return {"note": "covered by forward_small (arbitrary t) and loop_small (tabulated t)"}


@section("c1")
def _c1():
    g = np.load(os.path.join(ROOT, "tests/golden/c1.npz"))
    m, sd = make_model(8, 0)
    mel, noise = synth_inputs(1, 180, seed=0)
    with torch.no_grad():
        xp, xo = m.encode_music(mel.cuda(), "cuda")
    res = {"xf_out": err(xo, g["xf_out"]), "xf_proj": err(xp, g["xf_proj"])}
    d = diffusion(25)
    kw = dict(xf_proj=torch.from_numpy(g["xf_proj"]).cuda(), xf_out=torch.from_numpy(g["xf_out"]).cuda(), length=[180])
    x0s = torch.stack([o["pred_xstart"].cpu() for o in
                       d.ddim_sample_loop_progressive(m, noise.shape, noise=noise.cuda(), clip_denoised=False, model_kwargs=kw)])
    res["x0_all"] = err(x0s, g["ddim_x0"])
    res["x0_per_step_max"] = [round(float(np.abs(x0s[i].numpy() - g["ddim_x0"][i]).max()), 5) for i in range(25)]
    fin = d.ddim_sample_loop(m, noise.shape, noise=noise.cuda(), clip_denoised=False, model_kwargs=kw)
    res["final"] = err(fin, g["final"])
    return res


@section("timing")
def _timing():
    res = {}
    for (B, T, S, name) in [(64, 180, 50, "C2"), (32, 1800, 50, "C3"), (1, 180, 25, "C1")]:
        m, sd = make_model(8, 0)
        xf_proj, xf_out = synth_features(B, T, seed=1)
        _, noise = synth_inputs(B, T, seed=1)
        d = diffusion(S)
        kw = dict(xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda(), length=[T] * B)
        nz = noise.cuda()
        for _ in range(2):
            d.ddim_sample_loop(m, nz.shape, noise=nz, clip_denoised=False, model_kwargs=kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            d.ddim_sample_loop(m, nz.shape, noise=nz, clip_denoised=False, model_kwargs=kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        res[name] = {"ms_per_loop": round(ms, 3), "motion_s_per_s": round(B * T / 30 / (ms / 1e3), 1),
                     "frac_of_bf16_peak": round(B * T * S * 8500224 / (ms / 1e3) / 1624.7e12, 4)}
        del m
    return res


os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(OUT, open(os.path.join(ROOT, "gpurun_out", "diag_%s.json" % ("_".join(sorted(ONLY)) or "all")), "w"), indent=1)
print("DIAG DONE")
