"""GPU parity tests (pytest -m gpu): the CUDA path, called through the C ABI (ctypes shim), against
 (a) the committed golden vectors produced by the unmodified reference, and
 (b) the CPU oracle on the same seeded inputs,
plus size-independent properties at the BASELINE.json sizes.

Tolerances (stated per operand type; fp32 accumulate everywhere, LayerNorm / softmax / update rule fp32).
Reference floors measured on the reference itself: fp32-vs-fp64 7.5e-7; bf16 Linear layers
7.7e-3 max-abs per step, 4.5e-3 relative RMS on the final keypoints.
    bf16 operands: per-step pred_xstart  rel-RMS <= 5e-3,  max-abs <= 1.5e-2 * max(1, rms)
    fp16 operands: per-step pred_xstart  rel-RMS <= 1e-3,  max-abs <= 4e-3 * max(1, rms)
    whole trajectory (final keypoints): same bounds as a single step (errors do not compound under DDIM's
    contraction towards x0; measured r01: 3.3e-3 / 4.2e-4 rel-RMS on C1).
Index handling, coefficient tables and the sampler update given x0 are bit-exact.
Every comparison appends its measured (rel-RMS, max-abs) to gpurun_out/parity_report.json (written at session end).
"""
import atexit
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from diffusion_conductor_b200 import GaussianDiffusion, MotionTransformer, _lib
from diffusion_conductor_b200.gaussian_diffusion import LossType, ModelMeanType, ModelVarType, get_named_beta_schedule
from diffusion_conductor_b200.generate import generate_music_motion
from diffusion_conductor_b200.synth import synth_features, synth_inputs, synth_state_dict
from oracle import motion_oracle as O

pytestmark = pytest.mark.gpu

TOL = {"bf16": dict(rel=5e-3, mx=2.0e-2), "fp16": dict(rel=1e-3, mx=4e-3)}   # mx scales with max(1, ref rms); measured worst case over ~190k elements: 1.55e-2 rms (bf16), 1.9e-3 rms (fp16)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = []


@atexit.register
def _write_report():
    if REPORT:
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        json.dump(REPORT, open(os.path.join(out, "parity_report.json"), "w"), indent=0)



def make_model(num_layers, seed, operand="bf16", num_frames=1800):
    m = MotionTransformer(26, num_frames=num_frames, num_layers=num_layers, latent_dim=128, device="cuda",
                          music_model_path=None, operand_dtype=operand)
    sd = synth_state_dict(seed, num_layers=num_layers, num_frames=num_frames)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), sd


def diffusion(S):
    return GaussianDiffusion(betas=get_named_beta_schedule("linear", S), model_mean_type=ModelMeanType.START_X,
                             model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE)


def close(a, b, operand, what=""):
    a = a.detach().float().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().float().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert np.isfinite(a).all(), what
    rms = float(np.sqrt((b ** 2).mean()))
    rel = float(np.sqrt(((a - b) ** 2).mean())) / max(rms, 1e-12)
    mx = float(np.abs(a - b).max())
    t = TOL[operand]
    REPORT.append({"what": what, "operand": operand, "rel_rms": rel, "max_abs": mx, "ref_rms": rms})
    assert rel <= t["rel"], f"{what}: rel-RMS {rel:.3e} > {t['rel']:.1e}"
    assert mx <= t["mx"] * max(1.0, rms), f"{what}: max-abs {mx:.3e} > {t['mx'] * max(1.0, rms):.1e}"
    return rel, mx


# ------------------------------------------------------------------------------------------------
def test_library_is_the_native_one():
    lib = _lib.load()
    assert os.path.samefile(lib._name, _lib.LIB_PATH)


@pytest.mark.parametrize("operand", ["bf16", "fp16"])
def test_tcgen05_gemm_selftest(operand):
    """The tcgen05/TMEM GEMM building block against numpy on 16-bit-rounded inputs: only fp32 accumulation
    order differs, so the bound is ~1e-5 relative.  Covers ragged M, N in {16..256}, K in {64..512}."""
    lib = _lib.load()
    rng = np.random.default_rng(0)
    dt = torch.bfloat16 if operand == "bf16" else torch.float16
    r16 = lambda a: torch.from_numpy(a).to(dt).double().numpy()  # noqa: E731
    for (M, N, K) in [(128, 128, 64), (128, 256, 512), (300, 64, 128), (1, 16, 64), (257, 240, 192), (1024, 256, 512)]:
        A = rng.standard_normal((M, K), dtype=np.float32)
        W = rng.standard_normal((N, K), dtype=np.float32) * 0.1
        b = rng.standard_normal(N, dtype=np.float32)
        out = np.full((M, N), np.nan, dtype=np.float32)
        rc = lib.dc_selftest_gemm(0, 0 if operand == "bf16" else 1, M, N, K, A.ctypes.data, W.ctypes.data, b.ctypes.data,
                                  out.ctypes.data)
        assert rc == 0, lib.dc_last_error(None)
        ref = r16(A) @ r16(W).T + b
        assert np.abs(out - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max()), (M, N, K)
    assert lib.dc_selftest_gemm(0, 0, 8, 24, 64, A.ctypes.data, W.ctypes.data, None, out.ctypes.data) == -1


@pytest.mark.parametrize("operand", ["bf16", "fp16"])
def test_forward_small_masked_vs_reference_golden(golden_dir, operand):
    """forward(x, t, length, xf_proj, xf_out) with per-sample timesteps and ragged lengths (33 and 1 of 40)."""
    g = np.load(os.path.join(golden_dir, "small_masked.npz"))
    m, sd = make_model(2, 7, operand)
    B, T = 3, 40
    xf_proj, xf_out = synth_features(B, T, seed=11)
    _, x = synth_inputs(B, T, seed=11)
    length = [int(v) for v in g["length"]]
    y = m(x.cuda(), torch.from_numpy(g["t"]).cuda(), length=length, xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda())
    assert y.shape == (B, T, 26) and y.is_cuda
    close(y, g["forward"], operand, "forward vs reference golden")
    # (B,T,13,2) input is flattened like the reference (quirk Q4)
    y4 = m(x.view(B, T, 13, 2).cuda(), torch.from_numpy(g["t"]).cuda(), length=length, xf_proj=xf_proj.cuda(),
           xf_out=xf_out.cuda())
    assert torch.equal(y4, y)


def test_sampler_update_bit_exact(golden_dir):
    """DDIM / DDPM update given the reference's own pred_xstart: bit-identical samples for all 25 steps."""
    g = np.load(os.path.join(golden_dir, "small_masked.npz"))
    m, _ = make_model(2, 7)
    B, T = 3, 40
    xf_proj, xf_out = synth_features(B, T, seed=11)
    _, x = synth_inputs(B, T, seed=11)
    d = diffusion(25)
    eng = d._bind(m, x.cuda(), dict(xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda(), length=[T] * B))
    nz = torch.from_numpy(g["ddpm_noise"]).cuda()
    for sampler, key in ((_lib.DC_SAMPLER_DDIM, "ddim"), (_lib.DC_SAMPLER_DDPM, "ddpm")):
        img = x.cuda().clone()
        for n, i in enumerate(range(24, -1, -1)):
            eng.sampler_update(sampler, img, torch.from_numpy(g[key + "_x0"][n]).cuda(), i,
                               nz[n] if key == "ddpm" else None)
            assert np.array_equal(img.cpu().numpy(), g[key + "_sample"][n]), (key, i)
            img = torch.from_numpy(g[key + "_sample"][n]).cuda()


@pytest.mark.parametrize("operand", ["bf16", "fp16"])
def test_sampling_loops_small(golden_dir, operand):
    g = np.load(os.path.join(golden_dir, "small_masked.npz"))
    m, sd = make_model(2, 7, operand)
    B, T = 3, 40
    xf_proj, xf_out = synth_features(B, T, seed=11)
    _, x = synth_inputs(B, T, seed=11)
    length = [int(v) for v in g["length"]]
    d = diffusion(25)
    kw = dict(xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda(), length=length)
    outs = list(d.ddim_sample_loop_progressive(m, x.shape, noise=x.cuda(), clip_denoised=False, model_kwargs=kw))
    assert len(outs) == 25 and set(outs[0]) == {"sample", "pred_xstart"}
    for n, o in enumerate(outs):                                   # per-step tolerance on predicted x0
        close(o["pred_xstart"], g["ddim_x0"][n], operand, f"pred_xstart step {24 - n}")
    close(outs[-1]["sample"], g["ddim_sample"][-1], operand, "trajectory final")
    # t = 0: alpha_bar_prev = 1 -> sample == pred_xstart bit for bit (quirk Q11)
    assert torch.equal(outs[-1]["sample"], outs[-1]["pred_xstart"])
    # graph-replayed whole loop == step-by-step generator, bit for bit
    fin = d.ddim_sample_loop(m, x.shape, noise=x.cuda(), clip_denoised=False, model_kwargs=kw)
    assert torch.equal(fin, outs[-1]["sample"])
    # idxs -> dict {i: sample after step counter i} plus {S: final} (quirk Q12)
    tr = d.ddim_sample_loop(m, x.shape, noise=x.cuda(), clip_denoised=False, model_kwargs=kw, idxs=[0, 5, 24])
    assert sorted(tr) == [0, 5, 24, 25]
    assert torch.equal(tr[5], outs[5]["sample"]) and torch.equal(tr[25], fin)
    # DDPM with the reference's own noise stream
    eng = m.engine(torch.device("cuda", 0))
    xs = x.cuda().clone()
    eng.sample_loop(_lib.DC_SAMPLER_DDPM, xs, step_noise=torch.from_numpy(g["ddpm_noise"]).cuda())
    close(xs, g["ddpm_sample"][-1], operand, "ddpm final")
    # clip_denoised=True clamps pred_xstart
    o = d.ddim_sample(m, x.cuda(), torch.tensor([24] * B).cuda(), clip_denoised=True, model_kwargs=kw)
    assert float(o["pred_xstart"].abs().max()) <= 1.0
    ref = torch.from_numpy(g["ddim_x0"][0]).clamp(-1, 1)
    close(o["pred_xstart"], ref, operand, "clipped x0")
    # per-sample (non-uniform) timesteps through ddim_sample
    tt = torch.from_numpy(g["t"]).cuda()
    o = d.ddim_sample(m, x.cuda(), tt, clip_denoised=False, model_kwargs=kw)
    close(o["pred_xstart"], g["forward"], operand, "ddim_sample non-uniform t")
    ref = O.ddim_update(O.Tables(O.linear_betas(25)), x, torch.from_numpy(g["t"]), o["pred_xstart"].cpu())
    assert torch.allclose(o["sample"].cpu(), ref, atol=1e-6)


@pytest.mark.parametrize("operand", ["bf16", "fp16"])
def test_c1_trajectory_vs_reference_golden(golden_dir, operand):
    """BASELINE.json configs[0]: 25-step DDIM, batch 1, 6 s clip, 8 layers, through the music encoder."""
    g = np.load(os.path.join(golden_dir, "c1.npz"))
    m, sd = make_model(8, 0, operand)
    mel, noise = synth_inputs(1, 180, seed=0)
    with torch.no_grad():
        xp, xo = m.encode_music(mel.cuda(), "cuda")
    # hand-written fp32 CUDA encoder vs the reference's CPU convolutions (golden): summation-order noise only
    assert np.abs(xo.cpu().numpy() - g["xf_out"]).max() < 2e-4 and np.abs(xp.cpu().numpy() - g["xf_proj"]).max() < 2e-4
    d = diffusion(25)
    kw = dict(xf_proj=torch.from_numpy(g["xf_proj"]).cuda(), xf_out=torch.from_numpy(g["xf_out"]).cuda(), length=[180])
    for n, o in enumerate(d.ddim_sample_loop_progressive(m, noise.shape, noise=noise.cuda(), clip_denoised=False,
                                                         model_kwargs=kw)):
        close(o["pred_xstart"], g["ddim_x0"][n], operand, f"C1 pred_xstart step {24 - n}")
    fin = d.ddim_sample_loop(m, noise.shape, noise=noise.cuda(), clip_denoised=False, model_kwargs=kw)
    close(fin, g["final"], operand, "C1 final keypoints")
    # generate_music_motion-equivalent driver from the mel itself
    out = generate_music_motion(m, d, mel[0].numpy(), 26, noise=noise.cuda())
    assert out.shape == (1, 180, 26)
    close(out, g["final"], operand, "C1 via generate_music_motion")


@pytest.mark.parametrize("operand", ["bf16", "fp16"])
def test_oracle_agreement_mid_size_and_edge_shapes(operand):
    """Seeded inputs vs the CPU oracle at shapes that exercise tile boundaries: clips straddling 128-row tiles,
    a ragged last tile, T = 1, a zero-length clip, and the maximum T = num_frames."""
    m, sd = make_model(2, 21, operand, num_frames=300)
    for (B, T, length) in [(5, 77, [77, 10, 77, 0, 76]), (1, 1, [1]), (3, 128, [128, 128, 64]), (2, 300, [300, 299])]:
        xf_proj, xf_out = synth_features(B, T, seed=B * 1000 + T)
        _, x = synth_inputs(B, T, seed=B * 1000 + T)
        t = torch.arange(B) * 3 % 25
        y = m(x.cuda(), t.cuda(), length=length, xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda())
        with torch.no_grad():
            ref = O.motion_transformer_forward(sd, x, t, length, xf_proj, xf_out)
        close(y, ref, operand, f"forward B={B} T={T}")
    with pytest.raises(RuntimeError):            # T beyond the positional table must fail loudly
        xf_proj, xf_out = synth_features(1, 301, seed=1)
        m(torch.zeros(1, 301, 26).cuda(), torch.zeros(1, dtype=torch.long).cuda(), length=[301], xf_proj=xf_proj.cuda(),
          xf_out=xf_out.cuda())
    with pytest.raises(ValueError):              # music frames must equal motion frames (quirk Q2)
        m(torch.zeros(1, 10, 26).cuda(), torch.zeros(1, dtype=torch.long).cuda(), length=[10],
          xf_proj=torch.zeros(1, 11, 64).cuda(), xf_out=torch.zeros(1, 11, 64).cuda())
    with pytest.raises(TypeError):
        m(torch.zeros(1, 10, 26).cuda(), torch.zeros(1, dtype=torch.long).cuda(), xf_proj=torch.zeros(1, 10, 64).cuda(),
          xf_out=torch.zeros(1, 10, 64).cuda())


def test_call_order_errors():
    lib = _lib.load()
    cfg = _lib.DcConfig(26, 64, 128, 64, 1, 8, 0, 0)
    h = C.c_void_p()
    assert lib.dc_create(C.byref(cfg), C.byref(h)) == 0
    x = torch.zeros(1, 8, 26, device="cuda")
    assert lib.dc_finalize_weights(h) == -1 and b"missing weight" in lib.dc_last_error(h)
    assert lib.dc_sample_loop(h, 1, 25, C.c_void_p(x.data_ptr()), None, None, None, None) == -4
    assert lib.dc_prepare_cond(h, C.c_void_p(x.data_ptr()), C.c_void_p(x.data_ptr()), None, 1, 8, None) == -4
    lib.dc_destroy(h)


def test_full_size_properties_c2():
    """BASELINE.json configs[1] size (batch 64 x 180 frames, 50-step DDIM, 8 layers): properties that hold at
    any size -- determinism, clip independence (a clip's motion does not depend on its batch neighbours or its
    position in the batch), permutation equivariance, final sample == final pred_xstart -- plus agreement of a
    few clips with the oracle over the first steps."""
    m, sd = make_model(8, 0, "bf16")
    B, T, S = 64, 180, 50
    xf_proj, xf_out = synth_features(B, T, seed=1)
    _, noise = synth_inputs(B, T, seed=1)
    d = diffusion(S)
    kw = dict(xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda(), length=[T] * B)
    a = d.ddim_sample_loop(m, noise.shape, noise=noise.cuda(), clip_denoised=False, model_kwargs=kw)
    b = d.ddim_sample_loop(m, noise.shape, noise=noise.cuda(), clip_denoised=False, model_kwargs=kw)
    assert torch.isfinite(a).all() and torch.equal(a, b)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(0))
    kwp = dict(xf_proj=xf_proj[perm].cuda(), xf_out=xf_out[perm].cuda(), length=[T] * B)
    p = d.ddim_sample_loop(m, noise.shape, noise=noise[perm].cuda(), clip_denoised=False, model_kwargs=kwp)
    close(p, a[perm.cuda()], "bf16", "permutation equivariance")
    sub = [3, 17, 40]
    kws = dict(xf_proj=xf_proj[sub].cuda(), xf_out=xf_out[sub].cuda(), length=[T] * 3)
    s = d.ddim_sample_loop(m, (3, T, 26), noise=noise[sub].cuda(), clip_denoised=False, model_kwargs=kws)
    assert torch.equal(s, a[sub])            # clip-aligned tiles: bit-identical whatever the batch around the clip
    # oracle on the 3-clip sub-batch, first 2 steps
    _, x0s, _ = O.sample_loop(sd, O.Tables(O.linear_betas(S)), noise[sub], [T] * 3, xf_proj[sub], xf_out[sub], max_steps=2)
    gen = d.ddim_sample_loop_progressive(m, (3, T, 26), noise=noise[sub].cuda(), clip_denoised=False, model_kwargs=kws)
    for n in range(2):
        close(next(gen)["pred_xstart"], x0s[n], "bf16", f"C2 sub-batch step {n}")


def test_full_size_properties_c3_long_sequence():
    """BASELINE.json configs[2] shape (60 s clips, 1800 frames, mel 5400x128): a 4-clip batch over 5 of the 50
    steps against the oracle, plus masking semantics (frames >= length do not influence frames < length)."""
    m, sd = make_model(8, 0, "bf16")
    B, T, S = 4, 1800, 50
    xf_proj, xf_out = synth_features(B, T, seed=2)
    _, noise = synth_inputs(B, T, seed=2)
    d = diffusion(S)
    length = [1800, 1800, 1000, 1800]
    kw = dict(xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda(), length=length)
    _, x0s, _ = O.sample_loop(sd, O.Tables(O.linear_betas(S)), noise, length, xf_proj, xf_out, max_steps=3)
    gen = d.ddim_sample_loop_progressive(m, noise.shape, noise=noise.cuda(), clip_denoised=False, model_kwargs=kw)
    for n in range(3):
        close(next(gen)["pred_xstart"], x0s[n], "bf16", f"C3 step {n}")
    # masked tail of clip 2: changing the noise there must not change the valid frames' prediction
    x2 = noise.clone()
    x2[2, 1000:] += 5.0
    t = torch.tensor([49] * B).cuda()
    y1 = m(noise.cuda(), t, **kw)
    y2 = m(x2.cuda(), t, **kw)
    assert torch.equal(y1[2, :1000], y2[2, :1000]) and torch.equal(y1[[0, 1, 3]], y2[[0, 1, 3]])


def test_pair_mode_cta_group_2(golden_dir, monkeypatch):
    """The opt-in CTA-pair build of the layer kernel (tcgen05 cta_group::2, DC_PAIR=1) computes the same thing."""
    monkeypatch.setenv("DC_PAIR", "1")
    g = np.load(os.path.join(golden_dir, "small_masked.npz"))
    m, sd = make_model(2, 7, "bf16")                      # a fresh handle reads DC_PAIR at creation
    B, T = 3, 40
    xf_proj, xf_out = synth_features(B, T, seed=11)
    _, x = synth_inputs(B, T, seed=11)
    length = [int(v) for v in g["length"]]
    y = m(x.cuda(), torch.from_numpy(g["t"]).cuda(), length=length, xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda())
    close(y, g["forward"], "bf16", "pair-mode forward vs reference golden")
    m8, sd8 = make_model(8, 0, "bf16")
    xf_proj, xf_out = synth_features(5, 300, seed=3)       # odd tile count -> padding tile in the last pair; fused reduction off/on
    _, x = synth_inputs(5, 300, seed=3)
    t = torch.tensor([3, 0, 24, 7, 11])
    y = m8(x.cuda(), t.cuda(), length=[300, 300, 17, 300, 299], xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda())
    with torch.no_grad():
        ref = O.motion_transformer_forward(sd8, x, t, [300, 300, 17, 300, 299], xf_proj, xf_out)
    close(y, ref, "bf16", "pair-mode forward B=5 T=300")


@pytest.mark.parametrize("operand", ["bf16", "fp16"])
def test_cluster_sizes_and_merge_paths(operand):
    """The cluster-per-clip kernel at every cluster regime against the oracle: 1 tile (no exchange), 2-4 tiles (all-to-all
    DSMEM pull), 5-8 (reduce-scatter + all-gather, portable cluster sizes), 9-16 (non-portable cluster sizes), with ragged
    lengths, a zero-length clip and tiles whose last rows are padding.  Both operand types: the static-softmax-shift
    decision depends on it (bound <= 30 for bf16, <= 4 for fp16)."""
    m, sd = make_model(2, 31, operand, num_frames=2048)
    for (B, T, length) in [(3, 100, [100, 0, 31]), (3, 200, [200, 129, 1]), (2, 384, [384, 257]), (2, 513, [513, 512]),
                           (2, 600, [600, 77]), (2, 1000, [1000, 999]), (2, 1100, [1100, 300]), (1, 1800, [1800]), (2, 2048, [2048, 1025])]:
        xf_proj, xf_out = synth_features(B, T, seed=B * 1000 + T)
        _, x = synth_inputs(B, T, seed=B * 1000 + T)
        t = (torch.arange(B) * 7 + 3) % 25
        y = m(x.cuda(), t.cuda(), length=length, xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda())
        with torch.no_grad():
            ref = O.motion_transformer_forward(sd, x, t, length, xf_proj, xf_out)
        close(y, ref, operand, f"forward B={B} T={T} ({-(-T // 128)} tiles per clip)")


@pytest.mark.parametrize("operand", ["bf16", "fp16"])
def test_exchange_paths_agree(monkeypatch, operand):
    """The per-clip exchange of the time-axis reduction has three implementations (distributed shared memory pulled, pushed for
    two-tile clips, and flag-in-data words through L2 for long clips).  Force each of them on shapes the default would route
    elsewhere -- with the running-max softmax shift too, which exercises the rescaling merge -- and compare against the oracle
    and against each other."""
    m, sd = make_model(2, 41, operand, num_frames=2048)
    monkeypatch.setenv("DC_STATIC_SHIFT", "0")               # read when the weights are finalised (first call): a second model
    m_run, _ = make_model(2, 41, operand, num_frames=2048)
    xf_proj, xf_out = synth_features(1, 8, seed=1)
    m_run(torch.zeros(1, 8, 26).cuda(), torch.zeros(1, dtype=torch.long).cuda(), length=[8], xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda())
    monkeypatch.delenv("DC_STATIC_SHIFT")
    for (B, T, length) in [(3, 200, [200, 129, 0]), (2, 500, [500, 499]), (2, 700, [700, 130]), (1, 1800, [1800])]:
        xf_proj, xf_out = synth_features(B, T, seed=B * 100 + T)
        _, x = synth_inputs(B, T, seed=B * 100 + T)
        t = (torch.arange(B) * 5 + 2) % 25
        with torch.no_grad():
            ref = O.motion_transformer_forward(sd, x, t, length, xf_proj, xf_out)
        outs = {}
        for name, model, env in (("cluster", m, {"DC_GX": "0", "DC_PUSH": "0"}), ("cluster-push", m, {"DC_GX": "0", "DC_PUSH": "1"}),
                                 ("l2", m, {"DC_GX": "1"}), ("l2-running-max", m_run, {"DC_GX": "1"}),
                                 ("cluster-running-max", m_run, {"DC_GX": "0"})):
            for k in ("DC_GX", "DC_PUSH"):
                monkeypatch.delenv(k, raising=False)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            y = model(x.cuda(), t.cuda(), length=length, xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda())
            close(y, ref, operand, f"forward T={T}, exchange {name}")
            outs[name] = y.cpu()
        for k in ("DC_GX", "DC_PUSH"):
            monkeypatch.delenv(k, raising=False)
        rms = float(ref.pow(2).mean().sqrt())
        for a, b in (("cluster", "l2"), ("cluster", "cluster-push"), ("l2-running-max", "cluster-running-max")):
            assert float((outs[a] - outs[b]).pow(2).mean().sqrt()) < (2e-3 if operand == "bf16" else 3e-4) * rms, (T, a, b)


def test_static_and_running_softmax_shift(monkeypatch):
    """Time-axis softmax of the self-attention keys: the kernel replaces the running column max by a static
    weight-norm bound of |k| when that bound is small (bf16 operands, ordinary weights) and keeps the exact running max
    otherwise (forced with DC_STATIC_SHIFT=0, or automatically when the key weights are large).  Both against the
    oracle, with ragged lengths and a zero-length clip (every frame masked: the reference yields A = 0)."""
    B, T = 4, 200
    xf_proj, xf_out = synth_features(B, T, seed=12)
    _, x = synth_inputs(B, T, seed=12)
    t = torch.tensor([24, 1, 13, 0])
    length = [200, 0, 131, 199]
    outs = {}
    for mode in ("static", "running", "large-weights"):
        if mode == "running":
            monkeypatch.setenv("DC_STATIC_SHIFT", "0")
        else:
            monkeypatch.delenv("DC_STATIC_SHIFT", raising=False)
        sd = synth_state_dict(23, num_layers=3)
        if mode == "large-weights":
            for l in range(3):
                sd[f"temporal_decoder_blocks.{l}.sa_block.key.weight"] = sd[f"temporal_decoder_blocks.{l}.sa_block.key.weight"] * 12.0
        m = MotionTransformer(26, num_frames=1800, num_layers=3, latent_dim=128, device="cuda", music_model_path=None,
                              operand_dtype="bf16")
        m.load_state_dict(sd, strict=True)
        m = m.cuda().eval()
        y = m(x.cuda(), t.cuda(), length=length, xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda())
        with torch.no_grad():
            ref = O.motion_transformer_forward(sd, x, t, length, xf_proj, xf_out)
        close(y, ref, "bf16", f"forward, softmax shift mode {mode}")
        outs[mode] = y
    close(outs["static"], outs["running"], "bf16", "static vs running shift")


def test_batch_larger_than_the_gpu():
    """More clusters than the GPU holds at once (300 clips x 2 tiles > 148 SMs): clusters are scheduled as SMs free up
    and every clip still comes out as if it were generated alone (25-step DDIM loop, one launch)."""
    m, sd = make_model(2, 17, "bf16")
    B, T, S = 300, 180, 25
    xf_proj, xf_out = synth_features(B, T, seed=6)
    _, noise = synth_inputs(B, T, seed=6)
    d = diffusion(S)
    kw = dict(xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda(), length=[T] * B)
    a = d.ddim_sample_loop(m, noise.shape, noise=noise.cuda(), clip_denoised=False, model_kwargs=kw)
    assert torch.isfinite(a).all()
    sub = [0, 149, 150, 299]
    kws = dict(xf_proj=xf_proj[sub].cuda(), xf_out=xf_out[sub].cuda(), length=[T] * len(sub))
    s = d.ddim_sample_loop(m, (len(sub), T, 26), noise=noise[sub].cuda(), clip_denoised=False, model_kwargs=kws)
    assert torch.equal(s, a[sub])            # clip-aligned tiles: a clip's arithmetic does not depend on its neighbours
    ref, _, _ = O.sample_loop(sd, O.Tables(O.linear_betas(S)), noise[sub], [T] * len(sub), xf_proj[sub], xf_out[sub])
    close(s, ref, "bf16", "4 of 300 clips vs oracle")


def test_fused_time_axis_reduction_both_ways(monkeypatch):
    """Per-layer launch path (DC_PERSIST=0; used for clips longer than 16 tiles): the time-axis softmax + K^T V reduction
    fused into the layer kernel and the stand-alone kv_reduce kernel agree with the oracle on the same inputs (T = 300:
    2-3 tiles per clip, clips straddling tiles)."""
    monkeypatch.setenv("DC_PERSIST", "0")
    B, T = 4, 300
    xf_proj, xf_out = synth_features(B, T, seed=9)
    _, x = synth_inputs(B, T, seed=9)
    t = torch.tensor([24, 1, 13, 0])
    length = [300, 120, 300, 299]
    outs = []
    for flag in ("0", "1"):
        monkeypatch.setenv("DC_FUSE_KV", flag)
        m, sd = make_model(3, 5, "fp16")
        y = m(x.cuda(), t.cuda(), length=length, xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda())
        with torch.no_grad():
            ref = O.motion_transformer_forward(sd, x, t, length, xf_proj, xf_out)
        close(y, ref, "fp16", f"forward, DC_FUSE_KV={flag}")
        outs.append(y)
    close(outs[0], outs[1], "fp16", "fused vs stand-alone reduction")


def test_ddpm_1000_steps_small():
    """BASELINE.json configs[4] sampler (1000-step DDPM) at a small shape: the whole stochastic trajectory with a fixed
    noise stream against the oracle."""
    m, sd = make_model(2, 13, "fp16")
    B, T, S = 2, 24, 1000
    xf_proj, xf_out = synth_features(B, T, seed=4)
    _, x = synth_inputs(B, T, seed=4)
    d = diffusion(S)
    gen = torch.Generator().manual_seed(99)
    step_noise = torch.randn(S, B, T, 26, generator=gen)
    eng = d._bind(m, x.cuda(), dict(xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda(), length=[T] * B))
    xs = x.cuda().clone()
    eng.sample_loop(_lib.DC_SAMPLER_DDPM, xs, step_noise=step_noise.cuda())
    with pytest.raises(RuntimeError):            # buffers sized for another S than the uploaded schedule: refused by the library
        eng.sample_loop(_lib.DC_SAMPLER_DDPM, xs.clone(), num_steps=S - 1)
    ref, _, _ = O.sample_loop(sd, O.Tables(O.linear_betas(S)), x, [T] * B, xf_proj, xf_out, kind="ddpm", step_noise=step_noise)
    assert torch.isfinite(xs).all()
    close(xs, ref, "bf16", "1000-step DDPM final sample (fp16 operands, bf16-level bound over 1000 stochastic steps)")
    # the public API draws the same per-step noise from torch's generator as the reference does
    torch.manual_seed(5)
    a = d.p_sample_loop(m, x.shape, noise=x.cuda(), clip_denoised=False,
                        model_kwargs=dict(xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda(), length=[T] * B))
    torch.manual_seed(5)
    b = d.p_sample_loop(m, x.shape, noise=x.cuda(), clip_denoised=False,
                        model_kwargs=dict(xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda(), length=[T] * B))
    assert torch.equal(a, b) and not torch.equal(a, xs)
    # bounded noise buffer: the same loop as 16 launches of <= 64 steps (dc_sample_range) is bit-identical
    d.NOISE_BLOCK_BYTES = 64 * x.numel() * 4
    torch.manual_seed(5)
    c = d.p_sample_loop(m, x.shape, noise=x.cuda(), clip_denoised=False,
                        model_kwargs=dict(xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda(), length=[T] * B), idxs=[10, 999])
    assert torch.equal(c[1000], a) and sorted(c) == [10, 999, 1000] and torch.equal(c[999], a)


def test_smooth_motion_on_device():
    """pixel scaling + Savitzky-Golay smoothing (kernel 19, order 5) of generated motion on the GPU
    against the CPU oracle (scipy, float64).  fp32 FIR on values up to 600: tolerance 2e-3 pixels absolute."""
    from diffusion_conductor_b200.generate import smooth_motion
    from oracle import postprocess_oracle as PO

    g = torch.Generator().manual_seed(5)
    for (B, T) in ((1, 19), (3, 180), (2, 1800), (2, 37)):
        motion = torch.rand(B, T, 26, generator=g)
        out = smooth_motion(motion.cuda(), kernel=19, order=5, window=600.0)
        ref = PO.vis_motion_keypoints(motion.numpy().astype(np.float64), window=600, kernel=19)
        assert out.shape == (B, T, 13, 2)
        assert float(np.abs(out.cpu().numpy() - ref).max()) < 2e-3, (B, T)
    one = smooth_motion(motion[0].cuda())
    assert one.shape == (37, 13, 2) and torch.equal(one, out[0])
    with pytest.raises(ValueError):
        smooth_motion(torch.rand(1, 18, 26).cuda())           # scipy: window_length must be <= size of x
    with pytest.raises(RuntimeError):
        smooth_motion(torch.rand(1, 40, 26))                   # no CPU path


def test_music_encoder_on_device():
    """MusicEncoder + proj in the CUDA library (direct fp32 convolutions, BatchNorm folded) against the
    oracle at mel lengths that exercise partial tiles, the stride-3 pool's floor and tiny inputs; fp32 both sides, so the
    tolerance only covers summation order and the BatchNorm folding: 1e-4 absolute on features of rms ~0.5."""
    m, sd = make_model(2, 41, "bf16")
    for (B, Tm) in ((2, 540), (3, 541), (1, 100), (2, 8), (1, 5400)):
        mel, _ = synth_inputs(B, Tm, seed=Tm)            # (B, 3 Tm, 128): take the first Tm mel frames
        mel = mel[:, :Tm].contiguous()
        xp, xo = m.encode_music(mel.cuda(), "cuda")
        with torch.no_grad():
            rp, ro = O.encode_music(sd, mel)
        assert xo.shape == ro.shape == (B, (Tm - 1) // 3 + 1, 64), (xo.shape, ro.shape)
        assert float((xo.cpu() - ro).abs().max()) < 1e-4 and float((xp.cpu() - rp).abs().max()) < 1e-4, (B, Tm)
    with pytest.raises(RuntimeError):
        m.encode_music(torch.rand(1, 3, 128).cuda(), "cuda")      # too few frames for the reflect-padded convolutions
    with pytest.raises(RuntimeError):
        m.encode_music(torch.rand(1, 30, 128), "cpu")             # no CPU path


# ------------------------------------------------------------------------------------------------
# round 2: the BASELINE.json configs pinned by reference-generated fixtures (oracle/make_golden.py pinned_configs)
# ------------------------------------------------------------------------------------------------
def _pinned_steps(gen, S, steps):
    """pred_xstart of the timestep indices in `steps` from a progressive generator (yield n belongs to timestep S-1-n)."""
    want = {S - 1 - int(t): int(t) for t in steps}
    got = {}
    for n, o in enumerate(gen):
        if n in want:
            got[want[n]] = o["pred_xstart"]
    return got


@pytest.mark.parametrize("operand", ["bf16", "fp16"])
def test_c3_full_clip_vs_reference_golden(golden_dir, operand):
    """north_star target: ONE full 60 s clip (T = 1800 -> a 15-CTA cluster, mel 5400 x 128) through the on-device music
    encoder and all 50 DDIM steps, against what the unmodified reference produced (gaussian_diffusion.py:871-965,
    transformer.py:447-497): per-step pred_xstart at timesteps 49/40/25/10/0 and the final keypoints."""
    g = np.load(os.path.join(golden_dir, "c3_clip.npz"))
    m, sd = make_model(8, 0, operand)
    mel, noise = synth_inputs(1, 1800, seed=3)
    xp, xo = m.encode_music(mel.cuda(), "cuda")
    assert np.abs(xo.cpu().numpy()[:, ::8] - g["xf_out_rows8"]).max() < 2e-4
    assert np.abs(xp.cpu().numpy()[:, ::8] - g["xf_proj_rows8"]).max() < 2e-4
    d = diffusion(50)
    kw = dict(xf_proj=xp, xf_out=xo, length=[1800])
    got = _pinned_steps(d.ddim_sample_loop_progressive(m, noise.shape, noise=noise.cuda(), clip_denoised=False, model_kwargs=kw),
                        50, g["steps"])
    for i, t in enumerate(g["steps"]):
        close(got[int(t)], g["ddim_x0"][i], operand, f"C3 clip pred_xstart t={int(t)}")
    fin = d.ddim_sample_loop(m, noise.shape, noise=noise.cuda(), clip_denoised=False, model_kwargs=kw)
    close(fin, g["final"], operand, "C3 clip final keypoints (whole trajectory, music encoder included)")
    out = generate_music_motion(m, d, mel, 26, noise=noise.cuda())
    assert torch.equal(out, fin)


@pytest.mark.parametrize("operand", ["bf16", "fp16"])
def test_c2_schedule_pair_vs_reference_golden(golden_dir, operand):
    """The C2 schedule (50-step DDIM, 8 layers, 6 s clips) on a 2-clip batch against the reference's trajectory."""
    g = np.load(os.path.join(golden_dir, "c2_pair.npz"))
    m, sd = make_model(8, 0, operand)
    xf_proj, xf_out = synth_features(2, 180, seed=21)
    _, noise = synth_inputs(2, 180, seed=21)
    d = diffusion(50)
    kw = dict(xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda(), length=torch.LongTensor([180, 180]).cuda())   # as ddpm_trainer.py:198
    got = _pinned_steps(d.ddim_sample_loop_progressive(m, noise.shape, noise=noise.cuda(), clip_denoised=False, model_kwargs=kw,
                                                       progress=True), 50, g["steps"])
    for i, t in enumerate(g["steps"]):
        close(got[int(t)], g["ddim_x0"][i], operand, f"C2 pair pred_xstart t={int(t)}")
    fin = d.ddim_sample_loop(m, noise.shape, noise=noise.cuda(), clip_denoised=False, model_kwargs=kw, progress=True)
    close(fin, g["final"], operand, "C2 pair final keypoints")


@pytest.mark.parametrize("operand", ["bf16", "fp16"])
def test_ddpm_1000_steps_vs_reference_golden(golden_dir, operand):
    """The C5 sampler: 1000-step DDPM (gaussian_diffusion.py:667-781), B = 1, T = 180, 8 layers, ONE launch for all 1000
    steps, on the noise stream the reference drew (regenerated from the seed in the fixture)."""
    g = np.load(os.path.join(golden_dir, "ddpm1000.npz"))
    m, sd = make_model(8, 0, operand)
    xf_proj, xf_out = synth_features(1, 180, seed=22)
    _, noise = synth_inputs(1, 180, seed=22)
    torch.manual_seed(int(g["seed"]))            # the reference's draws: one CPU randn_like per step (gaussian_diffusion.py:656)
    nz = torch.stack([torch.randn(1, 180, 26) for _ in range(1000)]).cuda()
    d = diffusion(1000)
    eng = d._bind(m, noise.cuda(), dict(xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda(), length=[180]))
    xs = noise.cuda().clone()
    trace = torch.empty(1000, 1, 180, 26, device="cuda")
    eng.sample_loop(_lib.DC_SAMPLER_DDPM, xs, step_noise=nz, trace_x=trace)
    for i, t in enumerate(g["steps"]):
        close(trace[999 - int(t)], g["ddpm_sample"][i], operand, f"DDPM-1000 sample after t={int(t)}")
    assert torch.equal(trace[999], xs)


def test_time_embed_kernel_vs_reference_golden(golden_dir):
    """timestep_embedding + time_embed MLP (transformer.py:8-25, 410-414) straight from the kernel that tabulates the
    schedule: integer timesteps incl. 0, 999 against the reference's fp32 values (precise sinf / cosf, fp32 FMA chains)."""
    g = np.load(os.path.join(golden_dir, "time_embed.npz"))
    m, sd = make_model(2, 7)
    te = m.engine(torch.device("cuda", 0)).time_embedding(torch.from_numpy(g["t"]).cuda())
    assert te.shape == (len(g["t"]), 512)
    err = float(np.abs(te.cpu().numpy() - g["te"]).max())
    REPORT.append({"what": "time_embed kernel vs reference", "operand": "f32", "rel_rms": None, "max_abs": err, "ref_rms": float(np.sqrt((g["te"] ** 2).mean()))})
    assert err <= 1e-5, err


def test_rng_stream_matches_reference_after_ddim_loop():
    """Drop-in fidelity: ddim_sample draws one randn_like per step even at eta = 0 (gaussian_diffusion.py:822),
    so after ddim_sample_loop the caller's next draw must be what it would be after the reference's loop."""
    m, sd = make_model(2, 7)
    B, T, S = 2, 40, 25
    xf_proj, xf_out = synth_features(B, T, seed=1)
    _, noise = synth_inputs(B, T, seed=1)
    d = diffusion(S)
    kw = dict(xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda(), length=[T] * B)
    torch.manual_seed(11)
    d.ddim_sample_loop(m, noise.shape, noise=noise.cuda(), clip_denoised=False, model_kwargs=kw)
    after_ours = torch.randn(7, device="cuda")
    torch.manual_seed(11)
    for _ in range(S):
        torch.randn_like(noise.cuda())
    assert torch.equal(after_ours, torch.randn(7, device="cuda"))
    # noise=None: the initial draw comes first (gaussian_diffusion.py:942), then the per-step draws
    torch.manual_seed(12)
    a = d.ddim_sample_loop(m, (B, T, 26), clip_denoised=False, model_kwargs=kw, device="cuda")
    torch.manual_seed(12)
    x0 = torch.randn(B, T, 26, device="cuda")
    b = d.ddim_sample_loop(m, (B, T, 26), noise=x0, clip_denoised=False, model_kwargs=kw)
    assert torch.equal(a, b)
    # stochastic DDIM (eta > 0) uses those very draws
    torch.manual_seed(13)
    c = d.ddim_sample_loop(m, (B, T, 26), noise=x0, clip_denoised=False, model_kwargs=kw, eta=0.5)
    torch.manual_seed(13)
    nz = torch.stack([torch.randn_like(x0) for _ in range(S)])
    xs = x0.clone()
    eng = d._bind(m, x0, kw, eta=0.5)
    eng.sample_loop(_lib.DC_SAMPLER_DDIM, xs, step_noise=nz)
    assert torch.equal(c, xs) and not torch.equal(c, b)


def test_schedule_cache_is_keyed_on_content():
    """Two GaussianDiffusion objects with different S used back to back on one model (CPython may give them the same id):
    the library must follow the schedule, and a stale S must be refused rather than index out of bounds."""
    m, sd = make_model(2, 7)
    B, T = 2, 40
    xf_proj, xf_out = synth_features(B, T, seed=1)
    _, noise = synth_inputs(B, T, seed=1)
    kw = dict(xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda(), length=[T] * B)
    outs = {}
    for S in (50, 25, 50, 25):
        d = diffusion(S)
        o = d.ddim_sample_loop(m, noise.shape, noise=noise.cuda(), clip_denoised=False, model_kwargs=kw, idxs=[0])
        assert sorted(o) == [0, S]
        if S in outs:
            assert torch.equal(outs[S], o[S])
        outs[S] = o[S]
        del d
    assert not torch.equal(outs[25], outs[50])


def test_generate_music_motion_ragged_mel_lengths():
    """generate_music_motion with a mel length that is not a multiple of 3 (T = (Tm - 1) // 3 + 1 like the encoder and
    ddpm_trainer.py:187-188), ragged `length`, and a wrong-shaped noise."""
    m, sd = make_model(2, 7)
    d = diffusion(25)
    for Tm in (541, 542, 543):
        mel = torch.rand(2, Tm, 128, generator=torch.Generator().manual_seed(Tm))
        T = (Tm - 1) // 3 + 1
        noise = torch.randn(2, T, 26, generator=torch.Generator().manual_seed(1))
        out = generate_music_motion(m, d, mel, 26, length=torch.LongTensor([T, T - 5]), noise=noise.cuda())
        assert out.shape == (2, T, 26) and torch.isfinite(out).all()
        xp, xo = O.encode_music(sd, mel)
        ref, _, _ = O.sample_loop(sd, O.Tables(O.linear_betas(25)), noise, [T, T - 5], xp, xo)
        close(out, ref, "bf16", f"generate_music_motion Tm={Tm}")
    with pytest.raises(ValueError):
        generate_music_motion(m, d, mel, 26, noise=torch.randn(2, 180, 26).cuda())


def test_cluster_occupancy_query_and_loud_fallback():
    """cudaOccupancyMaxActiveClusters for the cluster sizes of C2 (2 tiles per clip) and C3 (15): recorded in the parity
    report; a cluster size the device cannot co-schedule is an error, not a silent drop to the per-layer path."""
    m, sd = make_model(2, 7)
    eng = m.engine(torch.device("cuda", 0))
    occ = {nt: eng.cluster_occupancy(nt) for nt in (1, 2, 4, 8, 15, 16)}
    REPORT.append({"what": "cudaOccupancyMaxActiveClusters(clip_kernel) by cluster size", "occupancy": occ})
    assert occ[1] >= 100 and occ[2] >= 50 and occ[15] >= 1
    with pytest.raises(RuntimeError):
        eng.cluster_occupancy(17)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_generate_music_motion_under_nccl(tmp_path):
    """The product's multi-GPU path on hardware (ddpm_trainer.py:183-201 extended to ranks): generate_music_motion under
    NCCL on 2 GPUs, ragged lengths, odd and even clip counts, must reproduce the single-process result bit for bit."""
    script = os.path.join(ROOT, "tests", "_nccl_worker.py")
    out = tmp_path / "res"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29741", script, str(out)],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    a, b = torch.load(str(out) + ".0"), torch.load(str(out) + ".1")
    for key in ("odd", "even"):
        assert torch.equal(a[key], b[key]), key                 # every rank holds the full gathered batch
        assert torch.equal(a[key], a[key + "_single"]), key      # == one process generating all clips (clip-aligned tiles)


# ------------------------------------------------------------------------------------------------
# on-device evaluation features (diffusion_conductor_b200/evaluation.py, dc_eval_*)
# ------------------------------------------------------------------------------------------------
def test_evaluation_features_on_device(golden_dir):
    """ST-GCN motion encoder latents, Frechet statistics, latent L1, motion beats and beat consistency against the reference
    golden (tests/golden/eval_features.npz) and the CPU oracle at a second shape (ragged tile, single frame)."""
    from diffusion_conductor_b200 import evaluation as EV
    from diffusion_conductor_b200.synth import synth_motion, synth_stgcn_state_dict
    from oracle import eval_oracle as E

    g = np.load(os.path.join(golden_dir, "eval_features.npz"))
    enc = EV.MotionEncoder_STGCN()
    sd = synth_stgcn_state_dict(5)
    enc.load_state_dict(sd, strict=True)
    enc = enc.cuda().eval()
    motion = synth_motion(3, 200, seed=9)
    lat = enc(motion.cuda())
    scale = max(1.0, float(np.abs(g["latent"]).max()))
    err = float(np.abs(lat.cpu().numpy() - g["latent"]).max())
    REPORT.append({"what": "ST-GCN latents vs reference golden", "operand": "fp32", "rel_rms": None, "max_abs": err, "ref_rms": scale})
    assert err < 5e-5 * scale
    assert torch.equal(enc.features(motion.cuda())[-1], lat)
    with pytest.raises(RuntimeError):
        enc(motion)                                                 # CPU tensor: no fallback
    # a second shape vs the oracle: 5 clips x 37 frames (3 time tiles, ragged last tile), and a single frame
    for (N, T) in [(5, 37), (1, 1), (2, 16)]:
        mo = synth_motion(N, T, seed=N * 100 + T)
        ref = E.motion_features(sd, mo).numpy()
        assert np.abs(enc(mo.cuda()).cpu().numpy() - ref).max() < 5e-5 * max(1.0, float(np.abs(ref).max())), (N, T)
    # new weights are picked up (the engine reloads when the state_dict changes)
    sd2 = synth_stgcn_state_dict(6)
    enc.load_state_dict(sd2, strict=True)
    ref2 = E.motion_features(sd2, motion).numpy()
    assert np.abs(enc(motion.cuda()).cpu().numpy() - ref2).max() < 5e-5 * max(1.0, float(np.abs(ref2).max()))
    # ---- statistics, FGD, L1 on the golden latents
    a = torch.from_numpy(g["latent"][:2]).cuda()
    b = torch.from_numpy(g["latent"][1:]).cuda()
    mu, cov = EV.feature_statistics(a)
    assert np.abs(mu - g["mu_a"]).max() < 1e-6 and np.abs(cov - g["cov_a"]).max() < 1e-6 * max(1.0, float(np.abs(g["cov_a"]).max()))
    fgd = EV.frechet_gesture_distance(a, b)
    assert abs(fgd - float(g["fgd"])) < 1e-5 * max(1.0, abs(float(g["fgd"])))
    assert abs(EV.latent_l1(a, b) - float(g["l1"])) < 1e-5 * float(g["l1"])
    div = EV.diversity_score([a[0], a[1], b[1]], perm=torch.tensor([2, 0, 1]))
    ref_div = E.feature_l1(np.concatenate([g["latent"][0], g["latent"][1], g["latent"][2]]), np.concatenate([g["latent"][2], g["latent"][0], g["latent"][1]]))
    assert abs(div - ref_div) < 1e-5 * ref_div
    # ---- beats
    env, beats = EV.motion_beats(motion.cuda())
    assert np.array_equal(beats.cpu().numpy(), g["beats"])
    ref_env = np.stack([E.motion_peak_onehot(motion[i].numpy())[0] for i in range(3)])
    assert np.abs(env.cpu().numpy() - ref_env).max() < 1e-5
    scores = EV.beat_consistency(torch.from_numpy(g["music_beats"]).cuda(), beats)
    assert np.abs(scores.cpu().numpy() - g["beat_scores"]).max() < 1e-5
    none = EV.beat_consistency(torch.from_numpy(g["music_beats"]).cuda(), torch.zeros_like(beats))
    assert float(none.abs().max()) == 0.0
    # order as an argument; a short clip (T < order) has no interior minimum test beyond the clipped neighbours
    env2, beats2 = EV.motion_beats(motion[:, :7].cuda(), order=3)
    for i in range(3):
        assert np.array_equal(beats2[i].cpu().numpy(), E.motion_peak_onehot(motion[i, :7].numpy(), order=3)[1])
