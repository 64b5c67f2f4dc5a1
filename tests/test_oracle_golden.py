"""The oracle restatement must reproduce what the unmodified reference produced
(fixtures written by oracle/make_golden.py in the build container)."""
import os

import numpy as np
import torch

from diffusion_conductor_b200.synth import synth_features, synth_inputs, synth_state_dict
from oracle import motion_oracle as O

TABLE_NAMES = ["betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod",
               "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
               "posterior_mean_coef1", "posterior_mean_coef2"]


def test_tables_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "tables.npz"))
    for S in (25, 50, 1000):
        tb = O.Tables(O.linear_betas(S))
        for n in TABLE_NAMES:
            a, b = getattr(tb, n), g[f"S{S}_{n}"]
            assert a.dtype == np.float64 and np.array_equal(a.view(np.uint64), b.view(np.uint64)), (S, n)


def test_time_embedding(golden_dir):
    g = np.load(os.path.join(golden_dir, "time_embed.npz"))
    sd = synth_state_dict(7, num_layers=2)
    t = torch.from_numpy(g["t"])
    s = O.timestep_embedding(t, 128)
    assert np.array_equal(s.numpy(), g["sinusoid"])          # bit-exact: same ops, same order
    te = O._lin(sd, "time_embed.2", torch.nn.functional.silu(O._lin(sd, "time_embed.0", s)))
    np.testing.assert_allclose(te.numpy(), g["te"], rtol=0, atol=1e-6)


def test_small_masked_forward_and_loops(golden_dir):
    g = np.load(os.path.join(golden_dir, "small_masked.npz"))
    sd = synth_state_dict(7, num_layers=2)
    B, T = 3, 40
    xf_proj, xf_out = synth_features(B, T, seed=11)
    _, x = synth_inputs(B, T, seed=11)
    length = [int(v) for v in g["length"]]
    with torch.no_grad():
        y = O.motion_transformer_forward(sd, x, torch.from_numpy(g["t"]), length, xf_proj, xf_out)
    np.testing.assert_allclose(y.numpy(), g["forward"], rtol=0, atol=2e-6)
    tb = O.Tables(O.linear_betas(25))
    final, x0s, smp = O.sample_loop(sd, tb, x, length, xf_proj, xf_out, kind="ddim")
    np.testing.assert_allclose(torch.stack(x0s).numpy(), g["ddim_x0"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(torch.stack(smp).numpy(), g["ddim_sample"], rtol=0, atol=2e-5)
    # last DDIM step: alpha_bar_prev == 1 -> sample == pred_xstart bit for bit
    assert np.array_equal(smp[-1].numpy(), x0s[-1].numpy())
    final, x0s, smp = O.sample_loop(sd, tb, x, length, xf_proj, xf_out, kind="ddpm",
                                    step_noise=torch.from_numpy(g["ddpm_noise"]))
    np.testing.assert_allclose(torch.stack(smp).numpy(), g["ddpm_sample"], rtol=0, atol=5e-5)


def test_update_rules_bit_exact_given_x0(golden_dir):
    """Index handling and coefficient arithmetic are bit-exact when fed the reference's own x0."""
    g = np.load(os.path.join(golden_dir, "small_masked.npz"))
    _, x = synth_inputs(3, 40, seed=11)
    tb = O.Tables(O.linear_betas(25))
    img = x
    for n, i in enumerate(range(24, -1, -1)):
        t = torch.tensor([i] * 3)
        img = O.ddim_update(tb, img, t, torch.from_numpy(g["ddim_x0"][n]))
        assert np.array_equal(img.numpy(), g["ddim_sample"][n]), i
        img = torch.from_numpy(g["ddim_sample"][n])
    img = x
    for n, i in enumerate(range(24, -1, -1)):
        t = torch.tensor([i] * 3)
        out = O.ddpm_update(tb, img, t, torch.from_numpy(g["ddpm_x0"][n]), torch.from_numpy(g["ddpm_noise"][n]))
        assert np.array_equal(out.numpy(), g["ddpm_sample"][n]), i
        img = torch.from_numpy(g["ddpm_sample"][n])


def test_c1_music_encoder_and_trajectory(golden_dir):
    g = np.load(os.path.join(golden_dir, "c1.npz"))
    sd = synth_state_dict(0, num_layers=8)
    mel, noise = synth_inputs(1, 180, seed=0)
    with torch.no_grad():
        xp, xo = O.encode_music(sd, mel)
    np.testing.assert_allclose(xo.numpy(), g["xf_out"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(xp.numpy(), g["xf_proj"], rtol=0, atol=2e-5)
    tb = O.Tables(O.linear_betas(25))
    final, x0s, _ = O.sample_loop(sd, tb, noise, [180], torch.from_numpy(g["xf_proj"]), torch.from_numpy(g["xf_out"]))
    np.testing.assert_allclose(torch.stack(x0s).numpy(), g["ddim_x0"], rtol=0, atol=5e-5)
    np.testing.assert_allclose(final.numpy(), g["final"], rtol=0, atol=5e-5)


def test_postprocess_oracle_reproduces_polynomials():
    """Known answers of Savitzky-Golay smoothing (visualization.py:20-26): a polynomial of degree <= order is a fixed
    point everywhere, edges included; white noise is attenuated."""
    from oracle import postprocess_oracle as PO

    T = 60
    t = np.linspace(0.0, 1.0, T)
    poly = 0.3 + 0.2 * t - 0.5 * t ** 2 + 0.1 * t ** 5
    motion = np.tile(poly[:, None], (1, 26)).astype(np.float64)[None]
    out = PO.vis_motion_keypoints(motion, window=600, kernel=19)
    assert out.shape == (1, T, 13, 2)
    assert np.allclose(out[0, :, 4, 1], 600 * poly, atol=1e-8)
    rng = np.random.default_rng(0)
    noise = rng.standard_normal((1, 200, 26))
    sm = PO.vis_motion_keypoints(noise, window=1.0, kernel=19)
    assert sm.std() < 0.75 * noise.std()


def test_savgol_matrices_match_scipy():
    """The host-side coefficient builder (numpy only) against scipy: interior FIR == savgol_coeffs, and the full filter
    assembled from (fir, edge) == savgol_filter(mode='interp') on random data."""
    from scipy.signal import savgol_coeffs, savgol_filter

    from diffusion_conductor_b200.generate import savgol_matrices

    for kernel, order in ((19, 5), (11, 5), (7, 2), (31, 3)):
        fir, edge = savgol_matrices(kernel, order)
        assert np.allclose(fir, savgol_coeffs(kernel, order)[::-1], atol=1e-12)
        half = kernel // 2
        x = np.random.default_rng(kernel).standard_normal(kernel + 23)
        y = np.convolve(x, fir[::-1], mode="same")
        y[:half] = edge @ x[:kernel]
        y[-half:] = (edge @ x[::-1][:kernel])[::-1]
        assert np.allclose(y, savgol_filter(x, kernel, order), atol=1e-10)


# ---- round-2 fixtures: the BASELINE.json configs pinned by the unmodified reference (oracle/make_golden.py pinned_configs)
def _pin(x0s, S, steps):
    return torch.stack([x0s[S - 1 - int(t)] for t in steps]).numpy()


def test_c3_full_clip_vs_reference(golden_dir):
    """north_star target: one 60 s clip (T = 1800, mel 5400 x 128), music encoder + 50-step DDIM."""
    g = np.load(os.path.join(golden_dir, "c3_clip.npz"))
    sd = synth_state_dict(0, num_layers=8)
    mel, noise = synth_inputs(1, 1800, seed=3)
    with torch.no_grad():
        xp, xo = O.encode_music(sd, mel)
    np.testing.assert_allclose(xo.numpy()[:, ::8], g["xf_out_rows8"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(xp.numpy()[:, ::8], g["xf_proj_rows8"], rtol=0, atol=2e-5)
    final, x0s, _ = O.sample_loop(sd, O.Tables(O.linear_betas(50)), noise, [1800], xp, xo)
    np.testing.assert_allclose(_pin(x0s, 50, g["steps"]), g["ddim_x0"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(final.numpy(), g["final"], rtol=0, atol=1e-4)


def test_c2_schedule_pair_vs_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "c2_pair.npz"))
    sd = synth_state_dict(0, num_layers=8)
    xf_proj, xf_out = synth_features(2, 180, seed=21)
    _, noise = synth_inputs(2, 180, seed=21)
    final, x0s, _ = O.sample_loop(sd, O.Tables(O.linear_betas(50)), noise, [180, 180], xf_proj, xf_out)
    np.testing.assert_allclose(_pin(x0s, 50, g["steps"]), g["ddim_x0"], rtol=0, atol=5e-5)
    np.testing.assert_allclose(final.numpy(), g["final"], rtol=0, atol=5e-5)


def ddpm_noise_stream(seed, S, shape):
    """The per-step noise the reference draws (gaussian_diffusion.py:656: one randn_like per step from torch's global CPU
    generator), regenerated from the seed stored in the fixture."""
    torch.manual_seed(int(seed))
    return torch.stack([torch.randn(*shape) for _ in range(S)])


def test_ddpm_1000_steps_vs_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "ddpm1000.npz"))
    sd = synth_state_dict(0, num_layers=8)
    xf_proj, xf_out = synth_features(1, 180, seed=22)
    _, noise = synth_inputs(1, 180, seed=22)
    nz = ddpm_noise_stream(g["seed"], 1000, (1, 180, 26))
    _, x0s, smp = O.sample_loop(sd, O.Tables(O.linear_betas(1000)), noise, [180], xf_proj, xf_out, kind="ddpm", step_noise=nz)
    got = torch.stack([smp[999 - int(t)] for t in g["steps"]]).numpy()
    np.testing.assert_allclose(got, g["ddpm_sample"], rtol=0, atol=2e-4)


# ------------------------------------------------------------------------------------------------
# evaluation features (oracle/eval_oracle.py) against what the unmodified reference ST_GCN module and the
# reference's own metric source produced (oracle/make_golden.py eval -> tests/golden/eval_features.npz)
# ------------------------------------------------------------------------------------------------
def test_eval_oracle_vs_reference_golden(golden_dir):
    import json

    from diffusion_conductor_b200.evaluation import MotionEncoder_STGCN, conductor_graph
    from diffusion_conductor_b200.synth import stgcn_shapes, synth_motion, synth_stgcn_state_dict
    from oracle import eval_oracle as E

    g = np.load(os.path.join(golden_dir, "eval_features.npz"))
    # the graph of the reference's Graph('ConductorMotionX', 'uniform') (restated twice: oracle and product constructor)
    assert np.array_equal(E.conductor_graph().astype(np.float32), g["graph_A"])
    assert np.array_equal(conductor_graph().astype(np.float32), g["graph_A"])
    # checkpoint contract: same keys, shapes, dtypes as the reference's MotionEncoder_STGCN
    layout = json.load(open(os.path.join(golden_dir, "stgcn_state_dict_layout.json")))
    ours = MotionEncoder_STGCN().state_dict()
    assert list(ours.keys()) == list(layout.keys())
    for k, (shape, dtype) in layout.items():
        assert list(ours[k].shape) == shape and str(ours[k].dtype) == dtype, k
    assert {k: list(v) for k, v in stgcn_shapes().items()} == {k: v[0] for k, v in layout.items()}
    sd = synth_stgcn_state_dict(5)
    motion = synth_motion(3, 200, seed=9)
    lat = E.motion_features(sd, motion).numpy()
    assert np.abs(lat - g["latent"]).max() < 2e-5 * max(1.0, np.abs(g["latent"]).max())
    for i in range(3):
        env, beats = E.motion_peak_onehot(motion[i].numpy())
        assert np.array_equal(beats, g["beats"][i]) and beats.sum() > 0
        assert abs(E.alignment_score(g["music_beats"][i], beats) - g["beat_scores"][i]) < 1e-6
    assert E.alignment_score(g["music_beats"][0], np.zeros(200, dtype=bool)) == 0.0
    a, b = g["latent"][:2].reshape(-1, 64), g["latent"][1:].reshape(-1, 64)
    mu, cov = E.feature_stats(a)
    assert np.array_equal(mu, g["mu_a"]) and np.array_equal(cov, g["cov_a"])
    assert abs(E.frechet_distance(mu, cov, *E.feature_stats(b)) - float(g["fgd"])) < 1e-9
    assert abs(E.feature_l1(a, b) - float(g["l1"])) < 1e-6
