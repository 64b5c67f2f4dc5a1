"""CPU tests of the host side: checkpoint layout, schedule/coefficients (bit-exact), C-ABI exports,
argument validation, and the rank-sharding logic under gloo (world_size 2)."""
import ctypes as C
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import diffusion_conductor_b200 as dcb
from diffusion_conductor_b200 import _lib
from diffusion_conductor_b200.gaussian_diffusion import (GaussianDiffusion, LossType, ModelMeanType, ModelVarType,
                                                         get_named_beta_schedule)
from diffusion_conductor_b200.generate import shard_range
from diffusion_conductor_b200.synth import reference_shapes, synth_state_dict
from oracle import motion_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="session")
def built_lib():
    if not os.path.exists(_lib.LIB_PATH):
        subprocess.run(["bash", os.path.join(ROOT, "build.sh")], check=True, cwd=ROOT)
    return _lib.load()


def diffusion(S):
    return GaussianDiffusion(betas=get_named_beta_schedule("linear", S), model_mean_type=ModelMeanType.START_X,
                             model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE)


def test_state_dict_layout_matches_reference(golden_dir):
    layout = json.load(open(os.path.join(golden_dir, "state_dict_layout.json")))
    m = dcb.MotionTransformer(26, num_frames=1800, num_layers=8, latent_dim=128, device="cpu", music_model_path=None)
    sd = m.state_dict()
    assert list(sd.keys()) == list(layout.keys())          # same keys, same order
    for k, (shape, dtype) in layout.items():
        assert list(sd[k].shape) == shape and str(sd[k].dtype) == dtype, k
    assert set(reference_shapes().keys()) == set(layout.keys())
    # a fresh model keeps the reference's zero-initialised output projections (quirk Q6)
    assert float(sd["out.weight"].abs().max()) == 0.0
    assert float(sd["temporal_decoder_blocks.0.ffn.linear2.weight"].abs().max()) == 0.0
    assert m.load_state_dict(synth_state_dict(0), strict=True).missing_keys == []


def test_constructor_surface():
    with pytest.raises(NotImplementedError):
        dcb.MotionTransformer(26, num_frames=8, latent_dim=128, device="cpu", music_model_path=None, no_eff=True)
    m = dcb.MotionTransformer(26, num_frames=64, num_layers=1, latent_dim=128, device="cpu", music_model_path=None,
                              no_clip=True)      # unknown kwargs are swallowed like the reference's **kargs
    assert (m.num_frames, m.latent_dim, m.time_embed_dim, m.cond_mask_prob) == (64, 128, 512, 0.1)
    mask = m.generate_src_mask(5, [5, 2, 0])
    assert mask.tolist() == [[1, 1, 1, 1, 1], [1, 1, 0, 0, 0], [0, 0, 0, 0, 0]]
    with pytest.raises(RuntimeError):          # CPU tensors must fail loudly, not fall back
        m(torch.zeros(1, 4, 26), torch.zeros(1, dtype=torch.long), length=[4], xf_proj=torch.zeros(1, 4, 64),
          xf_out=torch.zeros(1, 4, 64))


def test_tables_and_step_coefficients_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "tables.npz"))
    for S in (25, 50, 1000):
        d = diffusion(S)
        for name in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod",
                     "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
                     "posterior_mean_coef1", "posterior_mean_coef2"):
            assert np.array_equal(getattr(d, name).view(np.uint64), g[f"S{S}_{name}"].view(np.uint64)), (S, name)
        # the coefficient rows must reproduce the reference update rule bit for bit
        tb = O.Tables(O.linear_betas(S))
        gen = torch.Generator().manual_seed(S)
        x, x0, nz = (torch.randn(S, 7, 26, generator=gen) for _ in range(3))
        t = torch.arange(S)
        for eta in (0.0, 0.37):
            cf = d.step_coefficients(eta)
            assert cf.shape == (S, 8) and cf.dtype == torch.float32
            c = cf.view(S, 1, 1, 8)
            eps = (c[..., 0] * x - x0) / c[..., 1]
            mine = (x0 * c[..., 2] + c[..., 3] * eps) + c[..., 4] * nz
            assert torch.equal(mine, O.ddim_update(tb, x, t, x0, eta=eta, noise=nz)), (S, eta)
        c = d.step_coefficients(0.0).view(S, 1, 1, 8)
        mine = (c[..., 5] * x0 + c[..., 6] * x) + c[..., 7] * nz
        assert torch.equal(mine, O.ddpm_update(tb, x, t, x0, nz))
        assert float(d.step_coefficients(0.0)[0, 2]) == 1.0     # alpha_bar_prev(0) = 1 -> last sample == pred_xstart


def test_unsupported_configurations_raise():
    betas = get_named_beta_schedule("linear", 25)
    m = dcb.MotionTransformer(26, num_frames=16, num_layers=1, latent_dim=128, device="cpu", music_model_path=None)
    d = GaussianDiffusion(betas=betas, model_mean_type=ModelMeanType.EPSILON, model_var_type=ModelVarType.FIXED_SMALL,
                          loss_type=LossType.MSE)
    with pytest.raises(NotImplementedError):
        d.ddim_sample_loop(m, (1, 4, 26), model_kwargs={})
    d = diffusion(25)
    with pytest.raises(NotImplementedError):
        d.ddim_sample_loop(torch.nn.Linear(2, 2), (1, 4, 26))
    with pytest.raises(NotImplementedError):
        d.ddim_sample_loop(m, (1, 4, 26), cond_fn=lambda *a, **k: None, model_kwargs={})
    with pytest.raises(NotImplementedError):
        d.p_sample_loop(m, (1, 4, 26), pre_seq=torch.zeros(1), model_kwargs={})
    with pytest.raises(NotImplementedError):
        get_named_beta_schedule("nope", 10)
    assert len(get_named_beta_schedule("cosine", 10)) == 10


def test_cabi_exports_every_declared_symbol(built_lib):
    header = open(os.path.join(ROOT, "include", "dc_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(dc_[a-z_0-9]+)\s*\(", header)))
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(built_lib, name), f"{name} declared in include/dc_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback(built_lib):
    cfg = _lib.DcConfig(26, 1800, 128, 64, 8, 8, 0, 0)
    h = C.c_void_p()
    rc = built_lib.dc_create(C.byref(cfg), C.byref(h))
    assert rc == -3 and not h.value
    assert b"no CPU fallback" in built_lib.dc_last_error(None)
    cfg.latent_dim = 64
    assert built_lib.dc_create(C.byref(cfg), C.byref(h)) == -2      # unsupported specialisation is reported as such
    with pytest.raises(RuntimeError):
        _lib.check(-3)


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 512):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_sharded_sampling_world_size_2_gloo(tmp_path):
    """Two gloo ranks shard 5 ragged clips, run the (oracle) sampler on their shard, all_gather:
    every rank must end up with exactly the single-process result."""
    script = os.path.join(ROOT, "tests", "_gloo_worker.py")
    out = tmp_path / "res"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", script, str(out)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    a = torch.load(str(out) + ".0")
    b = torch.load(str(out) + ".1")
    assert torch.equal(a["gathered"], b["gathered"])
    assert torch.equal(a["gathered"], a["single"])
    assert torch.equal(a["even"], b["even"]) and torch.equal(a["even"], a["single"][:4])
