"""CPU oracle for the Diffusion-Conductor denoising hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file; only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s baseline legs (cpu_baseline,
torch_eager_gpu_baseline, `--impl reference`) may call it, and there only as the checker
(or the timed baseline), never as the thing shipped.

It is a functional restatement (plain torch-CPU tensor ops over a `state_dict`, no
nn.Module graph) of the reference algorithm.  Each function cites the reference lines
it follows (paths relative to /root/reference/Diffusion_Stage/models/):

  * beta schedule + coefficient tables ....... gaussian_diffusion.py:228-245, 328-379
  * timestep embedding ........................ transformer.py:8-25
  * StylizationBlock .......................... transformer.py:68-81
  * LinearTemporalSelfAttention ............... transformer.py:96-123
  * LinearTemporalCrossAttention .............. transformer.py:138-158
  * FFN ....................................... transformer.py:170-173
  * MotionTransformer.forward ................. transformer.py:469-497
  * MusicEncoder.forward / encode_music ....... transformer.py:289-340, 447-459
  * ddim_sample / p_sample + loops ............ gaussian_diffusion.py:605-665, 783-831, 917-965

Pinning: the reference ships no tests or golden vectors, so this oracle
is pinned against the reference *itself*, imported unmodified in the build container by
`oracle/make_golden.py`, which writes the fixtures under `tests/golden/`;
`tests/test_oracle_golden.py` replays them on every run (CPU).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]

LN_EPS = 1e-5
BN_EPS = 1e-5


# ----------------------------------------------------------------------------------------
# diffusion tables (host, float64)
# ----------------------------------------------------------------------------------------
def linear_betas(num_steps: int) -> np.ndarray:
    """gaussian_diffusion.py:237-245 -- 'linear' schedule scaled by 1000/S, fp64."""
    scale = 1000 / num_steps
    return np.linspace(scale * 0.0001, scale * 0.02, num_steps, dtype=np.float64)


class Tables:
    """gaussian_diffusion.py:343-379 -- every fp64 table the samplers gather from."""

    def __init__(self, betas: np.ndarray):
        betas = np.array(betas, dtype=np.float64)
        assert betas.ndim == 1 and (betas > 0).all() and (betas <= 1).all()
        self.betas = betas
        self.num_timesteps = int(betas.shape[0])
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.alphas_cumprod_next = np.append(self.alphas_cumprod[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(
            np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)


def gather(arr: np.ndarray, t: Tensor, ndim: int) -> Tensor:
    """gaussian_diffusion.py:1168-1181 -- fp64 table -> index -> .float() -> [B,1,1..]."""
    res = torch.from_numpy(arr)[t].float()
    while res.dim() < ndim:
        res = res[..., None]
    return res


# ----------------------------------------------------------------------------------------
# network pieces
# ----------------------------------------------------------------------------------------
def timestep_embedding(t: Tensor, dim: int, max_period: int = 10000) -> Tensor:
    """transformer.py:8-25 -- [cos | sin], fp32 frequencies."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(t.device)   # CPU exp, then moved (:19-20)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _lin(sd: StateDict, name: str, x: Tensor) -> Tensor:
    return F.linear(x, sd[name + ".weight"].to(x.dtype), sd[name + ".bias"].to(x.dtype))


def _ln(sd: StateDict, name: str, x: Tensor) -> Tensor:
    w = sd[name + ".weight"].to(x.dtype)
    return F.layer_norm(x, (w.shape[0],), w, sd[name + ".bias"].to(x.dtype), LN_EPS)


def stylization(sd: StateDict, p: str, h: Tensor, emb: Tensor) -> Tensor:
    """transformer.py:68-81 -- FiLM: LN(h)*(1+scale)+shift -> SiLU -> Linear."""
    emb_out = _lin(sd, p + ".emb_layers.1", F.silu(emb))
    scale, shift = torch.chunk(emb_out, 2, dim=2)
    h = _ln(sd, p + ".norm", h) * (1 + scale) + shift
    return _lin(sd, p + ".out_layers.2", F.silu(h))


def linear_self_attention(sd: StateDict, p: str, x: Tensor, emb: Tensor, src_mask: Tensor, H: int) -> Tensor:
    """transformer.py:96-123 -- softmax_hd(Q), softmax_T(K), K^T V, Q.A, stylization, residual."""
    B, T, D = x.shape
    n = _ln(sd, p + ".norm", x)
    q = _lin(sd, p + ".query", n)
    k = _lin(sd, p + ".key", n) + (1 - src_mask) * -1000000
    q = F.softmax(q.view(B, T, H, -1), dim=-1)
    k = F.softmax(k.view(B, T, H, -1), dim=1)
    v = (_lin(sd, p + ".value", n) * src_mask).view(B, T, H, -1)
    att = torch.einsum("bnhd,bnhl->bhdl", k, v)
    y = torch.einsum("bnhd,bhdl->bnhl", q, att).reshape(B, T, D)
    return x + stylization(sd, p + ".proj_out", y, emb)


def cross_attention_kv(sd: StateDict, p: str, xf: Tensor, H: int) -> Tensor:
    """transformer.py:149-155 -- the step-invariant half: softmax_N(K)^T V  -> (B,H,hd,hd)."""
    B, N, _ = xf.shape
    nt = _ln(sd, p + ".text_norm", xf)
    k = F.softmax(_lin(sd, p + ".key", nt).view(B, N, H, -1), dim=1)
    v = _lin(sd, p + ".value", nt).view(B, N, H, -1)
    return torch.einsum("bnhd,bnhl->bhdl", k, v)


def linear_cross_attention(sd: StateDict, p: str, x: Tensor, xf: Tensor, emb: Tensor, H: int) -> Tensor:
    """transformer.py:138-158."""
    B, T, D = x.shape
    q = F.softmax(_lin(sd, p + ".query", _ln(sd, p + ".norm", x)).view(B, T, H, -1), dim=-1)
    att = cross_attention_kv(sd, p, xf, H)
    y = torch.einsum("bnhd,bhdl->bnhl", q, att).reshape(B, T, D)
    return x + stylization(sd, p + ".proj_out", y, emb)


def ffn(sd: StateDict, p: str, x: Tensor, emb: Tensor) -> Tensor:
    """transformer.py:170-173 -- exact-erf GELU, no pre-norm."""
    y = _lin(sd, p + ".linear2", F.gelu(_lin(sd, p + ".linear1", x)))
    return x + stylization(sd, p + ".proj_out", y, emb)


def src_mask_from_length(T: int, length: Sequence[int]) -> Tensor:
    """transformer.py:461-467 -- ones, zero where j >= length[i]."""
    ar = torch.arange(T)[None, :]
    return (ar < torch.as_tensor(list(length))[:, None]).float()


def num_layers_of(sd: StateDict) -> int:
    return 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("temporal_decoder_blocks."))


def motion_transformer_forward(sd: StateDict, x: Tensor, timesteps: Tensor, length: Sequence[int],
                               xf_proj: Tensor, xf_out: Tensor, num_heads: int = 8,
                               dtype: torch.dtype = torch.float32,
                               collect: Optional[List[Tensor]] = None) -> Tensor:
    """transformer.py:469-497 (with xf_proj/xf_out already produced by encode_music)."""
    B, T = x.shape[0], x.shape[1]
    D = sd["joint_embed.weight"].shape[0]
    xp = _lin(sd, "linear", xf_proj.to(dtype))
    xf = _lin(sd, "linear", xf_out.to(dtype))
    te = timestep_embedding(timesteps, D).to(dtype)            # forced-fp32 sinusoid (Q7)
    te = _lin(sd, "time_embed.2", F.silu(_lin(sd, "time_embed.0", te)))
    emb = te.unsqueeze(1) + xp
    if x.dim() == 4:
        x = torch.flatten(x, start_dim=2, end_dim=3)
    h = _lin(sd, "joint_embed", x.to(dtype)) + sd["sequence_embedding"].to(dtype)[None, :T, :]
    mask = src_mask_from_length(T, length).to(device=x.device, dtype=dtype).unsqueeze(-1)
    for i in range(num_layers_of(sd)):
        p = f"temporal_decoder_blocks.{i}"
        h = linear_self_attention(sd, p + ".sa_block", h, emb, mask, num_heads)
        h = linear_cross_attention(sd, p + ".ca_block", h, xf, emb, num_heads)
        h = ffn(sd, p + ".ffn", h, emb)
        if collect is not None:
            collect.append(h.clone())
    return _lin(sd, "out", h).view(B, T, -1).contiguous()


# ----------------------------------------------------------------------------------------
# music encoder (conditioning front-end; once per clip)
# ----------------------------------------------------------------------------------------
def _bn(sd: StateDict, p: str, x: Tensor) -> Tensor:
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, BN_EPS)


def _res_layer(sd: StateDict, p: str, x: Tensor, residual: bool = True) -> Tensor:
    """transformer.py:289-311 -- reflect-padded 3x3 conv + BN + ReLU (+ identity / 1x1 residual)."""
    w = sd[p + ".conv2d_layer.0.weight"]
    y = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), w, sd[p + ".conv2d_layer.0.bias"])
    y = F.relu(_bn(sd, p + ".conv2d_layer.1", y))
    if not residual:
        return y
    if w.shape[0] == w.shape[1]:
        return y + x
    r = F.conv2d(x, sd[p + ".residual.0.weight"], sd[p + ".residual.0.bias"])
    return y + _bn(sd, p + ".residual.1", r)


def music_encoder(sd: StateDict, mel: Tensor, prefix: str = "music_encoder") -> Tensor:
    """transformer.py:330-340 (eval mode) -- mel (B,3T,128) -> (B,T,64)."""
    p = prefix
    h = mel.unsqueeze(1)
    h = _res_layer(sd, p + ".conv1.0", h, residual=False)
    h = _res_layer(sd, p + ".conv1.1", h)
    h = _res_layer(sd, p + ".conv1.2", h)
    h = F.max_pool2d(h, (5, 5), (1, 2), (2, 2))
    h = _res_layer(sd, p + ".conv2.0", h)
    h = _res_layer(sd, p + ".conv2.1", h)
    h = F.max_pool2d(h, (5, 5), (3, 2), (2, 2))
    h = _res_layer(sd, p + ".conv3.0", h)
    h = _res_layer(sd, p + ".conv3.1", h)
    h = F.max_pool2d(h, (3, 3), (1, 2), (1, 1))
    h = h.transpose(1, 2).flatten(start_dim=2).transpose(1, 2)      # (B, 512, T)
    h = F.conv1d(h, sd[p + ".conv4.0.weight"], sd[p + ".conv4.0.bias"])
    h = _bn(sd, p + ".conv4.1", h)
    return h.transpose(1, 2)


def encode_music(sd: StateDict, mel: Tensor):
    """transformer.py:447-459 in eval mode: (proj(x), x)."""
    x = music_encoder(sd, mel)
    return _lin(sd, "proj", x), x


# ----------------------------------------------------------------------------------------
# samplers
# ----------------------------------------------------------------------------------------
def ddim_update(tb: Tables, x: Tensor, t: Tensor, x0: Tensor, eta: float = 0.0,
                noise: Optional[Tensor] = None) -> Tensor:
    """gaussian_diffusion.py:812-830 given pred_xstart (START_X, no clamp)."""
    nd = x.dim()
    eps = (gather(tb.sqrt_recip_alphas_cumprod, t, nd) * x - x0) / gather(tb.sqrt_recipm1_alphas_cumprod, t, nd)
    ab = gather(tb.alphas_cumprod, t, nd)
    abp = gather(tb.alphas_cumprod_prev, t, nd)
    sigma = eta * torch.sqrt((1 - abp) / (1 - ab)) * torch.sqrt(1 - ab / abp)
    if noise is None:
        noise = torch.zeros_like(x)
    mean_pred = x0 * torch.sqrt(abp) + torch.sqrt(1 - abp - sigma ** 2) * eps
    nonzero = (t != 0).float().view(-1, *([1] * (nd - 1)))
    return mean_pred + nonzero * sigma * noise


def ddpm_update(tb: Tables, x: Tensor, t: Tensor, x0: Tensor, noise: Tensor) -> Tensor:
    """gaussian_diffusion.py:426-429, 495-501, 656-664 (FIXED_SMALL variance)."""
    nd = x.dim()
    mean = gather(tb.posterior_mean_coef1, t, nd) * x0 + gather(tb.posterior_mean_coef2, t, nd) * x
    logvar = gather(tb.posterior_log_variance_clipped, t, nd)
    nonzero = (t != 0).float().view(-1, *([1] * (nd - 1)))
    return mean + nonzero * torch.exp(0.5 * logvar) * noise


def sample_loop(sd: StateDict, tb: Tables, noise: Tensor, length: Sequence[int], xf_proj: Tensor, xf_out: Tensor,
                kind: str = "ddim", eta: float = 0.0, step_noise: Optional[Tensor] = None,
                dtype: torch.dtype = torch.float32, max_steps: Optional[int] = None):
    """gaussian_diffusion.py:917-965 / 730-781 -- returns (final sample, [pred_xstart per step], [sample per step])."""
    img = noise.to(dtype)
    B = img.shape[0]
    x0s, samples = [], []
    steps = list(range(tb.num_timesteps))[::-1]
    if max_steps is not None:
        steps = steps[:max_steps]
    with torch.no_grad():
        for n, i in enumerate(steps):
            t = torch.tensor([i] * B)
            x0 = motion_transformer_forward(sd, img, t, length, xf_proj, xf_out, dtype=dtype)
            nz = None if step_noise is None else step_noise[n].to(dtype)
            if kind == "ddim":
                img = _cast_tables_update(ddim_update, tb, img, t, x0, dtype, eta=eta, noise=nz)
            else:
                img = _cast_tables_update(ddpm_update, tb, img, t, x0, dtype, noise=nz)
            x0s.append(x0)
            samples.append(img)
    return img, x0s, samples


def _cast_tables_update(fn, tb, img, t, x0, dtype, **kw):
    # coefficients are always fp32-rounded (Q11); arithmetic follows the tensor dtype
    return fn(tb, img, t, x0, **kw).to(dtype)
