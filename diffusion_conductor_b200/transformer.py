"""Drop-in `MotionTransformer` for the denoising hot path.

Same constructor, attributes, `forward` / `encode_music` / `generate_src_mask` signatures and the
same 396-entry `state_dict` as reference Diffusion_Stage/models/transformer.py:360-497, so a
reference checkpoint (`checkpoint['encoder']`, ddpm_trainer.py:303-319) loads unchanged.  The nn
sub-modules below only *hold* parameters under the reference's names; the arithmetic of the
decoder runs in the sm_100a library through the C ABI (include/dc_b200.h).  The music-encoder CNN
(reference :289-357), the once-per-clip front-end outside the step loop, runs there too in eval mode
(dc_encode_music); the `MusicEncoder` module below holds its parameters and is only
executed by PyTorch in training mode (condition dropout), which is outside the sampling path.

Unsupported on purpose (raise, never fall back): `no_eff=True` (quadratic attention variant),
autograd through `forward`, CPU tensors.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence

import os

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib


def timestep_embedding(timesteps: torch.Tensor, dim: int, max_period: int = 10000) -> torch.Tensor:
    """Sinusoidal embedding, [cos | sin], frequencies built in fp32 on the CPU (reference :8-25)."""
    half = dim // 2
    freqs = timestep_frequencies(dim, max_period).to(device=timesteps.device)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def timestep_frequencies(dim: int, max_period: int = 10000) -> torch.Tensor:
    half = dim // 2
    return torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half)


def _as_int_list(length) -> list:
    """`length` as the reference callers pass it: a python list, or a (CUDA) LongTensor (ddpm_trainer.py:198) -- one
    device->host copy instead of one per element."""
    if isinstance(length, torch.Tensor):
        return [int(v) for v in length.detach().reshape(-1).tolist()]
    return [int(v) for v in length]


def _zero_(module: nn.Module) -> nn.Module:
    for p in module.parameters():
        p.detach().zero_()
    return module


# ------------------------------------------------------------------------------------------------
# parameter containers with the reference's attribute names (forward lives in the CUDA library)
# ------------------------------------------------------------------------------------------------
class _Stylization(nn.Module):
    def __init__(self, latent_dim: int, time_embed_dim: int, dropout: float):
        super().__init__()
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(time_embed_dim, 2 * latent_dim))
        self.norm = nn.LayerNorm(latent_dim)
        self.out_layers = nn.Sequential(nn.SiLU(), nn.Dropout(p=dropout), _zero_(nn.Linear(latent_dim, latent_dim)))


class _SelfAttention(nn.Module):
    def __init__(self, latent_dim, num_head, dropout, time_embed_dim):
        super().__init__()
        self.num_head = num_head
        self.norm = nn.LayerNorm(latent_dim)
        self.query = nn.Linear(latent_dim, latent_dim)
        self.key = nn.Linear(latent_dim, latent_dim)
        self.value = nn.Linear(latent_dim, latent_dim)
        self.dropout = nn.Dropout(dropout)
        self.proj_out = _Stylization(latent_dim, time_embed_dim, dropout)


class _CrossAttention(nn.Module):
    def __init__(self, latent_dim, text_latent_dim, num_head, dropout, time_embed_dim):
        super().__init__()
        self.num_head = num_head
        self.norm = nn.LayerNorm(latent_dim)
        self.text_norm = nn.LayerNorm(text_latent_dim)
        self.query = nn.Linear(latent_dim, latent_dim)
        self.key = nn.Linear(text_latent_dim, latent_dim)
        self.value = nn.Linear(text_latent_dim, latent_dim)
        self.dropout = nn.Dropout(dropout)
        self.proj_out = _Stylization(latent_dim, time_embed_dim, dropout)


class _FeedForward(nn.Module):
    def __init__(self, latent_dim, ffn_dim, dropout, time_embed_dim):
        super().__init__()
        self.linear1 = nn.Linear(latent_dim, ffn_dim)
        self.linear2 = _zero_(nn.Linear(ffn_dim, latent_dim))
        self.activation = nn.GELU()
        self.dropout = nn.Dropout(dropout)
        self.proj_out = _Stylization(latent_dim, time_embed_dim, dropout)


class _DecoderLayer(nn.Module):
    def __init__(self, latent_dim, text_latent_dim, time_embed_dim, ffn_dim, num_head, dropout):
        super().__init__()
        self.sa_block = _SelfAttention(latent_dim, num_head, dropout, time_embed_dim)
        self.ca_block = _CrossAttention(latent_dim, text_latent_dim, num_head, dropout, time_embed_dim)
        self.ffn = _FeedForward(latent_dim, ffn_dim, dropout, time_embed_dim)


class _ConvRes(nn.Module):
    """reflect-padded 3x3 conv + BN + ReLU with identity / 1x1-conv residual (reference :289-311)."""

    def __init__(self, cin, cout, residual=True):
        super().__init__()
        self.conv2d_layer = nn.Sequential(
            nn.Conv2d(cin, cout, kernel_size=(3, 3), stride=(1, 1), padding=(1, 1), padding_mode="reflect"),
            nn.BatchNorm2d(cout), nn.ReLU())
        self._mode = "none" if not residual else ("identity" if cin == cout else "conv")
        if self._mode == "conv":
            self.residual = nn.Sequential(nn.Conv2d(cin, cout, kernel_size=1, stride=1), nn.BatchNorm2d(cout))

    def forward(self, x):
        y = self.conv2d_layer(x)
        if self._mode == "identity":
            return y + x
        if self._mode == "conv":
            return y + self.residual(x)
        return y


class MusicEncoder(nn.Module):
    """M2SNet music encoder: mel (B, 3T, 128) -> (B, T, 64) (reference :313-340)."""

    def __init__(self, device):
        super().__init__()
        self.device = device
        self.conv1 = nn.Sequential(_ConvRes(1, 16, residual=False), _ConvRes(16, 16), _ConvRes(16, 16),
                                   nn.MaxPool2d(kernel_size=(5, 5), stride=(1, 2), padding=(2, 2)))
        self.conv2 = nn.Sequential(_ConvRes(16, 32), _ConvRes(32, 32),
                                   nn.MaxPool2d(kernel_size=(5, 5), stride=(3, 2), padding=(2, 2)))
        self.conv3 = nn.Sequential(_ConvRes(32, 32), _ConvRes(32, 32),
                                   nn.MaxPool2d(kernel_size=(3, 3), stride=(1, 2), padding=(1, 1)))
        self.conv4 = nn.Sequential(nn.Conv1d(32 * 16, 64, kernel_size=1, stride=1), nn.BatchNorm1d(64))

    def forward(self, x):
        mel = x.unsqueeze(1).to(self.device)
        # keep the CNN in true fp32 (no TF32 convolutions): the reference's features are fp32
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            h = self.conv3(self.conv2(self.conv1(mel)))
            h = h.transpose(1, 2).flatten(start_dim=2).transpose(1, 2)
            return self.conv4(h).transpose(1, 2)


# ------------------------------------------------------------------------------------------------
class MotionTransformer(nn.Module):
    def __init__(self, input_feats, num_frames=240, latent_dim=16, ff_size=64, num_layers=8, num_heads=8, dropout=0,
                 activation="gelu", device="cuda", text_num_heads=4,
                 music_model_path="/home/zhuoran/DiffuseConductor/Diffusion_Stage/stage_one_checkpoints/M2SNet_latest.pt",
                 no_eff=False, operand_dtype="bf16", **kargs):
        super().__init__()
        if no_eff:
            raise NotImplementedError("no_eff=True (quadratic TemporalSelfAttention) is outside the B200 hot path; "
                                      "only the Linear* blocks are implemented")
        if operand_dtype not in ("bf16", "fp16"):
            raise ValueError("operand_dtype must be 'bf16' or 'fp16'")
        self.num_frames = num_frames
        self.latent_dim = latent_dim
        self.ff_size = ff_size
        self.num_layers = num_layers
        self.num_heads = num_heads
        self.dropout = dropout
        self.activation = activation
        self.input_feats = input_feats
        self.time_embed_dim = latent_dim * 4
        self.operand_dtype = operand_dtype
        self.sequence_embedding = nn.Parameter(torch.randn(num_frames, latent_dim))
        self.device = device
        self.cond_mask_prob = 0.1

        self.music_encoder = MusicEncoder(device=device)
        if music_model_path is not None:     # warm start from an M2SNet checkpoint (reference :394-401)
            base = torch.load(music_model_path)
            sub = {k.replace("module.music_encoder.", ""): v for k, v in base.items()
                   if k.startswith("module.music_encoder")}
            self.music_encoder.load_state_dict(sub, strict=False)
        self.music_encoder.eval()
        self.linear = nn.Linear(64, 512)
        music_latent_dim = 512
        self.joint_embed = nn.Linear(self.input_feats, self.latent_dim)
        self.time_embed = nn.Sequential(nn.Linear(self.latent_dim, self.time_embed_dim), nn.SiLU(),
                                        nn.Linear(self.time_embed_dim, self.time_embed_dim))
        self.temporal_decoder_blocks = nn.ModuleList(
            _DecoderLayer(latent_dim, music_latent_dim, self.time_embed_dim, ff_size, num_heads, dropout)
            for _ in range(num_layers))
        self.out = _zero_(nn.Linear(self.latent_dim, self.input_feats))
        self.proj = nn.Linear(64, 64)

        self._engine: Optional[_Engine] = None
        self._weights_dirty = True

    # ---- weight bookkeeping -------------------------------------------------------------------
    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._weights_dirty = True
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._weights_dirty = True
        return out

    def refresh_weights(self):
        """Call after editing parameters in place so the packed device copies are rebuilt."""
        self._weights_dirty = True

    def engine(self, device: torch.device) -> "_Engine":
        if device.type != "cuda":
            raise RuntimeError("diffusion_conductor_b200 runs on CUDA (sm_100a) only; got a tensor on %s" % device)
        index = device.index if device.index is not None else torch.cuda.current_device()
        if self._engine is None or self._engine.device_index != index:
            self._engine = _Engine(self, index)
            self._weights_dirty = True
        if self._weights_dirty:
            self._engine.upload(self)
            self._weights_dirty = False
        return self._engine

    def engine_for(self, device: torch.device, B: int, T: int):
        """Engine for a (B, T) batch: one handle serves any batch size.  Clips of up to 2048 frames run on the
        cluster-per-clip persistent kernel (clusters are scheduled by the hardware as SMs free up); longer clips --
        possible only when the model was built with num_frames > 2048 -- use the per-layer launch path of the same
        handle (a captured CUDA graph per 5 steps)."""
        return self.engine(device)

    # ---- reference API ------------------------------------------------------------------------
    def encode_music(self, text, device):
        """mel (B, 3T, 128) -> (xf_proj, xf_out), each (B, T, 64) (reference :447-459).  Eval mode runs the M2SNet CNN and
        `proj` in the CUDA library (dc_encode_music, exact fp32).  Training mode (10 % per-frame condition dropout,
        reference :451-456) is outside the sampling path and keeps the plain PyTorch modules."""
        if self.training:
            with torch.no_grad():
                x = self.music_encoder(text)
            b, t, _ = x.shape
            mask = torch.bernoulli(torch.ones((b, t), device=device) * self.cond_mask_prob).view((b, t, 1))
            x = x * (1 - mask)
            return self.proj(x), x
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("encode_music runs on the CUDA device only (there is no CPU fallback)")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        return self.engine(dev).encode_music(torch.as_tensor(text))

    def generate_src_mask(self, T, length):
        ar = torch.arange(T)[None, :]
        return (ar < torch.as_tensor(_as_int_list(length))[:, None]).float()

    def forward(self, x, timesteps, length=None, text=None, xf_proj=None, xf_out=None):
        """x: (B,T,26) or (B,T,13,2); timesteps: (B,) integer; returns predicted x0 (B,T,26)."""
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            if x.requires_grad:
                raise NotImplementedError("autograd through the B200 denoiser is not implemented (sampling path only)")
        B, T = x.shape[0], x.shape[1]
        if text is not None and len(text) != B:      # DataParallel leftover (reference :474-476)
            index = x.device.index
            text = text[index * B: index * B + B]
        if xf_proj is None or xf_out is None:
            xf_proj, xf_out = self.encode_music(text, x.device)
        if x.dim() == 4:
            x = torch.flatten(x, start_dim=2, end_dim=3)
        if length is None:
            raise TypeError("length is required (the reference calls len(length), transformer.py:462)")
        eng = self.engine_for(x.device, B, T)
        with torch.no_grad():
            eng.prepare(xf_proj, xf_out, length, B, T)
            return eng.forward(x, timesteps)


class _Engine:
    """Owns one dc_handle on one GPU and keeps it in sync with a MotionTransformer's parameters."""

    def __init__(self, model: MotionTransformer, device_index: int):
        self.lib = _lib.load()
        self.device_index = device_index
        self.device = torch.device("cuda", device_index)
        cfg = _lib.DcConfig(model.input_feats, model.num_frames, model.latent_dim, model.ff_size, model.num_layers,
                            model.num_heads, device_index,
                            _lib.DC_OPERAND_BF16 if model.operand_dtype == "bf16" else _lib.DC_OPERAND_FP16)
        handle = C.c_void_p()
        _lib.check(self.lib.dc_create(C.byref(cfg), C.byref(handle)))
        self.handle = handle
        self._cond_key = None
        self._schedule_key = None
        self.B = self.T = self.S = 0

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.dc_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def _ck(self, rc):
        _lib.check(rc, self.handle)

    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def upload(self, model: MotionTransformer):
        sd = model.state_dict()
        sd = dict(sd)
        sd["aux.timestep_freqs"] = timestep_frequencies(model.latent_dim)
        for key, val in sd.items():
            if key.endswith("num_batches_tracked"):
                continue
            t = val.detach().to(torch.float32).contiguous()
            shape = (C.c_int64 * t.dim())(*t.shape)
            self._ck(self.lib.dc_set_weight(self.handle, key.encode(), C.c_void_p(t.data_ptr()), shape, t.dim()))
        self._ck(self.lib.dc_finalize_weights(self.handle))
        self._cond_key = None

    def set_schedule(self, coef: torch.Tensor):
        """Upload the [S, 8] coefficient table unless it is, value for value, the one already in the library.  The cache
        is keyed on the table CONTENT: object ids are reused by CPython, and a stale table with another S would make the
        library index noise / trace buffers that Python sized for the new S."""
        coef = coef.detach().to("cpu", torch.float32).contiguous()
        assert coef.dim() == 2 and coef.shape[1] == 8
        key = (int(coef.shape[0]), coef.numpy().tobytes())
        if key == self._schedule_key:
            return
        self._ck(self.lib.dc_set_schedule(self.handle, coef.shape[0], C.c_void_p(coef.data_ptr())))
        self._schedule_key = key
        self.S = int(coef.shape[0])

    @staticmethod
    def _f32(t: torch.Tensor, device) -> torch.Tensor:
        return t.detach().to(device=device, dtype=torch.float32).contiguous()

    def prepare(self, xf_proj, xf_out, length: Sequence[int], B: int, T: int):
        xf_proj = self._f32(xf_proj, self.device)
        xf_out = self._f32(xf_out, self.device)
        if tuple(xf_proj.shape) != (B, T, 64) or tuple(xf_out.shape) != (B, T, 64):
            raise ValueError(f"xf_proj/xf_out must be (B,T,64)=({B},{T},64); music frames must equal motion frames "
                             f"(got {tuple(xf_proj.shape)}, {tuple(xf_out.shape)})")
        length = _as_int_list(length)
        if len(length) != B:
            raise ValueError("len(length) must equal the batch size")
        key = (xf_proj.data_ptr(), xf_proj._version, xf_out.data_ptr(), xf_out._version, B, T, tuple(length))
        if key == self._cond_key:
            return
        arr = (C.c_int64 * B)(*length)
        self._ck(self.lib.dc_prepare_cond(self.handle, C.c_void_p(xf_proj.data_ptr()), C.c_void_p(xf_out.data_ptr()), arr,
                                          B, T, self.stream()))
        self._cond_key = key
        self._keepalive = (xf_proj, xf_out)
        self.B, self.T = B, T

    def encode_music(self, mel: torch.Tensor):
        """dc_encode_music: mel (B, Tm, 128) -> (xf_proj, xf_out) (B, T, 64) on this engine's device."""
        dev = torch.device("cuda", self.device_index)
        if mel.dim() != 3 or mel.shape[2] != 128:
            raise ValueError(f"mel must be (B, 3T, 128), got {tuple(mel.shape)}")
        m = self._f32(mel, dev)
        B, Tm = int(m.shape[0]), int(m.shape[1])
        T = (Tm - 1) // 3 + 1
        xf_proj = torch.empty(B, T, 64, device=dev)
        xf_out = torch.empty(B, T, 64, device=dev)
        with torch.cuda.device(dev):
            self._ck(self.lib.dc_encode_music(self.handle, m.data_ptr(), xf_proj.data_ptr(), xf_out.data_ptr(), B, Tm, self.stream()))
        return xf_proj, xf_out

    def forward(self, x: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        x = self._f32(x, self.device)
        t = timesteps.detach().to(device=self.device, dtype=torch.int64).contiguous()
        if tuple(x.shape) != (self.B, self.T, 26) or tuple(t.shape) != (self.B,):
            raise ValueError(f"x must be ({self.B},{self.T},26) and timesteps ({self.B},)")
        out = torch.empty_like(x)
        self._ck(self.lib.dc_forward(self.handle, C.c_void_p(x.data_ptr()), C.c_void_p(t.data_ptr()),
                                     C.c_void_p(out.data_ptr()), self.stream()))
        return out

    def sample_step(self, sampler: int, x: torch.Tensor, step: int, noise: Optional[torch.Tensor]):
        """In place on x; returns pred_xstart."""
        x0 = torch.empty_like(x)
        nz = C.c_void_p(noise.data_ptr()) if noise is not None else None
        self._ck(self.lib.dc_sample_step(self.handle, sampler, C.c_void_p(x.data_ptr()), C.c_void_p(x0.data_ptr()), step, nz,
                                         self.stream()))
        return x0

    def sample_loop(self, sampler: int, x: torch.Tensor, step_noise=None, trace_x0=None, trace_x=None, num_steps=None):
        """num_steps: the S the caller sized step_noise / traces for (defaults to their leading dimension, else the
        schedule last uploaded through this engine); the library refuses a mismatch with its own schedule."""
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
        if num_steps is None:
            sized = [t for t in (step_noise, trace_x0, trace_x) if t is not None]
            num_steps = int(sized[0].shape[0]) if sized else self.S
        for t in (step_noise, trace_x0, trace_x):
            if t is not None and (int(t.shape[0]) != num_steps or t.numel() != num_steps * x.numel()):
                raise ValueError(f"per-step buffers must be [{num_steps}, *x.shape]; got {tuple(t.shape)}")
        self._ck(self.lib.dc_sample_loop(self.handle, sampler, int(num_steps), p(x), p(step_noise), p(trace_x0), p(trace_x), self.stream()))

    def sample_range(self, sampler: int, x: torch.Tensor, step0: int, n_steps: int, step_noise=None, trace_x0=None, trace_x=None):
        """Steps step0, step0 - 1, ... (n_steps of them) in one launch; per-step buffers are [n_steps, *x.shape]."""
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
        for t in (step_noise, trace_x0, trace_x):
            if t is not None and t.numel() != n_steps * x.numel():
                raise ValueError(f"per-step buffers must be [{n_steps}, *x.shape]; got {tuple(t.shape)}")
        self._ck(self.lib.dc_sample_range(self.handle, sampler, int(step0), int(n_steps), p(x), p(step_noise), p(trace_x0), p(trace_x),
                                          self.stream()))

    def time_embedding(self, timesteps: torch.Tensor) -> torch.Tensor:
        """time_embed(timestep_embedding(t, latent_dim)) -> (n, 512) (reference transformer.py:8-25, 410-414, 482)."""
        t = timesteps.detach().to(device=self.device, dtype=torch.int64).contiguous()
        out = torch.empty(t.numel(), 512, device=self.device)
        self._ck(self.lib.dc_time_embedding(self.handle, C.c_void_p(t.data_ptr()), t.numel(), C.c_void_p(out.data_ptr()), self.stream()))
        return out

    def cluster_occupancy(self, tiles_per_clip: int) -> int:
        n = C.c_int(0)
        self._ck(self.lib.dc_cluster_occupancy(self.handle, int(tiles_per_clip), C.byref(n)))
        return int(n.value)

    def sampler_update(self, sampler: int, x: torch.Tensor, x0: torch.Tensor, step: int, noise=None):
        nz = C.c_void_p(noise.data_ptr()) if noise is not None else None
        self._ck(self.lib.dc_sampler_update(self.handle, sampler, C.c_void_p(x.data_ptr()), C.c_void_p(x0.data_ptr()), step, nz,
                                            x.numel(), self.stream()))

    def generate_host(self, sampler, xf_proj, xf_out, length, noise, out, B, T):
        arr = (C.c_int64 * B)(*_as_int_list(length)) if length is not None else None
        self._ck(self.lib.dc_generate_host(self.handle, sampler, C.c_void_p(xf_proj.data_ptr()), C.c_void_p(xf_out.data_ptr()),
                                           arr, C.c_void_p(noise.data_ptr()), C.c_void_p(out.data_ptr()), B, T, self.stream()))
        self._cond_key = None
        self.B, self.T = B, T

    def profile_step(self, sampler: int, x: torch.Tensor, step: int):
        """One denoise step with per-kernel-class event timing: ({class: ms}, {class: launches})."""
        ms = (C.c_float * 4)()
        cnt = (C.c_int * 4)()
        self._ck(self.lib.dc_profile_step(self.handle, sampler, C.c_void_p(x.data_ptr()), step, ms, cnt, self.stream()))
        names = ("step_begin", "layer", "kv_reduce", "out_update")
        return {n: float(ms[i]) for i, n in enumerate(names)}, {n: int(cnt[i]) for i, n in enumerate(names)}

    def kernel_launches(self) -> int:
        return int(self.lib.dc_kernel_launches(self.handle))

    def set_graphs(self, enabled: bool):
        self._ck(self.lib.dc_set_graphs(self.handle, 1 if enabled else 0))
