"""Drop-in `GaussianDiffusion` for the sampling path (DDIM / DDPM), backed by the sm_100a library.

Mirrors the public surface of reference Diffusion_Stage/models/gaussian_diffusion.py that
DDPMTrainer.generate_music_motion uses (ddpm_trainer.py:89-97,190-200): the constructor and its
public numpy tables (:328-379), `ddim_sample[_loop[_progressive]]` (:783-965) and
`p_sample[_loop[_progressive]]` (:605-781) with the same signatures and return types, the enums
(:275-308) and `get_named_beta_schedule` (:228-252).

Fast path = what the reference trainer configures: START_X + FIXED_SMALL, cond_fn=None,
denoised_fn=None, pre_seq=transl_req=None, `model` a diffusion_conductor_b200.MotionTransformer.
Anything else raises NotImplementedError -- there is no eager fallback.  Training-side utilities
(losses, VLB, schedule samplers) are out of scope.
"""
from __future__ import annotations

import enum
import math
from typing import Optional

import numpy as np
import torch as th

from . import _lib
from .transformer import MotionTransformer


def _advance_rng_like_randn(x: th.Tensor, draws: int) -> None:
    """Leave torch's generator for x.device exactly where `draws` consecutive `th.randn_like(x)` calls would leave it,
    without materialising the (unused) samples: the reference draws one randn_like per step in ddim_sample even when
    eta == 0 (gaussian_diffusion.py:822), so a drop-in must consume the same stream.  A CUDA generator advances its
    Philox offset by a fixed amount per call of a given size: one real draw measures it, set_offset does the rest."""
    if draws <= 0:
        return
    if x.is_cuda:
        gen = th.cuda.default_generators[x.device.index if x.device.index is not None else th.cuda.current_device()]
        if hasattr(gen, "get_offset") and hasattr(gen, "set_offset"):
            before = gen.get_offset()
            th.randn_like(x)
            step = gen.get_offset() - before
            gen.set_offset(before + step * draws)
            return
    for _ in range(draws):
        th.randn_like(x)


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps):
    if schedule_name == "linear":
        scale = 1000 / num_diffusion_timesteps
        return np.linspace(scale * 0.0001, scale * 0.02, num_diffusion_timesteps, dtype=np.float64)
    if schedule_name == "cosine":
        return betas_for_alpha_bar(num_diffusion_timesteps, lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2)
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def betas_for_alpha_bar(num_diffusion_timesteps, alpha_bar, max_beta=0.999):
    n = num_diffusion_timesteps
    return np.array([min(1 - alpha_bar((i + 1) / n) / alpha_bar(i / n), max_beta) for i in range(n)])


class ModelMeanType(enum.Enum):
    PREVIOUS_X = enum.auto()
    START_X = enum.auto()
    EPSILON = enum.auto()


class ModelVarType(enum.Enum):
    LEARNED = enum.auto()
    FIXED_SMALL = enum.auto()
    FIXED_LARGE = enum.auto()
    LEARNED_RANGE = enum.auto()


class LossType(enum.Enum):
    MSE = enum.auto()
    RESCALED_MSE = enum.auto()
    KL = enum.auto()
    RESCALED_KL = enum.auto()

    def is_vb(self):
        return self in (LossType.KL, LossType.RESCALED_KL)


def _extract_into_tensor(arr, timesteps, broadcast_shape):
    res = th.from_numpy(arr).to(device=timesteps.device)[timesteps].float()
    while len(res.shape) < len(broadcast_shape):
        res = res[..., None]
    return res.expand(broadcast_shape)


class GaussianDiffusion:
    NOISE_BLOCK_BYTES = 1 << 30      # pre-drawn step noise of stochastic loops is generated in blocks of at most this size

    def __init__(self, *, betas, model_mean_type, model_var_type, loss_type, rescale_timesteps=False):
        self.model_mean_type = model_mean_type
        self.model_var_type = model_var_type
        self.loss_type = loss_type
        self.rescale_timesteps = rescale_timesteps

        betas = np.array(betas, dtype=np.float64)
        self.betas = betas
        assert len(betas.shape) == 1, "betas must be 1-D"
        assert (betas > 0).all() and (betas <= 1).all()
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
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)
        self._coef_cache = {}

    # ------------------------------------------------------------------------------------------
    def step_coefficients(self, eta: float = 0.0) -> th.Tensor:
        """[S, 8] fp32 table handed to dc_set_schedule.  Built with the same torch fp32 expressions the
        reference evaluates per step on the gathered, fp32-rounded table entries
        (gaussian_diffusion.py:557-559, 814-830, 426-429, 495-501, 663), so index handling and
        coefficient values are bit-identical."""
        key = float(eta)
        if key in self._coef_cache:
            return self._coef_cache[key]
        f = lambda a: th.from_numpy(a).float()  # noqa: E731
        t = th.arange(self.num_timesteps)
        ab, abp = f(self.alphas_cumprod), f(self.alphas_cumprod_prev)
        sigma = eta * th.sqrt((1 - abp) / (1 - ab)) * th.sqrt(1 - ab / abp)
        nonzero = (t != 0).float()
        coef = th.stack([
            f(self.sqrt_recip_alphas_cumprod),
            f(self.sqrt_recipm1_alphas_cumprod),
            th.sqrt(abp),
            th.sqrt(1 - abp - sigma ** 2),
            nonzero * sigma,
            f(self.posterior_mean_coef1),
            f(self.posterior_mean_coef2),
            nonzero * th.exp(0.5 * f(self.posterior_log_variance_clipped)),
        ], dim=1).contiguous()
        self._coef_cache[key] = coef
        return coef

    def _scale_timesteps(self, t):
        if self.rescale_timesteps:
            return t.float() * (1000.0 / self.num_timesteps)
        return t

    def q_sample(self, x_start, t, noise=None):
        if noise is None:
            noise = th.randn_like(x_start)
        return (_extract_into_tensor(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start
                + _extract_into_tensor(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise)

    # ------------------------------------------------------------------------------------------
    def _require_fast_path(self, model, denoised_fn=None, cond_fn=None, pre_seq=None, transl_req=None):
        if not isinstance(model, MotionTransformer):
            raise NotImplementedError("the B200 samplers drive diffusion_conductor_b200.MotionTransformer only")
        if self.model_mean_type is not ModelMeanType.START_X or self.model_var_type is not ModelVarType.FIXED_SMALL:
            raise NotImplementedError("only START_X + FIXED_SMALL (the configuration of DDPMTrainer, "
                                      "ddpm_trainer.py:91-97) is implemented")
        if self.rescale_timesteps:
            raise NotImplementedError("rescale_timesteps=True is not implemented (the reference trainer uses False)")
        if denoised_fn is not None or cond_fn is not None:
            raise NotImplementedError("denoised_fn / cond_fn (classifier guidance) are not implemented")
        if pre_seq is not None or transl_req is not None:
            raise NotImplementedError("pre_seq / transl_req in-painting hooks are not implemented")

    def _bind(self, model, x, model_kwargs, eta=0.0):
        """Prepare the engine for (model, conditioning, schedule); returns it."""
        model_kwargs = dict(model_kwargs or {})
        B, T = x.shape[0], x.shape[1]
        length = model_kwargs.get("length")
        if length is None:
            raise TypeError("model_kwargs['length'] is required (reference transformer.py:462)")
        xf_proj, xf_out = model_kwargs.get("xf_proj"), model_kwargs.get("xf_out")
        if xf_proj is None or xf_out is None:
            text = model_kwargs.get("text")
            if text is None:
                raise TypeError("model_kwargs needs xf_proj/xf_out (encode_music outputs) or text (mel)")
            xf_proj, xf_out = model.encode_music(text, x.device)
        eng = model.engine_for(x.device, B, T)
        eng.prepare(xf_proj, xf_out, length, B, T)
        eng.set_schedule(self.step_coefficients(eta))
        return eng

    @staticmethod
    def _as_state(x, device):
        x = x.detach().to(device=device, dtype=th.float32)
        if x.dim() == 4:
            x = th.flatten(x, start_dim=2, end_dim=3)
        return x.contiguous()

    def _uniform_step(self, t) -> Optional[int]:
        tl = t.tolist() if isinstance(t, th.Tensor) else list(t)
        return int(tl[0]) if all(v == tl[0] for v in tl) else None

    def _single_step(self, sampler, model, x, t, clip_denoised, model_kwargs, eta, noise):
        assert t.shape == (x.shape[0],)
        eng = self._bind(model, x, model_kwargs, eta)
        step = self._uniform_step(t)
        flags = sampler | (_lib.DC_FLAG_CLIP if clip_denoised else 0)
        if step is not None:
            xs = self._as_state(x, eng.device).clone()
            x0 = eng.sample_step(flags, xs, step, noise)
            return {"sample": xs.view(x.shape), "pred_xstart": x0.view(x.shape)}
        # per-sample timesteps: network through the library, update rule with the same gathered coefficients
        x0 = eng.forward(self._as_state(x, eng.device), t)
        if clip_denoised:
            x0 = x0.clamp(-1, 1)
        cf = self.step_coefficients(eta).to(x0.device)[t.to(x0.device)]
        cf = cf.view(cf.shape[0], 1, 1, 8)
        xs = self._as_state(x, eng.device)
        nz = noise if noise is not None else th.zeros_like(xs)
        if sampler == _lib.DC_SAMPLER_DDIM:
            eps = (cf[..., 0] * xs - x0) / cf[..., 1]
            sample = x0 * cf[..., 2] + cf[..., 3] * eps + cf[..., 4] * nz
        else:
            sample = cf[..., 5] * x0 + cf[..., 6] * xs + cf[..., 7] * nz
        return {"sample": sample.view(x.shape), "pred_xstart": x0.view(x.shape)}

    # ---- DDIM ---------------------------------------------------------------------------------
    def ddim_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None, eta=0.0):
        self._require_fast_path(model, denoised_fn, cond_fn)
        noise = th.randn_like(x)          # drawn every step even when eta == 0 (reference :822)
        return self._single_step(_lib.DC_SAMPLER_DDIM, model, x, t, clip_denoised, model_kwargs, eta,
                                 self._as_state(noise, x.device) if eta != 0.0 else None)

    def ddim_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                         model_kwargs=None, device=None, progress=False, eta=0.0, idxs=[], match_rng_stream=True):
        """match_rng_stream (extension, default on): consume torch's generator like the reference does -- one randn_like
        per step even at eta == 0 (:822) -- so that whatever the caller draws next matches a run of the reference."""
        return self._loop(_lib.DC_SAMPLER_DDIM, model, shape, noise, clip_denoised, denoised_fn, cond_fn, model_kwargs,
                          device, progress, eta, idxs, match_rng_stream)

    def ddim_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                                     model_kwargs=None, device=None, progress=False, eta=0.0):
        self._require_fast_path(model, denoised_fn, cond_fn)
        yield from self._progressive(_lib.DC_SAMPLER_DDIM, model, shape, noise, clip_denoised, model_kwargs, device,
                                     progress, eta)

    # ---- DDPM ---------------------------------------------------------------------------------
    def p_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, pre_seq=None, transl_req=None,
                 model_kwargs=None):
        self._require_fast_path(model, denoised_fn, cond_fn, pre_seq, transl_req)
        noise = th.randn_like(x)
        return self._single_step(_lib.DC_SAMPLER_DDPM, model, x, t, clip_denoised, model_kwargs, 0.0,
                                 self._as_state(noise, x.device))

    def p_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                      model_kwargs=None, device=None, pre_seq=None, transl_req=None, progress=False, idxs=[]):
        self._require_fast_path(model, denoised_fn, cond_fn, pre_seq, transl_req)
        return self._loop(_lib.DC_SAMPLER_DDPM, model, shape, noise, clip_denoised, denoised_fn, cond_fn, model_kwargs,
                          device, progress, 0.0, idxs, True)

    def p_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                                  model_kwargs=None, device=None, pre_seq=None, transl_req=None, progress=False):
        self._require_fast_path(model, denoised_fn, cond_fn, pre_seq, transl_req)
        yield from self._progressive(_lib.DC_SAMPLER_DDPM, model, shape, noise, clip_denoised, model_kwargs, device,
                                     progress, 0.0)

    # ---- shared loop bodies -------------------------------------------------------------------
    def _initial(self, model, shape, noise, device):
        if device is None:
            device = next(model.parameters()).device
        assert isinstance(shape, (tuple, list))
        img = noise if noise is not None else th.randn(*shape, device=device)
        return img, th.device(device)

    def _progressive(self, sampler, model, shape, noise, clip_denoised, model_kwargs, device, progress, eta):
        """Step-at-a-time generator with the reference's yield protocol (:917-965, :730-781)."""
        img, device = self._initial(model, shape, noise, device)
        eng = self._bind(model, img, model_kwargs, eta)
        x = self._as_state(img, eng.device).clone()
        indices = list(range(self.num_timesteps))[::-1]
        if progress:
            from tqdm.auto import tqdm
            indices = tqdm(indices)
        flags = sampler | (_lib.DC_FLAG_CLIP if clip_denoised else 0)
        stochastic = sampler == _lib.DC_SAMPLER_DDPM or eta != 0.0
        for i in indices:
            nz = th.randn_like(x)             # the reference draws one randn_like per step in both samplers
            x0 = eng.sample_step(flags, x, i, nz if stochastic else None)
            yield {"sample": x.clone().view(img.shape), "pred_xstart": x0.view(img.shape)}

    def _loop(self, sampler, model, shape, noise, clip_denoised, denoised_fn, cond_fn, model_kwargs, device, progress,
              eta, idxs, match_rng_stream):
        """Whole-loop entry: one C call = one launch of the persistent kernel for all S steps (dc_sample_loop)."""
        self._require_fast_path(model, denoised_fn, cond_fn)
        img, device = self._initial(model, shape, noise, device)
        eng = self._bind(model, img, model_kwargs, eta)
        x = self._as_state(img, eng.device).clone()
        S = self.num_timesteps
        stochastic = sampler == _lib.DC_SAMPLER_DDPM or eta != 0.0
        trace = th.empty((S,) + tuple(x.shape), device=x.device, dtype=th.float32) if len(idxs) else None
        flags = sampler | (_lib.DC_FLAG_CLIP if clip_denoised else 0)
        if progress:
            from tqdm.auto import tqdm
            bar = tqdm(total=S)
        if stochastic:
            # The reference draws one randn_like per step from torch's generator; here the same draws, in the same order
            # (normal_ on a contiguous slice consumes the generator exactly like randn_like of that shape), land in a noise
            # buffer that is bounded to NOISE_BLOCK_BYTES: the loop runs as ceil(S / block) launches of `block` steps each.
            block = max(1, min(S, self.NOISE_BLOCK_BYTES // max(1, x.numel() * 4)))
            buf = th.empty((block,) + tuple(x.shape), device=x.device, dtype=th.float32)
            done = 0
            while done < S:
                n = min(block, S - done)
                for i in range(n):
                    buf[i].normal_()
                eng.sample_range(flags, x, S - 1 - done, n, step_noise=buf[:n],
                                 trace_x=None if trace is None else trace[done:done + n])
                done += n
                if progress:
                    bar.update(n)
        else:
            if match_rng_stream:
                _advance_rng_like_randn(x, S)
            eng.sample_loop(flags, x, trace_x=trace, num_steps=S)
            if progress:
                bar.update(S)
        if progress:
            bar.close()
        final = x.view(img.shape)
        if len(idxs) == 0:
            return final
        result = {i: trace[i].view(img.shape) for i in range(S) if i in idxs}
        result[S] = final
        return result
