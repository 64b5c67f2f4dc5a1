"""Batched, multi-GPU generation driver: the B200 counterpart of DDPMTrainer.generate_music_motion
(reference Diffusion_Stage/trainers/ddpm_trainer.py:183-201, which handles one clip, B = 1).

Clips are independent through the whole sampling loop (no cross-batch op in the reference path), so
the batch is split contiguously across ranks, every rank runs the captured loop on its shard with no
communication, and the generated motion is collected with ONE all_gather (NCCL over NVLink on GPUs,
gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of rank's items; the first n_items % world ranks get one extra."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_shards(local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """all_gather of per-rank shards with (possibly) unequal leading sizes -> (n_items, ...) on every rank."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    if world == 1:
        return local
    sizes = [shard_range(n_items, r, world) for r in range(world)]
    cap = max(b - a for a, b in sizes)
    if all(b - a == cap for a, b in sizes):
        # equal shards (the usual case): ONE collective straight into the (n_items, ...) result, no staging copies
        out = local.new_empty((n_items,) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = local.new_zeros((cap,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    buf = local.new_empty((world * cap,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(buf, pad, group=group)
    return torch.cat([buf[r * cap: r * cap + (b - a)] for r, (a, b) in enumerate(sizes)], dim=0)


def sharded_sample(sample_fn: Callable[[torch.Tensor, torch.Tensor, torch.Tensor, List[int]], torch.Tensor],
                   xf_proj: torch.Tensor, xf_out: torch.Tensor, noise: torch.Tensor, length: Sequence[int],
                   group=None) -> torch.Tensor:
    """Split (xf_proj, xf_out, noise, length) along the clip dimension, run `sample_fn` on this rank's
    shard, all_gather.  `sample_fn(xf_proj, xf_out, noise, length) -> (b, T, P)`."""
    B = noise.shape[0]
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    a, b = shard_range(B, rank, world)
    length = [int(v) for v in length]
    if b > a:
        local = sample_fn(xf_proj[a:b], xf_out[a:b], noise[a:b], length[a:b])
    else:
        local = noise.new_zeros((0,) + tuple(noise.shape[1:]))
    return gather_shards(local, B, group)


def generate_music_motion(model, diffusion, music_mel, dim_pose: int = 26, length: Optional[Sequence[int]] = None,
                          noise: Optional[torch.Tensor] = None, sampler: str = "ddim", eta: float = 0.0,
                          idxs: Sequence[int] = (), progress: bool = False, device=None, group=None):
    """mel (B, 3T, 128) or (3T, 128) -> motion (B, T, dim_pose).

    Mirrors ddpm_trainer.py:183-201 (encode_music, then ddim_sample_loop with clip_denoised=False and
    model_kwargs {xf_proj, xf_out, length}) and extends it to batches, ragged `length`, and sharding over
    the ranks of an initialised process group."""
    if device is None:
        device = next(model.parameters()).device
    mel = torch.as_tensor(music_mel)
    if mel.dim() == 2:
        mel = mel.unsqueeze(0)
    mel = mel.to(device=device, dtype=torch.float32)
    # motion frames = music-feature frames: the reference takes T from xf_proj.shape[1] (ddpm_trainer.py:187-188); the
    # encoder's stride-3 max-pool (kernel 5, padding 2) yields (Tm - 1) // 3 + 1 frames for Tm mel frames
    B, T = mel.shape[0], (mel.shape[1] - 1) // 3 + 1
    if length is None:
        length = [T] * B
    else:
        length = [int(v) for v in (length.tolist() if isinstance(length, torch.Tensor) else length)]
        if len(length) != B:
            raise ValueError(f"len(length)={len(length)} must equal the number of clips {B}")
    if noise is None:
        noise = torch.randn(B, T, dim_pose, device=device)
    elif tuple(noise.shape) != (B, T, dim_pose):
        raise ValueError(f"noise must be {(B, T, dim_pose)} for {mel.shape[1]} mel frames, got {tuple(noise.shape)}")
    if len(idxs) and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        raise NotImplementedError("idxs (intermediate samples) is single-process only")

    def sample_fn(mel_s, _unused, noise_s, length_s):
        with torch.no_grad():
            xf_proj, xf_out = model.encode_music(mel_s, device)
        kw = {"xf_proj": xf_proj, "xf_out": xf_out, "length": length_s}
        shape = tuple(noise_s.shape)
        if sampler == "ddim":
            return diffusion.ddim_sample_loop(model, shape, noise=noise_s, clip_denoised=False, progress=progress,
                                              model_kwargs=kw, eta=eta, idxs=list(idxs))
        return diffusion.p_sample_loop(model, shape, noise=noise_s, clip_denoised=False, progress=progress,
                                       model_kwargs=kw, idxs=list(idxs))

    if len(idxs):
        return sample_fn(mel, None, noise, length)
    return sharded_sample(sample_fn, mel, mel, noise, length, group)


# ------------------------------------------------------------------------------------------------
# Post-processing on device: reference tools/visualization.py:20-26 (smooth_motion) and
# :107-126 (vis_motion: reshape to (T, 13, 2), * window pixels, Savitzky-Golay kernel 19 / order 5).
# ------------------------------------------------------------------------------------------------
def savgol_matrices(kernel: int, order: int):
    """(fir [kernel], edge [kernel // 2, kernel]) of scipy.signal.savgol_filter(x, kernel, order) with its default
    mode='interp': the interior is a symmetric FIR (least-squares polynomial evaluated at the window centre); the first
    kernel // 2 frames are the polynomial fitted to the first `kernel` frames evaluated at frames 0 .. kernel // 2 - 1
    (the tail is the time-reversed mirror image).  Computed in float64 with numpy only."""
    import numpy as np

    if kernel % 2 != 1 or kernel < 3:
        raise ValueError("kernel must be an odd integer >= 3")
    if order >= kernel:
        raise ValueError("polyorder must be less than window_length.")       # scipy's message
    pos = np.arange(kernel, dtype=np.float64)
    A = np.vander(pos, order + 1, increasing=True)                            # [kernel, order + 1]
    H = A @ np.linalg.pinv(A)                                                 # hat matrix: fitted value at pos i <- data
    half = kernel // 2
    return H[half].copy(), H[:half].copy()


def smooth_motion(motion: torch.Tensor, kernel: int = 19, order: int = 5, window: float = 600.0) -> torch.Tensor:
    """Batched, on-device version of the reference post-processing: motion (B, T, 26) or (T, 26) keypoints in [0, 1]
    -> (B, T, 13, 2) pixel coordinates, Savitzky-Golay-smoothed along time (reference smooth_motion is called with
    kernel=19, order=5 after `motions[i] *= 600`, visualization.py:117-120).  Runs in the CUDA library; there is no
    CPU path."""
    from . import _lib

    squeeze = motion.dim() == 2
    m = motion.unsqueeze(0) if squeeze else motion
    if not m.is_cuda:
        raise RuntimeError("smooth_motion runs on the CUDA device only (there is no CPU fallback)")
    B, T, C_ = m.shape
    if T < kernel:
        raise ValueError("If mode is 'interp', window_length must be less than or equal to the size of x.")   # scipy's message
    fir, edge = savgol_matrices(kernel, order)
    import numpy as np

    fir32 = np.ascontiguousarray(fir, dtype=np.float32)
    edge32 = np.ascontiguousarray(edge, dtype=np.float32)
    x = m.detach().to(torch.float32).contiguous()
    out = torch.empty_like(x)
    lib = _lib.load()
    rc = lib.dc_smooth_motion(x.device.index or 0, x.data_ptr(), out.data_ptr(), B, T, C_, kernel, fir32.ctypes.data, edge32.ctypes.data,
                              float(window), torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(rc, None)
    out = out.view(B, T, C_ // 2, 2)
    return out[0] if squeeze else out
