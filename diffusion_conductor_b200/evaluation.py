"""On-device evaluation features for generated conducting motion.

Host-side mirror of the reference's evaluation script, Diffusion_Stage/tools/eval_new_metrics.py:

  * `MotionEncoder_STGCN` (:38-74)  -- same constructor, same `state_dict` keys (ST_GCN in mode 'M2S' with edge-importance
    weighting + `fc`), so the `module.motion_encoder.*` weights of the reference's stage-one checkpoint load unchanged
    (:83-90); `forward(x)` / `features(x)[-1]` return the 64-d latent per frame, computed by the CUDA library.
  * `Evaluator`-style metric functions: `frechet_gesture_distance` (get_scores + calculate_frechet_distance, :159-241),
    `latent_l1` (diversity :148-156 and latent MAE :179-185), `motion_beats` (motion_peak_onehot, :277-303) and
    `beat_consistency` (alignment_score, :243-267).

The arithmetic runs in libdc_b200.so through the C ABI (include/dc_b200.h, dc_eval_*); there is no CPU path.  Only the 64 x 64
matrix square root of the Frechet distance (scipy.linalg.sqrtm on fp64 statistics reduced on the device) and the reference's
music beat tracker (librosa, third party) stay on the host.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
from torch import nn

from . import _lib

NUM_NODE = 13


def conductor_graph() -> np.ndarray:
    """Graph(layout='ConductorMotionX', strategy='uniform', max_hop=1).A -> (1, 13, 13)
    (reference models/ST_GCN/st_gcn_utils/graph.py:41-56 edges, 138-151 hop distance, 154-162 column normalisation)."""
    n = NUM_NODE
    links = [(0, 1), (0, 2), (1, 3), (2, 4), (0, 5), (0, 6), (5, 6), (5, 7), (7, 9), (6, 8), (8, 10), (11, 12), (5, 11), (6, 12)]
    reach = np.eye(n)
    for i, j in links:
        reach[i, j] = reach[j, i] = 1.0                        # hop distance <= 1 (self links included)
    return (reach / reach.sum(0, keepdims=True))[None]         # A D^-1: every column sums to 1


class _ParamBag(nn.Module):
    """Holds tensors under the reference's parameter names; never executed."""


def _bn(c):
    return nn.BatchNorm1d(c)


class _StGcnBlock(nn.Module):
    """Parameter container of reference st_gcn (ST_GCN.py:146-215): gcn.conv, tcn.{0: BN, 2: Conv(3 x 1), 3: BN}."""

    def __init__(self, cin, cout):
        super().__init__()
        self.gcn = _ParamBag()
        self.gcn.conv = nn.Conv2d(cin, cout, kernel_size=(1, 1))
        self.tcn = nn.Sequential(nn.BatchNorm2d(cout), nn.ReLU(inplace=True), nn.Conv2d(cout, cout, (3, 1), (1, 1), (1, 0)),
                                 nn.BatchNorm2d(cout), nn.Dropout(0, inplace=True))


class _StGcn(nn.Module):
    """Parameter container of reference ST_GCN(in_channels=2, out_channels=32, mode='M2S', edge_importance_weighting=True)
    (ST_GCN.py:33-83)."""

    def __init__(self):
        super().__init__()
        self.register_buffer("A", torch.tensor(conductor_graph(), dtype=torch.float32))
        self.data_bn = nn.BatchNorm1d(2 * NUM_NODE)
        self.st_gcn_networks = nn.ModuleList([_StGcnBlock(2 if i == 0 else 32, 32) for i in range(10)])
        self.edge_importance = nn.ParameterList([nn.Parameter(torch.ones(1, NUM_NODE, NUM_NODE)) for _ in range(10)])
        self.fcn = nn.Conv2d(256, 32, kernel_size=1)           # present in the reference state_dict, unused by features()


class MotionEncoder_STGCN(nn.Module):
    """Drop-in for eval_new_metrics.py:38-74.  `forward(motion)`: (N, T, 13, 2) -> (N, T, 64) latents on the CUDA device."""

    def __init__(self):
        super().__init__()
        self.st_gcn = _StGcn()
        self.fc = nn.Sequential(nn.Conv1d(32 * NUM_NODE, 64, kernel_size=1), nn.BatchNorm1d(64))
        self._handle = None
        self._loaded_version = None

    # ---- engine -------------------------------------------------------------------------------
    def _engine(self, device: torch.device):
        lib = _lib.load()
        key = (device.index, tuple(int(p._version) for p in self.state_dict().values()))
        if self._handle is not None and self._loaded_version == key:
            return lib, self._handle
        if self._handle is not None:
            lib.dc_eval_destroy(self._handle)
            self._handle = None
        h = C.c_void_p()
        _lib.check(lib.dc_eval_create(device.index or 0, C.byref(h)))
        for name, t in self.state_dict().items():
            if not t.dtype.is_floating_point:
                continue
            host = t.detach().to("cpu", torch.float32).contiguous()
            _lib.check(lib.dc_eval_set_weight(h, name.encode(), C.c_void_p(host.data_ptr()), host.numel()))
        _lib.check(lib.dc_eval_finalize(h))
        self._handle, self._loaded_version = h, key
        return lib, h

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load().dc_eval_destroy(self._handle)
        except Exception:
            pass

    # ---- reference API ------------------------------------------------------------------------
    def forward(self, input: torch.Tensor) -> torch.Tensor:
        if self.training:
            raise NotImplementedError("MotionEncoder_STGCN runs in eval mode only (the reference evaluator calls .eval(), :91)")
        if not input.is_cuda:
            raise RuntimeError("MotionEncoder_STGCN runs on the CUDA device only (there is no CPU fallback)")
        if input.dim() != 4 or input.shape[2] != NUM_NODE or input.shape[3] != 2:
            raise ValueError(f"expected motion of shape (N, T, 13, 2), got {tuple(input.shape)}")
        x = input.detach().to(torch.float32).contiguous()
        N, T = x.shape[0], x.shape[1]
        lib, h = self._engine(x.device)
        out = torch.empty(N, T, 64, device=x.device, dtype=torch.float32)
        _lib.check(lib.dc_eval_motion_features(h, C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), N, T,
                                               C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)))
        return out

    def features(self, input: torch.Tensor):
        """The reference returns the per-layer feature maps with the latent last (:62-74); the evaluator only reads [-1]."""
        return [self.forward(input)]


def _stream(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _rows64(feats: torch.Tensor) -> torch.Tensor:
    if not feats.is_cuda:
        raise RuntimeError("evaluation features live on the CUDA device (there is no CPU fallback)")
    return feats.detach().to(torch.float32).reshape(-1, 64).contiguous()


def feature_statistics(feats: torch.Tensor):
    """(mean [64], covariance [64, 64]) in fp64 of (..., 64) latents: np.mean(axis=0) / np.cov(rowvar=False) (:164-168)."""
    f = _rows64(feats)
    rows = f.shape[0]
    s = torch.empty(64, device=f.device, dtype=torch.float64)
    m2 = torch.empty(64, 64, device=f.device, dtype=torch.float64)
    _lib.check(_lib.load().dc_eval_feature_stats(f.device.index or 0, C.c_void_p(f.data_ptr()), rows, C.c_void_p(s.data_ptr()),
                                                 C.c_void_p(m2.data_ptr()), _stream(f)))
    return (s / rows).cpu().numpy(), (m2 / max(rows - 1, 1)).cpu().numpy()


def frechet_distance(mu1, sigma1, mu2, sigma2, eps: float = 1e-6) -> float:
    """calculate_frechet_distance (:189-241, pytorch-fid's stable formula) on fp64 statistics; the 64 x 64 sqrtm runs on the host."""
    from scipy import linalg

    diff = np.atleast_1d(mu1) - np.atleast_1d(mu2)
    sigma1, sigma2 = np.atleast_2d(sigma1), np.atleast_2d(sigma2)
    covmean = linalg.sqrtm(sigma1.dot(sigma2))      # reference: sqrtm(..., disp=False)[0]; newer scipy has no `disp`
    if not np.isfinite(covmean).all():
        offset = np.eye(sigma1.shape[0]) * eps
        covmean = linalg.sqrtm((sigma1 + offset).dot(sigma2 + offset))
    if np.iscomplexobj(covmean):
        if not np.allclose(np.diagonal(covmean).imag, 0, atol=1e-3):
            raise ValueError("Imaginary component {}".format(np.max(np.abs(covmean.imag))))
        covmean = covmean.real
    return float(diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(covmean))


def frechet_gesture_distance(generated_feats: torch.Tensor, real_feats: torch.Tensor) -> float:
    """FGD of get_scores (:159-177); 1e10 when the matrix square root is not real, as the reference."""
    try:
        return frechet_distance(*feature_statistics(generated_feats), *feature_statistics(real_feats))
    except ValueError:
        return 1e10


def latent_l1(a: torch.Tensor, b: torch.Tensor) -> float:
    """np.mean(np.sum(np.absolute(a - b), axis=-1)) over (..., 64) latents: the reduction of get_diversity_scores (:148-156, a =
    the generated latents, b = the same list shuffled) and of the latent-space MAE (:179-185, a = real, b = generated)."""
    fa, fb = _rows64(a), _rows64(b)
    if fa.shape != fb.shape:
        raise ValueError(f"shape mismatch {tuple(fa.shape)} vs {tuple(fb.shape)}")
    out = torch.empty(1, device=fa.device, dtype=torch.float64)
    _lib.check(_lib.load().dc_eval_feature_l1(fa.device.index or 0, C.c_void_p(fa.data_ptr()), C.c_void_p(fb.data_ptr()), fa.shape[0],
                                              C.c_void_p(out.data_ptr()), _stream(fa)))
    return float(out.item()) / fa.shape[0]


def diversity_score(generated_latents, perm=None) -> float:
    """get_diversity_scores (:148-156): the first 500 clips' latents against a random selection of 500 clips."""
    n = len(generated_latents)
    if perm is None:
        perm = torch.randperm(n)[:500]
    first = torch.cat([generated_latents[i] for i in range(min(n, 500))], dim=0)
    other = torch.cat([generated_latents[int(i)] for i in perm], dim=0)
    return latent_l1(first, other)


def motion_beats(motion: torch.Tensor, order: int = 10):
    """motion_peak_onehot (:277-303), batched: motion (N, T, 13, 2) or (N, T, 26) on the device -> (envelope (N, T) float32,
    beats (N, T) bool)."""
    if not motion.is_cuda:
        raise RuntimeError("motion_beats runs on the CUDA device only (there is no CPU fallback)")
    x = motion.detach().to(torch.float32).reshape(motion.shape[0], motion.shape[1], -1).contiguous()
    if x.shape[2] != 2 * NUM_NODE:
        raise ValueError(f"expected 13 x 2 keypoints per frame, got {tuple(motion.shape)}")
    N, T = x.shape[0], x.shape[1]
    env = torch.empty(N, T, device=x.device, dtype=torch.float32)
    beats = torch.empty(N, T, device=x.device, dtype=torch.uint8)
    _lib.check(_lib.load().dc_eval_motion_beats(x.device.index or 0, C.c_void_p(x.data_ptr()), C.c_void_p(env.data_ptr()),
                                                C.c_void_p(beats.data_ptr()), N, T, int(order), _stream(x)))
    return env, beats.bool()


def beat_consistency(music_beats: torch.Tensor, motion_beat_onehot: torch.Tensor, sigma: float = 3.0) -> torch.Tensor:
    """alignment_score (:243-267) per clip: music_beats (N, Tm) and motion beats (N, T), any dtype (non-zero = beat) -> (N,) scores."""
    if not (music_beats.is_cuda and motion_beat_onehot.is_cuda):
        raise RuntimeError("beat_consistency runs on the CUDA device only (there is no CPU fallback)")
    mu = (music_beats != 0).to(torch.uint8).contiguous()
    mo = (motion_beat_onehot != 0).to(torch.uint8).contiguous()
    if mu.dim() != 2 or mo.dim() != 2 or mu.shape[0] != mo.shape[0]:
        raise ValueError("expected (N, Tm) music beats and (N, T) motion beats")
    out = torch.empty(mu.shape[0], device=mu.device, dtype=torch.float32)
    _lib.check(_lib.load().dc_eval_beat_alignment(mu.device.index or 0, C.c_void_p(mu.data_ptr()), mu.shape[1], C.c_void_p(mo.data_ptr()),
                                                  mo.shape[1], mu.shape[0], float(sigma), C.c_void_p(out.data_ptr()), _stream(mu)))
    return out
