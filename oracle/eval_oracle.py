"""CPU oracle for the evaluation features that follow the sampling path.

TEST INFRASTRUCTURE ONLY (same rule as oracle/motion_oracle.py: only tests/ may import it).

Functional restatement (torch-CPU / numpy over a `state_dict`) of

  * MotionEncoder_STGCN.features(x)[-1] ........ tools/eval_new_metrics.py:38-74
      ST_GCN.forward (mode 'M2S') ............... models/ST_GCN/ST_GCN.py:86-113
      st_gcn.forward ............................ models/ST_GCN/ST_GCN.py:146-228
      ConvTemporalGraphical.forward ............. models/ST_GCN/st_gcn_utils/tgcn.py:61-73
      Graph('ConductorMotionX', 'uniform') ...... models/ST_GCN/st_gcn_utils/graph.py:27-100, 138-162
  * Evaluator.get_scores / frechet distance .... tools/eval_new_metrics.py:159-241
  * Evaluator.get_diversity_scores ............. tools/eval_new_metrics.py:148-156
  * Evaluator.motion_peak_onehot ............... tools/eval_new_metrics.py:277-303
  * Evaluator.alignment_score .................. tools/eval_new_metrics.py:243-267
(paths relative to /root/reference/Diffusion_Stage/).

Pinned: `oracle/make_golden.py eval` runs the UNMODIFIED reference ST_GCN module and the reference's own
motion_peak_onehot / alignment_score / calculate_frechet_distance source (extracted from
tools/eval_new_metrics.py, whose module-level imports of mmcv / librosa are absent here) and writes
tests/golden/eval_features.npz; tests/test_oracle_golden.py replays it.
Music beat detection (get_music_beat, eval_new_metrics.py:313-340) is librosa (third party, not installed):
out of scope; alignment_score takes the music beat one-hot as an input.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]
BN_EPS = 1e-5
NUM_NODE = 13
NUM_LAYERS = 10          # mode 'M2S' (ST_GCN.py:60-72)


def conductor_graph() -> np.ndarray:
    """graph.py:41-56 (edges), 138-151 (hop distance), 154-162 (normalize_digraph), 92-100 ('uniform') -> A (1, 13, 13)."""
    n = NUM_NODE
    edge = [(i, i) for i in range(n)] + [(0, 1), (0, 2), (1, 3), (2, 4), (0, 5), (0, 6), (5, 6), (5, 7), (7, 9), (6, 8), (8, 10),
                                         (11, 12), (5, 11), (6, 12)]
    adj = np.zeros((n, n))
    for i, j in edge:
        adj[j, i] = 1
        adj[i, j] = 1
    hop = np.zeros((n, n)) + np.inf
    arrive = np.stack([np.linalg.matrix_power(adj, d) for d in range(2)]) > 0
    for d in (1, 0):
        hop[arrive[d]] = d
    a = np.zeros((n, n))
    a[hop == 0] = 1
    a[hop == 1] = 1
    dl = a.sum(0)
    dn = np.zeros((n, n))
    for i in range(n):
        if dl[i] > 0:
            dn[i, i] = dl[i] ** (-1)
    return (a @ dn)[None]


def _bn(sd: StateDict, p: str, x: Tensor) -> Tensor:
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, BN_EPS)


def st_gcn_layer(sd: StateDict, p: str, x: Tensor, A: Tensor, residual: bool) -> Tensor:
    """ST_GCN.py:217-228 with tgcn.py:61-73: x (N, C, T, V)."""
    res = x if residual else 0
    y = F.conv2d(x, sd[p + ".gcn.conv.weight"], sd[p + ".gcn.conv.bias"])
    n, kc, t, v = y.shape
    y = y.view(n, A.shape[0], kc // A.shape[0], t, v)
    y = torch.einsum("nkctv,kvw->nctw", y, A).contiguous()
    y = F.relu(_bn(sd, p + ".tcn.0", y))
    y = F.conv2d(y, sd[p + ".tcn.2.weight"], sd[p + ".tcn.2.bias"], padding=(1, 0))
    y = _bn(sd, p + ".tcn.3", y)
    return F.relu(y + res)


def motion_features(sd: StateDict, motion: Tensor) -> Tensor:
    """eval_new_metrics.py:62-74, last element of features(): motion (N, T, 13, 2) -> (N, T, 64)."""
    with torch.no_grad():
        x = motion.float().transpose(1, 2).transpose(1, 3).unsqueeze(4)            # (N, C, T, V, M)
        N, C, T, V, M = x.shape
        x = x.permute(0, 4, 3, 1, 2).contiguous().view(N * M, V * C, T)
        x = _bn(sd, "st_gcn.data_bn", x)
        x = x.view(N, M, V, C, T).permute(0, 1, 3, 4, 2).contiguous().view(N * M, C, T, V)
        A = sd["st_gcn.A"]
        for i in range(NUM_LAYERS):
            x = st_gcn_layer(sd, f"st_gcn.st_gcn_networks.{i}", x, A * sd[f"st_gcn.edge_importance.{i}"], residual=i > 0)
        out = torch.flatten(x.transpose(1, 2), start_dim=2)                         # (N, T, C * V)
        out = F.conv1d(out.transpose(1, 2), sd["fc.0.weight"], sd["fc.0.bias"])
        return _bn(sd, "fc.1", out).transpose(1, 2)


def feature_stats(feats: np.ndarray):
    """eval_new_metrics.py:164-168: mean and np.cov(rowvar=False) of (rows, 64) features."""
    return np.mean(feats, axis=0), np.cov(feats, rowvar=False)


def frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6) -> float:
    """eval_new_metrics.py:189-241 (pytorch-fid's formula): ||mu1 - mu2||^2 + Tr(C1 + C2 - 2 sqrt(C1 C2))."""
    from scipy import linalg

    mu1, mu2 = np.atleast_1d(mu1), np.atleast_1d(mu2)
    sigma1, sigma2 = np.atleast_2d(sigma1), np.atleast_2d(sigma2)
    diff = mu1 - mu2
    covmean = linalg.sqrtm(sigma1.dot(sigma2))      # reference: sqrtm(..., disp=False)[0]; newer scipy has no `disp`
    if not np.isfinite(covmean).all():
        offset = np.eye(sigma1.shape[0]) * eps
        covmean = linalg.sqrtm((sigma1 + offset).dot(sigma2 + offset))
    if np.iscomplexobj(covmean):
        if not np.allclose(np.diagonal(covmean).imag, 0, atol=1e-3):
            raise ValueError("Imaginary component {}".format(np.max(np.abs(covmean.imag))))
        covmean = covmean.real
    return float(diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(covmean))


def feature_l1(a: np.ndarray, b: np.ndarray) -> float:
    """eval_new_metrics.py:154 / 181-185: mean over rows of sum |a - b| (diversity and latent MAE use the same reduction)."""
    return float(np.mean(np.sum(np.absolute(a - b), axis=-1)))


def motion_peak_onehot(joints: np.ndarray, order: int = 10):
    """eval_new_metrics.py:277-303: joints (T, 13, 2) -> (envelope (T,), beats (T,) bool): strict local minima of the summed
    joint speed within +-order frames (scipy argrelextrema, mode='clip')."""
    velocity = np.zeros_like(joints, dtype=np.float32)
    velocity[1:] = joints[1:] - joints[:-1]
    envelope = np.sum(np.linalg.norm(velocity, axis=2), axis=1)
    T = envelope.shape[0]
    beats = np.ones(T, dtype=bool)
    idx = np.arange(T)
    for k in range(1, order + 1):
        beats &= envelope < envelope[np.clip(idx + k, 0, T - 1)]
        beats &= envelope < envelope[np.clip(idx - k, 0, T - 1)]
    return envelope, beats


def alignment_score(music_beats: np.ndarray, motion_beats: np.ndarray, sigma: float = 3) -> float:
    """eval_new_metrics.py:243-267 (beat consistency): mean over music beats of exp(-d^2 / 2 sigma^2), d = distance (in
    indices, as the reference compares them) to the nearest motion beat; 0 when there is no motion beat."""
    if motion_beats.sum() == 0:
        return 0.0
    mi = np.where(music_beats)[0]
    bi = np.where(motion_beats)[0]
    scores = []
    for m in mi:
        d = np.abs(m - bi).astype(np.float32)
        scores.append(np.exp(-d[np.argmin(d)] ** 2 / 2 / sigma ** 2))
    return float(sum(scores) / len(scores))
