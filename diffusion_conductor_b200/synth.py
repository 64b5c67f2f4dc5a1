"""Deterministic synthetic weights and inputs of the ConductorMotion100 shape.

There is no checkpoint or dataset in the build environment, so parity tests and the
benchmark run on random-init weights and synthetic mel/noise.  A fresh
reference model outputs exactly 0 because 33 weight tensors are zero-initialised
(reference transformer.py:44-50,65,165,443), so every tensor is drawn here from its own
seeded CPU generator -- independent of module construction order and identical on every
machine with the same torch build.
"""
from __future__ import annotations

import zlib
from typing import Dict, Tuple

import torch


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1_000_003 + zlib.crc32(key.encode())) % (2 ** 31 - 1))
    return g


def _uniform(shape, bound, g):
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound


def reference_shapes(num_layers: int = 8, latent_dim: int = 128, ff_size: int = 64, num_frames: int = 1800,
                     input_feats: int = 26) -> Dict[str, Tuple[int, ...]]:
    """The state_dict layout of reference MotionTransformer (transformer.py:361-445), key -> shape."""
    D, E, F_, P = latent_dim, 4 * latent_dim, ff_size, input_feats
    s: Dict[str, Tuple[int, ...]] = {"sequence_embedding": (num_frames, D)}

    def conv_bn(p, cin, cout, k):
        s[p + ".0.weight"] = (cout, cin) + k
        s[p + ".0.bias"] = (cout,)
        for n in ("weight", "bias", "running_mean", "running_var"):
            s[p + ".1." + n] = (cout,)
        s[p + ".1.num_batches_tracked"] = ()

    for blk, chans in (("conv1", [(1, 16), (16, 16), (16, 16)]), ("conv2", [(16, 32), (32, 32)]),
                       ("conv3", [(32, 32), (32, 32)])):
        for i, (ci, co) in enumerate(chans):
            conv_bn(f"music_encoder.{blk}.{i}.conv2d_layer", ci, co, (3, 3))
            if ci != co and not (blk == "conv1" and i == 0):
                conv_bn(f"music_encoder.{blk}.{i}.residual", ci, co, (1, 1))
    conv_bn("music_encoder.conv4", 512, 64, (1,))

    def lin(p, o, i):
        s[p + ".weight"] = (o, i)
        s[p + ".bias"] = (o,)

    def ln(p, d):
        s[p + ".weight"] = (d,)
        s[p + ".bias"] = (d,)

    lin("linear", E, 64)
    lin("joint_embed", D, P)
    lin("time_embed.0", E, D)
    lin("time_embed.2", E, E)
    for i in range(num_layers):
        b = f"temporal_decoder_blocks.{i}"

        def styl(p):
            lin(p + ".emb_layers.1", 2 * D, E)
            ln(p + ".norm", D)
            lin(p + ".out_layers.2", D, D)

        ln(b + ".sa_block.norm", D)
        for n in ("query", "key", "value"):
            lin(f"{b}.sa_block.{n}", D, D)
        styl(b + ".sa_block.proj_out")
        ln(b + ".ca_block.norm", D)
        ln(b + ".ca_block.text_norm", E)
        lin(b + ".ca_block.query", D, D)
        lin(b + ".ca_block.key", D, E)
        lin(b + ".ca_block.value", D, E)
        styl(b + ".ca_block.proj_out")
        lin(b + ".ffn.linear1", F_, D)
        lin(b + ".ffn.linear2", D, F_)
        styl(b + ".ffn.proj_out")
    lin("out", P, D)
    lin("proj", 64, 64)
    return s


def synth_state_dict(seed: int = 0, **dims) -> Dict[str, torch.Tensor]:
    """Random weights for every key.  Linear/conv ~ U(+-1/sqrt(fan_in)) like torch's default init
    (including the tensors the reference zero-initialises, so the model output is non-trivial);
    LayerNorm/BatchNorm affine terms are perturbed away from (1, 0) so that folding bugs show up."""
    sd: Dict[str, torch.Tensor] = {}
    for key, shape in reference_shapes(**dims).items():
        g = _gen(seed, key)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            sd[key] = torch.tensor(0, dtype=torch.int64)
        elif key == "sequence_embedding":
            sd[key] = torch.randn(shape, generator=g)
        elif leaf == "running_mean":
            sd[key] = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "running_var":
            sd[key] = 0.5 + torch.rand(shape, generator=g)
        elif len(shape) == 1 and leaf == "weight":          # LayerNorm / BatchNorm gamma
            sd[key] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 1:                                # any bias / beta
            sd[key] = 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            sd[key] = _uniform(shape, fan_in ** -0.5, g)
    return sd


def synth_inputs(B: int, T: int, seed: int = 0):
    """mel ~ U[0,1) (B,3T,128) (real mels are normalised to [0,1]); noise ~ N(0,1) (B,T,26)."""
    mel = torch.rand(B, 3 * T, 128, generator=_gen(seed, "mel"), dtype=torch.float32)
    noise = torch.randn(B, T, 26, generator=_gen(seed, "noise"), dtype=torch.float32)
    return mel, noise


def synth_features(B: int, T: int, seed: int = 0):
    """Stand-ins for encode_music outputs (xf_proj, xf_out), each (B,T,64), for tests that skip the CNN."""
    xf_out = torch.randn(B, T, 64, generator=_gen(seed, "xf_out"), dtype=torch.float32)
    xf_proj = torch.randn(B, T, 64, generator=_gen(seed, "xf_proj"), dtype=torch.float32)
    return xf_proj, xf_out


def stgcn_shapes() -> Dict[str, Tuple[int, ...]]:
    """state_dict layout of the reference's MotionEncoder_STGCN (tools/eval_new_metrics.py:38-49: ST_GCN(in 2, out 32, mode
    'M2S', edge importance) + fc = Conv1d(416, 64, 1) + BatchNorm1d(64)), key -> shape."""
    s: Dict[str, Tuple[int, ...]] = {"st_gcn.A": (1, 13, 13)}

    def bn(p, c):
        for n in ("weight", "bias", "running_mean", "running_var"):
            s[p + "." + n] = (c,)
        s[p + ".num_batches_tracked"] = ()

    bn("st_gcn.data_bn", 26)
    for i in range(10):
        cin = 2 if i == 0 else 32
        p = f"st_gcn.st_gcn_networks.{i}"
        s[p + ".gcn.conv.weight"] = (32, cin, 1, 1)
        s[p + ".gcn.conv.bias"] = (32,)
        bn(p + ".tcn.0", 32)
        s[p + ".tcn.2.weight"] = (32, 32, 3, 1)
        s[p + ".tcn.2.bias"] = (32,)
        bn(p + ".tcn.3", 32)
    for i in range(10):
        s[f"st_gcn.edge_importance.{i}"] = (1, 13, 13)
    s["st_gcn.fcn.weight"] = (32, 256, 1, 1)
    s["st_gcn.fcn.bias"] = (32,)
    s["fc.0.weight"] = (64, 416, 1)
    s["fc.0.bias"] = (64,)
    bn("fc.1", 64)
    return s


def synth_stgcn_state_dict(seed: int = 0, A=None) -> Dict[str, torch.Tensor]:
    """Random weights for the ST-GCN motion encoder (same recipe as synth_state_dict); `A` (1, 13, 13) is the graph buffer
    (defaults to the ConductorMotionX uniform-strategy adjacency), edge importances ~ 1 + 0.2 N(0, 1)."""
    sd: Dict[str, torch.Tensor] = {}
    for key, shape in stgcn_shapes().items():
        g = _gen(seed, "stgcn." + key)
        leaf = key.rsplit(".", 1)[-1]
        if key == "st_gcn.A":
            if A is None:
                from .evaluation import conductor_graph
                A = conductor_graph()
            sd[key] = torch.as_tensor(A, dtype=torch.float32).reshape(shape).clone()
        elif key.startswith("st_gcn.edge_importance"):
            sd[key] = 1.0 + 0.2 * torch.randn(shape, generator=g)
        elif leaf == "num_batches_tracked":
            sd[key] = torch.tensor(0, dtype=torch.int64)
        elif leaf == "running_mean":
            sd[key] = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "running_var":
            sd[key] = 0.5 + torch.rand(shape, generator=g)
        elif len(shape) == 1 and leaf == "weight":
            sd[key] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 1:
            sd[key] = 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            sd[key] = _uniform(shape, fan_in ** -0.5, g)
    return sd


def synth_motion(N: int, T: int, seed: int = 0) -> torch.Tensor:
    """Smooth keypoint tracks in [0, 1] (N, T, 13, 2): a few sinusoids per coordinate plus a little noise, so that the joint-speed
    envelope has genuine local minima."""
    g = _gen(seed, "motion")
    t = torch.arange(T, dtype=torch.float32)[None, :, None, None]
    f = 0.02 + 0.08 * torch.rand(N, 1, 13, 2, generator=g)
    ph = 6.2831853 * torch.rand(N, 1, 13, 2, generator=g)
    f2 = 0.2 + 0.3 * torch.rand(N, 1, 13, 2, generator=g)
    base = torch.rand(N, 1, 13, 2, generator=g)
    x = base + 0.15 * torch.sin(6.2831853 * f * t + ph) + 0.03 * torch.sin(6.2831853 * f2 * t) + 0.002 * torch.randn(N, T, 13, 2, generator=g)
    return x.clamp(0, 1)
