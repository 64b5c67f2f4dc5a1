"""B200-native (sm_100a) implementation of Diffusion-Conductor's DDIM/DDPM denoising hot path.

Drop-in surface (reference Diffusion_Stage/models/__init__.py:1-4):
    from diffusion_conductor_b200 import MotionTransformer, GaussianDiffusion
"""
from .gaussian_diffusion import (GaussianDiffusion, LossType, ModelMeanType, ModelVarType,  # noqa: F401
                                 get_named_beta_schedule)
from .transformer import MotionTransformer, MusicEncoder, timestep_embedding  # noqa: F401

__all__ = ["MotionTransformer", "GaussianDiffusion", "MusicEncoder", "ModelMeanType", "ModelVarType", "LossType",
           "get_named_beta_schedule", "timestep_embedding"]
