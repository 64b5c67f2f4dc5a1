"""ctypes binding of the C ABI in include/dc_b200.h (the in-tree libdc_b200.so).

There is deliberately no fallback: if the library is missing or a call fails, an exception is
raised -- the CUDA path is the product.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
# DC_B200_LIB: an alternative build of the same library (A/B experiments on one GPU box); the product is the in-tree file
LIB_PATH = os.environ.get("DC_B200_LIB") or os.path.join(_HERE, "libdc_b200.so")

DC_OPERAND_BF16, DC_OPERAND_FP16 = 0, 1
DC_SAMPLER_NONE, DC_SAMPLER_DDIM, DC_SAMPLER_DDPM = 0, 1, 2
DC_FLAG_CLIP = 0x10

EXPORTS = [
    "dc_last_error", "dc_create", "dc_destroy", "dc_set_weight", "dc_finalize_weights", "dc_set_schedule",
    "dc_prepare_cond", "dc_forward", "dc_sample_step", "dc_sampler_update", "dc_sample_loop", "dc_generate_host",
    "dc_kernel_launches", "dc_set_graphs", "dc_selftest_gemm", "dc_profile_step", "dc_debug_timeline", "dc_smooth_motion",
    "dc_encode_music", "dc_time_embedding", "dc_cluster_occupancy", "dc_sample_range",
    "dc_eval_create", "dc_eval_destroy", "dc_eval_set_weight", "dc_eval_finalize", "dc_eval_motion_features", "dc_eval_feature_stats",
    "dc_eval_feature_l1", "dc_eval_motion_beats", "dc_eval_beat_alignment",
]


class DcConfig(C.Structure):
    _fields_ = [("input_feats", C.c_int), ("num_frames", C.c_int), ("latent_dim", C.c_int), ("ff_size", C.c_int),
                ("num_layers", C.c_int), ("num_heads", C.c_int), ("device", C.c_int), ("operand", C.c_int)]


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """dlopen the library (once) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with ./build.sh (or __graft_entry__.build()). "
            "diffusion_conductor_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    fp = C.c_void_p   # raw float* (host or device address)
    lib.dc_last_error.restype = C.c_char_p
    lib.dc_last_error.argtypes = [vp]
    lib.dc_create.argtypes = [C.POINTER(DcConfig), C.POINTER(vp)]
    lib.dc_destroy.restype = None
    lib.dc_destroy.argtypes = [vp]
    lib.dc_set_weight.argtypes = [vp, C.c_char_p, fp, C.POINTER(i64), i32]
    lib.dc_finalize_weights.argtypes = [vp]
    lib.dc_set_schedule.argtypes = [vp, i32, fp]
    lib.dc_prepare_cond.argtypes = [vp, fp, fp, C.POINTER(i64), i32, i32, vp]
    lib.dc_forward.argtypes = [vp, fp, fp, fp, vp]
    lib.dc_sample_step.argtypes = [vp, i32, fp, fp, i32, fp, vp]
    lib.dc_sampler_update.argtypes = [vp, i32, fp, fp, i32, fp, i64, vp]
    lib.dc_sample_loop.argtypes = [vp, i32, i32, fp, fp, fp, fp, vp]
    lib.dc_sample_range.argtypes = [vp, i32, i32, i32, fp, fp, fp, fp, vp]
    lib.dc_generate_host.argtypes = [vp, i32, fp, fp, C.POINTER(i64), fp, fp, i32, i32, vp]
    lib.dc_profile_step.argtypes = [vp, i32, fp, i32, C.POINTER(C.c_float), C.POINTER(i32), vp]
    lib.dc_debug_timeline.argtypes = [vp, fp, i32, C.POINTER(C.c_uint64), i32]
    lib.dc_kernel_launches.restype = i64
    lib.dc_kernel_launches.argtypes = [vp]
    lib.dc_set_graphs.argtypes = [vp, i32]
    lib.dc_selftest_gemm.argtypes = [i32, i32, i32, i32, i32, fp, fp, fp, fp]
    lib.dc_encode_music.argtypes = [vp, fp, fp, fp, i32, i32, vp]
    lib.dc_time_embedding.argtypes = [vp, fp, i32, fp, vp]
    lib.dc_cluster_occupancy.argtypes = [vp, i32, C.POINTER(i32)]
    lib.dc_smooth_motion.argtypes = [i32, fp, fp, i32, i32, i32, i32, fp, fp, C.c_float, vp]
    lib.dc_eval_create.argtypes = [i32, C.POINTER(vp)]
    lib.dc_eval_destroy.argtypes = [vp]
    lib.dc_eval_destroy.restype = None
    lib.dc_eval_set_weight.argtypes = [vp, C.c_char_p, fp, i64]
    lib.dc_eval_finalize.argtypes = [vp]
    lib.dc_eval_motion_features.argtypes = [vp, fp, fp, i32, i32, vp]
    lib.dc_eval_feature_stats.argtypes = [i32, fp, i64, fp, fp, vp]
    lib.dc_eval_feature_l1.argtypes = [i32, fp, fp, i64, fp, vp]
    lib.dc_eval_motion_beats.argtypes = [i32, fp, fp, fp, i32, i32, i32, vp]
    lib.dc_eval_beat_alignment.argtypes = [i32, fp, i32, fp, i32, i32, C.c_float, fp, vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("dc_last_error", "dc_destroy", "dc_kernel_launches"):
            fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc: int, handle=None) -> None:
    """Translate a dc_status into the exception the reference-facing API raises."""
    if rc == 0:
        return
    msg = load().dc_last_error(handle).decode("utf-8", "replace")
    if rc == -2:
        raise NotImplementedError(msg)
    raise RuntimeError(f"dc_b200 error {rc}: {msg}")
