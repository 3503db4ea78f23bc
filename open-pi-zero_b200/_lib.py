"""ctypes binding of `include/pz_b200.h` (the thin C-ABI extension).

There is deliberately no fallback: if the CUDA library cannot be built or
loaded, importing the compute path raises.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

vp, fp = C.c_void_p, C.c_void_p   # device pointers travel as integers


class PzConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "dtype", "vocab_size", "pad_token_id", "image_token_index", "s_vlm", "n_img_tokens",
        "n_images", "cond_steps", "horizon", "action_dim", "proprio_dim", "n_steps")] + \
        [("clip", C.c_float)] + \
        [(n, C.c_int32) for n in (
            "n_layers", "n_heads", "n_kv_heads", "head_dim", "vlm_hidden", "vlm_inter",
            "act_hidden", "act_inter", "vit_hidden", "vit_inter", "vit_layers", "vit_heads",
            "image_size", "patch_size", "patch_k_pad", "max_batch", "flags")]


class PzVitLayer(C.Structure):
    _fields_ = [(n, vp) for n in ("ln1_w", "ln1_b", "w_qkv", "b_qkv", "w_o", "b_o", "ln2_w",
                                  "ln2_b", "w_fc1", "b_fc1", "w_fc2", "b_fc2")]


class PzMixLayer(C.Structure):
    _fields_ = [(n, vp) for n in ("norm_in", "w_qkv", "w_o", "norm_post", "w_gate_up", "w_down")]


class PzWeights(C.Structure):
    _fields_ = [("embed", vp), ("patch_w", vp), ("patch_b", vp), ("pos_emb", vp),
                ("vit", C.POINTER(PzVitLayer)), ("post_ln_w", vp), ("post_ln_b", vp),
                ("proj_w", vp), ("proj_b", vp),
                ("vlm", C.POINTER(PzMixLayer)), ("proprio", C.POINTER(PzMixLayer)),
                ("action", C.POINTER(PzMixLayer)), ("action_final_norm", vp),
                ("enc_w1", vp), ("enc_b1", vp), ("enc_w2a", vp), ("enc_time_bias", vp),
                ("enc_w3", vp), ("enc_b3", vp), ("prop_w", vp), ("prop_b", vp),
                ("dec_w", vp), ("dec_b", vp),
                ("rope_vlm_cos", vp), ("rope_vlm_sin", vp), ("rope_act_cos", vp),
                ("rope_act_sin", vp), ("small_k_pad", C.c_int32),
                ("enc_w2t", vp), ("enc_b2", vp), ("time_freq", vp),
                ("vlm_final_norm", vp), ("lm_head", vp), ("rope_vlm_rows", C.c_int32)]


class PzCapture(C.Structure):
    _fields_ = [(n, vp) for n in ("vit_out", "image_features", "prefix_embeds", "prefix_vlm",
                                  "prefix_proprio", "denoise_action", "velocities",
                                  "action_preclip")]


PZ_ABI_VERSION = 8
PZ_F32, PZ_BF16 = 0, 1
PZ_FLAG_SIMPLE_KERNELS = 1
PZ_FLAG_ALLOW_FALLBACK = 2
PZ_SAMPLER_AUTO, PZ_SAMPLER_KERNELS, PZ_SAMPLER_BARRIER, PZ_SAMPLER_STREAM = 0, 1, 2, 3
LIN_GELU, LIN_OUT_F32, LIN_ACCUM, LIN_GEGLU, LIN_SILU = 1, 2, 4, 8, 16

# every symbol include/pz_b200.h declares
EXPORTS = ["pz_abi_version", "pz_create", "pz_destroy", "pz_last_error", "pz_bind_weights",
           "pz_workspace_bytes", "pz_set_pixel_format", "pz_set_io_normalization", "pz_kv_layout", "pz_debug_trace_offset", "pz_sampler_stream_bytes", "pz_sampler_pack", "pz_set_sampler", "pz_infer_action", "pz_embed_prefix",
           "pz_prefill", "pz_denoise", "pz_text_prefill", "pz_text_decode", "pz_joint_prefix", "pz_joint_action", "pz_velocity", "pz_flow_matching_loss", "pz_train_workspace_bytes", "pz_flow_matching_step", "pz_grad_sumsq", "pz_adamw_step", "pz_average_update", "pz_write_packed",
           "pz_launch_count", "pz_fallback_count", "pz_timing_begin", "pz_timing_end", "pz_op_linear", "pz_op_linear_ex", "pz_op_attention"]

_lib = None


def lib_path() -> str:
    return _build.LIB


def load(build_if_needed: bool = True):
    """Load (building first if sources are newer) libpz_b200.so.  Raises on failure."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_needed and os.environ.get("PZ_NO_BUILD") != "1":
        _build.build()
    if not os.path.exists(_build.LIB):
        raise RuntimeError(f"{_build.LIB} is missing: run `python __graft_entry__.py` / build() "
                           "first -- there is no CPU fallback for the infer_action path")
    lib = C.CDLL(_build.LIB)
    hp = C.c_void_p
    lib.pz_abi_version.restype = C.c_int
    lib.pz_create.argtypes = [C.POINTER(PzConfig), C.POINTER(hp)]
    lib.pz_destroy.argtypes = [hp]
    lib.pz_destroy.restype = None
    lib.pz_last_error.argtypes = [hp]
    lib.pz_last_error.restype = C.c_char_p
    lib.pz_bind_weights.argtypes = [hp, C.POINTER(PzWeights)]
    lib.pz_set_pixel_format.argtypes = [hp, C.c_int]
    lib.pz_workspace_bytes.argtypes = [hp, C.c_int]
    lib.pz_workspace_bytes.restype = C.c_size_t
    lib.pz_debug_trace_offset.argtypes = [hp, C.c_int]
    lib.pz_debug_trace_offset.restype = C.c_size_t
    lib.pz_sampler_stream_bytes.argtypes = [hp, C.c_int]
    lib.pz_sampler_stream_bytes.restype = C.c_size_t
    lib.pz_sampler_pack.argtypes = [hp, C.c_int, vp, C.c_size_t, vp]
    lib.pz_set_sampler.argtypes = [hp, C.c_int]
    lib.pz_kv_layout.argtypes = [hp, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                 C.POINTER(C.c_size_t)]
    lib.pz_infer_action.argtypes = [hp, vp, vp, vp, vp, vp, vp, vp, C.c_size_t, C.c_int,
                                    C.POINTER(PzCapture), vp]
    lib.pz_embed_prefix.argtypes = [hp, vp, vp, vp, C.c_size_t, C.c_int, C.POINTER(PzCapture), vp]
    lib.pz_prefill.argtypes = [hp, vp, vp, vp, C.c_size_t, C.c_int, C.POINTER(PzCapture), vp]
    lib.pz_denoise.argtypes = [hp, vp, vp, vp, vp, C.c_size_t, C.c_int, C.POINTER(PzCapture), vp]
    lib.pz_joint_prefix.argtypes = [hp, vp, vp, vp, vp, C.c_size_t, C.c_int, vp]
    lib.pz_joint_action.argtypes = [hp, vp, vp, vp, vp, C.c_size_t, C.c_int, vp]
    lib.pz_velocity.argtypes = [hp, vp, vp, vp, vp, vp, C.c_size_t, C.c_int, vp]
    lib.pz_flow_matching_loss.argtypes = [hp, vp, vp, vp, vp, vp, vp, vp, C.c_float, vp, vp, vp, C.c_size_t, C.c_int, vp]
    lib.pz_set_io_normalization.argtypes = [hp, vp, vp, C.c_int, vp, vp]
    lib.pz_train_workspace_bytes.argtypes = [hp, C.c_int]
    lib.pz_train_workspace_bytes.restype = C.c_size_t
    lib.pz_flow_matching_step.argtypes = [hp, vp, vp, vp, vp, vp, vp, vp, C.c_float, vp, C.c_float, vp, vp, C.c_size_t, C.c_int, C.c_int, vp, C.c_int, vp]
    lib.pz_average_update.argtypes = [vp, vp, C.c_size_t, C.c_float, vp]
    lib.pz_write_packed.argtypes = [vp, C.c_size_t, C.c_size_t, vp, vp, vp, C.c_int, C.c_int, vp]
    lib.pz_grad_sumsq.argtypes = [vp, C.c_size_t, vp, vp]
    lib.pz_adamw_step.argtypes = [vp, vp, vp, vp, C.c_size_t, C.c_size_t, vp, vp, vp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                  C.c_float, C.c_float, C.c_int, vp, C.c_float, C.c_float, C.c_int, vp]
    lib.pz_text_prefill.argtypes = [hp, vp, vp, vp, C.c_int, C.c_int, vp, C.c_int, vp, vp, vp, C.c_size_t, C.c_int, vp]
    lib.pz_text_decode.argtypes = [hp, vp, vp, C.c_int, vp, vp, C.c_int, vp, vp, vp, C.c_size_t, C.c_int, vp]
    lib.pz_launch_count.argtypes = [hp]
    lib.pz_launch_count.restype = C.c_int64
    lib.pz_fallback_count.argtypes = [hp]
    lib.pz_fallback_count.restype = C.c_int64
    lib.pz_timing_begin.argtypes = [hp, C.c_int]
    lib.pz_timing_end.argtypes = [hp, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.pz_op_linear.argtypes = [C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_int, C.c_float, vp]
    lib.pz_op_linear_ex.argtypes = [C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, vp]
    lib.pz_op_attention.argtypes = [C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp] + \
        [C.c_int] * 9 + [C.c_float, C.c_float, vp, C.c_size_t, vp]
    if lib.pz_abi_version() != PZ_ABI_VERSION:
        raise RuntimeError("libpz_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib
