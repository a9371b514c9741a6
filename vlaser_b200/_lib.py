"""ctypes binding of the C ABI in include/vlaser_b200.h.

The product path has no fallback: if the CUDA library is missing or fails to load, importing
callers get an ImportError / RuntimeError here rather than a slower code path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VLASER_LIB") or os.path.join(HERE, "libvlaser_b200.so")  # env override: A/B of builds

F32, BF16 = 0, 1

# epilogue codes (csrc/gemm.cuh)
EPI_STORE_BF16, EPI_GELU_BF16, EPI_SILU_BF16, EPI_RESID_F32, EPI_SWIGLU_BF16, EPI_QKV_ROPE, EPI_PATCH_EMBED, EPI_STORE_F32 = range(8)


class VlaserCfg(C.Structure):
    _fields_ = [
        ("image_size", C.c_int32), ("patch_size", C.c_int32), ("vit_hidden", C.c_int32), ("vit_layers", C.c_int32),
        ("vit_heads", C.c_int32), ("vit_mlp", C.c_int32), ("vit_eps", C.c_float), ("proj_eps", C.c_float),
        ("hidden", C.c_int32), ("layers", C.c_int32), ("n_heads", C.c_int32), ("n_kv_heads", C.c_int32),
        ("head_dim", C.c_int32), ("mlp", C.c_int32), ("rms_eps", C.c_float), ("rope_theta", C.c_float),
        ("hidden_e", C.c_int32), ("mlp_e", C.c_int32),
        ("embed_rows", C.c_int32), ("image_token_index", C.c_int32), ("pad_token_id", C.c_int32),
        ("max_image_text_tokens", C.c_int32),
        ("cond_steps", C.c_int32), ("horizon_steps", C.c_int32), ("action_dim", C.c_int32), ("proprio_dim", C.c_int32),
        ("num_inference_steps", C.c_int32), ("clip_enabled", C.c_int32), ("final_action_clip_value", C.c_float),
        ("time_max_period", C.c_float), ("max_batch", C.c_int32),
    ]


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("W", C.c_void_p), ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("M_dev", C.c_void_p), ("epi", C.c_int32), ("block_n", C.c_int32), ("max_splits", C.c_int32),
        ("simt", C.c_int32), ("out", C.c_void_p), ("ldo", C.c_int32), ("bias", C.c_void_p), ("scale", C.c_void_p),
        ("row_map", C.c_void_p), ("row_pos", C.c_void_p), ("row_kvrow", C.c_void_p), ("cos_tab", C.c_void_p),
        ("sin_tab", C.c_void_p), ("k_cache", C.c_void_p), ("v_cache", C.c_void_p), ("kv_head_stride", C.c_int64),
        ("n_q_heads", C.c_int32), ("n_kv_heads", C.c_int32), ("pos_emb", C.c_void_p), ("patches", C.c_int32),
        ("gamma_next", C.c_void_p), ("a_next", C.c_void_p), ("lda_next", C.c_int32), ("ss_next", C.c_void_p),
        ("tile_counters", C.c_void_p), ("row_ss", C.c_void_p), ("inv_width", C.c_float), ("eps", C.c_float),
        ("stages", C.c_int32), ("pair", C.c_int32), ("tma_out", C.c_int32),
        ("partial", C.c_void_p), ("partial_stride", C.c_int64), ("partial_splits", C.c_void_p), ("ss_slots", C.c_int32),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("q_row_stride", C.c_int64), ("q2", C.c_void_p), ("q2_row_stride", C.c_int64),
        ("k", C.c_void_p), ("v", C.c_void_p), ("kv_row_stride", C.c_int64), ("kv_head_stride", C.c_int64),
        ("out", C.c_void_p), ("out_row_stride", C.c_int64), ("out2", C.c_void_p), ("out2_row_stride", C.c_int64),
        ("q_start", C.c_void_p), ("q_len", C.c_void_p), ("kv_start", C.c_void_p), ("kv_len", C.c_void_p),
        ("uni_q_len", C.c_int32), ("uni_kv_len", C.c_int32), ("max_q_len", C.c_int32), ("group", C.c_int32),
        ("n_kv_heads", C.c_int32), ("batch", C.c_int32), ("head_dim", C.c_int32), ("scale", C.c_float),
        ("causal", C.c_int32), ("small_rows", C.c_int32),
    ]


EXPORTS = [
    "vlaser_abi_version", "vlaser_last_error", "vlaser_create", "vlaser_set_weight", "vlaser_finalize",
    "vlaser_infer_action", "vlaser_infer_action_host", "vlaser_destroy", "vlaser_set_option", "vlaser_debug_copy",
    "vlaser_last_launch_count", "vlaser_flops_per_call", "vlaser_profile_read", "vlaser_op_gemm", "vlaser_op_attention",
    "vlaser_op_layernorm", "vlaser_op_rmsnorm", "vlaser_op_pixel_shuffle", "vlaser_op_im2col", "vlaser_op_euler_step",
    "vlaser_op_vit_attention", "vlaser_profile_read_classes", "vlaser_op_stage_frames", "vlaser_op_postprocess_actions",
    "vlaser_op_splitk_reduce", "vlaser_op_vit_attention_tc", "vlaser_op_normalize", "vlaser_op_prefill_attention_tc",
    "vlaser_flow_matching_loss", "vlaser_generate", "vlaser_op_decode_attention",
    "vlaser_op_prefill_attention_tc_long",
]

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Loads libvlaser_b200.so and declares the prototypes. Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m vlaser_b200.build` "
            "(vlaser_b200 has no CPU / PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    lib.vlaser_abi_version.restype = i32
    lib.vlaser_last_error.restype = C.c_char_p
    lib.vlaser_create.argtypes = [C.POINTER(VlaserCfg), C.POINTER(vp)]
    lib.vlaser_create.restype = i32
    lib.vlaser_set_weight.argtypes = [vp, C.c_char_p, vp, C.POINTER(i64), i32, i32]
    lib.vlaser_set_weight.restype = i32
    lib.vlaser_finalize.argtypes = [vp]
    lib.vlaser_finalize.restype = i32
    lib.vlaser_infer_action.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, vp]
    lib.vlaser_infer_action.restype = i32
    lib.vlaser_infer_action_host.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32]
    lib.vlaser_infer_action_host.restype = i32
    lib.vlaser_flow_matching_loss.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, f32, vp, vp, i32, i32, vp]
    lib.vlaser_flow_matching_loss.restype = i32
    lib.vlaser_generate.argtypes = [vp, vp, vp, vp, i32, i64, i64, vp, vp, i32, i32, vp]
    lib.vlaser_generate.restype = i32
    lib.vlaser_destroy.argtypes = [vp]
    lib.vlaser_destroy.restype = None
    lib.vlaser_set_option.argtypes = [vp, C.c_char_p, i64]
    lib.vlaser_set_option.restype = i32
    lib.vlaser_debug_copy.argtypes = [vp, C.c_char_p, vp, i64]
    lib.vlaser_debug_copy.restype = i64
    lib.vlaser_last_launch_count.argtypes = [vp]
    lib.vlaser_last_launch_count.restype = i64
    lib.vlaser_flops_per_call.argtypes = [vp, vp, i32]
    lib.vlaser_flops_per_call.restype = C.c_double
    lib.vlaser_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i64)]
    lib.vlaser_profile_read.restype = i32
    lib.vlaser_profile_read_classes.argtypes = [vp, C.POINTER(C.c_double)]
    lib.vlaser_profile_read_classes.restype = i32
    lib.vlaser_op_gemm.argtypes = [C.POINTER(GemmArgs), vp]
    lib.vlaser_op_gemm.restype = i32
    lib.vlaser_op_attention.argtypes = [C.POINTER(AttnArgs), vp]
    lib.vlaser_op_attention.restype = i32
    lib.vlaser_op_vit_attention.argtypes = [vp, vp, i32, i32, i32, f32, vp]
    lib.vlaser_op_vit_attention.restype = i32
    lib.vlaser_op_decode_attention.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i64, i32, i32, f32, vp, vp, vp, vp]
    lib.vlaser_op_decode_attention.restype = i32
    lib.vlaser_op_vit_attention_tc.argtypes = [vp, vp, i32, i32, i32, f32, vp]
    lib.vlaser_op_vit_attention_tc.restype = i32
    lib.vlaser_op_layernorm.argtypes = [vp, vp, vp, vp, i32, i32, f32, i32, i32, vp]
    lib.vlaser_op_layernorm.restype = i32
    lib.vlaser_op_rmsnorm.argtypes = [vp, vp, vp, i32, i32, f32, vp]
    lib.vlaser_op_rmsnorm.restype = i32
    lib.vlaser_op_splitk_reduce.argtypes = [vp, i32, i32, vp, i64, vp, vp, vp, f32, vp]
    lib.vlaser_op_splitk_reduce.restype = i32
    lib.vlaser_op_pixel_shuffle.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    lib.vlaser_op_pixel_shuffle.restype = i32
    lib.vlaser_op_im2col.argtypes = [vp, i32, vp, i32, i32, i32, i32, vp]
    lib.vlaser_op_im2col.restype = i32
    lib.vlaser_op_euler_step.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, f32, f32, i32, i32, f32, vp]
    lib.vlaser_op_euler_step.restype = i32
    lib.vlaser_op_stage_frames.argtypes = [vp, i32, i32, i32, i32, i32, vp, i32, i32, vp, vp, vp]
    lib.vlaser_op_stage_frames.restype = i32
    lib.vlaser_op_postprocess_actions.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, i32, i32, vp, vp, vp]
    lib.vlaser_op_postprocess_actions.restype = i32
    lib.vlaser_op_prefill_attention_tc.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i64, i32, f32, vp]
    lib.vlaser_op_prefill_attention_tc.restype = i32
    lib.vlaser_op_prefill_attention_tc_long.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i64, i32, i32, i32, i32, f32, vp]
    lib.vlaser_op_prefill_attention_tc_long.restype = i32
    lib.vlaser_op_normalize.argtypes = [vp, i32, vp, i32, i64, i32, i32, i32, vp, vp, vp]
    lib.vlaser_op_normalize.restype = i32
    _lib = lib
    return lib


def last_error() -> str:
    return load().vlaser_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    if rc < 0:
        raise RuntimeError(f"vlaser_b200 {what} failed (rc={rc}): {last_error()}")


def ptr(t) -> Optional[int]:
    """data pointer of a torch tensor (or None)."""
    return None if t is None else t.data_ptr()
