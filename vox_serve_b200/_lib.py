"""ctypes binding of libvoxb200.so (include/vb_api.h).  The product path has no CPU fallback: if the
library is missing or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libvoxb200.so")

c_void_p, c_int, c_int64, c_float, c_size_t, c_uint64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t, C.c_uint64
P = c_void_p

# name -> (restype, argtypes); mirrors include/vb_api.h one to one
SIGNATURES = {
    "vb_last_error": (C.c_char_p, []),
    "vb_version": (c_int, []),
    "vb_set_pdl": (c_int, [c_int]),
    "vb_device_info": (c_int, [P, P]),
    "vb_tensor_map_kv": (c_int, [P, P, c_int64, c_int, c_int, c_int, c_int]),
    "vb_tensor_map_2d_bf16": (c_int, [P, P, c_int64, c_int64, c_int64, c_int]),
    "vb_rmsnorm": (c_int, [P, P, P, c_int, c_int, c_float, c_int, P]),
    "vb_rope_freqs": (c_int, [P, c_int, c_int, c_float, c_float, c_int, c_float, c_float, c_float, P]),
    "vb_rope": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "vb_plan_rows": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, P]),
    "vb_decode_advance": (c_int, [P, P, P, c_int, P]),
    "vb_token_feedback": (c_int, [P, P, P, P, P, c_int, c_int, c_int, P]),
    "vb_gather_i32": (c_int, [P, P, P, c_int, P]),
    "vb_build_input_ids": (c_int, [P, P, P, P, c_int, P]),
    "vb_latest_window": (c_int, [P, P, P, c_int, c_int, P]),
    "vb_gather_windows": (c_int, [P, P, P, P, P, c_int, c_int, c_int, P]),
    "vb_kv_append": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, P]),
    "vb_add_layernorm": (c_int, [P, P, P, P, P, c_int, c_int, c_float, c_int, P]),
    "vb_gelu_add": (c_int, [P, P, P, c_int64, c_int, c_int, P]),
    "vb_chw_to_rows": (c_int, [P, P, c_int, c_int, c_int, P]),
    "vb_avgpool_rows": (c_int, [P, P, c_int, c_int, c_int, P]),
    "vb_vq_argmin": (c_int, [P, P, P, P, c_int, c_int, c_int, P]),
    "vb_copy_pages": (c_int, [P, P, P, c_int, c_int, c_int64, c_int64, c_int, P]),
    "vb_attn_tile_tokens": (c_int, [c_int, c_int]),
    "vb_paged_attn_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "vb_paged_attn": (c_int, [P, P, P, c_int64, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int,
                              c_float, P, c_size_t, c_int, c_int, c_int, P]),
    "vb_prefill_attn_tile_rows": (c_int, [c_int, c_int]),
    "vb_paged_prefill_attn": (c_int, [P, P, P, c_int64, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                      c_int, P]),
    "vb_prefill_attn_tc_tile_rows": (c_int, [c_int, c_int]),
    "vb_paged_prefill_attn_tc": (c_int, [P, P, P, c_int64, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                         c_int, P]),
    "vb_set_trace": (c_int, [P]),
    "vb_gemm_t_tile": (c_int, [c_int]),
    "vb_set_gemm_smem_kb": (c_int, [c_int, c_int]),
    "vb_weight_tiles_bytes": (c_size_t, [c_int, c_int, c_int]),
    "vb_pack_weight_tiles": (c_int, [P, P, c_int, c_int, c_int64, c_int, P]),
    "vb_gemm_bf16": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P]),
    "vb_proj_residual": (c_int, [P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "vb_proj_norm_gateup_silu": (c_int, [P, P, P, P, P, c_int, P, c_float, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "vb_proj_norm_qkv_rope_append": (c_int, [P, P, P, P, P, P, c_int, P, c_float, P, P, P, c_int, c_int, c_int, c_int,
                                             c_int, c_int, c_int, P]),
    "vb_norm_lmhead_workspace_bytes": (c_size_t, [c_int, c_int]),
    "vb_norm_lmhead": (c_int, [P, P, P, c_float, P, P, c_int, c_int, c_int, c_int, c_int, P, c_size_t, P]),
    "vb_rope_table": (c_int, [P, P, P, c_int, c_int, P]),
    "vb_row_ssq": (c_int, [P, P, c_int, c_int, P]),
    "vb_reduce_residual_rmsnorm": (c_int, [P, P, P, c_int, P, P, c_int, c_int, c_float, c_int, P]),
    "vb_qkv_rope_append": (c_int, [P, P, P, c_int, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P,
                                   c_float, P, P]),
    "vb_embedding": (c_int, [P, P, P, c_int, c_int, c_int, P]),
    "vb_gather_rows": (c_int, [P, P, P, c_int, c_int, c_int, P]),
    "vb_multi_embed_sum": (c_int, [P, c_int, P, c_int64, c_int64, P, P, c_int64, c_int64, c_int, c_int, P, c_int64, c_int,
                                   c_int, c_int, c_int, P]),
    "vb_talker_embed": (c_int, [P, c_int, P, c_int64, P, c_int64, P, c_int64, P, P, c_int64, c_int, c_int, P]),
    "vb_interleave_rows": (c_int, [P, P, P, c_int, c_int, P]),
    "vb_transpose_i64": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    "vb_sample_workspace_bytes": (c_size_t, [c_int, c_int]),
    "vb_sample": (c_int, [P, P, c_int, c_int, c_int, P, P, c_int, c_int, c_int, c_float, c_int, c_int, c_float,
                          c_float, c_float, c_uint64, c_uint64, P, c_int, P, c_size_t, P]),
    "vb_apply_repetition_penalty": (c_int, [P, P, P, c_int, c_int, c_int, c_float, c_int, c_int, P]),
    "vb_update_repetition_cache": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "vb_snac_from_codes": (c_int, [P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "vb_snac_dwconv7": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "vb_snac_pwconv": (c_int, [P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "vb_snac_convtr": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "vb_snac_final": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "vb_snac_tf32x3_bytes": (c_int64, [c_int, c_int, c_int]),
    "vb_snac_pack_tf32x3": (c_int, [P, P, c_int, c_int, c_int, P]),
    "vb_snac_pwconv_tc": (c_int, [P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "vb_snac_convtr_tc": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "vb_pcm16": (c_int, [P, P, c_int64, P]),
    "vb_randn": (c_int, [P, c_int64, c_uint64, c_uint64, P, P]),
    "vb_mimi_codes_sum": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "vb_mimi_conv": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "vb_mimi_convtr": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "vb_mimi_upsample": (c_int, [P, P, P, c_int, c_int, c_int, c_int, P]),
    "vb_mimi_layernorm": (c_int, [P, P, P, P, c_int, c_int, c_int, c_float, P]),
    "vb_mimi_attention": (c_int, [P, P, c_int, c_int, c_int, c_int, c_float, P]),
    "vb_codec_conv": (c_int, [P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "vb_codec_convtr": (c_int, [P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "vb_codec_conv_tc": (c_int, [P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "vb_codec_convtr_tc": (c_int, [P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "vb_codec_activate": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, P]),
    "vb_codec_cache_update": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "vb_codec_dwconv": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, P]),
    "vb_codec_rmsnorm": (c_int, [P, P, P, c_int, c_int, c_int, c_float, P]),
    "vb_codec_attn_chunk": (c_int, [P, P, P, c_int64, P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P]),
    "vb_orpheus_window_codes": (c_int, [P, P, P, P, c_int, c_int, P]),
}

_lock = threading.Lock()
_lib = None


class VoxB200Error(RuntimeError):
    pass


def load(build_if_missing: bool = False):
    """Load the shared library (once).  Never falls back to another implementation."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            if build_if_missing:
                from .build import build

                build()
            else:
                raise VoxB200Error(
                    f"{LIB_PATH} not found: build it with `python -m vox_serve_b200.build` "
                    "(no CPU / PyTorch fallback exists for this path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().vb_last_error()
        raise VoxB200Error(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


# kernels (+ memset nodes) each entry point enqueues; everything not listed launches exactly one
LAUNCHES = {"vb_set_pdl": 0, "vb_set_trace": 0, "vb_weight_tiles_bytes": 0, "vb_tensor_map_kv": 0,
            "vb_tensor_map_2d_bf16": 0, "vb_device_info": 0, "vb_sample": 1, "vb_update_repetition_cache": 1,
            "vb_set_gemm_smem_kb": 0}
launch_counter = [0]


def call(name: str, *args):
    """Invoke an int-returning entry point and raise on error."""
    check(getattr(load(), name)(*args), name)
    launch_counter[0] += LAUNCHES.get(name, 1)
