"""ctypes binding of libcontrastors_b200.so (the C ABI declared in include/contrastors_b200.h).

The library is the product path: there is no CPU or PyTorch fallback.  ``load()`` raises if the shared object has
not been built (``python -m contrastors_b200.build``); compute calls raise ``RuntimeError`` with the library's
thread-local message when they fail (e.g. no sm_100 device).
"""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcontrastors_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "contrastors_b200.h")

_lib = None

_vp, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t

SIGNATURES = {
    "cx_last_error": (C.c_char_p, []),
    "cx_version": (_i, []),
    "cx_launch_count": (C.c_ulonglong, []),
    "cx_gemm_select_cluster": (_i, [_i]),
    "cx_gemm_bf16": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i64, _i64, _i64, _i, _i, _f, _vp]),
    "cx_gemm_qkv_rope": (_i, [_vp, _vp, _vp, _i, _i, _i, _i64, _i64, _i64, _vp, _vp, _i, _vp]),
    "cx_gemm_swiglu_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i64, _i64, _i64, _vp]),
    "cx_gemm_swiglu": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i64, _i64, _i64, _vp]),
    "cx_infonce_workspace_bytes": (_sz, [_i, _i, _i]),
    "cx_infonce_fwd": (_i, [_vp, _i64, _vp, _i64, _i, _i, _i, _f, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cx_infonce_mat_workspace_bytes": (_sz, [_i, _i, _i]),
    "cx_infonce_mat_fwd": (_i, [_vp, _i64, _vp, _i64, _i, _i, _i, _vp, _f, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cx_infonce_mat_bwd_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "cx_infonce_mat_bwd": (_i, [_vp, _i64, _vp, _i64, _i, _i, _i, _vp, _vp, _f, _vp, _vp, _vp, _i, _i, _vp, _f, _vp, _vp, _vp, _i64, _vp,
                                _i64, _vp, _vp, _vp, _vp]),
    "cx_infonce_bwd": (_i, [_vp, _i64, _vp, _i64, _i, _i, _i, _f, _vp, _vp, _vp, _i, _i, _vp, _f, _vp, _vp, _i64, _vp,
                            _i64, _i, _vp, _vp, _vp]),
    "cx_rows_to_bf16": (_i, [_vp, _i64, _vp, _i64, _vp, _i, _i, _i, _vp]),
    "cx_l2norm_bwd": (_i, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _vp]),
    "cx_add_layernorm_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _vp, _f, C.c_uint64, _vp]),
    "cx_layernorm_bwd_workspace_bytes": (_sz, [_i]),
    "cx_add_layernorm_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _f, C.c_uint64, _vp, _vp]),
    "cx_embed_layernorm_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f, C.c_uint64, _vp]),
    "cx_embed_layernorm_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i64, _f, C.c_uint64, _vp]),
    "cx_token_positions": (_i, [_vp, _i, _vp, _vp, _vp]),
    "cx_rope_inplace": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "cx_dq_finalize_rope": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "cx_dq_finalize": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "cx_swiglu_fwd": (_i, [_vp, _vp, _i64, _i, _vp]),
    "cx_swiglu_bwd": (_i, [_vp, _vp, _vp, _i64, _i, _vp]),
    "cx_mean_pool_fwd": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "cx_mean_pool_bwd": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "cx_embed_head_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "cx_embed_head_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "cx_grad_clip_workspace_bytes": (_sz, []),
    "cx_grad_clip_coef": (_i, [_vp, _i64, _f, _vp, _vp, _vp]),
    "cx_adamw_step": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _f, _i, _vp, _f, _i, _vp]),
    "cx_cast_f32_bf16": (_i, [_vp, _vp, _i64, _vp]),
    "cx_linear_bias_bf16": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i64, _i64, _vp]),
    "cx_colsum_bf16": (_i, [_vp, _i64, _i, _vp, _vp]),
    "cx_act_fwd": (_i, [_vp, _vp, _i64, _i, _vp]),
    "cx_act_bwd": (_i, [_vp, _vp, _vp, _i64, _i, _vp]),
    "cx_patchify": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "cx_vit_assemble_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "cx_vit_assemble_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "cx_cls_select_fwd": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "cx_cls_select_bwd": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "cx_attn_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp]),
    "cx_attn_pool_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp]),
    "cx_attn_pool_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp]),
    "cx_attn_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp]),
}


def declared_symbols():
    """Every entry point include/contrastors_b200.h declares."""
    with open(HEADER) as f:
        return sorted(set(re.findall(r"CX_API[^;(]*?\b(cx_[a-z0-9_]+)\s*\(", f.read())))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not built: run `python -m contrastors_b200.build` (there is no CPU/PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().cx_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"contrastors_b200 {what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(load().cx_launch_count())
