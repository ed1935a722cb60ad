"""ctypes binding of ``libmolly_b200.so`` (C ABI declared in ``include/molly_b200.h``).

The library is built in-tree by ``molly_b200/csrc/Makefile`` (``__graft_entry__.build()``).  There is NO fallback: if the
shared object is missing or fails to load, importing the ops raises.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import List

_HERE = os.path.dirname(os.path.abspath(__file__))
# MOLLY_LIB=<path>: load another build of the same library (A/B runs of kernel variants, tools/build_variant.sh)
LIB_PATH = os.environ.get("MOLLY_LIB") or os.path.join(_HERE, "libmolly_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "molly_b200.h")

# status codes (enum molly_status)
OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_WORKSPACE = 0, 1, 2, 3, 4
# enum molly_dtype / molly_position_type / molly_ffn_type / molly_epilogue / molly_err_bits
DTYPE_BF16, DTYPE_F32 = 0, 1
POS_ROTARY, POS_ABSOLUTE = 0, 1
FFN_GELU, FFN_GLU = 0, 1
EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL, EPI_GLU, EPI_SCATTER, EPI_BIAS_ROPE = 0, 1, 2, 3, 4, 5
ERRBIT_OOV, ERRBIT_OVERFLOW, ERRBIT_POSITION, ERRBIT_LAYOUT, ERRBIT_TOKEN = 1, 2, 4, 8, 16

c_void_pp = C.POINTER(C.c_void_p)


class EncoderConfig(C.Structure):
    """struct molly_encoder_config"""
    _fields_ = [
        ("hidden_size", C.c_int32), ("num_layers", C.c_int32), ("num_heads", C.c_int32),
        ("intermediate_size", C.c_int32), ("vocab_size", C.c_int32), ("pad_token_id", C.c_int32),
        ("mask_token_id", C.c_int32), ("position_type", C.c_int32), ("max_positions", C.c_int32),
        ("ffn_type", C.c_int32), ("token_dropout", C.c_int32), ("emb_layer_norm_before", C.c_int32),
        ("layer_norm_eps", C.c_float), ("llm_hidden_size", C.c_int32), ("project_token_num", C.c_int32),
    ]


class EncoderWeights(C.Structure):
    """struct molly_encoder_weights"""
    _fields_ = [
        ("word_emb_dev", C.c_void_p), ("pos_emb_dev", C.c_void_p),
        ("emb_ln_w_dev", C.c_void_p), ("emb_ln_b_dev", C.c_void_p),
        ("rope_cos_dev", C.c_void_p), ("rope_sin_dev", C.c_void_p), ("rope_len", C.c_int32),
        ("rope_cos_t_dev", C.c_void_p), ("rope_sin_t_dev", C.c_void_p),
        ("ln1_w_dev", c_void_pp), ("ln1_b_dev", c_void_pp),
        ("w_qkv_dev", c_void_pp), ("b_qkv_dev", c_void_pp),
        ("w_attn_out_dev", c_void_pp), ("b_attn_out_dev", c_void_pp),
        ("ln2_w_dev", c_void_pp), ("ln2_b_dev", c_void_pp),
        ("w_ffn1_dev", c_void_pp), ("b_ffn1_dev", c_void_pp),
        ("w_ffn2_dev", c_void_pp), ("b_ffn2_dev", c_void_pp),
        ("final_ln_w_dev", C.c_void_p), ("final_ln_b_dev", C.c_void_p),
        ("w_proj_dev", C.c_void_p), ("b_proj_dev", C.c_void_p),
    ]


class ProfileEntry(C.Structure):
    """struct molly_profile_entry"""
    _fields_ = [("name", C.c_char_p), ("launches", C.c_int32), ("work_is_flops", C.c_int32),
                ("total_ms", C.c_double), ("work", C.c_double)]


PROFILE_FAMILIES = 14
# enum molly_grad_slot / molly_grad_tail_slot (include/molly_b200.h)
(GRAD_LN2_W, GRAD_LN2_B, GRAD_B_FFN2, GRAD_B_FFN1, GRAD_LN1_W, GRAD_LN1_B, GRAD_B_O, GRAD_B_QKV, GRAD_W_FFN2, GRAD_W_FFN1,
 GRAD_W_O, GRAD_W_QKV, GRAD_SLOTS) = range(13)
GRAD_TAIL_FINAL_LN_W, GRAD_TAIL_FINAL_LN_B, GRAD_TAIL_WORD_EMB, GRAD_TAIL_POS_EMB, GRAD_TAIL_SLOTS = range(5)
_i32, _i64p, _vp, _sz = C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t

# name -> (restype, argtypes); must list every function declared in include/molly_b200.h (tests/test_host_logic.py::test_abi_library_exports_every_declared_symbol checks)
SIGNATURES = {
    "molly_encoder_create": (C.c_int, [C.POINTER(EncoderConfig), C.POINTER(EncoderWeights), C.POINTER(C.c_void_p)]),
    "molly_encoder_destroy": (None, [C.c_void_p]),
    "molly_encoder_workspace_bytes": (C.c_size_t, [C.c_void_p, _i32, _i32]),
    "molly_encode_project_merge_fwd": (C.c_int, [C.c_void_p, _vp, _vp, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _vp,
                                                 _sz, _vp, _vp, _vp]),
    "molly_encode_fwd": (C.c_int, [C.c_void_p, _vp, _i32, _i32, _vp, _vp, _sz, _vp, _vp]),
    "molly_pool_fwd": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "molly_placeholder_scan": (C.c_int, [_vp, _i32, _i32, C.POINTER(C.c_int64), _vp, _vp, _vp, _vp]),
    "molly_project_bwd": (C.c_int, [C.c_void_p, _vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i32, _vp,
                                    _sz, _vp]),
    "molly_gemm_bf16": (C.c_int, [_vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i32, _i32, _vp, _i32,
                                  _i32, _i32, _i32, _vp, _i32, C.c_float, _vp, _vp, _i32, _i32, _i32, _vp]),
    "molly_layernorm": (C.c_int, [_vp, _vp, _vp, _i32, _i32, C.c_float, _vp, _i32, _vp]),
    "molly_embed": (C.c_int, [_vp, _i32, _i32, C.POINTER(EncoderConfig), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "molly_rotary": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "molly_attention": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "molly_placeholder_runs": (C.c_int, [_vp, _i32, _i32, C.POINTER(C.c_int64), _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "molly_placeholder_reject": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "molly_build_seq_table": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "molly_embed_tokens_skip": (C.c_int, [_vp, _vp, C.POINTER(C.c_int64), _i32, _i32, _vp, _i32, _i32, _i32, _vp, _i32,
                                          _i32, _vp, _vp]),
    "molly_attention_lse": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "molly_attention_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "molly_linear_wgrad": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, C.c_size_t, _vp]),
    "molly_gather_rows": (C.c_int, [_vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "molly_transpose_bf16": (C.c_int, [_vp, _i32, _i32, _vp, _vp]),
    "molly_layernorm_bwd": (C.c_int, [_vp, _vp, _vp, _i32, _i32, C.c_float, _vp, _i32, _vp, _vp, _vp, _vp]),
    "molly_act_fwd_bwd": (C.c_int, [_i32, _vp, _vp, C.c_int64, _i32, _vp, _vp, _vp]),
    "molly_cast_f32_bf16": (C.c_int, [_vp, C.c_int64, _vp, _vp]),
    "molly_scale_cols": (C.c_int, [_vp, _i32, _i32, _i32, C.c_float, _vp]),
    "molly_scatter_add_rows": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp, _vp]),
    "molly_encoder_train_sizes": (C.c_int, [C.c_void_p, _i32, _i32, _i32, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                            C.POINTER(C.c_int64)]),
    "molly_encoder_grad_layout": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                            C.POINTER(C.c_int64)]),
    "molly_encode_train_fwd": (C.c_int, [C.c_void_p, _vp, _i32, _i32, _vp, _vp, _sz, _i32, _vp, _vp]),
    "molly_encode_train_bwd": (C.c_int, [C.c_void_p, _i32, _i32, _vp, _sz, _i32, _vp, _vp, _vp, _sz, _i32, _i32, _vp]),
    "molly_attention_debug": (C.c_int, [_vp]),
    "molly_merge_rows": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "molly_profile_start": (C.c_int, []),
    "molly_profile_stop": (C.c_int, [C.POINTER(ProfileEntry), _i32]),
    "molly_last_error": (C.c_char_p, []),
    "molly_abi_version": (C.c_int, []),
    "molly_kernel_launch_count": (C.c_int, []),
    "molly_add_kernel_launches": (C.c_int, [_i32]),
}

_lib = None


class MollyLibraryError(RuntimeError):
    pass


def declared_symbols() -> List[str]:
    """Function names declared in include/molly_b200.h."""
    with open(HEADER_PATH) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(molly_[a-z0-9_]+)\s*\(", src)))


def load():
    """Load the CUDA library; raises (no CPU fallback) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise MollyLibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"(make -C molly_b200/csrc). molly_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)            # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().molly_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status: int, what: str) -> None:
    """Map a C status to the Python exception the reference's call site would see."""
    if status == OK:
        return
    msg = f"{what}: {last_error()} (status {status})"
    if status in (ERR_INVALID, ERR_UNSUPPORTED):
        raise ValueError(msg)
    raise RuntimeError(msg)
