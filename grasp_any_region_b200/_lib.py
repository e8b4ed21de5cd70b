"""ctypes binding of libgar_b200.so (include/gar_b200.h).  No torch types cross this boundary:
only raw device pointers, sizes and the stream handle.  There is no CPU fallback: if the library
is missing or no CUDA device exists, the product path raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# GAR_B200_LIB: load another build of the same C ABI (A/B runs of two library versions on one GPU box)
LIB_PATH = os.environ.get("GAR_B200_LIB") or os.path.join(HERE, "libgar_b200.so")

_vp, _i, _f, _ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong


class GarModelDesc(C.Structure):
    _fields_ = [
        ("vit_dim", _i), ("vit_depth", _i), ("vit_heads", _i), ("vit_mlp", _i), ("vit_use_cls", _i),
        ("img_size", _i), ("patch_size", _i), ("ln_eps", _f), ("pool_ratio", _i),
        ("llm_dim", _i), ("llm_layers", _i), ("llm_heads", _i), ("llm_kv_heads", _i), ("llm_head_dim", _i),
        ("llm_inter", _i), ("vocab", _i), ("vocab_padded", _i), ("rms_eps", _f),
        ("prompt_numbers", _i), ("image_token_id", _ll), ("crop_token_ids", _ll * 8),
        ("max_tiles", _i), ("max_seq", _i), ("max_batch", _i), ("max_new_tokens", _i), ("max_prefill_rows", _i),
    ]


# name -> (restype, argtypes); mirrors include/gar_b200.h declaration by declaration
SIGNATURES = {
    "gar_last_error": (C.c_char_p, []),
    "gar_abi_version": (_i, []),
    "gar_copy_d2d": (_i, [_vp, _vp, _ll, _vp]),
    "gar_set_global_option": (_i, [C.c_char_p, _i]),
    "gar_gemm_skinny_bf16": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "gar_gemm_bf16": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp]),
    "gar_gemm_raster_probe": (_i, [_i, _i, _i, _i, _i, _i, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "gar_gemm_bf16_rope2d": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _i, _vp]),
    "gar_gemv_decode": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _i, _f, _vp,
                             C.POINTER(_i), _vp]),
    "gar_gemv_debug_timeline": (_i, [_vp]),
    "gar_embed_gather_sumsq": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "gar_patchify": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "gar_vit_embed_ln": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp]),
    "gar_layernorm": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _f, _vp]),
    "gar_rmsnorm": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _f, _vp]),
    "gar_vit_rope": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "gar_llama_rope_kv": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _i, _i, _vp]),
    "gar_swiglu": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "gar_pool2x2": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "gar_embed_gather": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "gar_argmax_step": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "gar_argmax_step_eos": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, C.POINTER(_i), _i, _i, _i, _vp, _vp]),
    "gar_bump_step": (_i, [_vp, _vp]),
    "gar_assemble_embeds": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, C.POINTER(_ll), C.POINTER(_f), C.POINTER(_i), _i,
                                 _f, _i, _i, _i, _ll, _i, _vp, _vp, _vp, _vp]),
    "gar_assemble_embeds_batch": (_i, [_i, C.POINTER(_vp), _i, _vp, _i, _vp, _i, _i, C.POINTER(_ll), C.POINTER(_f),
                                       C.POINTER(_i), _i, _f, C.POINTER(_i), C.POINTER(_i), _i, _ll, _i, _vp, _vp, _vp,
                                       _vp]),
    "gar_assemble_plan_batch": (_i, [_i, C.POINTER(_vp), _i, _i, C.POINTER(_ll), C.POINTER(_i), _i, _i, _ll, _i, _vp,
                                     _vp, _vp]),
    "gar_assemble_rows_batch": (_i, [_i, C.POINTER(_vp), _i, _vp, _i, _vp, _i, _i, C.POINTER(_ll), C.POINTER(_f),
                                     C.POINTER(_i), _i, _f, C.POINTER(_i), C.POINTER(_i), _i, _ll, _i, _vp, _vp, _vp]),
    "gar_resize_aa_u8_pass": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "gar_u8_to_tiles": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "gar_idmap_to_tiles": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "gar_attn_fwd": (_i, [_vp, _ll, _ll, _i, _vp, _vp, _ll, _ll, _i, _i, _i, _i, _i, _i, _vp, _ll, _ll, _i, _i, _i,
                          _i, _i, _i, _i, _i, _i, _f, _vp]),
    "gar_decode_attn_rope": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp,
                                  _i, _vp]),
    "gar_attn_debug_timeline": (_i, [_vp]),
    "gar_decode_attn": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _i, _vp]),
    "gar_engine_create": (_i, [C.POINTER(GarModelDesc), C.POINTER(_vp)]),
    "gar_engine_destroy": (_i, [_vp]),
    "gar_engine_set_weight": (_i, [_vp, C.c_char_p, _vp]),
    "gar_engine_get_buffer": (_i, [_vp, C.c_char_p, C.POINTER(_vp), C.POINTER(_ll)]),
    "gar_engine_vit_encode": (_i, [_vp, _vp, _vp, _i, _vp]),
    "gar_engine_patchify_at": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "gar_engine_vit_encode_patches": (_i, [_vp, _i, _i, _vp]),
    "gar_engine_assemble": (_i, [_vp, _i, _vp, _i, _i, C.POINTER(_f), C.POINTER(_i), _f, _i, _i, _vp]),
    "gar_engine_prefill": (_i, [_vp, _i, _i, _vp]),
    "gar_engine_assemble_at": (_i, [_vp, _i, _vp, _i, _i, _i, _i, C.POINTER(_f), C.POINTER(_i), _f, _i, _i, _vp]),
    "gar_engine_assemble_batch": (_i, [_vp, _i, _i, C.POINTER(_vp), _i, _i, C.POINTER(_f), C.POINTER(_i), _f,
                                       C.POINTER(_i), C.POINTER(_i), _vp]),
    "gar_engine_plan_batch": (_i, [_vp, _i, _i, C.POINTER(_vp), _i, _i, C.POINTER(_i), C.POINTER(_i), _vp]),
    "gar_engine_prefill_batch": (_i, [_vp, _i, _i, _i, _vp]),
    "gar_engine_prefill_all_logits": (_i, [_vp, _i, _i, _vp, _vp]),
    "gar_engine_generate": (_i, [_vp, _i, _i, _i, C.POINTER(_i), _vp]),
    "gar_engine_generate_group": (_i, [_vp, _i, _i, _i, _i, C.POINTER(_i), _vp]),
    "gar_engine_generate_eos": (_i, [_vp, _i, _i, _i, C.POINTER(_i), _i, C.POINTER(_i), C.POINTER(_i), _vp]),
    "gar_engine_status": (_i, [_vp, _i, C.POINTER(_i)]),
    "gar_engine_launch_count": (_ll, [_vp]),
    "gar_engine_partition": (_i, [_vp, _i, C.POINTER(_vp), C.POINTER(_i), C.POINTER(_vp), C.POINTER(_i)]),
    "gar_engine_set_option": (_i, [_vp, C.c_char_p, _i]),
    "gar_engine_set_replay": (_i, [_vp, _i, C.POINTER(_ll), _i]),
    "gar_engine_profile": (_i, [_vp, _i]),
    "gar_engine_profile_read": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_ll)]),
}

_lib = None


class GarError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the C-ABI library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GarError(
                f"{LIB_PATH} is missing: build it with `python -m grasp_any_region_b200.build` "
                "(there is no CPU or PyTorch fallback for this path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().gar_last_error().decode(errors="replace")
        raise GarError(f"{what} failed (code {rc}): {msg}")
