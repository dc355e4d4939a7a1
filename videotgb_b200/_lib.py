"""ctypes loader for libvideotgb_b200.so (the C ABI of include/videotgb_b200.h).

The product path has no CPU or PyTorch fallback: if the library is missing (run `python -m videotgb_b200.build`
or `__graft_entry__.build()`), or the device is not sm_100, every op raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TGB_LIB") or os.path.join(_HERE, "libvideotgb_b200.so")    # TGB_LIB: experiment builds (tools/)

TGB_OK, TGB_EINVAL, TGB_EARCH, TGB_ECUDA = 0, -1, -2, -3

c_int, c_i64, c_f32, c_vp, c_sz = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t

# name -> argtypes; must list every symbol declared in include/videotgb_b200.h (tests/test_abi.py checks this)
SIGNATURES = {
    "tgb_last_error": ([], ctypes.c_char_p),
    "tgb_version": ([], c_int),
    "tgb_device_check": ([], c_int),
    "tgb_sm_count": ([], c_int),
    "tgb_normalize_frames_u8": ([c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp], c_int),
    "tgb_resize_bilinear_u8": ([c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp], c_int),
    "tgb_gumbel_argmax": ([c_vp, c_vp, c_f32, c_int, c_int, c_vp, c_vp], c_int),
    "tgb_additive_mask": ([c_vp, c_int, c_int, c_f32, c_vp, c_vp, c_vp], c_int),
    "tgb_span_draw": ([c_vp, c_vp, c_vp, c_int, c_f32, c_int, c_int, c_vp, c_vp, c_vp], c_int),
    "tgb_spans_to_indices": ([c_vp, c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp], c_int),
    "tgb_pseudo_span": ([c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp], c_int),
    "tgb_span_ce_loss": ([c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp], c_int),
    "tgb_gather_frames": ([c_vp, c_vp, c_vp, c_int, c_int, c_int, c_sz, c_vp], c_int),
    "tgb_patchify_f32": ([c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp], c_int),
    "tgb_patchify_u8": ([c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp], c_int),
    "tgb_fill_cls_rows": ([c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp], c_int),
    "tgb_gemm_bf16": ([c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_int, c_int, c_vp, c_vp, c_i64, c_vp, c_int, c_int,
                       c_vp, c_i64, c_int, c_vp], c_int),
    "tgb_gemm_bf16_ln": ([c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_int, c_int, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_i64,
                          c_vp, c_i64, c_vp, c_vp], c_int),
    "tgb_gemm_ln_stats_slots": ([c_int, c_int], c_int),
    "tgb_ln_stats_finalize": ([c_vp, c_int, c_int, c_int, c_f32, c_vp, c_vp, c_vp], c_int),
    "tgb_rowstats_cast_f32": ([c_vp, c_i64, c_int, c_int, c_f32, c_vp, c_i64, c_vp, c_vp], c_int),
    "tgb_layernorm_f32": ([c_vp, c_i64, c_vp, c_vp, c_f32, c_int, c_int, c_vp, c_i64, c_vp, c_i64, c_vp], c_int),
    "tgb_attention_bf16": ([c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_int, c_int, c_int, c_int,
                            c_int, c_f32, c_vp], c_int),
    "tgb_self_attention_2seg_bf16": ([c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_int, c_int, c_int, c_int, c_int, c_f32,
                                      c_vp], c_int),
    "tgb_of_patch_reduce": ([c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp], c_int),
    "tgb_sampler_assemble": ([c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_f32, c_int, c_int, c_int, c_vp, c_vp, c_vp], c_int),
    "tgb_embed_layernorm": ([c_vp, c_vp, c_i64, c_vp, c_int, c_vp, c_vp, c_vp, c_f32, c_i64, c_int, c_vp, c_i64, c_vp,
                             c_i64, c_vp], c_int),
    "tgb_rope_bf16": ([c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_int, c_vp], c_int),
    "tgb_linear_small_f32": ([c_vp, c_i64, c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_i64, c_vp], c_int),
    "tgb_vit_attention_bf16": ([c_vp, c_i64, c_vp, c_i64, c_int, c_f32, c_vp], c_int),
    "tgb_llm_inputs_f32": ([c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp], c_int),
    "tgb_replace_ids": ([c_vp, c_i64, c_i64, c_i64, c_vp], c_int),
    "tgb_im2col_nhwc_bf16": ([c_vp, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_i64, c_vp], c_int),
    "tgb_instance_norm_nhwc": ([c_vp, c_int, c_int, c_int, c_f32, c_int, c_vp, c_vp, c_vp, c_vp], c_int),
    "tgb_act_f32": ([c_vp, c_i64, c_vp, c_i64, c_i64, c_int, c_int, c_f32, c_vp, c_i64, c_vp, c_i64, c_vp], c_int),
    "tgb_gru_rh": ([c_vp, c_vp, c_i64, c_int, c_vp, c_i64, c_vp], c_int),
    "tgb_gru_update": ([c_vp, c_vp, c_vp, c_i64, c_int, c_vp, c_i64, c_vp], c_int),
    "tgb_corr_pool": ([c_vp, c_vp, c_i64, c_int, c_int, c_vp], c_int),
    "tgb_corr_lookup": ([c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_i64, c_vp], c_int),
    "tgb_raft_upsample": ([c_vp, c_i64, c_vp, c_int, c_int, c_int, c_vp, c_vp], c_int),
    "tgb_raft_prep_image": ([c_vp, c_int, c_int, c_int, c_vp, c_i64, c_vp], c_int),
    "tgb_cast_f32_bf16": ([c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_vp], c_int),
    "tgb_group_mean_f32": ([c_vp, c_vp, c_int, c_int, c_i64, c_vp], c_int),
}

# test / A-B knobs (csrc/tgb_debug.h): exported by the library, not part of the contract in include/videotgb_b200.h
DEBUG_SIGNATURES = {
    "tgb_resize_set_variant": ([c_int], c_int),
    "tgb_patchify_set_variant": ([c_int], c_int),
    "tgb_gemm_set_cta_group": ([c_int], c_int),
    "tgb_gemm_set_schedule": ([c_int], c_int),
    "tgb_gemm_set_epilogue_warps": ([c_int], c_int),
    "tgb_gemm_set_resid_piece": ([c_int], c_int),
    "tgb_vit_attention_set_variant": ([c_int], c_int),
    "tgb_vit_attention_set_prefetch": ([c_int], c_int),
    "tgb_attention_set_tc": ([c_int], c_int),
}

_lib = None


class TgbError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TgbError(f"{LIB_PATH} is missing -- build it with `python -m videotgb_b200.build`; there is no fallback path")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (args, res) in {**SIGNATURES, **DEBUG_SIGNATURES}.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    _lib = lib
    sched = {"roundrobin": 0, "balanced": 1, "rowblocks": 2, "groups": 3}.get(os.environ.get("TGB_GEMM_SCHEDULE", ""))   # A/B knob
    if sched is not None:
        lib.tgb_gemm_set_schedule(sched)
    if os.environ.get("TGB_GEMM_RESID_PIECE"):                                                                # A/B knob
        lib.tgb_gemm_set_resid_piece(int(os.environ["TGB_GEMM_RESID_PIECE"]))
    if os.environ.get("TGB_ATTN_VARIANT"):                                                                    # A/B knob
        lib.tgb_vit_attention_set_variant(int(os.environ["TGB_ATTN_VARIANT"]))
    if os.environ.get("TGB_ATTN_TC"):                                                                         # A/B knob
        lib.tgb_attention_set_tc(int(os.environ["TGB_ATTN_TC"]))
    if os.environ.get("TGB_ATTN_PREFETCH"):                                                                   # A/B knob
        lib.tgb_vit_attention_set_prefetch(int(os.environ["TGB_ATTN_PREFETCH"]))
    return lib


def check(status: int) -> None:
    if status != TGB_OK:
        msg = load().tgb_last_error().decode(errors="replace")
        kind = {TGB_EINVAL: "invalid argument", TGB_EARCH: "unsupported device", TGB_ECUDA: "CUDA error"}.get(status, "error")
        raise TgbError(f"videotgb_b200 {kind} ({status}): {msg}")
