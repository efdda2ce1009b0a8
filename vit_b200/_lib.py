"""ctypes binding of the C ABI declared in include/vit_b200.h (the drop-in boundary).

There is no CPU or PyTorch fallback: if `libvitb200.so` is missing this module raises, and every
wrapper raises `RuntimeError(vitb200_strerror(rc))` on a non-zero return code.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvitb200_tl.so" if os.environ.get("VITB200_TIMELINE") == "1" else "libvitb200.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_GELU = 0, 1
LOSS_MSE, LOSS_L1, LOSS_CE, LOSS_GIVEN = 0, 1, 2, 3
SITE_EMB = 0


def site_attn(l: int) -> int:
    return 1 + 3 * l


def site_proj(l: int) -> int:
    return 2 + 3 * l


def site_mlp(l: int) -> int:
    return 3 + 3 * l


_p, _i, _f, _u32, _sz = C.c_void_p, C.c_int, C.c_float, C.c_uint32, C.c_size_t


class EmbedFwdArgs(C.Structure):  # vitb200_embed_fwd_args
    _fields_ = [(n, _i) for n in ("B", "L", "P", "S", "Np", "n_valid", "H")] + [("eps", _f), ("p_drop", _f)] + \
               [(n, _p) for n in ("rng", "x", "w_p", "b_p", "cls", "pos", "ln_g", "ln_b", "w_qkv", "b_qkv", "z0", "u",
                                  "mean", "rstd", "qkv")]


class LayerFwdArgs(C.Structure):  # vitb200_layer_fwd_args
    _fields_ = [(n, _i) for n in ("B", "T", "H", "last")] + [("eps", _f), ("p_drop", _f), ("rng", _p),
                                                              ("site_proj", _u32), ("site_mlp", _u32)] + \
               [(n, _p) for n in ("ctx", "z_in", "w_o", "w_1", "w_2", "w_qkv", "b_o", "ln2_g", "ln2_b", "b_1", "b_2",
                                  "lnn_g", "lnn_b", "b_qkv", "hmid", "u2", "mean2", "rstd2", "a", "m", "z_out", "u_next",
                                  "mean_n", "rstd_n", "qkv_next")]

class LayerBwdUpperArgs(C.Structure):  # vitb200_layer_bwd_upper_args
    _fields_ = [(n, _i) for n in ("B", "T", "H")] + [("p_drop", _f), ("rng", _p), ("site_proj", _u32), ("site_mlp", _u32)] + \
               [(n, _p) for n in ("dz", "dz_cls", "m", "a", "u2", "ctx", "hmid", "mean2", "rstd2", "ln2_g", "w_2", "w_1", "w_o", "dh",
                                  "dctx", "gpart")] + \
               [(n, _i) for n in ("n_opt", "off_w2", "off_b2", "off_w1", "off_b1", "off_ln2g", "off_ln2b", "off_wo", "off_bo")]


class LayerBwdLowerArgs(C.Structure):  # vitb200_layer_bwd_lower_args
    _fields_ = [(n, _i) for n in ("B", "T", "H")] + \
               [(n, _p) for n in ("dqkv", "u", "z", "mean1", "rstd1", "ln1_g", "dh", "w_qkv", "dz", "gpart")] + \
               [(n, _i) for n in ("n_opt", "off_wqkv", "off_bqkv", "off_ln1g", "off_ln1b")]


class EmbedBwdArgs(C.Structure):  # vitb200_embed_bwd_args
    _fields_ = [(n, _i) for n in ("B", "L", "P", "S", "Np", "n_valid", "H")] + [("p_drop", _f)] + \
               [(n, _p) for n in ("rng", "dz0", "x", "gpart")] + [(n, _i) for n in ("n_opt", "off_wp", "off_bp", "off_cls")]


class MegaFwdArgs(C.Structure):  # vitb200_mega_fwd_args
    _fields_ = [(n, _i) for n in ("B", "L", "P", "S", "Np", "n_valid", "layers", "C", "loss_kind", "cluster", "cls_only")] + \
               [("eps", _f), ("p_hidden", _f), ("p_attn", _f)] + \
               [(n, _p) for n in ("rng", "x", "labels", "params", "shadow")] + \
               [(n, _i) for n in ("off_cls", "off_pos", "off_wp", "off_bp", "off_layer0", "layer_stride", "o_ln1g", "o_ln1b",
                                  "o_wqkv", "o_bqkv", "o_wo", "o_bo", "o_ln2g", "o_ln2b", "o_w1", "o_b1", "o_w2", "o_b2",
                                  "off_lnfg", "off_lnfb", "off_wh", "off_bh")] + \
               [(n, _p) for n in ("rope_cos", "rope_sin", "z", "hmid", "u", "u2", "qkv", "ctx", "a", "m", "stats", "lse",
                                  "s_cls", "logits", "loss", "ws", "rows", "rows_base", "loss_log")] + [("defer_loss", _i)]


class MegaBwdArgs(C.Structure):  # vitb200_mega_bwd_args
    _fields_ = [("f", MegaFwdArgs), ("labels", _p), ("gloss", _p), ("loss_kind", _i), ("n_opt", _i), ("gpart", _p), ("dz0", _p),
                ("done", _p)]


MAX_GROUPS = 20


class GradStream(C.Structure):  # vitb200_grad_stream
    _fields_ = [("done", _p), ("expect", C.c_uint), ("n_groups", _i), ("lo", C.c_uint * MAX_GROUPS), ("hi", C.c_uint * MAX_GROUPS)]


# name -> (restype, argtypes); order and meaning follow include/vit_b200.h exactly
SIGNATURES = {
    "vitb200_strerror": (C.c_char_p, [_i]),
    "vitb200_last_cuda_error": (C.c_char_p, []),
    "vitb200_version": (_i, []),
    "vitb200_init": (_i, [_i]),
    "vitb200_patch_embed_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _p, _u32, _i, _p]),
    "vitb200_patch_embed_bwd_ws_bytes": (_sz, [_i, _i, _i, _i]),
    "vitb200_patch_embed_bwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _p, _u32, _i, _i, _p, _p]),
    "vitb200_add_ln_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _f, _p, _u32, _i, _p]),
    "vitb200_add_ln_bwd_ws_bytes": (_sz, [_i, _i]),
    "vitb200_add_ln_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _p, _u32, _i, _i, _p, _p]),
    "vitb200_linear_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "vitb200_linear_dgrad": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "vitb200_linear_wgrad_ws_bytes": (_sz, [_i, _i, _i]),
    "vitb200_linear_wgrad": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p]),
    "vitb200_tc_supported": (_i, [_i, _i, _i]),
    "vitb200_tc_linear_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "vitb200_tc_linear_dgrad": (_i, [_p, _p, _p, _p, _i, _i, _i, _p]),
    "vitb200_tc_linear_wgrad_ws_bytes": (_sz, [_i, _i, _i]),
    "vitb200_tc_linear_wgrad": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p, _p]),
    "vitb200_set_gemm_mode": (_i, [_i]),
    "vitb200_fused_supported": (_i, [_i, _i]),
    "vitb200_fused_embed_fwd": (_i, [_p, _p]),
    "vitb200_fused_layer_fwd": (_i, [_p, _p]),
    "vitb200_fused_embed_bwd_supported": (_i, [_i, _i, _i]),
    "vitb200_fused_embed_bwd": (_i, [_p, _p]),
    "vitb200_fused_bwd_supported": (_i, [_i]),
    "vitb200_fused_bwd_grid": (_i, [_i]),
    "vitb200_fused_layer_bwd_upper": (_i, [_p, _p]),
    "vitb200_fused_layer_bwd_lower": (_i, [_p, _p]),
    "vitb200_grad_reduce": (_i, [_p, _i, _sz, _sz, _sz, _p, _p]),
    "vitb200_mega_supported": (_i, [_i, _i, _i, _i, _i, _i, _i]),
    "vitb200_mega_ws_bytes": (_sz, []),
    "vitb200_mega_fwd_smem_bytes": (_sz, [_i]),
    "vitb200_mega_grid": (_i, [_i, _i]),
    "vitb200_mega_fwd": (_i, [_p, _p]),
    "vitb200_mega_bwd_supported": (_i, [_i, _i, _i, _i, _i, _i, _i, _i]),
    "vitb200_mega_bwd_smem_bytes": (_sz, [_i]),
    "vitb200_mega_bwd": (_i, [_p, _p]),
    "vitb200_attn_fwd": (_i, [_p, _p, _p, _i, _p, _p, _p, _p, _i, _i, _i, _i, _f, _f, _p, _u32, _i, _p]),
    "vitb200_attn_bwd": (_i, [_p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _i, _i, _i, _i, _f, _f, _p, _u32, _i, _p]),
    "vitb200_set_attn_mode": (_i, [_i]),
    "vitb200_attn_tc_supported": (_i, [_i, _i, _i, _i]),
    "vitb200_attn_tc_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _f, _p, _u32, _p]),
    "vitb200_attn_tc_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _f, _p, _u32, _p]),
    "vitb200_attn_tc_blocked_supported": (_i, [_i, _i, _i, _i]),
    "vitb200_attn_tc_blocked_ws_bytes": (_sz, [_i, _i, _i, _i]),
    "vitb200_attn_tc_blocked_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _f, _p, _u32, _p, _p]),
    "vitb200_attn_flash_supported": (_i, [_i, _i, _i, _i]),
    "vitb200_attn_flash_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _f, _p, _u32, _p]),
    "vitb200_attn_flash_bwd_ws_bytes": (_sz, [_i, _i, _i, _i]),
    "vitb200_attn_flash_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _f, _p, _u32, _p, _p]),
    "vitb200_attn_probs": (_i, [_p, _p, _i, _p, _p, _p, _p, _i, _i, _i, _i, _f, _i, _p]),
    "vitb200_head_loss_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "vitb200_head_loss_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "vitb200_head_fused_supported": (_i, [_i, _i]),
    "vitb200_head_fused_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "vitb200_head_fused_bwd": (_i, [_p, _p, _p, _p, _p, _p, _sz, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "vitb200_head_fused_fwd_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _sz, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "vitb200_grad_norm_ws_bytes": (_sz, [_sz]),
    "vitb200_grad_norm": (_i, [_p, _sz, _p, _p, _p, _p]),
    "vitb200_adamw": (_i, [_p, _p, _p, _p, _p, _sz, _p, _p, _p, _p]),
    "vitb200_clip_adamw_fused_ws_bytes": (_sz, []),
    "vitb200_clip_adamw_fused": (_i, [_p, _p, _p, _p, _p, _sz, _p, _p, _p, _p, _i, _sz, _sz, _sz, _p, _p]),
    "vitb200_peer_buffer_bytes": (_sz, [_sz]),
    "vitb200_peer_alloc": (_i, [_sz, C.POINTER(C.c_void_p), C.c_char_p]),
    "vitb200_peer_open": (_i, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "vitb200_peer_close": (_i, [_p]),
    "vitb200_peer_free": (_i, [_p]),
    "vitb200_sumsq_accum": (_i, [_p, _sz, _p, _p, _p]),
    "vitb200_clip_adamw_fused_dp": (_i, [_p, _p, _p, _p, _p, _sz, _p, _p, _p, _p, _i, _sz, _sz, _sz, _p, _p, _i, _i, _p]),
    "vitb200_clip_adamw_fused_streamed": (_i, [_p, _p, _p, _p, _p, _sz, _p, _p, _p, _p, _i, _sz, _sz, _sz, _p, _p, _p, _i, _i, _p]),
    "vitb200_cast_bf16": (_i, [_p, _p, _sz, _p]),
    "vitb200_gelu_fwd": (_i, [_p, _p, _sz, _i, _p]),
    "vitb200_residual_add": (_i, [_p, _p, _p, _sz, _i, _p]),
    "vitb200_cast_f32": (_i, [_p, _p, _sz, _p]),
    "vitb200_tc_prelinear_ws_bytes": (_sz, [_i, _i, _i]),
    "vitb200_tc_prelinear_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _p, _p]),
    "vitb200_zca_lowrank_supported": (_i, [_i, _i]),
    "vitb200_zca_lowrank_fwd": (_i, [_p, _p, _p, _f, _p, _p, _i, _i, _i, _i, _p]),
    "vitb200_patch_embed_dgrad": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _p, _u32, _i, _p]),
    "vitb200_gather_batch": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, C.c_longlong, _f, _p, _p]),
    "vitb200_eval_metrics_accum": (_i, [_p, _p, _p, _p, _i, _i, _i, _p]),
    "vitb200_dropout_mask": (_i, [_p, _sz, _f, _p, _u32, _p]),
}

_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """Load libvitb200.so (once).  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m vit_b200.build` (nvcc, sm_100a). "
                "vit_b200 has no CPU/PyTorch fallback for its kernels."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        lib = load()
        msg = lib.vitb200_strerror(rc).decode()
        if rc == -1:
            msg += ": " + lib.vitb200_last_cuda_error().decode()
        raise RuntimeError(f"vit_b200 {what} failed ({rc}): {msg}")
