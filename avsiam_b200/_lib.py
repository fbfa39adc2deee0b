"""ctypes binding of libavsiam_b200.so (the C-ABI declared in include/avsiam_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, we raise.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_longlong, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# AVSIAM_B200_LIB: developer override (A/B and trace builds of the same C-ABI); the product default is the in-tree build
LIB_PATH = os.environ.get("AVSIAM_B200_LIB") or os.path.join(_HERE, "libavsiam_b200.so")


class GemmEpilogue(Structure):
    _fields_ = [
        ("flags", c_int),
        ("alpha", c_float),
        ("bias", c_void_p),
        ("resid", c_void_p),
        ("ld_resid", c_longlong),
        ("aux_in", c_void_p),
        ("aux_out", c_void_p),
        ("ld_aux", c_longlong),
        ("rowadd", c_void_p),
        ("rowidx", c_void_p),
        ("rowadd_rows", c_int),
        ("colsum", c_void_p),
    ]


EPI_GELU, EPI_DGELU, EPI_OUT_F32, EPI_OUT_ATOMIC, EPI_AUX_GRAD, EPI_MUL_AUX = 1, 2, 4, 8, 16, 32

_P, _I, _L, _F = c_void_p, c_int, c_longlong, c_float

# name -> argtypes (restype is int unless listed in _RESTYPES); must mirror include/avsiam_b200.h
SIGNATURES = {
    "avs_last_error": [],
    "avs_version": [],
    "avs_launch_count": [],
    "avs_reset_launch_count": [],
    "avs_reset": [],
    "avs_gemm_bf16": [_P, _L, _I, _P, _L, _I, _P, _L, _I, _I, _I, POINTER(GemmEpilogue), _I, _P],
    "avs_mask_argsort": [_P, _I, _I, _I, _P, _P, _P, _P],
    "avs_mask_force_noise": [_P, _I, _I, _I, _P, _I, _P, _I, _F, _P],
    "avs_gather_rows": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "avs_patchify_audio": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "avs_patchify_video": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "avs_scatter_add_rows": [_P, _P, _P, _I, _I, _I, _F, _P],
    "avs_decoder_restore_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "avs_decoder_restore_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "avs_layernorm_fwd": [_P, _P, _P, _F, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "avs_layernorm_bwd": [_P, _P, _F, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "avs_layernorm_fwd2": [_P, _P, _P, _I, _P, _P, _F, _P, _P, _P, _I, _I, _P],
    "avs_layernorm_bwd2": [_P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _I, _I, _P],
    "avs_seq_mean_fwd": [_P, _P, _I, _I, _I, _I, _I, _P],
    "avs_seq_mean_bwd": [_P, _P, _I, _I, _I, _I, _I, _P],
    "avs_eval_stats": [_P, _P, _I, _I, _P, _P, _P, _P, _P],
    "avs_bce_with_logits": [_P, _P, _P, _P, _L, _P],
    "avs_cross_entropy_prob": [_P, _P, _P, _P, _I, _I, _P],
    "avs_cosine_sim": [_P, _P, _I, _I, _I, _P, _P, _P],
    "avs_retrieval_ranks": [_P, _I, _P, _P, _P],
    "avs_fbank_augment": [_P, _P, _P, _P, _P, _I, _I, _I, _F, _F, _I, _P],
    "avs_frames_preprocess": [_P, _I, _I, _I, _I, _P, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    "avs_mix_frames": [_P, _P, _P, _I, _L, _P],
    "avs_fbank": [_P, _L, _I, _I, _I, _P, _P, _P, _P, _I, _F, _F, _P],
    "avs_head_fwd": [_P, _P, _P, _F, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "avs_head_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "avs_attention_fwd": [_P, _L, _P, _L, _P, _I, _I, _I, _I, _P],
    "avs_attention_bwd": [_P, _L, _P, _P, _L, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "avs_mae_loss_fwd": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _P, _P],
    "avs_mae_loss_bwd": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _P, _P, _P],
    "avs_infonce_workspace_bytes": [_I, _I],
    "avs_infonce_fwd": [_P, _P, _I, _I, _F, _I, _P, _P, _P, _P],
    "avs_infonce_bwd": [_I, _I, _F, _I, _F, _P, _P, _I, _I, _P, _P, _P, _P],
    "avs_adam_step": [_P, _P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _I, _I, _P, _P, _P, _P],
    "avs_adam_step_groups": [_P, _P, _P, _P, _P, _L, POINTER(c_float), POINTER(c_float), _I, _F, _F, _F, _I, _I, _P, _P,
                             _P, _P, _P],
    "avs_cast_f32_to_bf16": [_P, _P, _L, _P],
    "avs_colsum_bf16": [_P, _L, _P, _I, _I, _F, _P],
    "avs_found_inf": [_P, _L, _P, _P],
}
_RESTYPES = {
    "avs_last_error": c_char_p,
    "avs_launch_count": c_longlong,
    "avs_reset_launch_count": None,
    "avs_reset": None,
    "avs_infonce_workspace_bytes": c_size_t,
}

_lib = None


def lib() -> ctypes.CDLL:
    """Load (once) and return the CUDA library. Raises if it has not been built — there is no CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m avsiam_b200.build` (or __graft_entry__.build()). "
                "avsiam_b200 has no CPU / PyTorch fallback.")
        l = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if the symbol is missing
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, c_int)
        _lib = l
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().avs_last_error()
        raise RuntimeError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")
