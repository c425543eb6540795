"""ctypes binding of liblaud_b200.so (the C ABI declared in include/laud_b200.h).

There is no fallback: if the CUDA library has not been built, importing the
operators raises.  Build it with `python -m laudnet_b200.build` (or
`__graft_entry__.build()`); the .so is kept in-tree under laudnet_b200/lib/.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LAUD_LIB") or os.path.join(_HERE, "lib", "liblaud_b200.so")   # LAUD_LIB: diagnostic builds

CONV_AUTO, CONV_UMMA, CONV_HMMA, CONV_NAIVE = 0, 1, 2, 3
RELU_NONE, RELU_ALL, RELU_WHERE_GATE0 = 0, 1, 2
GAP_SPLITS = 8
PREBIAS_CLASSES = 16

_vp, _i, _fp, _u8p, _i32p, _i64 = C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64


class ConvDesc(C.Structure):
    """Mirror of `struct laud_conv_desc`."""
    _fields_ = [
        ("x", _vp), ("ldx", C.c_int32),
        ("w", _vp),
        ("y", _vp), ("ldy", C.c_int32),
        ("B", C.c_int32), ("H_in", C.c_int32), ("W_in", C.c_int32), ("C_in", C.c_int32),
        ("H_out", C.c_int32), ("W_out", C.c_int32), ("C_out", C.c_int32),
        ("ksize", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
        ("scale", _vp), ("shift", _vp),
        ("relu_mode", C.c_int32),
        ("residual", _vp), ("ldr", C.c_int32),
        ("k_idx", _vp), ("k_cnt", _vp), ("k_ld", C.c_int32), ("k_gran", C.c_int32),
        ("n_idx", _vp), ("n_cnt", _vp), ("n_ld", C.c_int32), ("n_gran", C.c_int32),
        ("pre_bias", _vp), ("pre_bias_classes", C.c_int32), ("pre_bias_ld", C.c_int32),
        ("out_mask", _vp), ("mask_groups", C.c_int32),
        ("sample_idx", _vp), ("sample_cnt", _vp),
        ("row_idx", _vp), ("row_cnt", _vp),
        ("n_pad_align", C.c_int32),
        ("gap_partial", _vp), ("gap_tiles", C.c_int32),
        ("w_t", _vp),
        ("bias_t", _vp), ("bias_ld", C.c_int32),
        ("n_mask", _vp), ("n_mask_gran", C.c_int32),
        ("n_expand", C.c_int32),
    ]


class TokGemmDesc(C.Structure):
    """Mirror of `struct laud_tok_gemm_desc` (include/laud_adavit.h)."""
    _fields_ = [
        ("a", _vp), ("lda", C.c_int32),
        ("w", _vp),
        ("bias", _vp),
        ("rows_max", C.c_int32), ("K", C.c_int32), ("N", C.c_int32),
        ("row_cnt", _vp),
        ("act", C.c_int32),
        ("out", _vp), ("ldo", C.c_int32),
        ("resid", _vp), ("ldres", C.c_int32),
        ("row_idx", _vp),
        ("col_gate", _vp), ("gate_ld", C.c_int32),
        ("row_sample", _vp),
        ("bn", C.c_int32),
        ("cta_pair", C.c_int32),
    ]


ACT_NONE, ACT_GELU = 0, 1

# name -> argtypes; every symbol include/*.h declares must be listed here
SIGNATURES = {
    "laud_abi_version": ([], C.c_int),
    "laud_last_error": ([], C.c_char_p),
    "laud_launch_count": ([], C.c_ulonglong),
    "laud_conv_path_counts": ([C.POINTER(C.c_ulonglong * 3)], None),
    "laud_conv_tma_launch_count": ([], C.c_ulonglong),
    "laud_conv_profile": ([_i], None),
    "laud_conv_profile_read": ([C.POINTER(C.c_float)], _i),
    "laud_conv_profile_read_all": ([C.POINTER(C.c_float), _i], _i),
    "laud_masker_channel_mlp": ([_vp, _i, _i, _i, _i, _fp, _fp, _i, _fp, _fp, _i, _fp, _fp, _fp, _u8p, _i32p, _i32p, _i32p, _vp], _i),
    "laud_masker_channel_from_pooled": ([_fp, _i, _i, _i, _fp, _fp, _i, _fp, _fp, _i, _fp, _u8p, _i32p, _i32p, _i32p, _vp], _i),
    "laud_masker_channel_from_partials": ([_fp, _i, _i, _i, _i, _i, _fp, _fp, _i, _fp, _fp, _i, _fp, _fp, _u8p, _i32p, _i32p, _i32p, _vp], _i),
    "laud_global_avg_pool": ([_vp, _i, _i, _i, _i, _fp, _fp, _vp], _i),
    "laud_masker_spatial": ([_vp, _i, _i, _i, _i, _fp, _fp, _i, _i, _fp, _u8p, _i32p, _vp], _i),
    "laud_expand_mask": ([_u8p, _i, _i, _i, _i, _i, _i, _u8p, _i32p, _vp], _i),
    "laud_resize_mask_nearest": ([_u8p, _i, _i, _i, _i, _u8p, _vp], _i),
    "laud_broadcast_gate": ([_u8p, _i, _i, _i, _u8p, _i32p, _vp], _i),
    "laud_spatial_masks": ([_u8p, _i, _i, _i, _i, _i, _u8p, _u8p, _u8p, _i32p, _i32p, _vp], _i),
    "laud_layer_gate_lists": ([_u8p, _i, _i, _i, _i32p, _i32p, _i32p, _vp], _i),
    "laud_compact_rows": ([_u8p, _i, _i, _i, _i32p, _i32p, _i32p, _vp], _i),
    "laud_conv_forward": ([C.POINTER(ConvDesc), _i, _vp], _i),
    "laud_conv_set_pdl": ([_i], None),
    "laud_gate_from_logits": ([_fp, _fp, _i, _i, _i, C.c_float, _u8p, _i32p, _i32p, _i32p, _vp], _i),
    "laud_gate_inactive": ([_u8p, _i, _i, _i, _vp, _vp], _i),
    "laud_channel_consts_fold": ([_vp, _i, _i, _i, _i32p, _i32p, _i, _i, _i, _fp, _fp, _vp], _i),
    "laud_regnet_stem_forward": ([_vp, _i, _i, _i, _vp, _i, _fp, _fp, _vp, _vp], _i),
    "laud_grouped_conv3x3_forward": ([_vp, _i, _i, _i, _i, _i, _vp, _i, _fp, _fp, _u8p, _i, _vp, _vp], _i),
    "laud_se_gate": ([_fp, _i, _i, _fp, _fp, _i, _fp, _fp, _u8p, _i, _fp, _vp], _i),
    "laud_scale_channels": ([_vp, _i, _i, _i, _fp, _vp], _i),
    "laud_stem_forward": ([_vp, _i, _i, _i, _vp, _i, _fp, _fp, _vp, _vp], _i),
    "laud_head_forward": ([_vp, _i, _i, _i, _vp, _fp, _i, _fp, _fp, _vp], _i),
    "laud_head_forward_from_partials": ([_fp, _i, _i, _i, _i, _vp, _fp, _i, _fp, _fp, _vp], _i),
    "laud_nchw_to_nhwc_f16": ([_vp, _i, _i, _i, _i, _i, _vp, _i, _vp], _i),
    "laud_nhwc_f16_to_nchw_f32": ([_vp, _i, _i, _i, _i, _i, _fp, _vp], _i),
    "laud_forward_stats": ([_i32p, _vp, _i, _i64, _i64, _i64, _fp, _vp], _i),
    # include/laud_adavit.h
    "laud_tok_gemm": ([C.POINTER(TokGemmDesc), _vp], _i),
    "laud_tok_gemm_launch_count": ([], C.c_ulonglong),
    "laud_adavit_mlp_fused": ([_vp, _i, _i, _i, _i32p, _vp, _fp, _vp, _fp, _fp, _i, _i32p, _vp], _i),
    "laud_vit_patchify": ([_vp, _i, _i, _i, _vp, _vp], _i),
    "laud_vit_init_tokens": ([_fp, _i, _i, _i, _fp, _fp, _vp], _i),
    "laud_adavit_policy": ([_fp, _i, _i, _i, _i, C.c_float, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp,
                            _u8p, _i32p, _u8p, _u8p, _fp, _fp, _fp, _vp], _i),
    "laud_adavit_lists": ([_i32p, _u8p, _i, _i32p, _i32p, _vp], _i),
    "laud_adavit_ln_gather": ([_fp, _i, _i, _i, C.c_float, _fp, _fp, _u8p, _i32p, _vp, _i32p, _i32p, _vp], _i),
    "laud_adavit_row_lists": ([_u8p, _i, _i, _i32p, _i32p, _i32p, _i32p, _i32p, _vp], _i),
    "laud_adavit_ln_rows": ([_fp, _i, C.c_float, _fp, _fp, _i32p, _i32p, _i, _vp, _vp], _i),
    "laud_adavit_attention": ([_vp, _i, _i32p, _u8p, _i, _i, _i, _vp, _vp], _i),
}

_lib: Optional[C.CDLL] = None


class LaudError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load the shared library (once).  Fails loudly when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LaudError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built. "
                "Run `python -m laudnet_b200.build` - there is no CPU/PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (argtypes, restype) in SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError if the symbol is not exported
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().laud_last_error().decode(errors="replace")
        raise LaudError(f"{what or 'laud call'} failed (code {rc}): {msg}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    """cudaStream_t of torch's current stream, so our launches order with torch's."""
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise LaudError(f"{what}: expected a CUDA tensor - the LAUD operators have no CPU path")


def launch_count() -> int:
    return int(lib().laud_launch_count())


def conv_path_counts() -> dict:
    """Launches of the mask-conditioned conv by implementation (evidence that the
    tcgen05 kernel is the one that runs)."""
    out = (C.c_ulonglong * 3)()
    lib().laud_conv_path_counts(C.byref(out))
    return {"umma_tcgen05": int(out[0]), "umma_tcgen05_tma": int(lib().laud_conv_tma_launch_count()),
            "hmma_legacy": int(out[1]), "naive_selftest": int(out[2])}
