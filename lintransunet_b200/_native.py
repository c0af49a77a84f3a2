"""ctypes binding of libltu_b200.so (C ABI declared in include/ltu_b200.h).

There is NO fallback: if the shared library is missing or a tensor is not on a CUDA
device every entry point raises.  The library is built in-tree by
``python -m lintransunet_b200.build`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_uint, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libltu_b200.so")

_lib = None
_lock = threading.Lock()

P, I, L, F, Z, U64 = c_void_p, c_int, c_int64, c_float, c_size_t, c_uint64

# name -> (restype, argtypes); mirrors include/ltu_b200.h one to one
SIGNATURES = {
    "ltu_last_error": (c_char_p, []),
    "ltu_version": (I, []),
    "ltu_launch_count": (L, []),
    "ltu_kv_reduce_workspace": (Z, [I, L, I]),
    "ltu_kv_reduce": (I, [P, P, L, P, P, Z, I, L, I, I, P]),
    "ltu_kv_reduce_project": (I, [P, P, L, P, P, Z, I, L, I, P, P, P]),
    "ltu_q_readout": (I, [P, L, P, P, L, I, L, I, I, P]),
    "ltu_add_layernorm": (I, [P, P, P, P, P, L, I, F, I, P]),
    "ltu_add_layernorm_split": (I, [P, P, P, P, P, P, P, L, I, F, P]),
    "ltu_gelu": (I, [P, L, I, P]),
    "ltu_loss_sums_workspace": (Z, [I, I, L]),
    "ltu_loss_sums": (I, [P, P, P, P, Z, I, I, L, P]),
    "ltu_loss_sums_bwd": (I, [P, P, P, P, I, I, L, P]),
    "ltu_label_pool": (I, [P, P, I, I, I, I, I, I, I, P]),
    "ltu_dropout": (I, [P, P, L, I, L, F, U64, U64, I, I, P]),
    "ltu_concat2": (I, [P, I, P, I, P, L, I, P]),
    "ltu_posenc_dwconv3": (I, [P, P, P, P, I, I, I, I, I, I, P]),
    "ltu_posenc_dwconv3_split": (I, [P, P, P, P, P, P, I, I, I, I, I, P]),
    "ltu_conv3d_tiles": (I, [L, I]),
    "ltu_conv3d": (I, [P, I, P, I, I, I, I, I, I, I, I, I, I, I, P, P, I, P, I, I, I, I, P, I, P]),
    "ltu_conv3d_tc_supported": (I, [I, I, I, I, I]),
    "ltu_conv3d_tc_tiles": (I, [L, I]),
    "ltu_conv3d_tc_kpad": (I, [I, I]),
    "ltu_conv3d_tc": (I, [P, I, P, I, I, I, I, I, I, I, I, I, I, I, P, P, I, P, I, I, I, I, P, I, P, P]),
    "ltu_conv3d_tc3_supported": (I, [I, I, I, I, I, I, I, I, I, I, I]),
    "ltu_conv3d_tc3_tiles": (I, [I, I, I, I, I, I, I]),
    "ltu_conv3d_tc3": (I, [P, I, P, I, I, I, I, I, I, P, I, I, P, I, P, P, I, P, P]),
    "ltu_conv3d_tc3_masked": (I, [P, I, P, I, I, I, I, I, P, I, I, P, I, P, P, I, P, P, P]),
    "ltu_linear_tc": (I, [P, I, L, P, P, I, P, I, I, P, P, P, F, P]),
    "ltu_linear_fused": (I, [P, L, I, P, P, I, I, P, P, P, P, F, P, P, P]),
    "ltu_linear_fused_ex": (I, [P, L, L, I, P, P, I, I, P, P, P, P, F, P, P, I, I, P]),
    "ltu_ctx_project": (I, [P, P, P, I, I, P]),
    "ltu_kv_project_reduce_supported": (I, [I, I, L]),
    "ltu_kv_project_reduce_workspace": (Z, [I, L]),
    "ltu_kv_project_reduce": (I, [P, P, P, P, P, Z, I, L, P, P, P]),
    "ltu_ffn_fused_supported": (I, [I]),
    "ltu_ffn_fused": (I, [P, L, I, P, P, P, P, P, P, F, P, P]),
    "ltu_ffn_fused_trace": (I, [P, L, I, P, P, P, P, P, P, F, P, P, P]),
    "ltu_attn_out_fused_supported": (I, [I, I]),
    "ltu_ctx_pack_bf16": (I, [P, P, I, I, P]),
    "ltu_attn_out_fused": (I, [P, I, L, I, I, P, P, P, P, P, P, P, F, P, P]),
    "ltu_attn_out_fused_w": (I, [P, I, L, I, I, P, P, P, P, P, P, F, P, P]),
    "ltu_conv3d_halo_supported": (I, [I, I, I, I, I, I, I, I, I]),
    "ltu_conv3d_halo_tiles": (I, [I, I, I, I]),
    "ltu_conv3d_halo": (I, [P, I, P, I, I, I, I, I, I, P, I, P, I, P, I, P, I, P, P]),
    "ltu_instnorm_finalize": (I, [P, P, I, I, I, L, F, P]),
    "ltu_chan_partials": (I, [P, P, I, L, I, I, I, P]),
    "ltu_instnorm_apply": (I, [P, P, P, P, I, L, I, I, I, P]),
    "ltu_s2d_input": (I, [P, P, I, I, I, I, I, I, P]),
    "ltu_upsample_trilinear": (I, [P, P, I, I, I, I, I, I, I, P]),
    "ltu_mask_softmax": (I, [P, P, P, I, L, I, P]),
    "ltu_gate_fused": (I, [P, P, P, P, P, P, P, P, I, L, I, I, P]),
    "ltu_roi_bbox_scratch": (Z, [I, I, I]),
    "ltu_roi_bbox": (I, [P, P, P, Z, I, I, I, I, I, I, F, P]),
    "ltu_roi_resample": (I, [P, P, P, I, I, I, I, I, I, I, I, I, I, I, P]),
    "ltu_head_d2s_softmax": (I, [P, P, P, P, I, I, I, I, I, P]),
    "ltu_vote_accumulate": (I, [P, P, P, I, I, I, I, I, I, I, I, P]),
    "ltu_vote_argmax": (I, [P, P, I, L, P]),
    "ltu_vote_fractions": (I, [P, P, I, L, P]),
    "ltu_gather_windows": (I, [P, P, P, I, I, I, I, I, I, I, P]),
    "ltu_vote_decide": (I, [P, P, I, L, I, F, P]),
    "ltu_keep_largest_component_workspace": (Z, [L]),
    "ltu_keep_largest_component": (I, [P, I, c_uint, I, I, I, I, P, Z, P]),
    "ltu_overlap_counts": (I, [P, P, I, I, I, I, P, P]),
    "ltu_add_layernorm_bwd_workspace": (Z, [L, I]),
    "ltu_add_layernorm_bwd": (I, [P, P, P, P, P, P, P, P, Z, L, I, F, I, P]),
    "ltu_gelu_bwd": (I, [P, P, P, L, I, P]),
    "ltu_posenc_wgrad_workspace": (Z, [I, I, I, I, I]),
    "ltu_posenc_wgrad": (I, [P, P, P, P, P, Z, I, I, I, I, I, I, P]),
    "ltu_instnorm_bwd_workspace": (Z, [I, L, I]),
    "ltu_instnorm_bwd": (I, [P, P, P, P, P, Z, I, L, I, I, I, P]),
    "ltu_conv3d_wgrad_workspace": (Z, [I, I, I, I, I, I, I]),
    "ltu_conv3d_wgrad": (I, [P, P, P, P, Z, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, P]),
    "ltu_zero_insert": (I, [P, P, I, I, I, I, I, I, I, I, I, I, I, I, P]),
    "ltu_sumpool2": (I, [P, P, I, I, I, I, I, I, P]),
    "ltu_upsample_trilinear_bwd": (I, [P, P, I, I, I, I, I, I, I, P]),
    "ltu_mask_softmax_bwd": (I, [P, P, P, I, L, I, P]),
    "ltu_head_d2s_softmax_bwd": (I, [P, P, P, I, I, I, I, I, P]),
    "ltu_gate_bwd_workspace": (Z, [I, L, I]),
    "ltu_gate_bwd": (I, [P, P, P, P, P, P, P, P, P, P, P, P, P, Z, I, L, I, I, P]),
    "ltu_roi_resample_bwd_workspace": (Z, [I, I, I, I, I]),
    "ltu_roi_resample_bwd": (I, [P, P, P, P, Z, I, I, I, I, I, I, I, I, I, I, I, P]),
    "ltu_attn_bwd_workspace": (Z, [I, L, I]),
    "ltu_attn_bwd": (I, [P, L, P, P, L, P, L, P, P, P, P, L, P, P, P, Z, I, L, I, I, P]),
}


class NativeLibraryMissing(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle; raises if the .so is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise NativeLibraryMissing(
                    f"{LIB_PATH} not found: build it with `python -m lintransunet_b200.build` "
                    "(there is no CPU / PyTorch fallback for the hot path)")
            h = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(h, name)          # AttributeError if the export is missing
                fn.restype = res
                fn.argtypes = args
            _lib = h
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().ltu_last_error()
        raise RuntimeError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")


_replayed = 0


def note_replayed(n: int) -> None:
    """Account for native kernels executed through CUDA-graph replays (the C counter only sees
    direct launches and capture-time recording)."""
    global _replayed
    _replayed += n


def launch_count() -> int:
    """Native (libltu_b200) kernels executed by this process: direct launches + graph replays."""
    return int(lib().ltu_launch_count()) + _replayed
