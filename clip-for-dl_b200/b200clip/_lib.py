"""ctypes binding of libb200clip.so (the C ABI declared in include/b200clip.h).

There is no fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200CLIP_LIB_PATH") or os.path.join(_HERE, "_lib", "libb200clip.so")   # env: experiment builds

vp, ll, i32, f32, f64, sz = C.c_void_p, C.c_longlong, C.c_int, C.c_float, C.c_double, C.c_size_t

# name -> (restype, argtypes); mirrors include/b200clip.h one to one
SIGNATURES = {
    "b200clip_version": (i32, []),
    "b200clip_last_error_string": (C.c_char_p, []),
    "b200clip_launch_count": (C.c_ulonglong, []),
    "b200clip_bce_heads_mma_workspace_bytes": (sz, [ll]),
    "b200clip_bce_heads_mma_fwd": (i32, [vp, vp, ll, i32, vp, i32, vp, vp, i32, vp, i32, ll, f32, vp, f64, f64, vp, vp, vp, vp, vp, sz, vp]),
    "b200clip_skinny_outer_mma": (i32, [vp, vp, ll, i32, i32, vp, vp, vp, vp, vp, sz, vp]),
    "b200clip_cast_f32_bf16_2d": (i32, [vp, ll, vp, ll, ll, i32, vp]),
    "b200clip_fusion_fwd": (i32, [vp, ll, i32, vp, vp, vp, vp, f32, C.c_uint, vp, vp, vp]),
    "b200clip_fusion_bwd_workspace_bytes": (sz, [ll, i32]),
    "b200clip_fusion_bwd": (i32, [vp, vp, ll, i32, vp, vp, vp, f32, vp, vp, vp, vp, vp, vp, sz, vp]),
    "b200clip_asl_workspace_bytes": (sz, []),
    "b200clip_asl_fwd_bwd": (i32, [vp, vp, ll, f32, f32, f32, f32, i32, vp, vp, vp, vp, vp, vp, vp, sz, vp]),
    "b200clip_attention_fwd": (i32, [vp, vp, ll, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "b200clip_attention_bwd_workspace_bytes": (sz, [ll, i32, i32]),
    "b200clip_attention_bwd": (i32, [vp, vp, vp, vp, ll, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]),
    "b200clip_softclip_workspace_bytes": (sz, [ll]),
    "b200clip_softclip_logits": (i32, [vp, vp, ll, i32, f32, vp, vp]),
    "b200clip_softclip_fwd_bwd": (i32, [vp, vp, ll, i32, f32, vp, vp, vp, vp, vp, sz, vp]),
    "b200clip_infonce_general_fwd_bwd": (i32, [vp, vp, ll, i32, f32, vp, vp, vp, vp, vp, sz, vp]),
    "b200clip_infonce_shift": (f64, [f32]),
    "b200clip_rows_unit_check": (i32, [vp, ll, i32, f32, vp, vp]),
    "b200clip_adamw_tick": (i32, [vp, vp]),
    "b200clip_adamw_step": (i32, [vp, vp, vp, vp, ll, f32, f32, f32, f32, f32, vp, vp]),
    "b200clip_zs_thresholds_workspace_bytes": (sz, [ll, i32]),
    "b200clip_zs_dynamic_thresholds": (i32, [vp, vp, ll, i32, vp, vp, vp, sz, vp]),
    "b200clip_zs_merge_views": (i32, [vp, vp, ll, i32, f64, f64, vp, vp, vp]),
    "b200clip_head_loss_finalize": (i32, [vp, vp, f32, f64, f64, f64, vp, vp, vp, vp]),
    "b200clip_debug_set_nce_prof": (None, [vp]),
    "b200clip_gemm_bf16": (i32, [vp, vp, i32, i32, i32, i32, i32, ll, ll, i32, f32, vp, ll, vp, ll, vp, vp, ll, vp, ll, i32, vp]),
    "b200clip_l2norm_fwd": (i32, [vp, i32, ll, vp, vp, vp, ll, i32, f32, vp]),
    "b200clip_l2norm_bwd": (i32, [vp, i32, vp, i32, ll, vp, vp, i32, ll, i32, f32, vp, vp, vp]),
    "b200clip_layernorm_fwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, ll, i32, f32, f32, vp]),
    "b200clip_layernorm_bwd_workspace_bytes": (sz, [ll, i32]),
    "b200clip_layernorm_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, ll, i32, f32, C.c_uint, vp, vp, sz, vp]),
    "b200clip_dropout_seed_advance": (i32, [vp, vp]),
    "b200clip_colsum_workspace_bytes": (sz, [ll, i32]),
    "b200clip_colsum": (i32, [vp, i32, ll, ll, i32, vp, i32, vp, sz, vp]),
    "b200clip_cast_f32_bf16": (i32, [vp, vp, ll, vp]),
    "b200clip_dropout_mask": (i32, [vp, ll, i32, f32, C.c_uint, vp]),
    "b200clip_sum_f32": (i32, [vp, ll, vp, vp]),
    "b200clip_proj_fwd": (i32, [vp, ll, i32, i32, vp, vp, vp, vp, vp, vp, f32, f32, C.c_uint, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "b200clip_proj_bwd_workspace_bytes": (sz, [ll, i32, i32]),
    "b200clip_proj_bwd": (i32, [vp, vp, i32, vp, vp, vp, vp, vp, ll, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, f32, C.c_uint, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]),
    "b200clip_layernorm_l2_bwd": (i32, [vp, i32, vp, vp, f32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, ll, i32, f32, C.c_uint, vp, vp, sz, vp]),
    "b200clip_infonce_workspace_bytes": (sz, [ll, ll]),
    "b200clip_infonce_fwd_stats": (i32, [vp, vp, i32, ll, ll, f32, vp, vp, vp, sz, vp]),
    "b200clip_infonce_loss": (i32, [vp, vp, i32, ll, ll, ll, f32, vp, vp, ll, ll, vp, vp, vp, vp, vp, sz, vp]),
    "b200clip_infonce_inv_stats": (i32, [vp, ll, vp, ll, vp, vp, vp]),
    "b200clip_infonce_bwd_splits": (i32, [ll, ll]),
    "b200clip_infonce_bwd": (i32, [vp, vp, i32, ll, ll, ll, f32, vp, vp, vp, vp, i32, vp, i32, vp]),
    "b200clip_smallc_workspace_bytes": (sz, [ll, i32, i32]),
    "b200clip_mlbce_fwd_bwd": (i32, [vp, ll, vp, vp, i32, ll, ll, i32, i32, f32, vp, f64, vp, vp, i32, vp, vp, vp, vp, vp, vp, sz, vp]),
    "b200clip_fc_bce_fwd_bwd": (i32, [vp, ll, vp, vp, vp, ll, ll, i32, i32, f64, f32, vp, vp, i32, vp, vp, vp, vp, vp, vp, sz, vp]),
    "b200clip_bce_heads_fwd_bwd": (i32, [vp, ll, vp, i32, vp, vp, i32, vp, i32, ll, ll, i32, f32, vp, f64, f64, vp, vp, i32, vp, vp, vp, vp, vp, vp, sz, vp]),
    "b200clip_skinny_outer": (i32, [vp, i32, vp, ll, vp, ll, i32, vp, vp, i32, vp, vp, sz, vp]),
    "b200clip_predict_multilabel": (i32, [vp, ll, vp, ll, i32, i32, f32, f32, vp, vp]),
    "b200clip_multilabel_metrics_workspace_bytes": (sz, [ll]),
    "b200clip_multilabel_metrics": (i32, [vp, ll, vp, ll, ll, i32, f32, vp, vp, sz, vp]),
    "b200clip_prompt_mean_pool": (i32, [vp, vp, i32, i32, f32, i32, vp, vp]),
    "b200clip_zeroshot_workspace_bytes": (sz, [ll]),
    "b200clip_zeroshot_score": (i32, [vp, ll, ll, vp, i32, i32, i32, i32, f32, C.POINTER(f32), i32, f32, i32, i32, vp, vp, i32, vp, vp, vp, vp, vp, sz, vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (building is __graft_entry__.build()'s job; a missing .so is a hard error)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"b200clip: CUDA library not built ({LIB_PATH}); run `python clip-for-dl_b200/build.py`. "
                "There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().b200clip_last_error_string()
        raise RuntimeError(f"b200clip.{what} failed (code {rc}): {msg.decode() if msg else '?'}")


def require_cuda(*tensors) -> None:
    """Every op launches on the CURRENT device's current stream (stream_ptr) and the C side keeps per-device caches keyed by
    cudaGetDevice: tensors on another GPU would be read through the wrong context.  One process per GPU is the supported
    layout (torchrun sets the device once); anything else is rejected here instead of faulting in a kernel."""
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("b200clip: tensors must live on a CUDA device (B200); there is no CPU path")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise RuntimeError(f"b200clip: tensor on cuda:{t.device.index} but the current device is cuda:{cur}; "
                               "call torch.cuda.set_device() (one process per GPU) or wrap the call in torch.cuda.device(...)")
