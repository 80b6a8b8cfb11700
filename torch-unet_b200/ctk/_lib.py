"""ctypes binding of libctk.so (the C ABI declared in include/ctk.h).

There is deliberately no fallback: if the shared library is missing or a call fails, the caller gets an
exception -- never a silent PyTorch/CPU path.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_longlong, c_size_t, c_ulonglong, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# CTK_LIB: another build of the same C ABI (tools/build_alt_lib.py), for same-box A/B measurements of a kernel change
LIB_PATH = os.environ.get("CTK_LIB") or os.path.join(_HERE, "libctk.so")

CONV_NO_POOL = 1
CONV_NO_ACT = 2
CONV_SINGLE_CTA = 4
ADAM_CHUNK = 65536


class CtkError(RuntimeError):
    pass


# name -> (restype, argtypes); mirrors include/ctk.h one to one
_SIGNATURES = {
    "ctk_abi_version": (c_int, []),
    "ctk_status_string": (c_char_p, [c_int]),
    "ctk_last_cuda_error": (c_int, []),
    "ctk_device_check": (c_int, []),
    "ctk_set_persistent_sm_reserve": (c_int, [c_int]),
    "ctk_pearson_workspace_bytes": (c_size_t, [c_int]),
    "ctk_pearson_f32": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ctk_prepare_tiles": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ctk_tile_nmi_workspace_bytes": (c_size_t, [c_int]),
    "ctk_tile_nmi_f32": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ctk_tile_ssim_workspace_bytes": (c_size_t, [c_int]),
    "ctk_tile_ssim_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ctk_tile_metrics_workspace_bytes": (c_size_t, [c_int]),
    "ctk_tile_metrics_f32": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                     c_void_p]),
    "ctk_fold_bn_eval": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_void_p, c_void_p,
                                 c_void_p]),
    "ctk_pack_conv_weight_bf16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "ctk_pack_first_weight": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "ctk_pack_fc1_weight_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ctk_conv_first_eval": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                    c_float, c_void_p, c_int, c_int, c_void_p]),
    "ctk_pack_conv_weight_split_bf16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "ctk_pack_fc1_weight_split_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "ctk_conv_first_eval_split": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                          c_float, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "ctk_conv3x3_tc_eval_split": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                          c_float, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "ctk_conv_first_pool_codes": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                          c_float, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "ctk_conv3x3_tc_eval": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_float,
                                    c_void_p, c_int, c_int, c_int, c_void_p]),
    "ctk_gemm_bf16_splitk": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ctk_head_eval": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_void_p, c_float, c_int, c_void_p, c_void_p]),
    "ctk_mse_loss": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "ctk_scale_by_scalar": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "ctk_adam_multi": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float,
                               c_float, c_float, c_float, c_float, c_int, c_float, c_void_p]),
    # ---- training path
    "ctk_conv_first_raw_workspace_bytes": (c_size_t, [c_int]),
    "ctk_conv_first_raw": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                                   c_void_p, c_void_p, c_size_t, c_void_p]),
    "ctk_conv3x3_tc_raw_workspace_bytes": (c_size_t, [c_int]),
    "ctk_conv3x3_tc_raw": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                   c_size_t, c_void_p]),
    "ctk_pack_conv_weight_dgrad_bf16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "ctk_pack_conv_weights_train": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ctk_bn_finalize": (c_int, [c_void_p, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                c_float, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ctk_bn_finalize_moments": (c_int, [c_void_p, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_float, c_float, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ctk_first_patch_gram_workspace_bytes": (c_size_t, [c_int]),
    "ctk_first_patch_gram": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t,
                                     c_void_p]),
    "ctk_first_moments": (c_int, [c_void_p, c_void_p, c_int, c_int, c_double, c_void_p, c_void_p]),
    "ctk_first_wgrad_codes_workspace_bytes": (c_size_t, [c_int, c_int]),
    "ctk_first_wgrad_codes": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                      c_float, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ctk_first_wgrad_finalize": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double,
                                         c_int, c_int, c_void_p, c_void_p]),
    "ctk_bn_act_pool_fwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_int,
                                    c_int, c_void_p]),
    "ctk_bn_bwd_reduce_workspace_bytes": (c_size_t, [c_int]),
    "ctk_bn_bwd_reduce": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ctk_bn_bwd_reduce_pooled": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_longlong, c_int, c_void_p,
                                         c_void_p, c_float, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ctk_bn_bwd_reduce_guarded": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_float,
                                          c_void_p, c_void_p, c_size_t, c_void_p]),
    "ctk_bn_bwd_apply": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p]),
    "ctk_conv3x3_wgrad_tc_workspace_bytes": (c_size_t, [c_int, c_int]),
    "ctk_conv3x3_wgrad_tc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t,
                                     c_void_p]),
    "ctk_conv_first_wgrad_workspace_bytes": (c_size_t, [c_int, c_int]),
    "ctk_conv_first_wgrad": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                     c_void_p, c_size_t, c_void_p]),
    "ctk_feat_transpose_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "ctk_pack_fc1_weight_t_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ctk_gemm_bf16_out_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ctk_gemm_bf16_bt_out_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ctk_colstat": (c_int, [c_void_p, c_int, c_longlong, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "ctk_bn1d_act_drop_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_int, c_int, c_void_p,
                                      c_void_p]),
    "ctk_dropout_masks": (c_int, [c_void_p, c_longlong, c_float, c_void_p, c_longlong, c_float, c_ulonglong, c_ulonglong,
                                  c_void_p]),
    "ctk_sgemm_strided": (c_int, [c_void_p, c_longlong, c_longlong, c_void_p, c_longlong, c_longlong, c_void_p, c_int,
                                  c_int, c_int, c_void_p, c_int, c_void_p]),
    "ctk_head_out_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ctk_head_out_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                 c_void_p, c_void_p]),
    "ctk_bn1d_bwd_reduce": (c_int, [c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_float, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "ctk_bn1d_bwd_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                   c_void_p, c_void_p, c_int, c_void_p]),
    # ---- fp32 training path (CUDA cores, the reference's own arithmetic)
    "ctk_pack_conv_weight_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ctk_conv3x3_f32": (c_int, [c_void_p, c_longlong, c_longlong, c_longlong, c_longlong, c_int, c_int, c_int, c_int,
                                c_void_p, c_int, c_void_p, c_void_p]),
    "ctk_conv3x3_wgrad_f32_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "ctk_conv3x3_wgrad_f32": (c_int, [c_void_p, c_void_p, c_longlong, c_longlong, c_longlong, c_longlong, c_int, c_int,
                                      c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ctk_channel_sums_f64_workspace_bytes": (c_size_t, [c_int]),
    "ctk_channel_stats_f32": (c_int, [c_void_p, c_longlong, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ctk_bn_finalize_f64": (c_int, [c_void_p, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_float, c_float, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ctk_bn_act_pool_fwd_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_float, c_void_p, c_longlong, c_longlong, c_longlong, c_void_p]),
    "ctk_bn_bwd_reduce_f32": (c_int, [c_void_p, c_void_p, c_longlong, c_longlong, c_longlong, c_int, c_int, c_int, c_int,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p,
                                      c_size_t, c_void_p]),
    "ctk_bn_bwd_apply_f32": (c_int, [c_void_p, c_void_p, c_longlong, c_longlong, c_longlong, c_int, c_int, c_int, c_int,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_float, c_void_p,
                                     c_void_p]),
    "ctk_gemm_f32": (c_int, [c_void_p, c_longlong, c_longlong, c_void_p, c_longlong, c_longlong, c_void_p, c_int, c_int,
                             c_int, c_void_p, c_longlong, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES.keys())

_lib = None


def load() -> ctypes.CDLL:
    """Load libctk.so (building nothing: run ``python torch-unet_b200/build.py`` or ``__graft_entry__.build()`` first)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CtkError(f"{LIB_PATH} is missing: build it with `python torch-unet_b200/build.py` "
                           "(there is no CPU or PyTorch fallback for the CUDA hot path)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        lib = load()
        msg = lib.ctk_status_string(status).decode()
        raise CtkError(f"{what}: {msg} (status {status}, cuda error {lib.ctk_last_cuda_error()})")


def ptr(t) -> c_void_p:
    return c_void_p(0) if t is None else c_void_p(t.data_ptr())


def stream() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def workspace(query: str, *args, device=None):
    """(tensor, pointer, byte count) of a caller-owned workspace sized by the entry point's ``*_workspace_bytes`` query.
    Allocated through torch's caching allocator on the current stream, so it is recycled like any other temporary."""
    nbytes = int(getattr(load(), query)(*args))
    t = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=device if device is not None else torch.cuda.current_device())
    return t, c_void_p(t.data_ptr()), c_size_t(nbytes)


# kernels launched per successful call (bench.py reports the sum as "gpu_launches")
KERNELS_PER_CALL = {"ctk_pearson_f32": 2, "ctk_tile_metrics_f32": 4, "ctk_tile_nmi_f32": 4, "ctk_tile_ssim_f32": 3, "ctk_prepare_tiles": 1, "ctk_fold_bn_eval": 1, "ctk_pack_conv_weight_bf16": 1, "ctk_pack_first_weight": 1,
                    "ctk_pack_fc1_weight_bf16": 1, "ctk_conv_first_eval": 1, "ctk_conv_first_pool_codes": 1, "ctk_pack_conv_weight_split_bf16": 1,
                    "ctk_pack_fc1_weight_split_bf16": 1, "ctk_conv_first_eval_split": 1, "ctk_conv3x3_tc_eval_split": 1, "ctk_conv3x3_tc_eval": 1,
                    "ctk_gemm_bf16_splitk": 1, "ctk_head_eval": 1, "ctk_mse_loss": 1, "ctk_scale_by_scalar": 1, "ctk_adam_multi": 1,
                    "ctk_conv_first_raw": 2, "ctk_conv3x3_tc_raw": 1, "ctk_pack_conv_weight_dgrad_bf16": 1, "ctk_pack_conv_weights_train": 1,
                    "ctk_bn_finalize": 1, "ctk_bn_finalize_moments": 1, "ctk_first_patch_gram": 2,
                    "ctk_first_moments": 1, "ctk_first_wgrad_codes": 3, "ctk_first_wgrad_finalize": 1, "ctk_bn_act_pool_fwd": 1, "ctk_bn_bwd_reduce": 2, "ctk_bn_bwd_reduce_pooled": 2, "ctk_bn_bwd_reduce_guarded": 3, "ctk_bn_bwd_apply": 1,
                    "ctk_conv3x3_wgrad_tc": 2, "ctk_conv_first_wgrad": 2, "ctk_feat_transpose_bf16": 1,
                    "ctk_pack_fc1_weight_t_bf16": 1, "ctk_gemm_bf16_out_bf16": 1, "ctk_gemm_bf16_bt_out_bf16": 1, "ctk_colstat": 1,
                    "ctk_bn1d_act_drop_fwd": 1, "ctk_dropout_masks": 1, "ctk_sgemm_strided": 1, "ctk_head_out_fwd": 1, "ctk_head_out_bwd": 1,
                    "ctk_bn1d_bwd_reduce": 1, "ctk_bn1d_bwd_apply": 1,
                    "ctk_pack_conv_weight_f32": 1, "ctk_conv3x3_f32": 1, "ctk_conv3x3_wgrad_f32": 2, "ctk_channel_stats_f32": 2,
                    "ctk_bn_finalize_f64": 1, "ctk_bn_act_pool_fwd_f32": 1, "ctk_bn_bwd_reduce_f32": 3,
                    "ctk_bn_bwd_apply_f32": 1, "ctk_gemm_f32": 1}
launch_count = 0
_timeline = None      # when a list: (name, start_event, end_event, meta) per call, for per-kernel timing in bench.py
_timeline_only = None # optional set of entry points to instrument (events around every call cost ~2 us each)


def call(name: str, *args, meta=None) -> None:
    global launch_count
    fn = getattr(load(), name)
    if _timeline is not None and name in KERNELS_PER_CALL and (_timeline_only is None or name in _timeline_only):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        check(fn(*args), name)
        e1.record()
        _timeline.append((name, e0, e1, meta))
    else:
        check(fn(*args), name)
    launch_count += (meta or {}).get("kernels", KERNELS_PER_CALL.get(name, 0))


def start_timeline(only=None) -> list:
    global _timeline, _timeline_only
    _timeline = []
    _timeline_only = set(only) if only is not None else None
    return _timeline


def stop_timeline() -> list:
    global _timeline
    t, _timeline = _timeline, None
    return t or []


def require_device(t: torch.Tensor, dtype=None, what: str = "tensor") -> None:
    if not t.is_cuda:
        raise CtkError(f"{what} must live on a CUDA device (the ctk hot path has no CPU fallback)")
    if not t.is_contiguous():
        raise CtkError(f"{what} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise CtkError(f"{what} must be {dtype}, got {t.dtype}")
