"""ctypes binding of libvnpcc.so (include/vnpcc.h).  There is NO fallback: if the library is missing or a call
returns non-zero, an exception is raised (the reference ignores its kernels' error returns,
extensions/chamfer_distance/chamfer_distance.py:52,68)."""
from __future__ import annotations

import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libvnpcc.so")

_p, _i, _ll, _f, _d, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double, C.c_size_t

# name -> (restype, argtypes); must list every symbol include/vnpcc.h and include/vnpcc_debug.h declare (tests check this)
SIGNATURES = {
    "vnpcc_abi_version": (_i, []),
    "vnpcc_launch_count": (C.c_ulonglong, []),
    "vnpcc_set_fast_math": (None, [_i]),
    "vnpcc_set_tuning": (None, [_i, _i]),
    "vnpcc_debug_chamfer_plan": (None, [_i, _i, _i, _p]),
    "vnpcc_debug_fold_geometry": (None, [_i, _i, _i, _i, _i, _p]),
    "vnpcc_debug_wgrad_plan": (None, [_ll, _i, _i, _i, _p]),
    "vnpcc_debug_plan_chunk_len": (_i, [_ll, _i, _ll, _i, _i]),
    "vnpcc_debug_rows_plan": (None, [_ll, _i, _i, _i, _i, _i, _p]),
    "vnpcc_chamfer_workspace_bytes": (_sz, [_i, _i, _i]),
    "vnpcc_chamfer_forward": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _sz, _p]),
    "vnpcc_chamfer_backward": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p]),
    "vnpcc_chamfer_set_packed_math": (None, [_i]),
    "vnpcc_debug_chamfer_slow_counts": (_i, [_p, _i, _i, _i, _p, _p]),
    "vnpcc_cd_reduce": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "vnpcc_cd_reduce_bwd": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "vnpcc_cd_persample_fwd": (_i, [_p, _p, _i, _i, _i, _p, _p, _p]),
    "vnpcc_cd_persample_bwd": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p]),
    "vnpcc_fscore_sq": (_i, [_p, _p, _i, _i, _i, _f, _p, _p]),
    "vnpcc_nn_counts": (_i, [_p, _i, _i, _i, _p, _p]),
    "vnpcc_dcd_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _f, _f, _f, _p, _p, _p]),
    "vnpcc_dcd_bwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _f, _f, _f, _p, _p, _p, _p]),
    "vnpcc_gemm_rows_fp32": (_i, [_p, _ll, _p, _ll, _i, _p, _ll, _ll, _i, _i, _p, _ll, _ll, _i, _p]),
    "vnpcc_gemm_wgrad_fp32": (_i, [_p, _ll, _p, _ll, _p, _ll, _ll, _i, _i, _i, _p]),
    "vnpcc_transpose": (_i, [_p, _ll, _p, _ll, _i, _i, _p]),
    "vnpcc_gemm_rows_tf32": (_i, [_p, _ll, _p, _ll, _p, _ll, _ll, _i, _i, _p, _ll, _ll, _p]),
    "vnpcc_gemm_rows_tf32_stats": (_i, [_p, _ll, _p, _ll, _p, _ll, _ll, _i, _i, _p, _ll, _ll, _p, _i, _p]),
    "vnpcc_split_tf32": (_i, [_p, _ll, _ll, _i, _p, _ll, _i, _p]),
    "vnpcc_gemm_wgrad_tf32": (_i, [_p, _ll, _p, _ll, _p, _ll, _ll, _i, _i, _p, _sz, _p]),
    "vnpcc_gemm_wgrad_tf32_workspace_bytes": (_sz, [_ll, _i, _i]),
    "vnpcc_gemm_vn_stats": (_i, [_p, _ll, _p, _ll, _ll, _i, _i, _p, _ll, _ll, _p, _p]),
    "vnpcc_gemm_vn_apply": (_i, [_p, _ll, _p, _ll, _p, _ll, _ll, _i, _i, _p, _ll, _ll, _p, _p, _p, _f, _p]),
    "vnpcc_gemm_vn_pool": (_i, [_p, _ll, _p, _ll, _ll, _i, _i, _ll, _p, _p]),
    "vnpcc_vn_maxpool_decode": (_i, [_p, _ll, _p, _p]),
    "vnpcc_pool_linear_gather": (_i, [_p, _ll, _p, _ll, _p, _i, _i, _i, _i, _p, _ll, _p]),
    "vnpcc_vn_norm_stats": (_i, [_p, _ll, _ll, _i, _p, _p]),
    "vnpcc_bn_finalize": (_i, [_p, _d, _i, _i, _p, _p, _f, _f, _p, _p]),
    "vnpcc_vn_bn_leaky_fwd": (_i, [_p, _ll, _p, _ll, _p, _ll, _ll, _i, _p, _p, _p, _f, _p]),
    "vnpcc_vn_bn_leaky_bwd1": (_i, [_p, _ll, _p, _ll, _p, _ll, _p, _ll, _p, _ll, _ll, _i, _p, _p, _p, _f, _p, _p]),
    "vnpcc_vn_bn_bwd2": (_i, [_p, _ll, _p, _ll, _ll, _i, _p, _p, _p, _p, _d, _i, _p, _p, _p]),
    "vnpcc_vn_bn_bwd2_sbias": (_i, [_p, _ll, _p, _ll, _ll, _i, _p, _p, _p, _p, _d, _i, _p, _p, _p, _ll, _p, _ll, _ll, _p]),
    "vnpcc_vn_maxpool_argmax": (_i, [_p, _ll, _p, _ll, _i, _i, _i, _p, _p, _p]),
    "vnpcc_vn_maxpool_gather": (_i, [_p, _ll, _p, _i, _i, _i, _p, _ll, _p]),
    "vnpcc_vn_maxpool_scatter_add": (_i, [_p, _ll, _p, _i, _i, _i, _p, _ll, _p]),
    "vnpcc_vn_frame_fwd": (_i, [_p, _ll, _p, _ll, _ll, _i, _i, _p, _ll, _p, _p]),
    "vnpcc_vn_frame_bwd": (_i, [_p, _ll, _p, _p, _ll, _p, _ll, _ll, _i, _i, _p, _ll, _p, _ll, _p]),
    "vnpcc_rows_add_sample_bias": (_i, [_p, _ll, _p, _ll, _i, _i, _i, _p]),
    "vnpcc_rows_sample_sum": (_i, [_p, _ll, _i, _i, _i, _p, _ll, _p]),
    "vnpcc_rows_dot": (_i, [_p, _ll, _p, _ll, _i, _p, _p, _p]),
    "vnpcc_rows_dot_bwd": (_i, [_p, _p, _ll, _p, _ll, _i, _p, _ll, _p, _p]),
    "vnpcc_bn_leaky_dot_fwd": (_i, [_p, _ll, _p, _ll, _ll, _i, _p, _p, _p, _f, _p, _p, _p, _p]),
    "vnpcc_bn_leaky_dot_bwd1": (_i, [_p, _p, _ll, _p, _ll, _p, _ll, _p, _ll, _ll, _i, _p, _p, _p, _f, _p, _p, _p, _p]),
    "vnpcc_tail_bwd_tf32": (_i, [_p, _p, _ll, _ll, _i, _p, _p, _p, _f, _p, _p, _ll, _i, _i, _p, _p, _p, _ll, _p, _ll, _p, _ll, _p, _ll, _p]),
    "vnpcc_double_to_float": (_i, [_p, _p, _i, _p]),
    "vnpcc_fold_stats": (_i, [_p, _ll, _p, _ll, _p, _ll, _i, _i, _i, _i, _p, _p]),
    "vnpcc_fold_fwd": (_i, [_p, _ll, _p, _ll, _p, _ll, _i, _i, _i, _i, _p, _p, _p, _f, _p, _ll, _p]),
    "vnpcc_fold_bwd": (_i, [_p, _ll, _p, _ll, _p, _ll, _p, _ll, _i, _i, _i, _i, _p, _p, _p, _f, _i, _p, _p, _ll, _i, _p, _ll, _p, _ll,
                            _p, _p, _p]),
    "vnpcc_pool_linear_bwd": (_i, [_p, _ll, _p, _p, _ll, _p, _ll, _i, _i, _i, _i, _p, _ll, _p, _ll, _p]),
    "vnpcc_smallk_fwd": (_i, [_p, _ll, _p, _ll, _p, _ll, _ll, _p, _ll, _ll, _i, _i, _p]),
    "vnpcc_smallk_dgrad": (_i, [_p, _ll, _p, _ll, _p, _ll, _ll, _i, _i, _p]),
    "vnpcc_smallk_wgrad": (_i, [_p, _ll, _p, _ll, _i, _i, _i, _i, _p, _ll, _p, _ll, _p]),
    "vnpcc_knn3d": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "vnpcc_fps": (_i, [_p, _i, _i, _i, _p, _p]),
    "vnpcc_points_gather": (_i, [_p, _ll, _p, _i, _i, _i, _i, _p, _ll, _p]),
    "vnpcc_points_scatter_add": (_i, [_p, _ll, _p, _i, _i, _i, _i, _p, _ll, _p]),
    "vnpcc_edge_feature_fwd": (_i, [_p, _ll, _p, _i, _i, _i, _i, _p, _ll, _p]),
    "vnpcc_edge_feature_bwd": (_i, [_p, _ll, _p, _i, _i, _i, _i, _p, _ll, _p]),
    "vnpcc_rows_group_mean": (_i, [_p, _ll, _ll, _i, _i, _p, _ll, _p]),
    "vnpcc_rows_group_mean_bwd": (_i, [_p, _ll, _ll, _i, _i, _p, _ll, _p]),
    "vnpcc_vn_layernorm_fwd": (_i, [_p, _ll, _ll, _i, _p, _p, _f, _p, _ll, _p, _p]),
    "vnpcc_vn_layernorm_bwd": (_i, [_p, _ll, _p, _ll, _ll, _i, _p, _p, _p, _p, _ll, _p, _p, _p]),
    "vnpcc_rows_add": (_i, [_p, _ll, _p, _ll, _p, _ll, _ll, _i, _p]),
    "vnpcc_vn_attention_fwd": (_i, [_p, _ll, _i, _i, _i, _i, _f, _p, _ll, _p, _p]),
    "vnpcc_vn_attention_fwd_tf32": (_i, [_p, _ll, _i, _i, _i, _i, _f, _p, _ll, _p, _p, _p]),
    "vnpcc_vn_attention_delta": (_i, [_p, _ll, _p, _ll, _i, _i, _i, _i, _p, _p]),
    "vnpcc_vn_attention_bwd_tf32": (_i, [_p, _ll, _p, _ll, _p, _ll, _p, _i, _i, _i, _i, _f, _p, _ll, _p, _p, _sz, _p, _p]),
    "vnpcc_vn_attention_bwd": (_i, [_p, _ll, _p, _ll, _p, _ll, _p, _i, _i, _i, _i, _f, _p, _ll, _p, _p]),
    "vnpcc_edge_conv_stats": (_i, [_p, _ll, _p, _i, _i, _i, _i, _p, _p]),
    "vnpcc_edge_conv_fwd": (_i, [_p, _ll, _p, _i, _i, _i, _i, _p, _p, _p, _f, _p, _ll, _p]),
    "vnpcc_edge_conv_bwd": (_i, [_p, _ll, _p, _ll, _p, _i, _i, _i, _i, _p, _p, _p, _f, _i, _p, _p, _ll, _p, _p, _p]),
    "vnpcc_fscore": (_i, [_p, _p, _i, _i, _i, _f, _p, _p]),
    "vnpcc_voxel_occupancy": (_i, [_p, _i, _i, _i, _p, _p]),
    "vnpcc_voxel_iou": (_i, [_p, _p, _i, _i, _p, _p]),
    "vnpcc_adam_step": (_i, [_p, _p, _p, _p, _ll, _f, _f, _f, _f, _f, _i, _f, _p]),
    "vnpcc_adam_step_dev": (_i, [_p, _p, _p, _p, _ll, _p, _f, _f, _f, _f, _f, _p]),
    "vnpcc_measure_fp32_peak": (_i, [_i, _i, _p, _p, _p, _p]),
}

_lib = None


class VnpccError(RuntimeError):
    pass


def load():
    """dlopen libvnpcc.so and bind every symbol; raises if the library or a symbol is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VnpccError(f"{LIB_PATH} is missing: build it with `python -m vn_pointcloudcompletion_b200.build` "
                         "(or __graft_entry__.build()).  There is no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class _StreamArg:
    """placeholder for the `void* stream` argument: resolved inside call()/raw() to the CURRENT stream OF THE DEVICE THE TENSOR
    ARGUMENTS LIVE ON (not of the current device)"""
    __slots__ = ()


_STREAM = _StreamArg()


def ptr(t):
    """device-pointer argument of a tensor (None -> NULL).  The tensor itself is handed on: call()/raw() take its data_ptr() and use its
    device to pick the stream and the device guard, and reject operands that live on different devices."""
    return t


def stream():
    return _STREAM


def _invoke(name, args):
    fn = getattr(load(), name)
    dev = None
    n = len(args)
    out = [None] * n
    spos = -1
    for i in range(n):
        a = args[i]
        if isinstance(a, torch.Tensor):
            if not a.is_cuda:
                raise VnpccError(f"{name}: CPU tensor passed to a B200 kernel (there is no CPU fallback)")
            if dev is None:
                dev = a.device
            elif a.device != dev:
                raise VnpccError(f"{name}: operands on different devices ({dev} and {a.device})")
            out[i] = a.data_ptr()
        elif a is _STREAM:
            spos = i
        else:
            out[i] = a
    if dev is None or dev.index == torch.cuda.current_device():
        if spos >= 0:
            out[spos] = torch.cuda.current_stream().cuda_stream
        return fn(*out)
    # the reference lets a model live on config.device without torch.cuda.set_device (models/model.py:14-20): launch on the
    # tensors' device and on ITS current stream
    with torch.cuda.device(dev):
        if spos >= 0:
            out[spos] = torch.cuda.current_stream(dev).cuda_stream
        return fn(*out)


def call(name, *args):
    """call an int-returning entry point and raise on a non-zero return"""
    rc = _invoke(name, args)
    if rc != 0:
        raise VnpccError(f"{name} failed with code {rc}")
    return rc


def raw(name, *args):
    return _invoke(name, args)


def launch_count():
    return int(load().vnpcc_launch_count())
