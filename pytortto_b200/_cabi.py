"""ctypes binding of libtortto_b200.so (the C ABI declared in include/tortto_b200.h).

The library is built in-tree by `python -m pytortto_b200.build`.  There is NO fallback: if the shared object is
missing, or a call returns non-zero, a RuntimeError is raised (the product path never routes through the CPU
oracle or through library kernels).
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint8, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtortto_b200.so")
if os.environ.get("TORTTO_B200_LIB") == "tuning":  # the -DTTB_TUNING build (experiment knobs), never the default
    LIB_PATH = os.path.join(HERE, "libtortto_b200_tuning.so")

TTB_MATH_FP32, TTB_MATH_TF32, TTB_MATH_BF16 = 0, 1, 2


class ConvDesc(ctypes.Structure):
    """struct ttb_conv_desc"""
    _fields_ = [(n, c_int32) for n in ("n", "c", "h", "w", "k", "r", "s", "stride_h", "stride_w", "pad_h", "pad_w",
                                       "dil_h", "dil_w", "groups", "p", "q", "math_mode")]


class PoolDesc(ctypes.Structure):
    """struct ttb_pool_desc"""
    _fields_ = [(n, c_int32) for n in ("n", "c", "h", "w", "kh", "kw", "stride_h", "stride_w", "pad_h", "pad_w",
                                       "dil_h", "dil_w", "p", "q")]


class ConvEpilogue(ctypes.Structure):
    """struct ttb_conv_epilogue (device pointers as plain integers, None = NULL)"""
    _fields_ = [("scale", c_void_p), ("bias", c_void_p), ("residual", c_void_p), ("relu", c_int32), ("stats", c_void_p)]


class DgradBnStats(ctypes.Structure):
    """struct ttb_dgrad_bn_stats"""
    _fields_ = [("x", c_void_p), ("mean", c_void_p), ("rscale", c_void_p), ("rshift", c_void_p), ("partials", c_void_p)]


_F = c_void_p  # device pointers travel as plain integers
_PROTOS = {
    "ttb_last_error": (c_char_p, []),
    "ttb_version": (c_int, []),
    "ttb_device_sm_count": (c_int, [POINTER(c_int)]),
    "ttb_conv2d_tensor_path_supported": (c_int, [POINTER(ConvDesc), c_int]),
    "ttb_conv2d_kernel_variant": (c_int, [POINTER(ConvDesc), c_int]),
    "ttb_nchw_to_nhwc": (c_int, [_F, _F, c_int, c_int, c_int, c_int, c_void_p]),
    "ttb_nhwc_to_nchw": (c_int, [_F, _F, c_int, c_int, c_int, c_int, c_void_p]),
    "ttb_conv2d_workspace_size": (c_size_t, [POINTER(ConvDesc), c_int]),
    "ttb_conv2d_fprop": (c_int, [POINTER(ConvDesc), _F, _F, _F, _F, _F, c_size_t, c_void_p]),
    "ttb_conv2d_dgrad": (c_int, [POINTER(ConvDesc), _F, _F, _F, _F, c_size_t, c_void_p]),
    "ttb_conv2d_wgrad": (c_int, [POINTER(ConvDesc), _F, _F, _F, _F, c_size_t, c_void_p]),
    "ttb_conv2d_fused_epilogue_supported": (c_int, [POINTER(ConvDesc)]),
    "ttb_conv2d_fprop_stats_chunks": (c_int, [POINTER(ConvDesc)]),
    "ttb_conv2d_fprop_fused": (c_int, [POINTER(ConvDesc), _F, _F, POINTER(ConvEpilogue), _F, _F, c_size_t, c_void_p]),
    "ttb_bias_grad": (c_int, [_F, _F, c_int64, c_int, c_void_p]),
    "ttb_conv2d_dgrad_prepacked_supported": (c_int, [POINTER(ConvDesc)]),
    "ttb_conv2d_dgrad_pack_weights": (c_int, [c_int, POINTER(POINTER(ConvDesc)), POINTER(c_void_p), POINTER(c_void_p), c_void_p]),
    "ttb_conv2d_dgrad_prepacked": (c_int, [POINTER(ConvDesc), _F, _F, _F, _F, c_void_p]),
    "ttb_conv2d_wgrad_partial": (c_int, [POINTER(ConvDesc), _F, _F, _F, _F, c_size_t, POINTER(c_int), POINTER(c_void_p), c_void_p]),
    "ttb_sum_splits_multi": (c_int, [c_int, POINTER(c_void_p), POINTER(c_int), POINTER(c_int64), POINTER(c_void_p), c_void_p]),
    "ttb_conv2d_bf16_supported": (c_int, [POINTER(ConvDesc), c_int]),
    "ttb_conv2d_workspace_size_bf16": (c_size_t, [POINTER(ConvDesc), c_int]),
    "ttb_conv2d_fprop_bf16": (c_int, [POINTER(ConvDesc), _F, _F, POINTER(ConvEpilogue), _F, c_void_p]),
    "ttb_conv2d_dgrad_bf16": (c_int, [POINTER(ConvDesc), _F, _F, _F, _F, c_void_p]),
    "ttb_conv2d_wgrad_bf16": (c_int, [POINTER(ConvDesc), _F, _F, _F, _F, c_size_t, c_void_p]),
    "ttb_conv2d_dgrad_bn_stats_chunks": (c_int, [POINTER(ConvDesc)]),
    "ttb_conv2d_dgrad_bn": (c_int, [POINTER(ConvDesc), _F, _F, _F, _F, POINTER(DgradBnStats), c_void_p]),
    "ttb_conv2d_pack_weights_bf16": (c_int, [c_int, POINTER(POINTER(ConvDesc)), POINTER(c_void_p), POINTER(c_void_p),
                                             POINTER(c_void_p), c_void_p]),
    "ttb_to_bf16": (c_int, [_F, _F, c_int64, c_void_p]),
    "ttb_bn_num_chunks": (c_int, [c_int64, c_int]),
    "ttb_bn_stats": (c_int, [_F, c_int64, c_int, _F, c_int, c_void_p]),
    "ttb_add_bn_stats": (c_int, [_F, _F, _F, c_int64, c_int, _F, c_int, c_void_p]),
    "ttb_bn_reduce_partials": (c_int, [_F, c_int, c_int, _F, c_void_p]),
    "ttb_bn_finalize": (c_int, [_F, c_int, c_int64, c_int, c_float, c_float, _F, _F, _F, _F, _F, _F, _F, _F, _F, c_void_p]),
    "ttb_bn_prepare_eval": (c_int, [_F, _F, c_int, c_float, _F, _F, _F, _F, _F, _F, _F, c_void_p]),
    "ttb_bn_fold_eval": (c_int, [_F, _F, c_int, c_float, _F, _F, _F, _F, _F, c_void_p]),
    "ttb_bn_apply": (c_int, [_F, _F, c_int64, c_int, _F, _F, _F, c_int, _F, c_void_p]),
    "ttb_bn_apply_add": (c_int, [_F, _F, _F, c_int64, c_int, _F, _F, _F, c_int, _F, c_void_p]),
    "ttb_bn_bwd_reduce": (c_int, [_F, _F, _F, _F, _F, _F, c_int64, c_int, _F, c_int, c_void_p]),
    "ttb_bn_bwd_finalize": (c_int, [_F, c_int, c_int64, c_int, _F, _F, _F, _F, _F, _F, c_void_p]),
    "ttb_bn_bwd_apply": (c_int, [_F, _F, _F, _F, _F, _F, _F, _F, _F, c_int64, c_int, _F, c_void_p]),
    "ttb_relu_fwd": (c_int, [_F, _F, c_int64, _F, c_void_p]),
    "ttb_relu_bwd": (c_int, [_F, _F, _F, c_int64, c_void_p]),
    "ttb_add": (c_int, [_F, _F, _F, c_int64, c_void_p]),
    "ttb_axpy": (c_int, [c_float, _F, _F, c_int64, c_void_p]),
    "ttb_scale": (c_int, [c_float, _F, c_int64, c_void_p]),
    "ttb_fill": (c_int, [c_float, _F, c_int64, c_void_p]),
    "ttb_sgd_step": (c_int, [_F, _F, _F, c_int64, c_float, c_float, c_float, c_float, c_int, c_int, c_void_p]),
    "ttb_sgd_step_multi": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float,
                                   c_float, c_int, c_void_p]),
    "ttb_adam_advance": (c_int, [_F, c_float, c_float, c_void_p]),
    "ttb_adam_step_multi": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, _F, c_float, c_float,
                                    c_float, c_float, c_float, c_int, c_void_p]),
    "ttb_mean_hw_fwd": (c_int, [_F, _F, c_int, c_int, c_int, c_void_p]),
    "ttb_mean_hw_bwd": (c_int, [_F, _F, c_int, c_int, c_int, c_void_p]),
    "ttb_matmul": (c_int, [_F, _F, _F, _F, c_int, c_int, c_int, c_int64, c_int64, c_int64, c_int64, c_void_p]),
    "ttb_log_softmax_fwd": (c_int, [_F, _F, c_int, c_int, c_void_p]),
    "ttb_log_softmax_bwd": (c_int, [_F, _F, _F, c_int, c_int, c_void_p]),
    "ttb_nll_loss_fwd": (c_int, [_F, _F, c_int, c_int, c_int64, c_int, _F, _F, c_void_p]),
    "ttb_nll_loss_bwd": (c_int, [_F, _F, c_int, c_int, c_int64, c_int, _F, _F, c_void_p]),
    "ttb_bce_logits_workspace_size": (c_size_t, []),
    "ttb_bce_logits_fwd": (c_int, [_F, _F, c_int64, c_int, _F, _F, c_void_p]),
    "ttb_bce_logits_bwd": (c_int, [_F, _F, _F, c_int, c_float, c_int64, _F, c_void_p]),
    "ttb_copy_channels": (c_int, [_F, _F, c_int64, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ttb_add_bias": (c_int, [_F, c_int64, c_int, c_void_p]) if False else (c_int, [_F, _F, c_int64, c_int, c_void_p]),
    "ttb_comm_alloc": (c_int, [c_size_t, POINTER(c_void_p), c_void_p]),
    "ttb_comm_open": (c_int, [c_void_p, POINTER(c_void_p)]),
    "ttb_comm_close": (c_int, [c_void_p]),
    "ttb_comm_free": (c_int, [c_void_p]),
    "ttb_comm_slot_bytes": (c_size_t, [c_int]),
    "ttb_comm_allreduce": (c_int, [_F, c_int, c_int, c_void_p, c_int, c_int, c_size_t, _F, c_void_p]),
    "ttb_comm_bn_finalize": (c_int, [_F, c_int, _F, c_int, c_int, c_size_t, c_int64, c_int, c_float, c_float, _F, _F, _F, _F,
                                     _F, _F, _F, _F, _F, c_void_p]),
    "ttb_comm_bn_bwd_finalize": (c_int, [_F, c_int, _F, c_int, c_int, c_size_t, c_int64, c_int, _F, _F, _F, _F, _F, _F,
                                         c_void_p]),
    "ttb_maxpool2d_fwd": (c_int, [POINTER(PoolDesc), _F, _F, _F, c_void_p]),
    "ttb_maxpool2d_bwd": (c_int, [POINTER(PoolDesc), _F, _F, _F, c_int, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_PROTOS)

_lib = None
launch_count = 0  # kernels-launching C-ABI calls made so far (bench.py reports the per-step delta)


def load():
    """dlopen the library (no CUDA context is created by loading)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not found: build it with `python -m pytortto_b200.build` "
                               "(there is no CPU / library fallback for the CUDA path)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error():
    return load().ttb_last_error().decode("utf-8", "replace")


def call(name, *args):
    """Invoke an int-returning entry point; raise on a non-zero status."""
    global launch_count
    rc = getattr(load(), name)(*args)
    launch_count += 1
    if rc != 0:
        raise RuntimeError(f"{name} failed (status {rc}): {last_error()}")
