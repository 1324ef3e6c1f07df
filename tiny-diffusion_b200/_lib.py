"""ctypes binding of libtinydiff.so (the C ABI declared in include/tinydiff.h).

There is no CPU fallback: if the shared library is missing, or the device is not an sm_100-class
GPU, every compute call raises.  PyTorch is used only for device memory, streams and
torch.distributed; all arithmetic on the hot path happens inside libtinydiff.so.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libtinydiff.so")

TD_F32, TD_BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_SILU, ACT_GELU, ACT_SIGMOID = 0, 1, 2, 3, 4
CONV_SIMT, CONV_TC, CONV_DIRECT = 0, 1, 2

_P = C.c_void_p


class EmbedArgs(C.Structure):
    _fields_ = [("batch", C.c_int), ("dim", C.c_int), ("in_mode", C.c_int), ("proj_out", C.c_int),
                ("t", _P), ("t_dev", _P), ("w0", _P), ("b0", _P), ("w2", _P), ("b2", _P),
                ("y", _P), ("class_table", _P), ("text", _P), ("proj_w", _P), ("proj_b", _P),
                ("saved", _P), ("proj_out_ptr", _P)]


class EmbedGrads(C.Structure):
    _fields_ = [("d_proj", _P), ("scratch", _P), ("d_w0", _P), ("d_b0", _P), ("d_w2", _P), ("d_b2", _P),
                ("d_class_table", _P), ("num_classes", C.c_int), ("d_proj_w", _P), ("d_proj_b", _P)]


class WgradDesc(C.Structure):
    _fields_ = [("batch", C.c_int), ("height", C.c_int), ("width", C.c_int), ("cin", C.c_int), ("cout", C.c_int),
                ("x_dtype", C.c_int), ("dy_dtype", C.c_int),
                ("x", _P), ("ldx", C.c_int), ("x_coff", C.c_int), ("x_nchw", C.c_int),
                ("dy", _P), ("lddy", C.c_int), ("dy_coff", C.c_int), ("dy_nchw", C.c_int),
                ("dw", _P), ("workspace", _P)]


class GemmArgs(C.Structure):
    _fields_ = [("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
                ("A", _P), ("a_rs", C.c_int64), ("a_cs", C.c_int64),
                ("B", _P), ("b_rs", C.c_int64), ("b_cs", C.c_int64),
                ("C", _P), ("ldc", C.c_int64), ("alpha", C.c_float), ("bias", _P), ("act", C.c_int),
                ("pre_out", _P), ("ld_pre", C.c_int64), ("residual", _P), ("ldr", C.c_int64),
                ("gather_idx", _P), ("gather_table", _P), ("ld_table", C.c_int64),
                ("accumulate", C.c_int), ("splitk_ws", _P), ("allow_tf32", C.c_int)]


class ConvDesc(C.Structure):
    _fields_ = [("batch", C.c_int), ("height", C.c_int), ("width", C.c_int),
                ("cin", C.c_int), ("cout", C.c_int), ("x_dtype", C.c_int), ("y_dtype", C.c_int),
                ("x", _P), ("ldx", C.c_int), ("x_coff", C.c_int),
                ("y", _P), ("ldy", C.c_int), ("y_coff", C.c_int),
                ("w", _P), ("scale", _P), ("shift", _P), ("relu", C.c_int),
                ("stats", _P), ("x_nchw", C.c_int), ("y_nchw", C.c_int), ("splitk_ws", _P),
                ("pool_y", _P), ("pool_ceil", C.c_int)]


_SIGS = {
    "td_version": (C.c_int, []),
    "td_last_error_string": (C.c_char_p, []),
    "td_device_check": (C.c_int, [C.c_int]),
    "td_set_pdl": (C.c_int, [C.c_int]),
    "td_set_sm_budget": (C.c_int, [C.c_int]),
    "td_launch_count": (C.c_int64, []),
    "td_qsample": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int64, C.c_int, _P, _P]),
    "td_mse_num_partials": (C.c_int64, [C.c_int64]),
    "td_mse_grad": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int64, C.c_float, _P]),
    "td_psample_step": (C.c_int, [_P, _P, _P, C.c_int64, _P, _P, C.c_int64, C.c_int, _P, _P]),
    "td_psample_step_cfg": (C.c_int, [_P, _P, C.c_int64, C.c_float, _P, C.c_int64, _P, _P, C.c_int, _P, _P]),
    "td_psample_step_advance": (C.c_int, [_P, _P, _P, C.c_int64, _P, _P, C.c_int64, C.c_int, _P, _P, _P]),
    "td_psample_step_cfg_advance": (C.c_int, [_P, _P, C.c_int64, C.c_float, _P, C.c_int64, _P, _P, C.c_int, _P, _P, _P]),
    "td_bn_fold": (C.c_int, [_P, _P, _P, _P, _P, C.c_float, _P, _P, C.c_int, _P]),
    "td_counter_add": (C.c_int, [_P, C.c_int32, _P]),
    "td_adam_multi": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int64, _P, C.c_float, _P, C.c_double,
                                C.c_double, C.c_float, _P, _P, _P]),
    "td_grad_clip_num_partials": (C.c_int64, [C.c_int64]),
    "td_grad_clip_scale": (C.c_int, [_P, C.c_int64, C.c_float, C.c_float, _P, _P, _P, _P, _P]),
    "td_randint": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int, _P, _P]),
    "td_seed_advance": (C.c_int, [_P, C.c_int64, _P]),
    "td_fill_f32": (C.c_int, [_P, C.c_int64, C.c_float, _P]),
    "td_embed_head_saved_floats": (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    "td_embed_head_fwd": (C.c_int, [C.POINTER(EmbedArgs), _P]),
    "td_embed_head_bwd": (C.c_int, [C.POINTER(EmbedArgs), C.POINTER(EmbedGrads), _P]),
    "td_conv3x3_splitk_workspace": (C.c_int64, [C.POINTER(ConvDesc)]),
    "td_conv3x3_plan_create": (C.c_int, [C.POINTER(_P), C.POINTER(ConvDesc), C.c_int]),
    "td_conv3x3_run": (C.c_int, [_P, _P]),
    "td_conv3x3_pool_fused": (C.c_int, [_P]),
    "td_conv3x3_stats_rows": (C.c_int, [_P]),
    "td_conv3x3_plan_destroy": (None, [_P]),
    "td_conv3x3_flops": (C.c_double, [_P]),
    "td_conv3x3_debug_counters": (C.c_int, [_P, C.c_int]),
    "td_maxpool2_fwd": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "td_upcat_fwd": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.c_int, C.c_int, C.c_int, _P]),
    "td_resize_bilinear_fwd": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "td_final_resize_conv": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int,
                                       C.c_int, _P, _P]),
    "td_cast_f32_to_bf16": (C.c_int, [_P, _P, C.c_int64, _P]),
    "td_pack_conv_weights_multi": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "td_pack_conv_weight": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "td_pack_conv_weight_dgrad": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "td_conv3x3_wgrad_workspace": (C.c_int64, [C.POINTER(WgradDesc), C.c_int]),
    "td_conv3x3_wgrad_plan_create": (C.c_int, [C.POINTER(_P), C.POINTER(WgradDesc), C.c_int]),
    "td_conv3x3_wgrad_run": (C.c_int, [_P, _P]),
    "td_conv3x3_wgrad_plan_destroy": (None, [_P]),
    "td_chan_reduce_rows": (C.c_int, [C.c_int, C.c_int64, C.c_int]),
    "td_bn_stats": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int, C.c_int64, C.c_int, _P, C.c_int, _P]),
    "td_bn_finalize": (C.c_int, [_P, C.c_int, C.c_int, C.c_int64, _P, _P, _P, C.c_float, C.c_float, _P, _P, _P, _P,
                                 _P, _P, _P, _P]),
    "td_bn_relu_apply": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int64, C.c_int, C.c_int64, C.c_int, C.c_int, _P]),
    "td_bn_relu_bwd_reduce": (C.c_int, [_P, C.c_int64, C.c_int, _P, C.c_int, _P, _P, _P, C.c_int64, C.c_int, _P, _P]),
    "td_bn_bwd_finalize": (C.c_int, [_P, C.c_int, C.c_int, C.c_int64, _P, _P, _P, _P, _P, _P, _P]),
    "td_bn_relu_bwd_apply": (C.c_int, [_P, C.c_int64, C.c_int, _P, C.c_int, _P, _P, _P, _P, C.c_int64, C.c_int, _P]),
    "td_bn_apply_fused": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int64, _P, _P, _P, C.c_float, C.c_float, _P, _P, _P, _P, _P, _P, _P,
                                    _P, C.c_int, C.c_int64, C.c_int, C.c_int64, C.c_int, C.c_int, _P]),
    "td_bn_bwd_reduce_rows": (C.c_int, [C.c_int, C.c_int64, C.c_int]),
    "td_bn_bwd_reduce": (C.c_int, [_P, C.c_int64, C.c_int, _P, C.c_int, C.c_int, _P, _P, _P, C.c_int64, C.c_int, _P, _P]),
    "td_bn_bwd_apply_fused": (C.c_int, [_P, C.c_int64, C.c_int, _P, C.c_int, C.c_int, _P, C.c_int, C.c_int64, _P, _P, _P, _P, _P, _P,
                                        _P, C.c_int64, C.c_int, _P]),
    "td_spectral_sigma": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int, C.c_float, _P, _P, _P]),
    "td_scale_by_inv_sigma": (C.c_int, [_P, _P, C.c_int, _P]),
    "td_pack_conv4x4_weight": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "td_conv4x4s2_fwd": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "td_convT4x4s2_fwd": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "td_self_attention_fwd": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "td_dense_tape_op_bytes": (C.c_int, []),
    "td_dense_tape_run": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int, _P]),
    "td_dense_cluster_op_bytes": (C.c_int, []),
    "td_dense_cluster_limits": (C.c_int, [C.c_int, _P, _P, _P, _P, _P, _P]),
    "td_dense_cluster_run": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "td_dense_cluster_debug_counters": (C.c_int, [_P, C.c_int]),
    "td_dense_cluster_step_fused": (C.c_int, []),
    "td_dense_cluster_step": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_int64, _P, _P, _P, C.c_int, _P]),
    "td_dense_cluster_weight_map": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P]),
    "td_maxpool2_bwd": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "td_resize_bilinear_bwd": (C.c_int, [_P, C.c_int64, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_int, _P]),
    "td_upcat_bwd": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.c_int, C.c_int, _P]),
    "td_partial_sum": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "td_nchw_chansum": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "td_gemm_f32_workspace": (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    "td_gemm_f32": (C.c_int, [C.POINTER(GemmArgs), _P]),
    "td_gemm_f32_path": (C.c_int, [_P]),
    "td_colsum_f32": (C.c_int, [_P, C.c_int64, _P, C.c_int, C.c_int, C.c_int, _P]),
    "td_act_bwd_f32": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int, _P]),
    "td_add2d_f32": (C.c_int, [_P, C.c_int64, _P, C.c_int64, C.c_int, C.c_int, C.c_int, _P]),
    "td_transpose_f32": (C.c_int, [_P, C.c_int64, _P, C.c_int64, C.c_int, C.c_int, _P]),
    "td_dropout_f32": (C.c_int, [_P, C.c_int64, _P, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P]),
    "td_embedding_bwd": (C.c_int, [_P, C.c_int64, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int64, _P]),
    "td_time_features": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "td_layernorm_fwd": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_float, _P]),
    "td_layernorm_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, _P, C.c_int64, _P]),
    "td_bn1d_fwd": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int, C.c_int, C.c_float,
                              C.c_float, C.c_int, C.c_int, _P]),
    "td_bn1d_bwd": (C.c_int, [_P, C.c_int64, _P, C.c_int64, _P, C.c_int64, _P, _P, _P, _P, C.c_int64, _P, _P, C.c_int,
                              C.c_int, C.c_int, _P]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)

_lib = None


def load() -> C.CDLL:
    """Load libtinydiff.so (no CUDA call is made by loading)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"libtinydiff.so not found at {LIB_PATH}; build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (or `make -C tiny-diffusion_b200/csrc`). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().td_last_error_string().decode("utf-8", "replace")
        raise RuntimeError(f"libtinydiff {what} failed (status {status}): {msg}")


def require_device(device) -> torch.device:
    """Raise unless `device` is a CUDA device libtinydiff can run on (no CPU fallback)."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"tinydiff runs on sm_100a GPUs only (got device '{device}'); there is no CPU path")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    check(load().td_device_check(idx), "td_device_check")
    return torch.device("cuda", idx)


def ptr(t) -> int:
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return TD_F32
    if dt == torch.bfloat16:
        return TD_BF16
    raise ValueError(f"unsupported dtype {dt}")
