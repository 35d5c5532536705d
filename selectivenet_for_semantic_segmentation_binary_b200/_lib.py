"""ctypes binding of ``libsunet_b200.so`` (C ABI declared in ``include/sunet_b200.h``).

There is deliberately no fallback: if the shared library is missing, or a call returns a
non-zero status, this module raises.  The CUDA path is the only product path.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SUNET_LIB") or os.path.join(_HERE, "libsunet_b200.so")     # SUNET_LIB: experiments only

A_CONV3X3, A_PLAIN, A_GATHER2X2 = 0, 1, 2
D_NHWC, D_SCATTER2X2 = 0, 1
WORKSPACE_BYTES = 8 << 20
SCALE_MODES = {"None": 0, "clip": 1, "minmax": 2, "sigmoid": 3}


class SunetError(RuntimeError):
    pass


class ConvGemmArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int), ("height", C.c_int), ("width", C.c_int),
        ("a_mode", C.c_int),
        ("src0", C.c_void_p), ("src0_channels", C.c_int), ("src0_pix_stride", C.c_int),
        ("src1", C.c_void_p), ("src1_channels", C.c_int), ("src1_pix_stride", C.c_int),
        ("weights", C.c_void_p), ("n_total", C.c_int), ("k_total", C.c_int),
        ("bias", C.c_void_p),
        ("dst", C.c_void_p), ("dst_pix_stride", C.c_int), ("d_mode", C.c_int),
        ("stats", C.c_void_p),
        ("bnb_y", C.c_void_p), ("bnb_y_pix_stride", C.c_int),
        ("bnb_scale", C.c_void_p), ("bnb_shift", C.c_void_p), ("bnb_mean", C.c_void_p), ("bnb_invstd", C.c_void_p),
        ("ep_scale", C.c_void_p), ("ep_shift", C.c_void_p),
        ("bnb_col0", C.c_int),
        ("pro_scale", C.c_void_p), ("pro_shift", C.c_void_p), ("pro_mask", C.c_int),
    ]


class WgradGemmArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int), ("height", C.c_int), ("width", C.c_int),
        ("a", C.c_void_p), ("a_channels", C.c_int), ("a_pix_stride", C.c_int),
        ("b_mode", C.c_int),
        ("b0", C.c_void_p), ("b0_channels", C.c_int), ("b0_pix_stride", C.c_int),
        ("b1", C.c_void_p), ("b1_channels", C.c_int), ("b1_pix_stride", C.c_int),
        ("partials", C.c_void_p), ("partials_bytes", C.c_size_t),
        ("b_pro_scale", C.c_void_p), ("b_pro_shift", C.c_void_p),
    ]


class PackJob(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("a", C.c_int), ("b", C.c_int), ("tile_start", C.c_int),
        ("w", C.c_void_p), ("bias", C.c_void_p), ("wf", C.c_void_p), ("wd", C.c_void_p), ("bias4", C.c_void_p),
    ]


class AdamTensor(C.Structure):
    _fields_ = [
        ("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
        ("numel", C.c_longlong),
    ]


_vp, _i, _ll, _f, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t

# name -> argtypes (restype is int unless listed in _RESTYPES); this table is also what
# tests/test_abi.py checks against the header.
SIGNATURES = {
    "sunet_abi_version": [],
    "sunet_last_error": [],
    "sunet_launch_count": [],
    "sunet_conv_gemm":[C.POINTER(ConvGemmArgs), _vp],
    "sunet_conv_gemm_stat_rows": [C.POINTER(ConvGemmArgs)],
    "sunet_conv_gemm_bnb_supported": [C.POINTER(ConvGemmArgs)],
    "sunet_conv_gemm_pro_supported": [C.POINTER(ConvGemmArgs)],
    "sunet_wgrad_gemm_pro_supported": [C.POINTER(WgradGemmArgs)],
    "sunet_wgrad_gemm": [C.POINTER(WgradGemmArgs), _vp],
    "sunet_wgrad_gemm_splits": [C.POINTER(WgradGemmArgs)],
    "sunet_wgrad_reduce": [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "sunet_pack_input_im2col": [_vp, _vp, _i, _i, _i, _i, _vp],
    "sunet_pack_input_im2col32": [_vp, _vp, _i, _i, _i, _i, _vp],
    "sunet_pack_conv1_pair_weights": [_vp, _vp, _i, _i, _vp],
    "sunet_pack_input_u8_im2col32": [_vp, _vp, _vp, _vp, _i, _i, _i, _vp],
    "sunet_pack_label_u8": [_vp, _vp, _vp, _i, _i, _i, _vp],
    "sunet_pack_conv3x3_weights": [_vp, _vp, _vp, _i, _i, _vp],
    "sunet_pack_conv1_weights": [_vp, _vp, _i, _i, _vp],
    "sunet_pack_convT_weights": [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "sunet_pack_weights_table": [_vp, _i, _i, _vp],
    "sunet_bn_finalize": [_vp, _i, _i, _ll, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp],
    "sunet_bn_eval_affine": [_vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _i, _vp],
    "sunet_bn_relu_pool": [_vp, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp],
    "sunet_bn_relu_pool_ywin": [_vp, _i, _vp, _vp, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp],
    "sunet_bn_pool_bwd_apply": [_vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _i, _i, _i, _vp,
                                _vp, _vp, _i, _i, _i, _i, _i, _vp, _sz, _vp],
    "sunet_maxpool2x2": [_vp, _i, _vp, _i, _i, _i, _i, _i, _vp],
    "sunet_bn_relu_pool_bwd": [_vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i,
                               _vp, _sz, _vp],
    "sunet_bn_bwd_apply": [_vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _sz,
                           _vp],
    "sunet_colsum_finalize": [_vp, _i, _i, _i, _i, _vp, _vp],
    "sunet_heads_fwd": [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _ll, _vp],
    "sunet_bn_relu_heads": [_vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _ll, _vp],
    "sunet_heads_bwd": [_vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _vp, _sz, _vp],
    "sunet_heads_bwd_bn_rows": [_ll],
    "sunet_heads_bwd_bn": [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp,
                           _vp, _vp, _i, _ll, _vp, _sz, _vp],
    "sunet_loss_sums": [_vp, _vp, _vp, _vp, _ll, _vp, _vp, _vp, _sz, _vp],
    "sunet_loss_finalize": [_vp, _ll, _vp, _f, _f, _vp, _vp],
    "sunet_loss_bwd": [_vp, _vp, _vp, _vp, _ll, _vp, _ll, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp],
    "sunet_metric_hist": [_vp, _vp, _vp, _i, _ll, _f, _f, _i, _vp, _vp],
    "sunet_minmax_f32": [_vp, _ll, _vp, _vp, _sz, _vp],
    "sunet_ensemble_mean": [_vp, _vp, _i, _ll, _i, _vp, _vp],
    "sunet_f32_conv3x3_fwd": [_vp, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "sunet_f32_conv3x3_dgrad": [_vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp],
    "sunet_f32_conv3x3_wgrad": [_vp, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _vp],
    "sunet_f32_convT_fwd": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "sunet_f32_convT_dgrad": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "sunet_f32_convT_wgrad": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "sunet_f32_bn_stats": [_vp, _ll, _i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp],
    "sunet_f32_bn_relu_pool": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "sunet_f32_bn_relu_pool_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp],
    "sunet_f32_heads_fwd": [_vp, _vp, _vp, _i, _vp, _ll, _i, _vp],
    "sunet_f32_heads_bwd": [_vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _ll, _i, _vp],
    "sunet_adam_step": [_vp, _i, _ll, _f, _f, _f, _f, _f, _i, _vp, _vp, _vp],
}
_RESTYPES = {"sunet_last_error": C.c_char_p, "sunet_launch_count": C.c_longlong}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises SunetError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SunetError(
            f"{LIB_PATH} not found: the CUDA extension is not built (run `make` or "
            f"`python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError = symbol missing: fail loudly
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().sunet_last_error()
        raise SunetError(f"{what} failed (status {status}): {msg.decode() if msg else '?'}")
