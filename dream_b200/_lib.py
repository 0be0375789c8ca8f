"""ctypes binding of libdreamb200.so (the C-ABI declared in include/dreamb200.h).

The product path has no CPU fallback: if the library is missing this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DREAMB200_LIB: another build of the same library (A/B measurements of kernel variants, tools/layer_bench.py)
LIB_PATH = os.environ.get("DREAMB200_LIB") or os.path.join(_HERE, "libdreamb200.so")
MAX_TAPS = 16
OUT_NHWC_F16 = 0
OUT_NCHW_F32 = 1


class ConvDesc(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32),
        ("in_stride", C.c_int32),
        ("w", C.c_void_p), ("bias", C.c_void_p),
        ("taps", C.c_int32), ("Cout_pad", C.c_int32),
        ("tap_dy", C.c_int8 * MAX_TAPS), ("tap_dx", C.c_int8 * MAX_TAPS),
        ("y", C.c_void_p), ("Ho", C.c_int32), ("Wo", C.c_int32),
        ("y_stride_w", C.c_int64), ("y_stride_h", C.c_int64), ("y_stride_b", C.c_int64),
        ("out_mode", C.c_int32), ("cout_real", C.c_int32),
        ("residual", C.c_void_p),
        ("relu", C.c_int32),
        ("residual_f32", C.c_void_p), ("y_f32", C.c_void_p), ("y_pool", C.c_void_p),
        ("absmax", C.c_void_p),
        ("gate", C.c_void_p), ("out_scale", C.c_void_p), ("colsum", C.c_void_p),
    ]


class DreamB200Error(RuntimeError):
    pass


_lib = None

_PROTOS = {
    "dreamb200_last_error": (C.c_char_p, []),
    "dreamb200_version": (C.c_int, []),
    "dreamb200_launch_count": (C.c_int64, []),
    "dreamb200_conv_tile_utilization": (C.c_double, [C.c_int] * 4),
    "dreamb200_conv2d_fwd": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p]),
    "dreamb200_conv2d_fwd_phases": (C.c_int, [C.POINTER(ConvDesc), C.c_int, C.c_void_p]),
    "dreamb200_first_conv3x3": (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 3 + [C.c_void_p]),
    "dreamb200_first_conv3x3_u8": (C.c_int, [C.c_void_p] * 6 + [C.c_int] * 3 + [C.c_void_p]),
    "dreamb200_normalize_u8": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 3 + [C.c_void_p] * 3),
    "dreamb200_belief_targets": (C.c_int, [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p] * 3),
    "dreamb200_im2col_first": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 10 + [C.c_void_p]),
    "dreamb200_maxpool_nhwc": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 9 + [C.c_void_p]),
    "dreamb200_upsample2_nhwc": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p]),
    "dreamb200_add_f16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "dreamb200_nhwc_f16_to_nchw_f32": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]),
    "dreamb200_nchw_f32_to_nhwc_f16": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]),
    "dreamb200_peaks": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_double,
                                  C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "dreamb200_peaks_plan": (C.c_int, [C.c_int] * 5 + [C.POINTER(C.c_longlong), C.POINTER(C.c_int)]),
    "dreamb200_gaussian_smooth": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                            C.c_void_p, C.c_void_p]),
    "dreamb200_wgrad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 6 +
                        [C.c_void_p, C.c_void_p, C.c_void_p]),
    "dreamb200_wgrad_deconv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 6 +
                               [C.c_void_p, C.c_void_p, C.c_void_p]),
    "dreamb200_wgrad_strided": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 8 +
                                [C.c_void_p, C.c_void_p, C.c_void_p]),
    "dreamb200_bn_reduce_workspace": (C.c_int, [C.c_longlong, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_int)]),
    "dreamb200_bn_stats_f16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_void_p]),
    "dreamb200_bn_apply_f16": (C.c_int, [C.c_void_p] * 5 + [C.c_longlong, C.c_int, C.c_int, C.c_void_p]),
    "dreamb200_bn_bwd_reduce_f16": (C.c_int, [C.c_void_p] * 4 + [C.c_longlong, C.c_int, C.c_void_p, C.c_void_p,
                                                               C.c_void_p]),
    "dreamb200_bn_bwd_apply_f16": (C.c_int, [C.c_void_p] * 5 + [C.c_longlong, C.c_int, C.c_void_p]),
    "dreamb200_maxpool3_bwd_nhwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p]),
    "dreamb200_scale_mask_f16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "dreamb200_scale_mask_bias_f16": (C.c_int, [C.c_void_p] * 4 + [C.c_longlong, C.c_int, C.c_void_p]),
    "dreamb200_wgrad_first3x3": (C.c_int, [C.c_void_p] * 3 + [C.c_int] * 3 + [C.c_void_p]),
    "dreamb200_loss_scale_step": (C.c_int, [C.c_void_p] * 4 + [C.c_float, C.c_void_p]),
    "dreamb200_absmax_f16": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]),
    "dreamb200_maxpool2_bwd_nhwc": (C.c_int, [C.c_void_p] * 3 + [C.c_int] * 5 + [C.c_void_p]),
    "dreamb200_upsample2_bwd_nhwc": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p]),
    "dreamb200_bias_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]),
    "dreamb200_softargmax": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 4 +
                             [C.c_void_p, C.c_void_p]),
}

# every symbol include/dreamb200.h declares (checked by tests/test_host_logic.py::test_capi_exports_every_declared_symbol)
DECLARED_SYMBOLS = tuple(_PROTOS.keys())


def lib():
    """Load (once) and return the C-ABI library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DreamB200Error(
            "libdreamb200.so is missing (%s). Run `python -c 'import __graft_entry__ as g; g.build()'`; "
            "there is no CPU fallback for the hot path." % LIB_PATH)
    l = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        if not hasattr(l, name):
            continue
        fn = getattr(l, name)
        fn.restype = res
        fn.argtypes = args
    _lib = l
    return l


def check(rc, what):
    if rc != 0:
        msg = lib().dreamb200_last_error()
        raise DreamB200Error("%s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else "?"))


def launch_count():
    return int(lib().dreamb200_launch_count())
