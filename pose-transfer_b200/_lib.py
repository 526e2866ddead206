"""ctypes binding of csrc/libptk.so (the C ABI declared in include/ptk.h)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libptk.so")
_lib = None

i32, i64, f32, vp = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p


class ConvGeom(ctypes.Structure):
    """Mirror of ptk_conv_geom (include/ptk.h)."""
    _fields_ = [(n, i32) for n in ("N", "H", "W", "Cin", "ldx", "OH", "OW", "Cout", "ldy", "k", "stride", "pad",
                                   "transposed", "impl")]


class WarpLevel(ctypes.Structure):
    """Mirror of ptk_warp_level (include/ptk.h)."""
    _fields_ = [("x", vp), ("ldx", i32), ("mask", vp), ("y", vp), ("ldy", i32), ("argk", vp), ("dy", vp), ("lddy", i32),
                ("dx", vp), ("C", i32), ("h", i32), ("w", i32)]


# name -> argtypes (everything returns int unless listed in _RESTYPES)
SIGNATURES = {
    "ptk_version": [],
    "ptk_last_error": [],
    "ptk_launch_count": [],
    "ptk_nchw_to_nhwc": [vp, i32, i32, vp, i32, i32, i32, i32, i32, i32, i32, vp],
    "ptk_gather_nhwc": [ctypes.POINTER(vp), ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.POINTER(i32), i32, vp, i32,
                        i32, i32, i32, i32, i32, vp],
    "ptk_nhwc_to_nchw": [vp, i32, i32, vp, i32, i32, i32, i32, vp],
    "ptk_pack_weight": [vp, vp, i32, i32, i32, i32, i32, i32, vp],
    "ptk_pack_weight_dual": [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp],
    "ptk_unpack_weight_grad": [vp, vp, i32, i32, i32, i32, i32, vp],
    "ptk_fill": [vp, i64, f32, vp],
    "ptk_conv_tc_supported": [ctypes.POINTER(ConvGeom)],
    "ptk_conv_forward": [ctypes.POINTER(ConvGeom), vp, vp, vp, vp, i32, vp, vp, vp, vp],
    "ptk_conv_forward_ws": [ctypes.POINTER(ConvGeom), vp, vp, vp, vp, i32, vp, vp, vp, vp, i64, vp],
    "ptk_conv_wgrad": [ctypes.POINTER(ConvGeom), vp, vp, vp, vp],
    "ptk_conv_wgrad_parts": [ctypes.POINTER(ConvGeom), vp, vp, vp, i64, ctypes.POINTER(i32), vp],
    "ptk_unpack_weight_grad_parts": [vp, i32, i64, vp, i32, i32, i32, i32, i32, vp],
    "ptk_conv_wgrad_plan": [ctypes.POINTER(ConvGeom), i64, ctypes.POINTER(i32)],
    "ptk_transpose_weight": [vp, vp, i32, i32, i32, vp],
    "ptk_sum_parts": [vp, i32, i64, vp, i64, i32, vp],
    "ptk_head_pack_weights": [vp, i32, i32, vp, vp, vp],
    "ptk_head_shift_add": [vp, vp, i32, i32, i32, i32, i32, vp, vp, i32, vp],
    "ptk_head_shift_gather": [vp, i32, i32, i32, i32, i32, vp, vp],
    "ptk_head_wgrad_scatter": [vp, i32, i64, i32, i32, vp, i32, vp],
    "ptk_bias_grad": [vp, i32, i64, i32, vp, vp],
    "ptk_gn_stats": [vp, i32, i32, i64, i32, vp, vp],
    "ptk_gn_apply": [vp, i32, vp, vp, vp, vp, i32, i64, i32, vp, i32, i32, vp, i32, i32, vp],
    "ptk_gn_bwd_reduce": [vp, i32, vp, i32, i32, vp, i32, vp, i32, i32, vp, vp, i32, vp, i32, i64, i32, vp, vp, vp],
    "ptk_gn_bwd_apply": [vp, vp, i32, vp, vp, vp, i32, i64, i32, vp, vp, vp],
    "ptk_mask_pyramid": [vp, i32, i32, i32, i32, vp, i32, i32, vp],
    "ptk_mask_pyramid_levels": [vp, i32, i32, i32, i32, ctypes.POINTER(vp), ctypes.POINTER(i32), ctypes.POINTER(i32), i32, vp],
    "ptk_warp_forward": [vp, i32, vp, vp, vp, i32, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp],
    "ptk_warp_backward": [vp, i32, vp, i32, i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp],
    "ptk_warp_forward_levels": [ctypes.POINTER(WarpLevel), i32, vp, i32, i32, i32, i32, i32, vp],
    "ptk_warp_backward_levels": [ctypes.POINTER(WarpLevel), i32, vp, i32, i32, i32, i32, i32, i32, vp],
    "ptk_adv_loss": [vp, i32, i32, i32, f32, vp, vp, i32, vp],
    "ptk_l1_loss": [vp, vp, i64, f32, vp, vp, vp],
    "ptk_nnloss_forward": [vp, vp, vp, vp, i32, i32, i32, i32, f32, vp, vp, vp],
    "ptk_nnloss_backward": [vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, vp, vp],
    "ptk_nnloss_features_forward": [vp, vp, i32, i32, i32, i32, i32, f32, vp, vp, vp],
    "ptk_nnloss_features_backward": [vp, vp, vp, i32, i32, i32, i32, i32, f32, vp, vp],
    "ptk_tanh_bwd_combine": [vp, vp, i32, vp, vp, i32, i32, i32, i32, i32, vp],
    "ptk_vgg_preprocess": [vp, vp, i32, i32, i32, i32, i32, vp],
    "ptk_maxpool2_forward": [vp, i32, vp, i32, i32, i32, i32, i32, vp],
    "ptk_maxpool2_backward": [vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, vp],
    "ptk_relu_backward": [vp, i32, vp, i32, i64, i32, vp],
    "ptk_pose_heatmaps": [vp, i32, i32, i32, i32, f32, vp, i32, i32, vp],
    "ptk_pose_masks": [vp, i32, i32, i32, i32, vp, vp],
    "ptk_adam_step": [vp, vp, vp, vp, i64, f32, f32, f32, f32, i32, f32, vp],
}
_RESTYPES = {"ptk_last_error": ctypes.c_char_p, "ptk_launch_count": i64}


def lib():
    """Load libptk.so (once).  Raises if it was not built -- there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError("pose_transfer_b200: %s is missing -- run `python -c 'import __graft_entry__ as g; "
                               "g.build()'` (or `make -C pose-transfer_b200/csrc`) first; there is no CPU fallback"
                               % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, i32)
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, lib().ptk_last_error().decode()))


def launch_count():
    return int(lib().ptk_launch_count())
