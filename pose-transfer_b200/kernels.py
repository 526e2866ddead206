"""Thin typed wrappers over the C ABI (include/ptk.h).  Inputs are CUDA torch tensors (memory owners);
every call enqueues on torch's current stream.  No arithmetic happens in Python."""
import torch

from . import _lib
from ._lib import ConvGeom, check

ACT_NONE, ACT_LEAKY, ACT_RELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3, 4
IMPL_AUTO, IMPL_SIMT, IMPL_TC = 0, 1, 2


class Slice:
    """A channel slice [c0, c0+C) of an NHWC buffer `t` ([..., ld])."""
    __slots__ = ("t", "c0", "C")

    def __init__(self, t, c0=0, C=None):
        self.t, self.c0 = t, c0
        self.C = t.shape[-1] - c0 if C is None else C

    @property
    def ld(self):
        return self.t.shape[-1]

    @property
    def ptr(self):
        return self.t.data_ptr() + 4 * self.c0


def _p(x):
    if x is None:
        return None
    if isinstance(x, Slice):
        return x.ptr
    assert x.is_cuda and x.is_contiguous(), "kernel operands must be contiguous CUDA tensors"
    return x.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


# Optional per-kernel-family timing with CUDA events on the launching stream (used by bench.py for the
# roofline numbers; disabled => zero overhead beyond one `is None` test).
PROFILE = None


def profile_start():
    global PROFILE
    PROFILE = {}
    PROFILE_BYTES.clear()


def profile_stop():
    """Returns {family: (launches, total_ms)}; synchronises."""
    global PROFILE
    prof, PROFILE = PROFILE, None
    torch.cuda.synchronize()
    return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in (prof or {}).items()}


PROFILE_DETAIL = False   # also key conv launches by geometry (bench.py --layers)


def _conv_label(name, g):
    pix = g.H * g.W if g.transposed else g.OH * g.OW
    flop = 2.0 * g.N * pix * g.Cin * g.Cout * g.k * g.k
    return "%s|N%d %dx%d c%d->%d k%ds%dp%d %s|%.4g" % (name, g.N, g.H, g.W, g.Cin, g.Cout, g.k, g.stride, g.pad,
                                                   "T" if g.transposed else "C", flop)


PROFILE_BYTES = {}       # family -> algorithmic bytes moved by the profiled launches (HBM-bound families only)


def _timed(name, nbytes=None):
    """nbytes(*args, **kw) -> algorithmic bytes of the call (what the kernel must read + write once), for the rooflines."""
    def deco(fn):
        def wrapper(*a, **kw):
            if PROFILE is None:
                return fn(*a, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **kw)
            e1.record()
            PROFILE.setdefault(name, []).append((e0, e1))
            if nbytes is not None:
                PROFILE_BYTES[name] = PROFILE_BYTES.get(name, 0) + nbytes(*a, **kw)
            if PROFILE_DETAIL and name in ("conv_forward", "conv_wgrad"):
                PROFILE.setdefault(_conv_label(name, a[0]), []).append((e0, e1))
            return r
        wrapper.__name__ = fn.__name__
        wrapper.__doc__ = fn.__doc__
        return wrapper
    return deco


def _as_slice(x):
    return x if isinstance(x, Slice) else Slice(x)


def nchw_to_nhwc(src, c_src0, C, dst, act=ACT_NONE):
    """dst slice <- act(src[:, c_src0:c_src0+C]) ; src [N,Cs,H,W] fp32, dst Slice of [N,H,W,ld]."""
    N, Cs, H, W = src.shape
    dst = _as_slice(dst)
    check(_lib.lib().ptk_nchw_to_nhwc(_p(src), Cs, c_src0, dst.t.data_ptr(), dst.ld, dst.c0, N, C, H, W, act, _stream()),
          "ptk_nchw_to_nhwc")


def gather_nhwc(segs, dst, c_total):
    """dst slice rows [c0, c0 + c_total) <- the NCHW channel ranges segs = [(src [N,Cs,H,W], c_src0, C, c_dst_rel)], zeros in
    between: ONE launch that writes whole sectors (see ptk_gather_nhwc)."""
    import ctypes
    dst = _as_slice(dst)
    n = len(segs)
    N, _, H, W = segs[0][0].shape
    ptrs = (ctypes.c_void_p * n)(*[_p(s[0]) for s in segs])
    Cs = (ctypes.c_int * n)(*[s[0].shape[1] for s in segs])
    c0 = (ctypes.c_int * n)(*[s[1] for s in segs])
    C = (ctypes.c_int * n)(*[s[2] for s in segs])
    cd = (ctypes.c_int * n)(*[s[3] for s in segs])
    check(_lib.lib().ptk_gather_nhwc(ptrs, Cs, c0, C, cd, n, dst.t.data_ptr(), dst.ld, dst.c0, c_total, N, H, W, _stream()), "ptk_gather_nhwc")


def nhwc_to_nchw(src, dst):
    """dst [N,C,H,W] <- src slice."""
    src = _as_slice(src)
    N, C, H, W = dst.shape
    check(_lib.lib().ptk_nhwc_to_nchw(src.t.data_ptr(), src.ld, src.c0, _p(dst), N, C, H, W, _stream()), "ptk_nhwc_to_nchw")


@_timed("pack")
def pack_weight(src, dst, A, B, taps, A_pad, B_pad, transpose):
    check(_lib.lib().ptk_pack_weight(_p(src), _p(dst), A, B, taps, A_pad, B_pad, int(transpose), _stream()), "ptk_pack_weight")


@_timed("pack")
def pack_weight_dual(src, dst0, dst1, A, B, taps, rows0, cols0, rows1, cols1):
    check(_lib.lib().ptk_pack_weight_dual(_p(src), _p(dst0), _p(dst1), A, B, taps, rows0, cols0, rows1, cols1, _stream()),
          "ptk_pack_weight_dual")


@_timed("pack")
def unpack_weight_grad(src, grad, A, B, taps, B_pad, accumulate=True):
    check(_lib.lib().ptk_unpack_weight_grad(_p(src), _p(grad), A, B, taps, B_pad, int(accumulate), _stream()),
          "ptk_unpack_weight_grad")


def fill(dst, value=0.0):
    check(_lib.lib().ptk_fill(_p(dst), dst.numel(), float(value), _stream()), "ptk_fill")


def conv_geom(N, H, W, Cin, ldx, OH, OW, Cout, ldy, k, stride, pad, transposed=False, impl=IMPL_AUTO):
    return ConvGeom(N, H, W, Cin, ldx, OH, OW, Cout, ldy, k, stride, pad, int(transposed), impl)


def conv_tc_supported(g):
    return bool(_lib.lib().ptk_conv_tc_supported(g))


@_timed("conv_forward")
def conv_forward(g, x, w_t, w_k, bias, act, y, y_nchw=None, stats=None, scratch=None):
    """scratch: optional fp32 buffer for the deterministic split-K of small layers (partial tiles + fixed-order reduction)."""
    check(_lib.lib().ptk_conv_forward_ws(g, _p(x), _p(w_t), _p(w_k), _p(bias), act, _p(y), _p(y_nchw), _p(stats), _p(scratch),
                                         scratch.numel() if scratch is not None else 0, _stream()), "ptk_conv_forward")


@_timed("conv_wgrad")
def conv_wgrad(g, x, dy, dw):
    check(_lib.lib().ptk_conv_wgrad(g, _p(x), _p(dy), _p(dw), _stream()), "ptk_conv_wgrad")


@_timed("conv_wgrad")
def conv_wgrad_parts(g, x, dy, dw):
    """Split-K weight gradient without atomics: returns the number of partial gradients written back to back into dw."""
    import ctypes
    n = ctypes.c_int(0)
    check(_lib.lib().ptk_conv_wgrad_parts(g, _p(x), _p(dy), _p(dw), dw.numel(), ctypes.byref(n), _stream()), "ptk_conv_wgrad_parts")
    return n.value


def conv_wgrad_plan(g, capacity):
    """Number of split-K partial gradients conv_wgrad_parts will write into a scratch of `capacity` floats."""
    import ctypes
    n = ctypes.c_int(0)
    check(_lib.lib().ptk_conv_wgrad_plan(g, int(capacity), ctypes.byref(n)), "ptk_conv_wgrad_plan")
    return n.value


@_timed("pack")
def transpose_weight(src, dst, taps, A, Bp):
    """dst[tap][b][a] = src[tap][a][b] (A, Bp multiples of 32)."""
    check(_lib.lib().ptk_transpose_weight(_p(src), _p(dst), taps, A, Bp, _stream()), "ptk_transpose_weight")


@_timed("pack")
def sum_parts(src, nparts, part_stride, dst, n, accumulate=False):
    check(_lib.lib().ptk_sum_parts(_p(src), nparts, part_stride, _p(dst), n, int(accumulate), _stream()), "ptk_sum_parts")


@_timed("pack")
def unpack_weight_grad_parts(src, nparts, part_stride, grad, A, B, taps, B_pad, accumulate=True):
    check(_lib.lib().ptk_unpack_weight_grad_parts(_p(src), nparts, part_stride, _p(grad), A, B, taps, B_pad, int(accumulate),
                                                  _stream()), "ptk_unpack_weight_grad_parts")


@_timed("pack")
def head_pack_weights(w, wk, wd):
    Co, Cin = w.shape[0], w.shape[1]
    check(_lib.lib().ptk_head_pack_weights(_p(w), Co, Cin, _p(wk), _p(wd), _stream()), "ptk_head_pack_weights")


@_timed("head")
def head_shift_add(z, bias, Co, act, y_nchw, y_nhwc=None):
    N, H, W, _ = z.shape
    o2 = _as_slice(y_nhwc) if y_nhwc is not None else None
    check(_lib.lib().ptk_head_shift_add(_p(z), _p(bias), N, Co, H, W, act, _p(y_nchw), o2.ptr if o2 else None,
                                        o2.ld if o2 else 0, _stream()), "ptk_head_shift_add")


@_timed("head")
def head_shift_gather(dz, Co, dzs):
    N, H, W, ldz = dz.shape
    check(_lib.lib().ptk_head_shift_gather(_p(dz), ldz, N, Co, H, W, _p(dzs), _stream()), "ptk_head_shift_gather")


@_timed("head")
def head_wgrad_scatter(dwT, nparts, part_stride, Co, Cin, grad, accumulate=True):
    check(_lib.lib().ptk_head_wgrad_scatter(_p(dwT), nparts, part_stride, Co, Cin, _p(grad), int(accumulate), _stream()),
          "ptk_head_wgrad_scatter")


def bias_grad(dy, ld, pixels, C, dbias):
    check(_lib.lib().ptk_bias_grad(_p(dy), ld, pixels, C, _p(dbias), _stream()), "ptk_bias_grad")


@_timed("gn", lambda z, N, HW, C, stats: 4 * N * HW * C)
def gn_stats(z, N, HW, C, stats):
    z = _as_slice(z)
    check(_lib.lib().ptk_gn_stats(z.ptr, z.ld, N, HW, C, _p(stats), _stream()), "ptk_gn_stats")


@_timed("gn", lambda z, stats, gamma, beta, drop, N, HW, C, out1, act1, out2=None, act2=ACT_NONE:
        4 * N * HW * C * (2 + (out2 is not None)))
def gn_apply(z, stats, gamma, beta, drop, N, HW, C, out1, act1, out2=None, act2=ACT_NONE):
    z, out1 = _as_slice(z), _as_slice(out1)
    o2 = _as_slice(out2) if out2 is not None else None
    check(_lib.lib().ptk_gn_apply(z.ptr, z.ld, _p(stats), _p(gamma), _p(beta), _p(drop), N, HW, C, out1.ptr, out1.ld, act1,
                                  o2.ptr if o2 else None, o2.ld if o2 else 0, act2, _stream()), "ptk_gn_apply")


@_timed("gn", lambda g1, a1, act1, g2, a2, act2, drop, z, stats, N, HW, C, dy, sums:
        4 * N * HW * C * (2 + (a1 is not None) + (g2 is not None) + (a2 is not None) + (sums is not None)))
def gn_bwd_reduce(g1, a1, act1, g2, a2, act2, drop, z, stats, N, HW, C, dy, sums):
    g1 = _as_slice(g1)
    a1 = _as_slice(a1) if a1 is not None else None
    g2 = _as_slice(g2) if g2 is not None else None
    a2 = _as_slice(a2) if a2 is not None else None
    zz = _as_slice(z) if z is not None else None
    check(_lib.lib().ptk_gn_bwd_reduce(g1.ptr, g1.ld, a1.ptr if a1 else None, a1.ld if a1 else 0, act1,
                                       g2.ptr if g2 else None, g2.ld if g2 else 0, a2.ptr if a2 else None,
                                       a2.ld if a2 else 0, act2, _p(drop), zz.ptr if zz else None, zz.ld if zz else 0,
                                       _p(stats), N, HW, C, _p(dy), _p(sums), _stream()), "ptk_gn_bwd_reduce")


@_timed("gn", lambda dy, z, stats, sums, gamma, N, HW, C, dgamma, dbeta: 4 * N * HW * C * 3)
def gn_bwd_apply(dy, z, stats, sums, gamma, N, HW, C, dgamma, dbeta):
    z = _as_slice(z)
    check(_lib.lib().ptk_gn_bwd_apply(_p(dy), z.ptr, z.ld, _p(stats), _p(sums), _p(gamma), N, HW, C, _p(dgamma), _p(dbeta),
                                      _stream()), "ptk_gn_bwd_apply")


@_timed("mask_pyramid")
def mask_pyramid_levels(masks, outs):
    """f64 masks [N,K,H0,W0] -> every f32 [N,h,w,K] tensor of `outs` (cv2.resize INTER_LINEAR) in one launch."""
    import ctypes
    N, K, H0, W0 = masks.shape
    n = len(outs)
    ptrs = (ctypes.c_void_p * n)(*[_p(o) for o in outs])
    hs = (ctypes.c_int * n)(*[o.shape[1] for o in outs])
    ws = (ctypes.c_int * n)(*[o.shape[2] for o in outs])
    check(_lib.lib().ptk_mask_pyramid_levels(_p(masks), N, K, H0, W0, ptrs, hs, ws, n, _stream()), "ptk_mask_pyramid_levels")


@_timed("mask_pyramid")
def mask_pyramid(masks, out):
    """masks [N,K,H0,W0] f64 -> out [N,h,w,K] f32."""
    N, K, H0, W0 = masks.shape
    _, h, w, _ = out.shape
    assert masks.dtype == torch.float64 and out.dtype == torch.float32
    check(_lib.lib().ptk_mask_pyramid(_p(masks), N, K, H0, W0, _p(out), h, w, _stream()), "ptk_mask_pyramid")


@_timed("warp_forward", lambda x, warps, mask_lvl, y, argk, N, C, h, w, K, H0, W0, act=ACT_NONE, align_corners=False:
        N * h * w * (4 * C * 2 + 4 * K))      # read X + write Y + read masks
def warp_forward(x, warps, mask_lvl, y, argk, N, C, h, w, K, H0, W0, act=ACT_NONE, align_corners=False):
    x, y = _as_slice(x), _as_slice(y)
    check(_lib.lib().ptk_warp_forward(x.ptr, x.ld, _p(warps), _p(mask_lvl), y.ptr, y.ld, _p(argk), N, C, h, w, K, H0, W0,
                                      int(align_corners), act, _stream()), "ptk_warp_forward")


@_timed("warp_backward", lambda dy, y, act, warps, mask_lvl, argk, dx, N, C, h, w, K, H0, W0, align_corners=False:
        N * h * w * (4 * C * 2 + 4 * K))      # read dY + write dX + read masks (argk and y excluded)
def warp_backward(dy, y, act, warps, mask_lvl, argk, dx, N, C, h, w, K, H0, W0, align_corners=False):
    dy = _as_slice(dy)
    yy = _as_slice(y) if y is not None else None
    check(_lib.lib().ptk_warp_backward(dy.ptr, dy.ld, yy.ptr if yy else None, yy.ld if yy else 0, act, _p(warps),
                                       _p(mask_lvl), _p(argk), _p(dx), N, C, h, w, K, H0, W0, int(align_corners), _stream()),
          "ptk_warp_backward")


def _warp_levels(levels):
    arr = (_lib.WarpLevel * len(levels))()
    for a, lv in zip(arr, levels):
        for k in ("x", "y", "dy"):
            s = lv.get(k)
            if s is not None:
                s = _as_slice(s)
                setattr(a, k, s.ptr)
                setattr(a, {"x": "ldx", "y": "ldy", "dy": "lddy"}[k], s.ld)
        a.mask, a.argk = _p(lv["mask"]), _p(lv["argk"])
        a.dx = _p(lv.get("dx"))
        a.C, a.h, a.w = lv["C"], lv["h"], lv["w"]
    return arr


def _warp_levels_bytes(levels, warps, N, K, *a, **kw):
    return sum(N * lv["h"] * lv["w"] * (8 * lv["C"] + 4 * K) for lv in levels)


@_timed("warp_forward", _warp_levels_bytes)
def warp_forward_levels(levels, warps, N, K, H0, W0, act=ACT_NONE):
    """All warped levels of a generator forward in one launch.  levels: list of dict(x=Slice, mask=t, y=Slice, argk=t, C, h, w)."""
    arr = _warp_levels(levels)
    check(_lib.lib().ptk_warp_forward_levels(arr, len(levels), _p(warps), N, K, H0, W0, act, _stream()), "ptk_warp_forward_levels")


@_timed("warp_backward", _warp_levels_bytes)
def warp_backward_levels(levels, warps, N, K, H0, W0, act=ACT_NONE, zero_dx=True):
    """levels: list of dict(dy=Slice, mask=t, argk=t, dx=t (dense [N,h,w,C]), y=Slice (LeakyReLU only), C, h, w)."""
    arr = _warp_levels(levels)
    check(_lib.lib().ptk_warp_backward_levels(arr, len(levels), _p(warps), N, K, H0, W0, act, int(zero_dx), _stream()),
          "ptk_warp_backward_levels")


def adv_loss(logits, rows, J, n_true, scale, loss, dlogits=None, ldd=1):
    check(_lib.lib().ptk_adv_loss(_p(logits), rows, J, n_true, float(scale), _p(loss), _p(dlogits), ldd, _stream()),
          "ptk_adv_loss")


def l1_loss(a, b, scale, loss, grad=None):
    check(_lib.lib().ptk_l1_loss(_p(a), _p(b), a.numel(), float(scale), _p(loss), _p(grad), _stream()), "ptk_l1_loss")


@_timed("nnloss")
def nnloss_forward(pred, target, vgg_w, vgg_b, area, scale, loss, argmin):
    N, _, H, W = pred.shape
    check(_lib.lib().ptk_nnloss_forward(_p(pred), _p(target), _p(vgg_w), _p(vgg_b), N, H, W, area, float(scale), _p(loss),
                                        _p(argmin), _stream()), "ptk_nnloss_forward")


@_timed("nnloss")
def nnloss_backward(pred, target, vgg_w, vgg_b, argmin, area, scale, dpred):
    N, _, H, W = pred.shape
    check(_lib.lib().ptk_nnloss_backward(_p(pred), _p(target), _p(vgg_w), _p(vgg_b), _p(argmin), N, H, W, area,
                                         float(scale), _p(dpred), _stream()), "ptk_nnloss_backward")


def nnloss_features_forward(pred, gt, area, scale, loss, argmin):
    """nn_loss on materialised NCHW features (DeformablePose_GAN.nn_loss, pose_gan.py:173-199)."""
    N, C, H, W = pred.shape
    check(_lib.lib().ptk_nnloss_features_forward(_p(pred), _p(gt), N, C, H, W, area, float(scale), _p(loss), _p(argmin),
                                                 _stream()), "ptk_nnloss_features_forward")


def nnloss_features_backward(pred, gt, argmin, area, scale, dpred):
    N, C, H, W = pred.shape
    check(_lib.lib().ptk_nnloss_features_backward(_p(pred), _p(gt), _p(argmin), N, C, H, W, area, float(scale), _p(dpred),
                                                  _stream()), "ptk_nnloss_features_backward")


def tanh_bwd_combine(g_nchw, g_nhwc, out_nchw, dz, ld, N, C, H, W):
    g2 = _as_slice(g_nhwc) if g_nhwc is not None else None
    check(_lib.lib().ptk_tanh_bwd_combine(_p(g_nchw), g2.ptr if g2 else None, g2.ld if g2 else 0, _p(out_nchw), _p(dz), ld,
                                          N, C, H, W, _stream()), "ptk_tanh_bwd_combine")


@_timed("vgg_misc")
def vgg_preprocess(x_nchw, out_nhwc):
    """out_nhwc[..., 0:3] = preprocess_for_vgg(x_nchw) (utils/pose_utils.py:324-331, view-based normalisation)."""
    N, _, H, W = x_nchw.shape
    check(_lib.lib().ptk_vgg_preprocess(_p(x_nchw), _p(out_nhwc), out_nhwc.shape[-1], N, H, W, 0, _stream()), "ptk_vgg_preprocess")


@_timed("vgg_misc")
def vgg_preprocess_backward(g_nhwc, dx_nchw):
    N, _, H, W = dx_nchw.shape
    check(_lib.lib().ptk_vgg_preprocess(_p(g_nhwc), _p(dx_nchw), g_nhwc.shape[-1], N, H, W, 1, _stream()), "ptk_vgg_preprocess")


@_timed("vgg_misc")
def maxpool2_forward(x, y, N, H, W, C):
    x, y = _as_slice(x), _as_slice(y)
    check(_lib.lib().ptk_maxpool2_forward(x.ptr, x.ld, y.ptr, y.ld, N, H, W, C, _stream()), "ptk_maxpool2_forward")


@_timed("vgg_misc")
def maxpool2_backward(dy, x, dx, N, H, W, C):
    dy, x, dx = _as_slice(dy), _as_slice(x), _as_slice(dx)
    check(_lib.lib().ptk_maxpool2_backward(dy.ptr, dy.ld, x.ptr, x.ld, dx.ptr, dx.ld, N, H, W, C, _stream()), "ptk_maxpool2_backward")


@_timed("vgg_misc")
def relu_backward(y, dy, pixels, C):
    y, dy = _as_slice(y), _as_slice(dy)
    check(_lib.lib().ptk_relu_backward(y.ptr, y.ld, dy.ptr, dy.ld, pixels, C, _stream()), "ptk_relu_backward")


@_timed("pose_data")
def pose_heatmaps(kp, out, c0, sigma=6.0):
    """kp int32 [N,P,2] (y, x; -1 missing) -> out[:, c0:c0+P] (NCHW fp32) Gaussian heat-maps (pose_utils.py:79-86)."""
    N, P, _ = kp.shape
    _, C_total, H, W = out.shape
    assert kp.dtype == torch.int32 and out.dtype == torch.float32
    check(_lib.lib().ptk_pose_heatmaps(_p(kp), N, P, H, W, float(sigma), _p(out), C_total, c0, _stream()), "ptk_pose_heatmaps")


@_timed("pose_data")
def pose_masks(kp, masks):
    """kp int32 [N,P,2] -> masks [N,10,H,W] float64 (pose_transform.py:143-214)."""
    N, P, _ = kp.shape
    _, K10, H, W = masks.shape
    assert kp.dtype == torch.int32 and masks.dtype == torch.float64 and K10 == 10
    check(_lib.lib().ptk_pose_masks(_p(kp), N, P, H, W, _p(masks), _stream()), "ptk_pose_masks")


@_timed("adam", lambda p, g, m, v, *a, **k: 28 * p.numel())     # read p, g, m, v; write p, m, v
def adam_step(p, g, m, v, lr, beta1, beta2, eps, step, grad_scale=1.0):
    check(_lib.lib().ptk_adam_step(_p(p), _p(g), _p(m), _p(v), p.numel(), lr, beta1, beta2, eps, step, grad_scale,
                                   _stream()), "ptk_adam_step")
