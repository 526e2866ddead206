"""TEST INFRASTRUCTURE: a torch-CPU emulation of every wrapper in pose_transfer_b200/kernels.py with the
same signatures and buffer semantics (channel slices, padded layouts, accumulate-vs-overwrite).  It lets the
CPU test-suite execute the HOST logic of the product (engine schedules, concat-slice bookkeeping, trainer)
against the oracle without a GPU.  It is never importable from the product package."""
import contextlib

import torch
import torch.nn.functional as F

from oracle import restate


def _act(v, act):
    if act == 1:
        return F.leaky_relu(v, 0.2)
    if act == 2:
        return F.relu(v)
    if act == 3:
        return torch.tanh(v)
    if act == 4:
        return torch.sigmoid(v)
    return v


def _dact(a, act):
    if act == 1:
        return torch.where(a > 0, torch.ones_like(a), torch.full_like(a, 0.2))
    if act == 2:
        return (a > 0).float()
    return torch.ones_like(a)


def install(K):
    """Monkeypatch module `K` (pose_transfer_b200.kernels); returns a context manager restoring it."""
    Slice = K.Slice

    def S(x):
        return x if isinstance(x, Slice) else Slice(x)

    def view(x, C=None):
        x = S(x)
        C = x.C if C is None else C
        return x.t[..., x.c0:x.c0 + C]

    def nchw(v):
        return v.permute(0, 3, 1, 2)

    def nhwc(v):
        return v.permute(0, 2, 3, 1)

    def nchw_to_nhwc(src, c_src0, C, dst, act=0):
        view(dst, C).copy_(nhwc(_act(src[:, c_src0:c_src0 + C], act)))

    def gather_nhwc(segs, dst, c_total):
        v = view(dst, c_total)
        v.zero_()
        for src, c_src0, C, c_dst in segs:
            v[..., c_dst:c_dst + C].copy_(nhwc(src[:, c_src0:c_src0 + C]))

    def nhwc_to_nchw(src, dst):
        dst.copy_(nchw(view(src, dst.shape[1])))

    def pack_weight(src, dst, A, B, taps, A_pad, B_pad, transpose):
        w = src.reshape(A, B, taps)
        full = torch.zeros(taps, A_pad, B_pad)
        full[:, :A, :B] = w.permute(2, 0, 1)
        if transpose:
            full = full.permute(0, 2, 1)
        dst[:full.numel()].copy_(full.reshape(-1))

    def pack_weight_dual(src, dst0, dst1, A, B, taps, rows0, cols0, rows1, cols1):
        w = src.reshape(A, B, taps).permute(2, 0, 1)
        d0 = dst0[:taps * rows0 * cols0].view(taps, rows0, cols0)
        d1 = dst1[:taps * rows1 * cols1].view(taps, rows1, cols1)
        d0[:, :A, :B] = w
        d1[:, :B, :A] = w.permute(0, 2, 1)

    def unpack_weight_grad(src, grad, A, B, taps, B_pad, accumulate=True):
        v = src[:taps * A * B_pad].reshape(taps, A, B_pad)[:, :, :B].permute(1, 2, 0).reshape(grad.shape)
        if accumulate:
            grad.add_(v)
        else:
            grad.copy_(v)

    def unpack_weight_grad_parts(src, nparts, part_stride, grad, A, B, taps, B_pad, accumulate=True):
        n = taps * A * B_pad
        total = sum(src[q * part_stride:q * part_stride + n] for q in range(nparts))
        unpack_weight_grad(total, grad, A, B, taps, B_pad, accumulate)

    def transpose_weight(src, dst, taps, A, Bp):
        dst[:taps * A * Bp].copy_(src[:taps * A * Bp].reshape(taps, A, Bp).permute(0, 2, 1).reshape(-1))

    def sum_parts(src, nparts, part_stride, dst, n, accumulate=False):
        total = sum(src[q * part_stride:q * part_stride + n] for q in range(nparts))
        if accumulate:
            dst[:n].add_(total)
        else:
            dst[:n].copy_(total)

    def conv_wgrad_plan(g, capacity):
        return 2 if capacity >= 2 * g.k * g.k * g.Cin * g.Cout else 1

    def fill(dst, value=0.0):
        dst.fill_(value)

    def _weights(g, w_t, w_k=None):
        cout_pad = (g.Cout + 3) // 4 * 4
        k = g.k
        if w_t is None:     # only the K-major tensor-core operand [tap][Cout][Cin] was supplied
            return w_k[:k * k * g.Cout * g.Cin].reshape(k, k, g.Cout, g.Cin).permute(0, 1, 3, 2)
        w = w_t[:k * k * g.Cin * cout_pad].reshape(k, k, g.Cin, cout_pad)[..., :g.Cout]
        return w   # [kh][kw][ci][co]

    def head_pack_weights(w, wk, wd):
        Co, Cin = w.shape[0], w.shape[1]
        full = torch.zeros(32, Cin)
        full[:9 * Co] = w.reshape(Co, Cin, 9).permute(2, 0, 1).reshape(9 * Co, Cin)     # row j = tap*Co + co
        wk[:32 * Cin].copy_(full.reshape(-1))
        wd[:32 * Cin].copy_(full.t().reshape(-1))

    def _shift(t, dy, dx):
        """out[n, y, x] = t[n, y + dy, x + dx] with zero fill; t [N,H,W,C]."""
        N, H, W, C = t.shape
        out = torch.zeros_like(t)
        ys, ye = max(0, -dy), min(H, H - dy)
        xs_, xe = max(0, -dx), min(W, W - dx)
        if ys < ye and xs_ < xe:
            out[:, ys:ye, xs_:xe] = t[:, ys + dy:ye + dy, xs_ + dx:xe + dx]
        return out

    def head_shift_add(z, bias, Co, act, y_nchw, y_nhwc=None):
        acc = 0
        for kh in range(3):
            for kw in range(3):
                t = kh * 3 + kw
                acc = acc + _shift(z[..., t * Co:(t + 1) * Co], kh - 1, kw - 1)
        v = _act(acc + (bias if bias is not None else 0), act)
        if y_nchw is not None:
            y_nchw.copy_(nchw(v))
        if y_nhwc is not None:
            view(y_nhwc, Co).copy_(v)

    def head_shift_gather(dz, Co, dzs):
        dzs.zero_()
        for kh in range(3):
            for kw in range(3):
                t = kh * 3 + kw
                dzs[..., t * Co:(t + 1) * Co] = _shift(dz[..., :Co], -(kh - 1), -(kw - 1))

    def head_wgrad_scatter(dwT, nparts, part_stride, Co, Cin, grad, accumulate=True):
        tot = sum(dwT[q * part_stride:q * part_stride + Cin * 32] for q in range(nparts)).reshape(Cin, 32)[:, :9 * Co]
        v = tot.reshape(Cin, 9, Co).permute(2, 0, 1).reshape(grad.shape)
        if accumulate:
            grad.add_(v)
        else:
            grad.copy_(v)

    def conv_forward(g, x, w_t, w_k, bias, act, y, y_nchw=None, stats=None, scratch=None):
        xs = nchw(view(x, g.Cin)).reshape(g.N, g.Cin, g.H, g.W)
        w = _weights(g, w_t, w_k)
        if g.transposed:
            # a strided conv's input extent is not determined by its output extent: dgrad needs output_padding
            op = (g.OH - ((g.H - 1) * g.stride - 2 * g.pad + g.k), g.OW - ((g.W - 1) * g.stride - 2 * g.pad + g.k))
            z = F.conv_transpose2d(xs, w.permute(2, 3, 0, 1), bias, stride=g.stride, padding=g.pad, output_padding=op)
        else:
            z = F.conv2d(xs, w.permute(3, 2, 0, 1), bias, stride=g.stride, padding=g.pad)
        assert tuple(z.shape) == (g.N, g.Cout, g.OH, g.OW), (tuple(z.shape), (g.N, g.Cout, g.OH, g.OW))
        if stats is not None:
            stats[:, 0] += z.double().reshape(g.N, -1).sum(1)
            stats[:, 1] += (z.double() ** 2).reshape(g.N, -1).sum(1)
        o = _act(z, act)
        if y is not None:
            view(y, g.Cout).reshape(g.N, g.OH, g.OW, g.Cout).copy_(nhwc(o))
        if y_nchw is not None:
            y_nchw.copy_(o)

    def conv_wgrad(g, x, dy, dw):
        xs = nchw(view(x, g.Cin)).reshape(g.N, g.Cin, g.H, g.W)
        dys = nchw(view(dy, g.Cout)).reshape(g.N, g.Cout, g.OH, g.OW)
        k = g.k
        if g.transposed:
            w = torch.zeros(g.Cin, g.Cout, k, k, requires_grad=True)
            z = F.conv_transpose2d(xs, w, None, stride=g.stride, padding=g.pad)
            gw, = torch.autograd.grad(z, w, dys)
            res = gw.permute(2, 3, 0, 1).reshape(k * k, g.Cin, g.Cout)       # [tap][Cin][Cout]
        else:
            w = torch.zeros(g.Cout, g.Cin, k, k, requires_grad=True)
            z = F.conv2d(xs, w, None, stride=g.stride, padding=g.pad)
            gw, = torch.autograd.grad(z, w, dys)
            res = gw.permute(2, 3, 0, 1).reshape(k * k, g.Cout, g.Cin)       # [tap][Cout][Cin]
        dw[:res.numel()].copy_(res.reshape(-1))   # scratch is overwritten (include/ptk.h)

    def conv_wgrad_parts(g, x, dy, dw):
        """Emulates a 2-way split: two partial buffers whose sum is the gradient (exercises the reduction)."""
        n = g.k * g.k * g.Cin * g.Cout
        conv_wgrad(g, x, dy, dw)
        if dw.numel() < 2 * n:
            return 1
        dw[n:2 * n].copy_(dw[:n] * 0.25)
        dw[:n].mul_(0.75)
        return 2

    def bias_grad(dy, ld, pixels, C, dbias):
        dbias.add_(dy.reshape(-1, ld)[:pixels, :C].sum(0))

    def gn_stats(z, N, HW, C, stats):
        v = view(z, C).reshape(N, -1).double()
        stats[:, 0] += v.sum(1)
        stats[:, 1] += (v ** 2).sum(1)

    def _mean_rstd(stats, count):
        mean = stats[:, 0] / count
        var = (stats[:, 1] / count - mean ** 2).clamp_min(0)
        return mean.float(), (1.0 / torch.sqrt(var + 1e-3)).float()

    def gn_apply(z, stats, gamma, beta, drop, N, HW, C, out1, act1, out2=None, act2=0):
        v = view(z, C).reshape(N, HW, C)
        if gamma is not None:
            mean, rstd = _mean_rstd(stats, HW * C)
            v = (v - mean.view(N, 1, 1)) * rstd.view(N, 1, 1) * gamma + beta
        if drop is not None:
            v = v * drop.reshape(N, 1, C)
        view(out1, C).reshape(N, HW, C).copy_(_act(v, act1))
        if out2 is not None:
            view(out2, C).reshape(N, HW, C).copy_(_act(v, act2))

    def gn_bwd_reduce(g1, a1, act1, g2, a2, act2, drop, z, stats, N, HW, C, dy, sums):
        g = view(g1, C).reshape(N, HW, C).clone()
        if a1 is not None:
            g = g * _dact(view(a1, C).reshape(N, HW, C), act1)
        if g2 is not None:
            h = view(g2, C).reshape(N, HW, C)
            if a2 is not None:
                h = h * _dact(view(a2, C).reshape(N, HW, C), act2)
            g = g + h
        if drop is not None:
            g = g * drop.reshape(N, 1, C)
        dy.reshape(N, HW, C).copy_(g)
        if sums is not None:
            mean, rstd = _mean_rstd(stats, HW * C)
            xhat = (view(z, C).reshape(N, HW, C) - mean.view(N, 1, 1)) * rstd.view(N, 1, 1)
            sums[:, 0] += g.double().reshape(N, -1).sum(1)
            sums[:, 1] += (g * xhat).double().reshape(N, -1).sum(1)

    def gn_bwd_apply(dy, z, stats, sums, gamma, N, HW, C, dgamma, dbeta):
        mean, rstd = _mean_rstd(stats, HW * C)
        xhat = (view(z, C).reshape(N, HW, C) - mean.view(N, 1, 1)) * rstd.view(N, 1, 1)
        m1 = (sums[:, 0] / (HW * C)).float().view(N, 1, 1)
        m2 = (sums[:, 1] / (HW * C)).float().view(N, 1, 1)
        d = dy.reshape(N, HW, C)
        d.copy_(gamma * rstd.view(N, 1, 1) * (d - m1 - xhat * m2))
        dgamma += sums[:, 1].sum().float()
        dbeta += sums[:, 0].sum().float()

    def mask_pyramid(masks, out):
        _, h, w, _ = out.shape
        out.copy_(nhwc(restate.mask_pyramid_level(masks, h, w)))

    saved_warp = {}

    def mask_pyramid_levels(masks, outs):
        for o in outs:
            mask_pyramid(masks, o)

    def warp_forward(x, warps, mask_lvl, y, argk, N, C, h, w, Kp, H0, W0, act=0, align_corners=False):
        xs = nchw(view(x, C)).reshape(N, C, h, w).clone()
        # masks are already at level resolution: restate.mask_pyramid_level passes them through unchanged
        v = restate.affine_warp(xs, warps, nchw(mask_lvl).double(), (H0, W0), align_corners)
        view(y, C).reshape(N, h, w, C).copy_(nhwc(_act(v, act)))
        saved_warp[argk.data_ptr()] = (xs, warps.clone(), mask_lvl.clone(), H0, W0, align_corners)

    def warp_backward(dy, y, act, warps, mask_lvl, argk, dx, N, C, h, w, Kp, H0, W0, align_corners=False):
        xs, wr, ml, H0s, W0s, ac = saved_warp[argk.data_ptr()]
        xr = xs.clone().requires_grad_(True)
        v = restate.affine_warp(xr, wr, nchw(ml).double(), (H0s, W0s), ac)
        g = nchw(view(dy, C)).reshape(N, C, h, w)
        if act != 0:
            g = g * _dact(nchw(view(y, C)).reshape(N, C, h, w), act)
        gx, = torch.autograd.grad(v, xr, g)
        dx.reshape(N, h, w, C).add_(nhwc(gx))

    def warp_forward_levels(levels, warps, N, Kp, H0, W0, act=0):
        for lv in levels:
            warp_forward(lv["x"], warps, lv["mask"], lv["y"], lv["argk"], N, lv["C"], lv["h"], lv["w"], Kp, H0, W0, act)

    def warp_backward_levels(levels, warps, N, Kp, H0, W0, act=0, zero_dx=True):
        for lv in levels:
            if zero_dx:
                lv["dx"].zero_()
            warp_backward(lv["dy"], lv.get("y"), act, warps, lv["mask"], lv["argk"], lv["dx"], N, lv["C"], lv["h"], lv["w"], Kp, H0, W0)

    def adv_loss(logits, rows, J, n_true, scale, loss, dlogits=None, ldd=1):
        z = logits.reshape(rows, J).clone().requires_grad_(True)
        p = torch.sigmoid(z)
        lt = restate.adv_true(p[:n_true]) * scale if n_true > 0 else torch.zeros(())
        lf = restate.adv_fake(p[n_true:]) * scale if n_true < rows else torch.zeros(())
        (lt + lf).backward()
        loss[0] += lt.detach()
        loss[1] += lf.detach()
        if dlogits is not None:
            dlogits.reshape(-1, ldd)[:, 0].copy_(z.grad.reshape(-1))

    def l1_loss(a, b, scale, loss, grad=None):
        ar = a.clone().requires_grad_(True)
        l = (ar - b).abs().mean() * scale
        l.backward()
        loss[0] += l.detach()
        if grad is not None:
            grad.copy_(ar.grad)

    def nnloss_forward(pred, target, vgg_w, vgg_b, area, scale, loss, argmin):
        with torch.no_grad():
            l = restate.nn_loss(restate.feature_extractor(vgg_w, vgg_b, pred), restate.feature_extractor(vgg_w, vgg_b, target),
                                area, area) * scale
        loss[0] += l

    def nnloss_backward(pred, target, vgg_w, vgg_b, argmin, area, scale, dpred):
        pr = pred.clone().requires_grad_(True)
        l = restate.nn_loss(restate.feature_extractor(vgg_w, vgg_b, pr), restate.feature_extractor(vgg_w, vgg_b, target),
                            area, area) * scale
        l.backward()
        dpred.copy_(pr.grad)

    def nnloss_features_forward(pred, gt, area, scale, loss, argmin):
        with torch.no_grad():
            loss[0] += restate.nn_loss(pred, gt, area, area) * scale

    def nnloss_features_backward(pred, gt, argmin, area, scale, dpred):
        pr = pred.clone().requires_grad_(True)
        (restate.nn_loss(pr, gt, area, area) * scale).backward()
        dpred.copy_(pr.grad)

    def _vgg_consts(C, H, W):
        idx = torch.arange(C * H * W) % 3
        mean = torch.tensor([0.485, 0.456, 0.406])[idx].view(1, C, H, W)
        std = torch.tensor([0.229, 0.224, 0.225])[idx].view(1, C, H, W)
        return mean, std

    def vgg_preprocess(x_nchw, out_nhwc):
        mean, std = _vgg_consts(*x_nchw.shape[1:])
        out_nhwc[..., :3].copy_(nhwc((x_nchw - mean) / std))

    def vgg_preprocess_backward(g_nhwc, dx_nchw):
        _, std = _vgg_consts(*dx_nchw.shape[1:])
        dx_nchw.copy_(nchw(g_nhwc[..., :3]) / std)

    def maxpool2_forward(x, y, N, H, W, C):
        view(y, C).copy_(nhwc(F.max_pool2d(nchw(view(x, C)), 2, 2)))

    def maxpool2_backward(dy, x, dx, N, H, W, C):
        xi = nchw(view(x, C)).clone().requires_grad_(True)
        F.max_pool2d(xi, 2, 2).backward(nchw(view(dy, C)))
        view(dx, C).copy_(nhwc(xi.grad))

    def relu_backward(y, dy, pixels, C):
        g = view(dy, C)
        g.mul_((view(y, C) > 0).float())

    def tanh_bwd_combine(g_nchw, g_nhwc, out_nchw, dz, ld, N, C, H, W):
        g = torch.zeros(N, C, H, W)
        if g_nchw is not None:
            g = g + g_nchw
        if g_nhwc is not None:
            g = g + nchw(view(g_nhwc, C))
        dz.reshape(N, H, W, ld)[..., :C].copy_(nhwc(g * (1 - out_nchw ** 2)))

    def adam_step(p, g, m, v, lr, beta1, beta2, eps, step, grad_scale=1.0):
        gi = g * grad_scale
        m.lerp_(gi, 1 - beta1)
        v.mul_(beta2).addcmul_(gi, gi, value=1 - beta2)
        bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
        p.sub_((lr / bc1) * m / (v.sqrt() / bc2 ** 0.5 + eps))

    table = dict(nchw_to_nhwc=nchw_to_nhwc, gather_nhwc=gather_nhwc, nhwc_to_nchw=nhwc_to_nchw, pack_weight=pack_weight, pack_weight_dual=pack_weight_dual,
                 unpack_weight_grad=unpack_weight_grad, unpack_weight_grad_parts=unpack_weight_grad_parts, fill=fill,
                 transpose_weight=transpose_weight, sum_parts=sum_parts, conv_wgrad_plan=conv_wgrad_plan,
                 head_pack_weights=head_pack_weights, head_shift_add=head_shift_add, head_shift_gather=head_shift_gather,
                 head_wgrad_scatter=head_wgrad_scatter,
                 conv_forward=conv_forward, conv_wgrad=conv_wgrad, conv_wgrad_parts=conv_wgrad_parts,
                 bias_grad=bias_grad, gn_stats=gn_stats, gn_apply=gn_apply, gn_bwd_reduce=gn_bwd_reduce,
                 gn_bwd_apply=gn_bwd_apply, mask_pyramid=mask_pyramid, mask_pyramid_levels=mask_pyramid_levels, warp_forward=warp_forward,
                 warp_backward=warp_backward, warp_forward_levels=warp_forward_levels, warp_backward_levels=warp_backward_levels,
                 adv_loss=adv_loss, l1_loss=l1_loss, nnloss_forward=nnloss_forward,
                 nnloss_backward=nnloss_backward, nnloss_features_forward=nnloss_features_forward,
                 nnloss_features_backward=nnloss_features_backward, vgg_preprocess=vgg_preprocess,
                 vgg_preprocess_backward=vgg_preprocess_backward, maxpool2_forward=maxpool2_forward,
                 maxpool2_backward=maxpool2_backward, relu_backward=relu_backward, tanh_bwd_combine=tanh_bwd_combine, adam_step=adam_step)

    @contextlib.contextmanager
    def ctx():
        old = {k: getattr(K, k) for k in table}
        for k, f in table.items():
            setattr(K, k, f)
        try:
            yield
        finally:
            for k, f in old.items():
                setattr(K, k, f)
    return ctx()
