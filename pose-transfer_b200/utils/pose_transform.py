"""Drop-in for ``AffineTransformLayer`` / ``AffineLayer`` of the reference
(src_deformable/utils/pose_transform.py:16-92) on the fused ptk warp kernels.  The CPU pose-geometry helpers
of the reference file (affine_transforms, pose_masks, ... :94-327) are data preparation and out of scope."""
import torch
import torch.nn as nn

from .. import kernels as K


class _WarpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, warps, masks, init_size, use_mask):
        N, C, h, w = x.shape
        Kp = warps.shape[1]
        dev = x.device
        xin = torch.empty(N, h, w, C, device=dev)
        K.nchw_to_nhwc(x.contiguous(), 0, C, K.Slice(xin))
        mlv = torch.empty(N, h, w, Kp, device=dev)
        if use_mask:
            K.mask_pyramid(masks.contiguous().double(), mlv)
        else:
            mlv.fill_(1.0)   # warp_skip != 'mask': no mask multiply (pose_transform.py:78-88)
        y = torch.empty(N, h, w, C, device=dev)
        argk = torch.empty(N, h, w, C, dtype=torch.uint8, device=dev)
        wr = warps.contiguous().float()
        K.warp_forward(xin, wr, mlv, y, argk, N, C, h, w, Kp, init_size[0], init_size[1])
        out = torch.empty(N, C, h, w, device=dev)
        K.nhwc_to_nchw(K.Slice(y), out)
        ctx.save_for_backward(wr, mlv, argk)
        ctx.geom = (N, C, h, w, Kp, init_size)
        return out

    @staticmethod
    def backward(ctx, dout):
        wr, mlv, argk = ctx.saved_tensors
        N, C, h, w, Kp, init_size = ctx.geom
        dev = dout.device
        dy = torch.empty(N, h, w, C, device=dev)
        K.nchw_to_nhwc(dout.contiguous(), 0, C, K.Slice(dy))
        dx = torch.zeros(N, h, w, C, device=dev)
        K.warp_backward(dy, None, K.ACT_NONE, wr, mlv, argk, dx, N, C, h, w, Kp, init_size[0], init_size[1])
        out = torch.empty(N, C, h, w, device=dev)
        K.nhwc_to_nchw(K.Slice(dx), out)
        return out, None, None, None, None


class AffineLayer(nn.Module):
    """Kept for import compatibility (pose_transform.py:16-58); the warp runs fused in AffineTransformLayer."""

    def forward(self, input, transforms):
        raise RuntimeError("AffineLayer is fused into AffineTransformLayer on the B200 path")


class AffineTransformLayer(nn.Module):
    """AffineTransformLayer(number_of_transforms, init_image_size, warp_skip)(input, warps, masks)
    (pose_transform.py:60-92).  input [N,C,h,w] f32 with C % 4 == 0, warps [N,K,8], masks [N,K,H0,W0] (f64 ok).
    Gradient flows to `input` only, like the reference."""

    def __init__(self, number_of_transforms, init_image_size, warp_skip):
        super(AffineTransformLayer, self).__init__()
        self.number_of_transforms = number_of_transforms
        self.init_image_size = tuple(init_image_size)
        self.affine_layer = AffineLayer()
        self.warp_skip = warp_skip

    def forward(self, input, warps, masks):
        if not input.is_cuda:
            raise RuntimeError("AffineTransformLayer: CUDA tensors required (no CPU fallback)")
        if input.shape[1] % 4 != 0:
            raise RuntimeError("AffineTransformLayer: channel count must be a multiple of 4")
        return _WarpFn.apply(input, warps, masks, self.init_image_size, self.warp_skip == 'mask')
