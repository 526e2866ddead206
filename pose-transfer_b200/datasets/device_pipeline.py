"""Device-side replacement of the per-sample preprocessing in the reference's datasets/PoseTransfer_Dataset.py
(SURVEY 8f-2).  The reference builds every training sample on the host with numpy / skimage (``__getitem__`` :163-189:
two images, 2 x P Gaussian heat-map planes, ten float64 mask planes, ten affine maps) and ships the finished tensors over
PCIe: 14.6 MB per image at 256 x 256.  Here a batch crosses the bus as its two images, the 2 x P key-points and the 80
transform coefficients per sample (0.8 MB per image); heat-maps and masks are produced in HBM by csrc/pose_data.cu in the
tensor contract the trainer expects:

    input  [N, 3 + 2P, H, W] float32   (image | source-pose heat-maps | target-pose heat-maps)   PoseTransfer_Dataset.py:168-186
    target [N, 3, H, W]      float32
    warps  [N, 10, 8]        float32   (main.py:83 casts to float)
    masks  [N, 10, H, W]     float64   (the reference's dtype)
"""
import numpy as np
import torch

from .. import kernels as K
from ..utils import pose_geometry


class DevicePoseBatcher:
    """Turns (images, key-points) into the trainer's batch on the GPU.  ``__call__`` takes host tensors (pinned memory is
    used as is, so the copies are asynchronous) or device tensors."""

    def __init__(self, image_size, pose_dim, device=None):
        self.H, self.W = image_size
        self.P = pose_dim
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)

    @staticmethod
    def warps_on_host(kp_from, kp_to, pose_dim):
        """[N, 10, 8] float32 numpy: pose_transform.affine_transforms per sample (80 numbers each; the ten small
        least-squares problems stay on the CPU)."""
        kf, kt = np.asarray(kp_from), np.asarray(kp_to)
        return np.stack([pose_geometry.affine_transforms(kf[n], kt[n], pose_dim) for n in range(kf.shape[0])]).astype(np.float32)

    def __call__(self, img_from, img_to, kp_from, kp_to, warps=None, need_masks=True, checked=False):
        """img_* [N,3,H,W] float32, kp_* [N,P,2] integer (y, x; -1 = missing).  warps: optional precomputed [N,10,8]
        (a loader worker can run warps_on_host ahead of time).  need_masks=False: the batch only feeds the discriminator's
        'real' half (main.py:82), which uses neither warps nor masks.  checked=True: the caller already validated the torso
        key-points (require_torso)."""
        dev = self.device
        N, P, H, W = img_from.shape[0], self.P, self.H, self.W
        kf = torch.as_tensor(kp_from).to(torch.int32)
        kt = torch.as_tensor(kp_to).to(torch.int32)
        if need_masks and not checked:
            for n in range(N):                     # the reference raises KeyError on a missing torso key-point
                pose_geometry.require_torso(kt[n].tolist() if not kt.is_cuda else kt[n].cpu().tolist(), P)
        if warps is None and need_masks:
            warps = torch.from_numpy(self.warps_on_host(kf.cpu().numpy(), kt.cpu().numpy(), P))
        inp = torch.empty(N, 3 + 2 * P, H, W, device=dev)
        inp[:, :3].copy_(img_from, non_blocking=True)
        target = img_to.to(dev, non_blocking=True)
        kfd, ktd = kf.to(dev, non_blocking=True).contiguous(), kt.to(dev, non_blocking=True).contiguous()
        K.pose_heatmaps(kfd, inp, 3)
        K.pose_heatmaps(ktd, inp, 3 + P)
        if not need_masks:
            return {"input": inp, "target": target}
        masks = torch.empty(N, 10, H, W, dtype=torch.float64, device=dev)
        K.pose_masks(ktd, masks)
        return {"input": inp, "target": target, "warps": torch.as_tensor(warps).to(dev, torch.float32, non_blocking=True),
                "masks": masks}
