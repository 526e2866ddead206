"""GPU: the device data path (csrc/pose_data.cu + datasets/device_pipeline.py) against the oracle restatement
of the reference's host preprocessing: body-part masks BIT-EXACT, Gaussian heat-maps <= 1e-6 (double exp + one rounding on
both sides), and the assembled batch in the trainer's tensor contract driving a real training step."""
import argparse
import contextlib
import io

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
CASES = [(18, 256, 256, 3), (16, 224, 224, 4), (18, 128, 64, 5), (18, 37, 53, 6)]


@pytest.mark.parametrize("P,H,W,seed", CASES)
def test_heatmaps_and_masks_match_oracle(P, H, W, seed):
    from oracle import pose_data as od, synth
    from pose_transfer_b200 import kernels as K
    N = 4
    kp = synth.make_keypoints(N, H, W, P, seed=seed, missing=0.2)
    kpd = kp.to(torch.int32).cuda()
    out = torch.full((N, 3 + 2 * P, H, W), 7.0, device="cuda")
    K.pose_heatmaps(kpd, out, 3 + P)
    masks = torch.full((N, 10, H, W), 7.0, dtype=torch.float64, device="cuda")
    K.pose_masks(kpd, masks)
    assert float((out[:, :3 + P] - 7.0).abs().max()) == 0          # wrote only its channel slice
    nonempty = 0
    for n in range(N):
        want = np.transpose(od.cords_to_map(kp[n].numpy(), (H, W)), [2, 0, 1])
        got = out[n, 3 + P:].cpu().numpy()
        assert np.abs(got - want).max() <= 1e-6
        wm = od.pose_masks(kp[n].numpy(), (H, W), P).astype(np.float64)
        gm = masks[n].cpu().numpy()
        assert np.array_equal(gm, wm), "masks differ in %d pixels" % int((gm != wm).sum())
        nonempty += int((wm[1:].reshape(9, -1).sum(1) > 0).sum())
    assert nonempty > 0 or P == 16


def test_device_batcher_feeds_a_training_step():
    from oracle import pose_data as od, synth
    from pose_transfer_b200.datasets.device_pipeline import DevicePoseBatcher
    from pose_transfer_b200.models import pose_gan
    H = W = 64
    P, N = 18, 2
    kf, kt = synth.make_keypoints(N, H, W, P, seed=1), synth.make_keypoints(N, H, W, P, seed=2)
    g = torch.Generator().manual_seed(0)
    img_from, img_to = torch.rand(N, 3, H, W, generator=g) * 2 - 1, torch.rand(N, 3, H, W, generator=g) * 2 - 1
    batch = DevicePoseBatcher((H, W), P)(img_from.pin_memory(), img_to.pin_memory(), kf, kt)
    assert tuple(batch["input"].shape) == (N, 3 + 2 * P, H, W) and batch["input"].dtype == torch.float32
    assert tuple(batch["masks"].shape) == (N, 10, H, W) and batch["masks"].dtype == torch.float64
    assert tuple(batch["warps"].shape) == (N, 10, 8) and batch["warps"].dtype == torch.float32
    # the reference's own sample for the same key-points (PoseTransfer_Dataset.__getitem__ :163-189)
    for n in range(N):
        want_in = np.concatenate([img_from[n].numpy(), np.transpose(od.cords_to_map(kf[n].numpy(), (H, W)), [2, 0, 1]),
                                  np.transpose(od.cords_to_map(kt[n].numpy(), (H, W)), [2, 0, 1])], 0)
        assert np.abs(batch["input"][n].cpu().numpy() - want_in).max() <= 1e-6
        assert np.array_equal(batch["masks"][n].cpu().numpy(), od.pose_masks(kt[n].numpy(), (H, W), P).astype(np.float64))
        assert np.abs(batch["warps"][n].cpu().numpy() - od.affine_transforms(kf[n].numpy(), kt[n].numpy(), P)).max() <= 1e-3
    opt = argparse.Namespace(image_size=(H, W), use_input_pose=True, pose_dim=P, batch_size=N, num_stacks=4,
                             gen_type="baseline", warp_skip="mask", dataset="fasion", learning_rate=2e-4,
                             content_loss_layer="block1_conv2", nn_loss_area_size=5, gan_penalty_weight=1.0, l1_penalty_weight=0.01)
    with contextlib.redirect_stdout(io.StringIO()):
        model = pose_gan.DeformablePose_GAN(opt).cuda()
    io_ = {"warps": batch["warps"], "masks": batch["masks"]}
    d = model.dis_update(batch["input"], batch["target"], io_, batch["input"], batch["target"], vars(opt))
    out, _, gl = model.gen_update(batch["input"], batch["target"], io_, vars(opt))
    assert all(np.isfinite(d)) and all(np.isfinite(gl)) and tuple(out.shape) == (N, 3, H, W)
