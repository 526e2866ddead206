#!/usr/bin/env python
"""How much of the content loss (fused VGG conv1_1 + 5x5 NN loss, 1.1 ms of kernel time on a side stream) is EXPOSED in the
step?  Times the resident step with the benchmark objective, with nn_loss_area_size = 1 and with the plain L1 loss.
    python tools/bench_loss_exposure.py"""
import argparse
import contextlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import pose_transfer_b200  # noqa: E402,F401
from pose_transfer_b200.models import pose_gan  # noqa: E402
from oracle import synth  # noqa: E402

H = W = 256
P, N = 18, 8


def run(content, area, l1_w):
    opt = argparse.Namespace(image_size=(H, W), use_input_pose=True, pose_dim=P, batch_size=N, num_stacks=4, gen_type="baseline",
                             warp_skip="mask", dataset="fasion", learning_rate=2e-4, content_loss_layer=content,
                             nn_loss_area_size=area, gan_penalty_weight=1.0, l1_penalty_weight=l1_w)
    with contextlib.redirect_stdout(sys.stderr):
        model = pose_gan.DeformablePose_GAN(opt).cuda()
    model.gen.load_state_dict(synth.fill_state_dict(synth.generator_shapes(P, (H, W)), 0))
    model.disc.load_state_dict(synth.fill_state_dict(synth.discriminator_shapes(3 + 2 * P + 3), 1))
    od = vars(opt)
    bs = [{k: v.cuda() for k, v in synth.make_batch(N, H, W, P, seed=s).items()} for s in range(3)]
    bs = [dict(b, warps=b["warps"].float()) for b in bs]

    def step():
        b, r, b2 = bs
        model.dis_update(b["input"], b["target"], {"warps": b["warps"], "masks": b["masks"]}, r["input"], r["target"], od)
        model.gen_update(b2["input"], b2["target"], {"warps": b2["warps"], "masks": b2["masks"]}, od)
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20


def main():
    for name, cfg in (("nn5 (benchmark objective)", ("block1_conv2", 5, 0.01)), ("nn1", ("block1_conv2", 1, 0.01)),
                      ("plain L1", ("none", 1, 100.0)), ("nn5 again", ("block1_conv2", 5, 0.01))):
        print("%-28s %.3f ms/step" % (name, run(*cfg)))


if __name__ == "__main__":
    main()
