#!/usr/bin/env python
"""Micro-benchmark of the fused VGG conv1_1 + NN-loss kernels at the benchmark geometry (8 x 3 x 256 x 256, 5x5 window).
    python tools/bench_nnloss.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import pose_transfer_b200  # noqa: E402,F401
from pose_transfer_b200 import kernels as K  # noqa: E402
from oracle import synth  # noqa: E402


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    N, H, W = 8, 256, 256
    g = torch.Generator().manual_seed(0)
    pred = (torch.rand(N, 3, H, W, generator=g) * 2 - 1).cuda()
    tgt = (torch.rand(N, 3, H, W, generator=g) * 2 - 1).cuda()
    vw, vb = synth.vgg_conv1_1(0)
    vw, vb = vw.cuda().contiguous(), vb.cuda().contiguous()
    loss = torch.zeros(1, device="cuda")
    argmin = torch.zeros(N, H, W, dtype=torch.uint8, device="cuda")
    dpred = torch.zeros(N, 3, H, W, device="cuda")
    for area in (5, 3, 1):
        tf = timeit(lambda: K.nnloss_forward(pred, tgt, vw, vb, area, 0.01, loss, argmin))
        tb = timeit(lambda: K.nnloss_backward(pred, tgt, vw, vb, argmin, area, 0.01, dpred))
        print("area=%d  forward %.4f ms  backward %.4f ms" % (area, tf, tb))


if __name__ == "__main__":
    main()
