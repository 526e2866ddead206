#!/usr/bin/env python
"""Micro-benchmark of the fused warp launches (forward: the BENCH line's roofline kernel; backward) at the benchmark
geometry (256x256, P=18, batch 8: levels (64,256) (128,128) (256,64) (512,32)), forward, backward and each level alone.  CUDA events around 20 back-to-back launch sets (each moves > 4x the L2 capacity).  Usage on the GPU box:
    python tools/bench_warp.py [N]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import pose_transfer_b200  # noqa: E402,F401
from pose_transfer_b200 import kernels as K  # noqa: E402
from oracle import synth  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    H = 256
    b = synth.make_batch(N, H, H, 18, seed=0)
    wr = b["warps"].float().cuda().contiguous()
    masks = b["masks"].cuda()
    shapes = [(64, 256), (128, 128), (256, 64), (512, 32)]
    lv = []
    for C, h in shapes:
        ml = torch.empty(N, h, h, 10, device="cuda")
        K.mask_pyramid(masks, ml)
        x = torch.randn(N, h, h, C, device="cuda")
        cat = torch.zeros(N, h, h, 2 * C + (0 if C == 64 else C), device="cuda")      # written into a channel slice
        lv.append(dict(x=K.Slice(x), mask=ml, y=K.Slice(cat, 0, C), argk=torch.zeros(N, h, h, C, dtype=torch.uint8, device="cuda"),
                       dy=K.Slice(torch.randn_like(cat), 0, C), dx=torch.zeros(N, h, h, C, device="cuda"), C=C, h=h, w=h))
    alg = sum(N * h * h * (8 * C + 40) for C, h in shapes)
    peak = 6454.9
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        peak = json.load(open(p)).get("hbm_gbs", peak)

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

    print("algorithmic bytes per launch set: %.1f MB (N=%d); HBM peak %.0f GB/s" % (alg / 1e6, N, peak))
    ms = timeit(lambda: K.warp_forward_levels(lv, wr, N, 10, H, H, K.ACT_RELU))
    print("forward  %.4f ms  %.0f GB/s  frac %.3f" % (ms, alg / ms / 1e6, alg / ms / 1e6 / peak))
    ms = timeit(lambda: K.warp_backward_levels(lv, wr, N, 10, H, H, K.ACT_RELU, True))
    print("backward (incl. zero fill)  %.4f ms  %.0f GB/s  frac %.3f" % (ms, alg / ms / 1e6, alg / ms / 1e6 / peak))
    ms = timeit(lambda: K.warp_backward_levels(lv, wr, N, 10, H, H, K.ACT_RELU, False))
    print("backward (no fill; dx accumulates)  %.4f ms" % ms)
    # per-level forward times (single-level launches)
    for d in lv:
        ms = timeit(lambda: K.warp_forward_levels([d], wr, N, 10, H, H, K.ACT_RELU))
        a = N * d["h"] * d["w"] * (8 * d["C"] + 40)
        print("forward level C=%-3d h=%-3d alone: %.4f ms  %.0f GB/s" % (d["C"], d["h"], ms, a / ms / 1e6))


if __name__ == "__main__":
    main()
