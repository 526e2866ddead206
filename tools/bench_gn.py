#!/usr/bin/env python
"""Micro-benchmark of the norm-family streaming kernels at the benchmark geometries (batch 8): GB/s by algorithmic bytes.
    python tools/bench_gn.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import pose_transfer_b200  # noqa: E402,F401
from pose_transfer_b200 import kernels as K  # noqa: E402


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
    peak = 6454.9
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        peak = json.load(open(p)).get("hbm_gbs", peak)
    N = 8
    for C, h in ((64, 256), (128, 128), (256, 64), (512, 32)):
        HW = h * h
        z = torch.randn(N, h, h, C, device="cuda")
        a = torch.randn(N, h, h, C, device="cuda")
        g = torch.randn(N, h, h, C, device="cuda")
        g2 = torch.randn(N, h, h, 2 * C, device="cuda")
        a2 = torch.randn(N, h, h, 2 * C, device="cuda")
        o1, o2 = torch.empty_like(z), torch.empty(N, h, h, 2 * C, device="cuda")
        dy = torch.empty_like(z)
        stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
        K.gn_stats(z, N, HW, C, stats)
        gamma, beta = torch.ones(1, device="cuda"), torch.zeros(1, device="cuda")
        sums = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
        dg, db = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
        el = N * HW * C * 4
        rows = [("apply 1 out", lambda: K.gn_apply(z, stats, gamma, beta, None, N, HW, C, o1, K.ACT_LEAKY), 2 * el),
                ("apply 2 out", lambda: K.gn_apply(z, stats, gamma, beta, None, N, HW, C, o1, K.ACT_LEAKY, K.Slice(o2, C, C), K.ACT_RELU), 3 * el)]
        rows.append(("bwd_reduce 1 grad", lambda: K.gn_bwd_reduce(g, a, K.ACT_LEAKY, None, None, K.ACT_NONE, None, z, stats, N, HW, C, dy, sums), 4 * el))
        rows.append(("bwd_reduce 2 grads", lambda: K.gn_bwd_reduce(g, a, K.ACT_LEAKY, K.Slice(g2, C, C), K.Slice(a2, C, C), K.ACT_RELU, None, z,
                                                                    stats, N, HW, C, dy, sums), 6 * el))
        rows.append(("bwd_apply", lambda: K.gn_bwd_apply(dy, z, stats, sums, gamma, N, HW, C, dg, db), 3 * el))
        for name, fn, nbytes in rows:
            ms = timeit(fn)
            print("C=%-3d h=%-3d %-28s %.4f ms  %.0f GB/s  frac %.3f" % (C, h, name, ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / peak))


if __name__ == "__main__":
    main()
