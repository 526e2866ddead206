#!/usr/bin/env python
"""Micro-benchmark of single conv layers of the step (fprop / dgrad / wgrad through ConvLayer) under forced tile shapes
(PTK_TC_TILE / PTK_WG_TILE = "m_halves,block_n"; "auto" = the cost model's own choice), to check the cost model's picks.
    python tools/bench_conv.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import pose_transfer_b200  # noqa: E402,F401
from pose_transfer_b200 import kernels as K  # noqa: E402
from pose_transfer_b200.engine import ConvLayer, SPLITK_SCRATCH  # noqa: E402

# name, transposed, k, stride, pad, Cin, Cout, N, H, W, tiles to try
LAYERS = [
    ("stem 3x3 21->64 @256", False, 3, 1, 1, 21, 64, 8, 256, 256, ["auto", "1,64", "2,64", "1,32", "2,32"]),
    ("enc1 64->128 @256", False, 4, 2, 1, 64, 128, 8, 256, 256, ["auto", "1,128", "2,128", "1,64", "2,64"]),
    ("enc2 128->256 @128", False, 4, 2, 1, 128, 256, 8, 128, 128, ["auto", "1,128", "2,128", "1,256", "2,256"]),
    ("enc3 256->512 @64", False, 4, 2, 1, 256, 512, 8, 64, 64, ["auto", "1,128", "2,128", "1,256", "2,256"]),
    ("enc4 512->512 @32", False, 4, 2, 1, 512, 512, 8, 32, 32, ["auto", "1,128", "2,128", "1,256", "2,256"]),
    ("dec6 128->64 T @128", True, 4, 2, 1, 128, 64, 8, 128, 128, ["auto", "1,64", "2,64", "1,32"]),
    ("dec5 512->128 T @128", True, 4, 2, 1, 512, 128, 8, 128, 128, ["auto", "1,128", "2,128"]),
    ("D1 64->128 @127 N16", False, 4, 2, 1, 64, 128, 16, 127, 127, ["auto", "1,128", "2,128", "1,64"]),
    ("D0 42->64 p0 @256 N16", False, 4, 2, 0, 42, 64, 16, 256, 256, ["auto", "1,64", "2,64", "1,32"]),
    ("dec4 1024->256 T @64", True, 4, 2, 1, 1024, 256, 8, 64, 64, ["auto", "1,256", "2,256", "2,128"]),
    ("dec3 1536->512 T @32", True, 4, 2, 1, 1536, 512, 8, 32, 32, ["auto", "1,256", "2,256", "2,128"]),
    ("dec2 1536->512 T @16", True, 4, 2, 1, 1536, 512, 8, 16, 16, ["auto", "1,256", "2,256", "2,128", "1,128"]),
    ("enc5 512->512 @16", False, 4, 2, 1, 512, 512, 8, 16, 16, ["auto", "1,256", "2,256", "2,128", "1,128"]),
]


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
    scratch = torch.empty(SPLITK_SCRATCH, device="cuda")
    flush = torch.empty(64 << 20, device="cuda")      # 256 MB: larger than L2, written between layers
    for name, tr, k, s, p, Cin, Cout, N, H, W, tiles in LAYERS:
        wshape = (Cin, Cout, k, k) if tr else (Cout, Cin, k, k)
        w = torch.nn.Parameter(torch.randn(wshape, device="cuda") * 0.05)
        layer = ConvLayer(w, None, tr, k, s, p)
        layer.impl = K.IMPL_TC
        layer.pack_forward()
        OH, OW = layer.out_hw(H, W)
        x = torch.randn(N, H, W, layer.cin_pad, device="cuda")
        y = torch.empty(N, OH, OW, Cout, device="cuda")
        dy = torch.randn(N, OH, OW, layer.dy_pad, device="cuda")
        dx = torch.empty(N, H, W, layer.cin_pad, device="cuda")
        stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
        gflop = 2.0 * N * OH * OW * Cout * Cin * k * k / (s * s if tr else 1) / 1e9
        for tile in tiles + ["auto"]:          # "auto" again last: the first rows of a layer run at cooler clocks
            for var in ("PTK_TC_TILE", "PTK_WG_TILE"):
                if tile == "auto":
                    os.environ.pop(var, None)
                else:
                    os.environ[var] = tile
            try:
                flush.zero_()
                tf = timeit(lambda: layer.forward(K.Slice(x), N, H, W, K.Slice(y), K.ACT_NONE, stats, scratch=scratch))
                td = timeit(lambda: layer.dgrad(K.Slice(dy), N, H, W, K.Slice(dx), dx_channels=layer.cin_pad, scratch=scratch))
                gw = torch.zeros(wshape, device="cuda")
                wscratch = torch.empty(max(16 * layer.taps * layer.cin_pad * layer.cout_pad, 1 << 22), device="cuda")
                tw = timeit(lambda: layer.wgrad(K.Slice(x), K.Slice(dy), N, H, W, wscratch, gw))
                print("%-24s tile %-6s fprop %.4f ms %6.1f TF/s | dgrad %.4f ms %6.1f TF/s | wgrad(+unpack) %.4f ms %6.1f TF/s" %
                      (name, tile, tf, gflop / tf, td, gflop / td, tw, gflop / tw))
            except RuntimeError as e:
                print("%-24s tile %-6s unsupported (%s)" % (name, tile, str(e)[:60]))
        os.environ.pop("PTK_TC_TILE", None)
        os.environ.pop("PTK_WG_TILE", None)


if __name__ == "__main__":
    main()
