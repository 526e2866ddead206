"""GPU parity tests of the tcgen05 TF32 implicit-GEMM convolution path (csrc/conv_tc.cu) through the C ABI,
against torch CPU fp32.  Tolerance: TF32 operands (10-bit mantissa) with fp32 accumulation => relative L2
error <= 2e-3 per layer."""
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_l2

pytestmark = pytest.mark.gpu
TF32_TOL = 2e-3


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


TC_CASES = [
    # name, transposed, k, s, p, Cin, Cout, N, H, W
    ("down_64to128_16", False, 4, 2, 1, 64, 128, 2, 16, 16),
    ("down_odd_128to256", False, 4, 2, 1, 128, 256, 3, 15, 31),
    ("down_wide_64to128", False, 4, 2, 1, 64, 128, 2, 32, 256),
    ("down_deep_512to512_8", False, 4, 2, 1, 512, 512, 2, 8, 8),
    ("up_256to64", True, 4, 2, 1, 256, 64, 2, 6, 10),
    ("up_1536to512_4", True, 4, 2, 1, 1536, 512, 2, 4, 4),
    ("up_wide_512to128", True, 4, 2, 1, 512, 128, 2, 16, 128),
    ("k3_128to128", False, 3, 1, 1, 128, 128, 2, 12, 20),
]


def describe(got, ref, what):
    d = (got - ref).abs()
    idx = torch.nonzero(d > 1e-2 * ref.abs().max())
    return "%s rel_l2=%.3e max|d|=%.3e ref_max=%.3e bad=%d/%d first_bad=%s" % (
        what, rel_l2(got, ref), float(d.max()), float(ref.abs().max()), idx.shape[0], d.numel(), idx[:6].tolist())


@pytest.mark.parametrize("case", TC_CASES, ids=[c[0] for c in TC_CASES])
def test_conv_tc_fprop_stats_dgrad(case):
    import pose_transfer_b200  # noqa: F401
    from pose_transfer_b200 import kernels as K
    from pose_transfer_b200.engine import ConvLayer
    name, tr, k, s, p, Cin, Cout, N, H, W = case
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    wshape = (Cin, Cout, k, k) if tr else (Cout, Cin, k, k)
    w = (torch.rand(wshape, generator=g) * 2 - 1) / (Cin * k * k / (s * s if tr else 1)) ** 0.5
    x = torch.randn(N, Cin, H, W, generator=g)
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    z = F.conv_transpose2d(xr, wr, None, stride=s, padding=p) if tr else F.conv2d(xr, wr, None, stride=s, padding=p)
    dz = torch.randn(z.shape, generator=g)
    z.backward(dz)
    zd = z.detach()

    layer = ConvLayer(torch.nn.Parameter(w.cuda()), None, tr, k, s, p)
    layer.impl = K.IMPL_TC
    layer.pack_forward()
    xin = torch.zeros(N, H, W, Cin + 32, device="cuda")          # channel slice of a wider buffer
    xin[..., 32:] = nhwc(x).cuda()
    OH, OW = layer.out_hw(H, W)
    y = torch.full((N, OH, OW, Cout + 64), 7.0, device="cuda")
    stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    layer.forward(K.Slice(xin, 32, Cin), N, H, W, K.Slice(y, 64, Cout), K.ACT_NONE, stats)
    torch.cuda.synchronize()
    got = nchw(y[..., 64:]).cpu()
    assert rel_l2(got, zd) < TF32_TOL, describe(got, zd, "fprop")
    assert float((y[..., :64] - 7.0).abs().max()) == 0, "wrote outside its channel slice"
    ref_stats = torch.stack([zd.double().reshape(N, -1).sum(1), (zd.double() ** 2).reshape(N, -1).sum(1)], 1)
    assert rel_l2(stats[:, 1], ref_stats[:, 1]) < 2e-3, (stats.cpu(), ref_stats)
    rms = float(ref_stats[:, 1].max().sqrt())
    assert float((stats[:, 0].cpu() - ref_stats[:, 0]).abs().max()) < 2e-3 * rms * (zd[0].numel() ** 0.5) + 1e-3

    # dgrad (runs the other kind of gather on the tensor cores)
    dzd = nhwc(dz).cuda()
    dx = torch.zeros(N, H, W, Cin, device="cuda")
    layer.dgrad(K.Slice(dzd), N, H, W, K.Slice(dx))
    torch.cuda.synchronize()
    gotdx = nchw(dx).cpu()
    assert rel_l2(gotdx, xr.grad) < TF32_TOL, describe(gotdx, xr.grad, "dgrad")

    # wgrad on the tensor cores (MN-major operands, split-K over pixel tiles)
    if k == 4:
        gw = torch.zeros(wshape, device="cuda")
        scratch = torch.full((layer.taps * layer.cin_pad * layer.cout_pad,), 3.0, device="cuda")
        layer.wgrad(K.Slice(xin, 32, Cin), K.Slice(dzd), N, H, W, scratch, gw)
        torch.cuda.synchronize()
        assert rel_l2(gw, wr.grad) < TF32_TOL, describe(gw.cpu(), wr.grad, "wgrad")


def test_conv_tc_matches_simt_full_size():
    """Decoder level 5 geometry of BASELINE configs[1] (ConvT 512->128, 128x128 -> 256x256), N=2: tensor-core
    result vs the fp32 CUDA-core path on the same device."""
    import pose_transfer_b200  # noqa: F401
    from pose_transfer_b200 import kernels as K
    from pose_transfer_b200.engine import ConvLayer
    g = torch.Generator().manual_seed(5)
    w = ((torch.rand(512, 128, 4, 4, generator=g) * 2 - 1) / (512 * 4) ** 0.5).cuda()
    x = torch.randn(2, 128, 128, 512, generator=g).cuda()
    outs = []
    for impl in (K.IMPL_SIMT, K.IMPL_TC):
        layer = ConvLayer(torch.nn.Parameter(w), None, True, 4, 2, 1)
        layer.impl = impl
        layer.pack_forward()
        y = torch.zeros(2, 256, 256, 128, device="cuda")
        layer.forward(K.Slice(x), 2, 128, 128, K.Slice(y), K.ACT_NONE, None)
        outs.append(y)
    torch.cuda.synchronize()
    assert rel_l2(outs[1], outs[0]) < TF32_TOL


def _network_layer_geometries():
    """(name, transposed, pad, Cin, Cout, N, H, W) of every k4/s2 layer of G and D at 64x64 (N=2), the deep (small
    spatial extent) layers at 256x256 (N=8) and the odd-extent PatchGAN layers."""
    out = []
    enc64 = (64, 128, 256, 512, 512, 512)
    dec64 = (512, 512, 512, 256, 128, 3)
    hs = [64, 32, 16, 8, 4, 2]
    for i in range(1, 6):
        out.append(("enc64_%d" % i, False, 1, enc64[i - 1], enc64[i], 2, hs[i - 1], hs[i - 1]))
    for j in range(5):
        i = 5 - j
        cin = 2 * enc64[5] if j == 0 else 2 * enc64[i] + dec64[j - 1]
        out.append(("dec64_%d" % j, True, 1, cin, dec64[j], 2, hs[i], hs[i]))
    for i, (ci, co, h) in enumerate(((64, 128, 31), (128, 256, 15), (256, 512, 7))):
        out.append(("disc64_%d" % (i + 1), False, 1, ci, co, 4, h, h))
    for (ci, co, h) in ((512, 512, 16), (512, 512, 8)):
        out.append(("enc256_%dto%d" % (h, h // 2), False, 1, ci, co, 8, h, h))
    for (ci, co, h) in ((1024, 512, 4), (1536, 512, 8)):
        out.append(("dec256_%d" % h, True, 1, ci, co, 8, h, h))
    for i, (ci, co, h) in enumerate(((64, 128, 127), (128, 256, 63), (256, 512, 31))):
        out.append(("disc256_%d" % (i + 1), False, 1, ci, co, 4, h, h))
    return out


@pytest.mark.parametrize("geo", _network_layer_geometries(), ids=[g[0] for g in _network_layer_geometries()])
def test_conv_tc_vs_simt_on_network_geometries(geo):
    """Every op the engines may route to the tensor cores, on the exact layer geometries of the networks:
    tcgen05 result vs the fp32 CUDA-core kernels on the same device (no CPU reference needed)."""
    import pose_transfer_b200  # noqa: F401
    from pose_transfer_b200 import kernels as K
    from pose_transfer_b200.engine import ConvLayer
    name, tr, p, Cin, Cout, N, H, W = geo
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    wshape = (Cin, Cout, 4, 4) if tr else (Cout, Cin, 4, 4)
    w = ((torch.rand(wshape, generator=g) * 2 - 1) / (Cin * 4) ** 0.5).cuda()
    x = torch.randn(N, H, W, Cin, generator=g).cuda()
    res = {}
    for impl in (K.IMPL_SIMT, K.IMPL_AUTO):
        layer = ConvLayer(torch.nn.Parameter(w), None, tr, 4, 2, p)
        layer.impl = impl
        layer.pack_forward()
        OH, OW = layer.out_hw(H, W)
        y = torch.zeros(N, OH, OW, Cout, device="cuda")
        stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
        layer.forward(K.Slice(x), N, H, W, K.Slice(y), K.ACT_NONE, stats)
        dy = torch.randn(N, OH, OW, Cout, generator=torch.Generator().manual_seed(7)).cuda()
        dx = torch.zeros(N, H, W, Cin, device="cuda")
        layer.dgrad(K.Slice(dy), N, H, W, K.Slice(dx))
        gw = torch.zeros(wshape, device="cuda")
        scratch = torch.full((layer.taps * layer.cin_pad * layer.cout_pad,), 3.0, device="cuda")
        layer.wgrad(K.Slice(x), K.Slice(dy), N, H, W, scratch, gw)
        torch.cuda.synchronize()
        res[impl] = (y, stats, dx, gw)
    for what, a, b in zip(("fprop", "stats", "dgrad", "wgrad"), res[K.IMPL_AUTO], res[K.IMPL_SIMT]):
        assert rel_l2(a, b) < TF32_TOL, describe(a.float().cpu(), b.float().cpu(), "%s %s" % (name, what))


STEM_CASES = [
    # name, k, s, p, Cin, Cout, N, H, W, act
    ("stem_k3_21to64", 3, 1, 1, 21, 64, 2, 20, 36, 0),
    ("stem_k3_18to64_wide", 3, 1, 1, 18, 64, 2, 8, 256, 0),
    ("dstem_k4_p0_42to64_leaky", 4, 2, 0, 42, 64, 3, 34, 62, 1),
]


@pytest.mark.parametrize("case", STEM_CASES, ids=[c[0] for c in STEM_CASES])
def test_conv_tc_stems_with_bias_and_padding(case):
    """The small-Cin stems (models/networks.py:186,341) on the tensor cores: input channels zero-padded to 32/64,
    bias (+LeakyReLU for the PatchGAN stem) in the epilogue, wgrad with a half-filled M tile (Cout = 64)."""
    import pose_transfer_b200  # noqa: F401
    from pose_transfer_b200 import kernels as K
    from pose_transfer_b200.engine import ConvLayer
    name, k, s, p, Cin, Cout, N, H, W, act = case
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    w = (torch.rand(Cout, Cin, k, k, generator=g) * 2 - 1) / (Cin * k * k) ** 0.5
    b = torch.rand(Cout, generator=g) - 0.5
    x = torch.randn(N, Cin, H, W, generator=g)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    z = F.conv2d(xr, wr, br, stride=s, padding=p)
    yref = F.leaky_relu(z, 0.2) if act == 1 else z
    dz = torch.randn(z.shape, generator=g)
    z.backward(dz)
    layer = ConvLayer(torch.nn.Parameter(w.cuda()), torch.nn.Parameter(b.cuda()), False, k, s, p)
    layer.impl = K.IMPL_TC
    layer.pack_forward()
    assert layer.cin_pad % 32 == 0
    xin = torch.zeros(N, H, W, layer.cin_pad, device="cuda")
    xin[..., :Cin] = nhwc(x).cuda()
    OH, OW = layer.out_hw(H, W)
    y = torch.zeros(N, OH, OW, Cout, device="cuda")
    layer.forward(K.Slice(xin), N, H, W, K.Slice(y), act, None)
    torch.cuda.synchronize()
    assert rel_l2(nchw(y).cpu(), yref.detach()) < TF32_TOL, describe(nchw(y).cpu(), yref.detach(), "fprop")
    dzd = nhwc(dz).cuda()
    gw = torch.zeros(Cout, Cin, k, k, device="cuda")
    scratch = torch.full((layer.taps * layer.cin_pad * layer.cout_pad,), 3.0, device="cuda")
    layer.wgrad(K.Slice(xin), K.Slice(dzd), N, H, W, scratch, gw)
    torch.cuda.synchronize()
    assert rel_l2(gw, wr.grad) < TF32_TOL, describe(gw.cpu(), wr.grad, "wgrad")
    if layer.cin_pad % 64 == 0:   # only the PatchGAN stem's input gradient is ever needed (gen_update), Cin_pad = 64
        dx = torch.zeros(N, H, W, layer.cin_pad, device="cuda")
        layer.dgrad(K.Slice(dzd), N, H, W, K.Slice(dx), dx_channels=layer.cin_pad)
        torch.cuda.synchronize()
        assert rel_l2(nchw(dx[..., :Cin]).cpu(), xr.grad) < TF32_TOL, describe(nchw(dx[..., :Cin]).cpu(), xr.grad, "dgrad")


@pytest.mark.parametrize("name,k,s,p,Cin,Cout,N,H,W", [("final_dgrad", 3, 1, 1, 256, 3, 2, 24, 40),
                                                         ("dhead_dgrad", 4, 2, 1, 512, 1, 4, 15, 15)])
def test_conv_tc_dgrad_of_narrow_heads(name, k, s, p, Cin, Cout, N, H, W):
    """dgrad of the 3-channel generator head and the 1-channel PatchGAN head on the tensor cores: the gradient
    buffer is zero-padded to 32 channels (one K chunk)."""
    import pose_transfer_b200  # noqa: F401
    from pose_transfer_b200 import kernels as K
    from pose_transfer_b200.engine import ConvLayer
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    w = (torch.rand(Cout, Cin, k, k, generator=g) * 2 - 1) / (Cin * k * k) ** 0.5
    x = torch.randn(N, Cin, H, W, generator=g, requires_grad=True)
    z = F.conv2d(x, w, None, stride=s, padding=p)
    dz = torch.randn(z.shape, generator=g)
    z.backward(dz)
    layer = ConvLayer(torch.nn.Parameter(w.cuda()), None, False, k, s, p)
    layer.impl = K.IMPL_TC
    layer.pack_forward()
    OH, OW = layer.out_hw(H, W)
    dzd = torch.zeros(N, OH, OW, layer.dy_pad, device="cuda")
    dzd[..., :Cout] = nhwc(dz).cuda()
    dx = torch.zeros(N, H, W, Cin, device="cuda")
    layer.dgrad(K.Slice(dzd), N, H, W, K.Slice(dx))
    torch.cuda.synchronize()
    assert rel_l2(nchw(dx).cpu(), x.grad) < TF32_TOL, describe(nchw(dx).cpu(), x.grad, "dgrad")


TILE_CFGS = ["1,128", "1,256", "2,64", "2,128", "2,256"]
TILE_GEOS = [
    # name, transposed, Cin, Cout, N, H, W  (k4 s2 p1): ragged tiles in x / y / batch, phases, multi N-tiles
    ("down_64to256_ragged", False, 64, 256, 3, 22, 38),
    ("up_128to512_ragged", True, 128, 512, 3, 9, 21),
    ("down_128to256_tiny", False, 128, 256, 2, 8, 8),
    ("down_256to256_ragged", False, 256, 256, 3, 12, 20),
    ("up_256to256_ragged", True, 256, 256, 2, 7, 11),
]


@pytest.mark.parametrize("tile", TILE_CFGS)
@pytest.mark.parametrize("geo", TILE_GEOS, ids=[g[0] for g in TILE_GEOS])
def test_conv_tc_tile_shapes(geo, tile, monkeypatch):
    """Every (M-halves, BLOCK_N) tile shape of conv_tc_kernel, forced through PTK_TC_TILE, against torch CPU fp32:
    fprop + fused statistics + dgrad on geometries whose tiles are ragged in x, y and batch."""
    import pose_transfer_b200  # noqa: F401
    from pose_transfer_b200 import kernels as K
    from pose_transfer_b200.engine import ConvLayer
    name, tr, Cin, Cout, N, H, W = geo
    monkeypatch.setenv("PTK_TC_TILE", tile)
    monkeypatch.setenv("PTK_WG_TILE", tile)
    g = torch.Generator().manual_seed(sum(map(ord, name + tile)))
    wshape = (Cin, Cout, 4, 4) if tr else (Cout, Cin, 4, 4)
    w = (torch.rand(wshape, generator=g) * 2 - 1) / (Cin * 16 / (4 if tr else 1)) ** 0.5
    x = torch.randn(N, Cin, H, W, generator=g)
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    z = F.conv_transpose2d(xr, wr, None, stride=2, padding=1) if tr else F.conv2d(xr, wr, None, stride=2, padding=1)
    dz = torch.randn(z.shape, generator=g)
    z.backward(dz)
    zd = z.detach()
    layer = ConvLayer(torch.nn.Parameter(w.cuda()), None, tr, 4, 2, 1)
    layer.impl = K.IMPL_TC
    layer.pack_forward()
    xin = nhwc(x).cuda()
    OH, OW = layer.out_hw(H, W)
    y = torch.full((N, OH, OW, Cout + 32), 7.0, device="cuda")
    stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    layer.forward(K.Slice(xin), N, H, W, K.Slice(y, 32, Cout), K.ACT_NONE, stats)
    torch.cuda.synchronize()
    got = nchw(y[..., 32:]).cpu()
    assert rel_l2(got, zd) < TF32_TOL, describe(got, zd, "fprop " + tile)
    assert float((y[..., :32] - 7.0).abs().max()) == 0, "wrote outside its channel slice"
    ref_sq = (zd.double() ** 2).reshape(N, -1).sum(1)
    assert rel_l2(stats[:, 1].cpu(), ref_sq) < 2e-3
    dx = torch.zeros(N, H, W, Cin, device="cuda")
    layer.dgrad(K.Slice(nhwc(dz).cuda()), N, H, W, K.Slice(dx))
    torch.cuda.synchronize()
    assert rel_l2(nchw(dx).cpu(), xr.grad) < TF32_TOL, describe(nchw(dx).cpu(), xr.grad, "dgrad " + tile)
    # wgrad with the same forced (channel-halves, BLOCK_N) shape where the channel counts allow it
    gw = torch.zeros(wshape, device="cuda")
    scratch = torch.full((6 * layer.taps * layer.cin_pad * layer.cout_pad,), 3.0, device="cuda")   # room for split-K parts
    layer.wgrad(K.Slice(xin), K.Slice(nhwc(dz).cuda()), N, H, W, scratch, gw)
    torch.cuda.synchronize()
    assert rel_l2(gw.cpu(), wr.grad) < TF32_TOL, describe(gw.cpu(), wr.grad, "wgrad " + tile)
    gw2 = torch.zeros(wshape, device="cuda")                                                        # deterministic reduction
    layer.wgrad(K.Slice(xin), K.Slice(nhwc(dz).cuda()), N, H, W, scratch, gw2)
    torch.cuda.synchronize()
    assert torch.equal(gw, gw2), "split-K weight gradient is not bit-reproducible"


@pytest.mark.parametrize("shape", [(2, 64, 20, 36), (3, 256, 17, 9)], ids=["c64_20x36", "c256_17x9"])
def test_head_as_1x1_gemms(shape):
    """csrc/head.cu: Conv2d(C -> 3, k3, p1) + tanh re-expressed as 1x1 tensor-core GEMMs over a 32-column
    'tap x channel' tensor -- forward, weight gradient and input gradient against torch CPU fp32 autograd."""
    import pose_transfer_b200  # noqa: F401
    from pose_transfer_b200 import kernels as K
    N, C, H, W = shape
    g = torch.Generator().manual_seed(N * 1000 + C)
    w = (torch.rand(3, C, 3, 3, generator=g) * 2 - 1) / (9 * C) ** 0.5
    b = torch.rand(3, generator=g) - 0.5
    x = torch.randn(N, C, H, W, generator=g)
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    y = torch.tanh(F.conv2d(xr, wr, b, padding=1))
    gy = torch.randn(y.shape, generator=g)
    y.backward(gy)

    wk = torch.zeros(32 * C, device="cuda")
    wd = torch.zeros(C * 32, device="cuda")
    K.head_pack_weights(w.cuda(), wk, wd)
    xin = torch.zeros(N, H, W, C + 32, device="cuda")           # channel slice of a wider buffer
    xin[..., 32:] = nhwc(x).cuda()
    z27 = torch.empty(N, H, W, 32, device="cuda")
    g1 = K.conv_geom(N, H, W, C, C + 32, H, W, 32, 32, 1, 1, 0, False, K.IMPL_TC)
    K.conv_forward(g1, K.Slice(xin, 32, C), None, wk, None, K.ACT_NONE, K.Slice(z27), None, None)
    out = torch.empty(N, 3, H, W, device="cuda")
    out_nhwc = torch.full((N, H, W, 8), 7.0, device="cuda")
    K.head_shift_add(z27, b.cuda(), 3, K.ACT_TANH, out, K.Slice(out_nhwc, 2, 3))
    torch.cuda.synchronize()
    assert float((out.cpu() - y.detach()).abs().max()) < 2e-3
    assert torch.equal(nchw(out_nhwc[..., 2:5]), out) and float((out_nhwc[..., :2] - 7).abs().max()) == 0

    dz4 = torch.zeros(N, H, W, 4, device="cuda")
    dz4[..., :3] = nhwc(gy * (1 - y.detach() ** 2)).cuda()
    dzs = torch.empty(N, H, W, 32, device="cuda")
    K.head_shift_gather(dz4, 3, dzs)
    scratch = torch.empty(64 * C * 32, device="cuda")
    gw = K.conv_geom(N, H, W, C, C + 32, H, W, 32, 32, 1, 1, 0, True, K.IMPL_TC)
    nparts = K.conv_wgrad_parts(gw, K.Slice(xin, 32, C), K.Slice(dzs), scratch)
    grad = torch.zeros(3, C, 3, 3, device="cuda")
    K.head_wgrad_scatter(scratch, nparts, C * 32, 3, C, grad, True)
    dx = torch.full((N, H, W, C + 32), 5.0, device="cuda")
    gd = K.conv_geom(N, H, W, 32, 32, H, W, C, C + 32, 1, 1, 0, False, K.IMPL_TC)
    K.conv_forward(gd, K.Slice(dzs), None, wd, None, K.ACT_NONE, K.Slice(dx, 32, C), None, None)
    torch.cuda.synchronize()
    assert rel_l2(grad.cpu(), wr.grad) < TF32_TOL, describe(grad.cpu(), wr.grad, "head wgrad")
    assert rel_l2(nchw(dx[..., 32:]).cpu(), xr.grad) < TF32_TOL, describe(nchw(dx[..., 32:]).cpu(), xr.grad, "head dgrad")
    assert float((dx[..., :32] - 5).abs().max()) == 0


SPLITK_CASES = [
    # name, transposed, Cin, Cout, N, H, W   (k4 s2 p1: the bottleneck layers of the U-Net at batch 8 / 2)
    ("enc_512to512_16", False, 512, 512, 8, 16, 16),
    ("enc_512to512_8", False, 512, 512, 8, 8, 8),
    ("dec_1024to512_4", True, 1024, 512, 8, 4, 4),
    ("dec_1536to512_8", True, 1536, 512, 2, 8, 8),
    ("dgrad_like_512to1536_16", False, 512, 1536, 2, 16, 16),
]


@pytest.mark.parametrize("case", SPLITK_CASES, ids=[c[0] for c in SPLITK_CASES])
def test_conv_tc_deterministic_split_k(case):
    """Layers with few output tiles split their K loop over CTAs.  With a scratch buffer every split stores its partial
    tile and splitk_reduce_kernel sums them in a fixed order with the norm statistics fused: correct vs torch, statistics
    correct, and BIT-IDENTICAL from run to run (the atomic fallback without scratch is only equal to rounding)."""
    import pose_transfer_b200  # noqa: F401
    from pose_transfer_b200 import kernels as K
    from pose_transfer_b200.engine import ConvLayer
    name, tr, Cin, Cout, N, H, W = case
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    wshape = (Cin, Cout, 4, 4) if tr else (Cout, Cin, 4, 4)
    w = (torch.rand(wshape, generator=g) * 2 - 1) / (Cin * 16 / (4 if tr else 1)) ** 0.5
    x = torch.randn(N, Cin, H, W, generator=g)
    z = F.conv_transpose2d(x, w, None, stride=2, padding=1) if tr else F.conv2d(x, w, None, stride=2, padding=1)
    layer = ConvLayer(torch.nn.Parameter(w.cuda()), None, tr, 4, 2, 1)
    layer.impl = K.IMPL_TC
    layer.pack_forward()
    xin = nhwc(x).cuda()
    OH, OW = layer.out_hw(H, W)
    scratch = torch.full((1 << 24,), float("nan"), device="cuda")
    outs = []
    for rep in range(3):
        y = torch.full((N, OH, OW, Cout), 7.0, device="cuda")
        stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
        layer.forward(K.Slice(xin), N, H, W, K.Slice(y), K.ACT_NONE, stats, scratch=scratch if rep < 2 else None)
        torch.cuda.synchronize()
        outs.append((y, stats))
    got = nchw(outs[0][0]).cpu()
    assert rel_l2(got, z) < TF32_TOL, describe(got, z, "fprop split-K")
    assert torch.equal(outs[0][0], outs[1][0]), "deterministic split-K differs from run to run"
    assert rel_l2(outs[2][0], outs[0][0]) < 1e-5           # atomic fallback: same numbers up to summation order
    ref_stats = torch.stack([z.double().reshape(N, -1).sum(1), (z.double() ** 2).reshape(N, -1).sum(1)], 1)
    for y, stats in outs:
        assert rel_l2(stats[:, 1], ref_stats[:, 1]) < 2e-3
        mine = torch.stack([y.double().reshape(N, -1).sum(1), (y.double() ** 2).reshape(N, -1).sum(1)], 1)
        assert rel_l2(stats, mine) < 1e-6                  # the fused statistics describe exactly what was written


PERSIST_CASES = [
    # name, transposed, k, s, p, Cin, Cout, N, H, W, bias, act
    ("enc1_like", False, 4, 2, 1, 64, 128, 2, 64, 96, False, "none"),
    ("dec_like_T", True, 4, 2, 1, 128, 64, 2, 24, 40, False, "none"),
    ("head_1x1_to32", False, 1, 1, 0, 256, 32, 2, 40, 56, False, "none"),
    ("head_1x1_from32", False, 1, 1, 0, 32, 256, 2, 40, 56, False, "none"),
    ("k3_bias_leaky", False, 3, 1, 1, 32, 64, 3, 33, 47, True, "leaky"),
    ("odd_extent_p0", False, 4, 2, 0, 64, 64, 3, 31, 45, True, "leaky"),
]


@pytest.mark.parametrize("case", PERSIST_CASES, ids=[c[0] for c in PERSIST_CASES])
def test_conv_tc_persistent_double_buffered(case, monkeypatch):
    """conv_tc_persist_kernel (resident CTAs walking several tiles, two TMEM accumulators: the epilogue of tile i overlaps
    the MMAs of tile i + 1) forced on (PTK_TC_PERSIST=2) must reproduce the one-tile-per-CTA kernel BIT FOR BIT (same
    tile arithmetic, same accumulation order) -- output, bias / activation epilogue and the fused statistics."""
    import pose_transfer_b200  # noqa: F401
    from pose_transfer_b200 import kernels as K
    from pose_transfer_b200.engine import ConvLayer
    name, tr, k, s, p, Cin, Cout, N, H, W, with_bias, act = case
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    wshape = (Cin, Cout, k, k) if tr else (Cout, Cin, k, k)
    w = (torch.rand(wshape, generator=g) * 2 - 1) / (Cin * k * k) ** 0.5
    b = (torch.rand(Cout, generator=g) - 0.5) if with_bias else None
    x = torch.randn(N, Cin, H, W, generator=g)
    z = F.conv_transpose2d(x, w, b, stride=s, padding=p) if tr else F.conv2d(x, w, b, stride=s, padding=p)
    layer = ConvLayer(torch.nn.Parameter(w.cuda()), torch.nn.Parameter(b.cuda()) if with_bias else None, tr, k, s, p)
    layer.impl = K.IMPL_TC
    layer.pack_forward()
    xin = nhwc(x).cuda()
    OH, OW = layer.out_hw(H, W)
    code = K.ACT_LEAKY if act == "leaky" else K.ACT_NONE
    outs = []
    for mode in ("0", "2"):
        monkeypatch.setenv("PTK_TC_PERSIST", mode)
        y = torch.full((N, OH, OW, Cout + 32), 7.0, device="cuda")
        stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
        layer.forward(K.Slice(xin), N, H, W, K.Slice(y, 32, Cout), code, stats if not with_bias else None)
        torch.cuda.synchronize()
        outs.append((y, stats))
    ref = F.leaky_relu(z, 0.2) if act == "leaky" else z
    got = nchw(outs[1][0][..., 32:]).cpu()
    assert rel_l2(got, ref) < TF32_TOL, describe(got, ref, "persistent fprop")
    assert torch.equal(outs[0][0], outs[1][0]), "persistent kernel differs from the one-tile kernel"
    assert float((outs[1][0][..., :32] - 7.0).abs().max()) == 0
    if not with_bias:
        assert rel_l2(outs[1][1], outs[0][1]) < 1e-12        # fp64 atomics: equal up to summation order


AUTOTUNE_CASES = [
    # name, transposed, Cin, Cout, N, H, W
    ("enc_256to512_32", False, 256, 512, 4, 32, 32),
    ("dec_512to128_T_32", True, 512, 128, 4, 32, 32),
    ("enc_64to128_96", False, 64, 128, 2, 96, 96),
]


@pytest.mark.parametrize("case", AUTOTUNE_CASES, ids=[c[0] for c in AUTOTUNE_CASES])
def test_conv_tc_first_use_autotune(case, monkeypatch):
    """First-use autotuning (several tile shapes / split counts / persistent-or-not timed on the real operands, winner
    cached per geometry): whatever it picks is correct vs torch for fprop, dgrad and wgrad, agrees with the static cost
    model's choice up to fp32 summation order, and the cached choice is bit-reproducible from call to call."""
    import pose_transfer_b200  # noqa: F401
    from pose_transfer_b200 import kernels as K
    from pose_transfer_b200.engine import ConvLayer
    name, tr, Cin, Cout, N, H, W = case
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    wshape = (Cin, Cout, 4, 4) if tr else (Cout, Cin, 4, 4)
    w = (torch.rand(wshape, generator=g) * 2 - 1) / (Cin * 16 / (4 if tr else 1)) ** 0.5
    x = torch.randn(N, Cin, H, W, generator=g)
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    z = F.conv_transpose2d(xr, wr, None, stride=2, padding=1) if tr else F.conv2d(xr, wr, None, stride=2, padding=1)
    dz = torch.randn(z.shape, generator=g)
    z.backward(dz)
    xin, dzd = nhwc(x).cuda(), nhwc(dz).cuda()
    res = {}
    for mode in ("0", "1", "1"):
        monkeypatch.setenv("PTK_TC_AUTOTUNE", mode)
        layer = ConvLayer(torch.nn.Parameter(w.cuda()), None, tr, 4, 2, 1)
        layer.impl = K.IMPL_TC
        layer.pack_forward()
        OH, OW = layer.out_hw(H, W)
        scratch = torch.empty(1 << 24, device="cuda")
        y = torch.empty(N, OH, OW, Cout, device="cuda")
        stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
        layer.forward(K.Slice(xin), N, H, W, K.Slice(y), K.ACT_NONE, stats, scratch=scratch)
        dx = torch.empty(N, H, W, Cin, device="cuda")
        layer.dgrad(K.Slice(dzd), N, H, W, K.Slice(dx), scratch=scratch)
        gw = torch.zeros(wshape, device="cuda")
        layer.wgrad(K.Slice(xin), K.Slice(dzd), N, H, W, scratch, gw)
        torch.cuda.synchronize()
        res.setdefault(mode, []).append((y, dx, gw, stats))
    (ym, dxm, gwm, stm), = res["0"]
    (y1, dx1, gw1, st1), (y2, dx2, gw2, st2) = res["1"]
    assert rel_l2(nchw(y1).cpu(), z.detach()) < TF32_TOL and rel_l2(nchw(dx1).cpu(), xr.grad) < TF32_TOL and rel_l2(gw1.cpu(), wr.grad) < TF32_TOL
    assert rel_l2(y1, ym) < 1e-5 and rel_l2(dx1, dxm) < 1e-5 and rel_l2(gw1, gwm) < 1e-5 and rel_l2(st1[:, 1], stm[:, 1]) < 1e-5   # (sum of squares: no cancellation)
    assert torch.equal(y1, y2) and torch.equal(dx1, dx2) and torch.equal(gw1, gw2), "cached tile choice must be reproducible"

