"""GPU parity tests, kernel level: every ptk kernel family through the C ABI vs the CPU oracle
(torch fp32 CPU ops / oracle.restate) and vs the golden outputs of the unmodified reference."""
import pytest
import torch
import torch.nn.functional as F

from helpers import golden, max_abs, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    import pose_transfer_b200  # noqa: F401
    from pose_transfer_b200 import kernels
    return kernels


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def gen(seed):
    return torch.Generator().manual_seed(seed)


# ----------------------------------------------------------------------------- layout
def test_layout_roundtrip(K):
    x = torch.randn(3, 39, 17, 13, generator=gen(0))
    xd = x.cuda()
    dst = torch.zeros(3, 17, 13, 44, device="cuda")
    K.nchw_to_nhwc(xd, 0, 21, K.Slice(dst, 0, 21))
    K.nchw_to_nhwc(xd, 21, 18, K.Slice(dst, 24, 18), act=K.ACT_LEAKY)
    ref = torch.zeros(3, 17, 13, 44)
    ref[..., :21] = nhwc(x[:, :21])
    ref[..., 24:42] = nhwc(F.leaky_relu(x[:, 21:], 0.2))
    assert torch.equal(dst.cpu(), ref)
    back = torch.empty(3, 18, 17, 13, device="cuda")
    K.nhwc_to_nchw(K.Slice(dst, 24, 18), back)
    assert torch.equal(back.cpu(), F.leaky_relu(x[:, 21:], 0.2))


# ----------------------------------------------------------------------------- convolutions
CONV_CASES = [
    # name, transposed, k, s, p, Cin, Cout, N, H, W, bias, act
    ("stem_k3_21to64", False, 3, 1, 1, 21, 64, 2, 20, 12, True, 0),
    ("down_64to128", False, 4, 2, 1, 64, 128, 2, 16, 16, False, 0),
    ("down_odd_128to256", False, 4, 2, 1, 128, 256, 3, 15, 31, False, 0),
    ("dstem_p0_42to64", False, 4, 2, 0, 42, 64, 2, 34, 34, True, 1),
    ("up_256to64", True, 4, 2, 1, 256, 64, 2, 6, 10, False, 0),
    ("up_1536to512_tiny", True, 4, 2, 1, 1536, 512, 2, 2, 2, False, 0),
    ("final_k3_256to3_tanh", False, 3, 1, 1, 256, 3, 2, 12, 20, True, 3),
    ("dhead_512to1", False, 4, 2, 1, 512, 1, 3, 15, 15, False, 0),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_simt_fprop_dgrad_wgrad(K, case):
    from pose_transfer_b200.engine import ConvLayer
    name, tr, k, s, p, Cin, Cout, N, H, W, use_bias, act = case
    g = gen(hash(name) % 1000)
    wshape = (Cin, Cout, k, k) if tr else (Cout, Cin, k, k)
    w = (torch.rand(wshape, generator=g) * 2 - 1) / (Cin * k * k / (s * s if tr else 1)) ** 0.5
    b = torch.rand(Cout, generator=g) - 0.5 if use_bias else None
    x = torch.randn(N, Cin, H, W, generator=g)
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True) if use_bias else None
    if tr:
        z = F.conv_transpose2d(xr, wr, br, stride=s, padding=p)
    else:
        z = F.conv2d(xr, wr, br, stride=s, padding=p)
    y_ref = {0: z, 1: F.leaky_relu(z, 0.2), 3: torch.tanh(z)}[act]
    dz = torch.randn(z.shape, generator=g)
    z.backward(dz)

    layer = ConvLayer(torch.nn.Parameter(w.cuda()), torch.nn.Parameter(b.cuda()) if use_bias else None, tr, k, s, p)
    layer.impl = K.IMPL_SIMT
    layer.pack_forward()
    layer.pack_backward()
    cin_pad, cout_pad = layer.cin_pad, layer.dy_pad
    xin = torch.zeros(N, H, W, cin_pad + 4, device="cuda")   # wider ld: exercise channel-slice addressing
    xin[..., :Cin] = nhwc(x).cuda()
    OH, OW = layer.out_hw(H, W)
    assert (OH, OW) == tuple(z.shape[2:])
    y = torch.zeros(N, OH, OW, Cout + 8, device="cuda")
    y_nchw = torch.zeros(N, Cout, OH, OW, device="cuda") if Cout <= 4 else None
    layer.forward(K.Slice(xin, 0, cin_pad), N, H, W, K.Slice(y, 4, Cout), act, None, y_nchw)
    got = nchw(y[..., 4:4 + Cout]).cpu()
    assert rel_l2(got, y_ref) < 2e-6, "fprop"
    assert float(y[..., :4].abs().max()) == 0 and float(y[..., 4 + Cout:].abs().max()) == 0, "wrote outside its slice"
    if y_nchw is not None:
        assert rel_l2(y_nchw, y_ref) < 2e-6

    # dgrad (gradient w.r.t. x) and wgrad with the same upstream gradient dz
    dzd = torch.zeros(N, OH, OW, cout_pad, device="cuda")
    dzd[..., :Cout] = nhwc(dz).cuda()
    dx = torch.zeros(N, H, W, cin_pad, device="cuda")
    layer.dgrad(K.Slice(dzd), N, H, W, K.Slice(dx), dx_channels=cin_pad)
    assert rel_l2(nchw(dx[..., :Cin]), xr.grad) < 2e-6, "dgrad"
    gw = torch.zeros(wshape, device="cuda")
    scratch = torch.empty(layer.taps * cin_pad * cout_pad, device="cuda")
    layer.wgrad(K.Slice(xin, 0, cin_pad), K.Slice(dzd), N, H, W, scratch, gw)
    assert rel_l2(gw, wr.grad) < 5e-6, "wgrad"
    if use_bias:
        gb = torch.zeros(Cout, device="cuda")
        K.bias_grad(dzd, cout_pad, N * OH * OW, Cout, gb)
        assert rel_l2(gb, br.grad) < 5e-6, "bias grad"


# ----------------------------------------------------------------------------- norm
@pytest.mark.parametrize("N,C,H,W,with_drop", [(2, 64, 16, 16, False), (3, 128, 9, 7, True), (2, 512, 4, 4, True)])
def test_gn_forward_backward(K, N, C, H, W, with_drop):
    from oracle import restate
    g = gen(N * 1000 + C)
    z = torch.randn(N, C, H, W, generator=g) * 1.7 + 0.3
    gamma = torch.tensor([1.3])
    beta = torch.tensor([-0.2])
    drop = ((torch.rand(N, C, 1, 1, generator=g) < 0.5).float() * 2) if with_drop else None
    zr = z.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = restate.block_norm(zr, gr, br)
    if drop is not None:
        y = y * drop
    o1, o2 = F.leaky_relu(y, 0.2), F.relu(y)
    g1, g2 = torch.randn(o1.shape, generator=g), torch.randn(o2.shape, generator=g)
    (o1 * g1).sum().backward(retain_graph=True)
    (o2 * g2).sum().backward()

    zd = nhwc(z).cuda()
    stats = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    K.gn_stats(zd, N, H * W, C, stats)
    ref_stats = torch.stack([z.double().reshape(N, -1).sum(1), (z.double() ** 2).reshape(N, -1).sum(1)], 1)
    assert rel_l2(stats, ref_stats) < 1e-6
    out1 = torch.zeros(N, H, W, C, device="cuda")
    out2 = torch.zeros(N, H, W, C + 12, device="cuda")
    dd = drop.reshape(N, C).contiguous().cuda() if drop is not None else None
    K.gn_apply(zd, stats, gamma.cuda(), beta.cuda(), dd, N, H * W, C, out1, K.ACT_LEAKY, K.Slice(out2, 8, C), K.ACT_RELU)
    assert max_abs(nchw(out1), o1) < 2e-5
    assert max_abs(nchw(out2[..., 8:8 + C]), o2) < 2e-5
    # backward: two gradient sources through their activations
    dy = torch.empty(N, H, W, C, device="cuda")
    sums = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
    g2w = torch.zeros(N, H, W, C + 12, device="cuda")
    g2w[..., 8:8 + C] = nhwc(g2).cuda()
    K.gn_bwd_reduce(nhwc(g1).cuda(), out1, K.ACT_LEAKY, K.Slice(g2w, 8, C), K.Slice(out2, 8, C), K.ACT_RELU, dd, zd, stats,
                    N, H * W, C, dy, sums)
    dgam, dbet = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    K.gn_bwd_apply(dy, zd, stats, sums, gamma.cuda(), N, H * W, C, dgam, dbet)
    assert rel_l2(nchw(dy), zr.grad) < 2e-5
    assert abs(float(dgam) - float(gr.grad)) < 2e-4 * max(1.0, abs(float(gr.grad)))
    assert abs(float(dbet) - float(br.grad)) < 2e-4 * max(1.0, abs(float(br.grad)))


def test_gn_identity_mode_and_act_backward(K):
    g = gen(5)
    z = torch.randn(2, 8, 5, 6, generator=g)
    zd = nhwc(z).cuda()
    out = torch.zeros_like(zd)
    K.gn_apply(zd, None, None, None, None, 2, 30, 8, out, K.ACT_LEAKY)
    assert torch.equal(nchw(out).cpu(), F.leaky_relu(z, 0.2))
    gu = torch.randn(2, 8, 5, 6, generator=g)
    dy = torch.empty_like(zd)
    K.gn_bwd_reduce(nhwc(gu).cuda(), out, K.ACT_LEAKY, None, None, K.ACT_NONE, None, None, None, 2, 30, 8, dy, None)
    assert max_abs(nchw(dy), gu * torch.where(z > 0, 1.0, 0.2)) < 1e-7


def test_gather_nhwc_rows(K):
    """ptk_gather_nhwc: rows assembled from several NCHW channel ranges in one pass, zeros in the gaps and in the channel
    padding, nothing written outside [c0, c0 + c_total); ragged pixel count (HW % 128 != 0)."""
    g = gen(31)
    N, H, W = 3, 19, 23
    a = torch.randn(N, 39, H, W, generator=g)
    b = torch.randn(N, 3, H, W, generator=g)
    dst = torch.full((N, H, W, 72), 7.0, device="cuda")
    segs = [(a.cuda(), 0, 21, 0), (a.cuda(), 21, 18, 24), (b.cuda(), 0, 3, 21)]
    K.gather_nhwc(segs, K.Slice(dst, 4, 64), 64)
    want = torch.zeros(N, H, W, 64)
    want[..., 0:21] = nhwc(a[:, 0:21]); want[..., 24:42] = nhwc(a[:, 21:39]); want[..., 21:24] = nhwc(b)
    assert torch.equal(dst[..., 4:68].cpu(), want)
    assert float((dst[..., :4] - 7.0).abs().max()) == 0 and float((dst[..., 68:] - 7.0).abs().max()) == 0
    # a hole (the generator's slot) and a single narrow segment
    dst2 = torch.full((N, H, W, 32), 7.0, device="cuda")
    K.gather_nhwc([(a.cuda(), 21, 18, 0)], K.Slice(dst2, 0, 32), 32)
    want2 = torch.zeros(N, H, W, 32)
    want2[..., :18] = nhwc(a[:, 21:39])
    assert torch.equal(dst2.cpu(), want2)


# ----------------------------------------------------------------------------- warp
def _run_warp(K, x, warps, masks, H0, W0, gy=None, act=0):
    N, C, h, w = x.shape
    Kp = warps.shape[1]
    mlv = torch.empty(N, h, w, Kp, device="cuda")
    K.mask_pyramid(masks.cuda().contiguous(), mlv)
    xd = nhwc(x).cuda()
    y = torch.empty(N, h, w, C, device="cuda")
    argk = torch.empty(N, h, w, C, dtype=torch.uint8, device="cuda")
    wr = warps.float().cuda().contiguous()
    K.warp_forward(xd, wr, mlv, y, argk, N, C, h, w, Kp, H0, W0, act)
    dx = None
    if gy is not None:
        dx = torch.zeros(N, h, w, C, device="cuda")
        K.warp_backward(nhwc(gy).cuda(), y, act, wr, mlv, argk, dx, N, C, h, w, Kp, H0, W0)
        dx = nchw(dx).cpu()
    return nchw(y).cpu(), dx, mlv


def test_warp_matches_reference_golden(K):
    from oracle.make_golden import WARP_CASES, warp_inputs
    for name, N, C, h, w, H0, W0, seed in WARP_CASES:
        x, warps, masks, gy = warp_inputs(N, C, h, w, H0, W0, seed)
        y, dx, _ = _run_warp(K, x, warps, masks, H0, W0, gy)
        gold = golden(name)
        # tolerance: fp32 coordinate rounding under +-30 px translations
        assert max_abs(y, gold["y"]) <= 2e-4, name
        assert max_abs(dx, gold["dx"]) <= 2e-4, name


def test_mask_pyramid_matches_half_pixel_bilinear(K):
    from oracle import synth
    b = synth.make_batch(2, 64, 32, 2, seed=3)
    for div in (1, 2, 4, 8):
        h, w = 64 // div, 32 // div
        out = torch.empty(2, h, w, 10, device="cuda")
        K.mask_pyramid(b["masks"].cuda(), out)
        ref = b["masks"] if div == 1 else F.interpolate(b["masks"], size=(h, w), mode="bilinear", align_corners=False)
        assert torch.equal(nchw(out).cpu(), ref.float())   # dyadic values: exact


def test_warp_full_size_vs_oracle_and_relu_epilogue(K):
    from oracle import restate, synth
    b = synth.make_batch(2, 256, 256, 2, seed=11)
    for C, h in ((64, 256), (128, 64), (512, 32)):
        x = torch.randn(2, C, h, h, generator=gen(C))
        gy = torch.randn(2, C, h, h, generator=gen(C + 1))
        xr = x.clone().requires_grad_(True)
        ref = F.relu(restate.affine_warp(xr, b["warps"], b["masks"], (256, 256)))
        ref.backward(gy)
        y, dx, _ = _run_warp(K, x, b["warps"], b["masks"], 256, 256, gy, act=K.ACT_RELU)
        assert max_abs(y, ref) <= 3e-4
        assert rel_l2(dx, xr.grad) <= 4e-3   # a handful of arg-max flips at near-ties + atomics order


@pytest.mark.parametrize("act", ["none", "relu", "leaky"])
@pytest.mark.parametrize("C,h,w,H0,W0", [(64, 37, 53, 74, 106), (128, 48, 24, 96, 48), (256, 19, 33, 152, 264), (512, 16, 16, 128, 128),
                                         (64, 64, 64, 64, 64), (1024, 9, 7, 72, 56)])
def test_warp_fast_path_vs_oracle(K, C, h, w, H0, W0, act):
    """The record-staged tile kernels (C = 64 / 128 / 256 / 512 / 1024), ragged extents (tile edges in x and y),
    non-square images (the H != W translation quirk), with the folded ReLU / LeakyReLU or none, forward and backward vs
    the oracle."""
    from oracle import restate, synth
    N = 2
    b = synth.make_batch(N, H0, W0, 2, seed=C + h)
    x = torch.randn(N, C, h, w, generator=gen(C + w))
    gy = torch.randn(N, C, h, w, generator=gen(C + w + 1))
    xr = x.clone().requires_grad_(True)
    ref = restate.affine_warp(xr, b["warps"], b["masks"], (H0, W0))
    if act == "relu":
        ref = F.relu(ref)
    elif act == "leaky":
        ref = F.leaky_relu(ref, 0.2)
    ref.backward(gy)
    y, dx, _ = _run_warp(K, x, b["warps"], b["masks"], H0, W0, gy, act={"relu": K.ACT_RELU, "leaky": K.ACT_LEAKY, "none": K.ACT_NONE}[act])
    assert max_abs(y, ref) <= 3e-4
    assert rel_l2(dx, xr.grad) <= 2e-3


@pytest.mark.parametrize("C,h", [(64, 40), (256, 20)])
def test_warp_more_active_parts_than_records(K, C, h):
    """Every pixel covered by ALL ten part masks (fractional values): more live candidates per pixel than the six
    shared-memory records of the tile kernels -> the inline slow path of the forward gather and the backward scatter."""
    from oracle import restate, synth
    N, H0 = 2, 80
    b = synth.make_batch(N, H0, H0, 2, seed=9)
    g = gen(77)
    masks = (torch.rand(N, 10, H0, H0, generator=g) * 0.75 + 0.25).double()
    x = torch.randn(N, C, h, h, generator=g)
    gy = torch.randn(N, C, h, h, generator=g)
    xr = x.clone().requires_grad_(True)
    ref = restate.affine_warp(xr, b["warps"], masks, (H0, H0))
    ref.backward(gy)
    y, dx, _ = _run_warp(K, x, b["warps"], masks, H0, H0, gy)
    assert max_abs(y, ref) <= 3e-4
    assert rel_l2(dx, xr.grad) <= 4e-3


def test_warp_levels_api_matches_single_level_calls(K, monkeypatch):
    """ptk_warp_forward_levels / _backward_levels (one launch for the 4 skip levels, writes into channel slices of wider
    buffers) == four single-level calls, bit for bit on the forward (twice: run-to-run identical)."""
    from oracle import synth
    N, H0 = 2, 64
    b = synth.make_batch(N, H0, H0, 2, seed=3)
    wr = b["warps"].float().cuda().contiguous()
    shapes = [(64, 64), (128, 32), (256, 16), (512, 8)]
    xs, mls, gys = [], [], []
    for C, h in shapes:
        xs.append(torch.randn(N, h, h, C, generator=gen(C)).cuda())
        gys.append(torch.randn(N, h, h, C + 32, generator=gen(C + 7)).cuda())
        ml = torch.empty(N, h, h, 10, device="cuda")
        K.mask_pyramid(b["masks"].cuda(), ml)
        mls.append(ml)
    single = []
    for (C, h), x, ml, gy in zip(shapes, xs, mls, gys):
        y = torch.zeros(N, h, h, C + 32, device="cuda")
        argk = torch.zeros(N, h, h, C, dtype=torch.uint8, device="cuda")
        K.warp_forward(x, wr, ml, K.Slice(y, 32, C), argk, N, C, h, h, 10, H0, H0, K.ACT_RELU)
        dx = torch.zeros(N, h, h, C, device="cuda")
        K.warp_backward(K.Slice(gy, 32, C), K.Slice(y, 32, C), K.ACT_RELU, wr, ml, argk, dx, N, C, h, h, 10, H0, H0)
        single.append((y, dx))
    for rep in range(2):
        lv = []
        for (C, h), x, ml, gy in zip(shapes, xs, mls, gys):
            lv.append(dict(x=K.Slice(x), mask=ml, y=K.Slice(torch.zeros(N, h, h, C + 32, device="cuda"), 32, C),
                           argk=torch.zeros(N, h, h, C, dtype=torch.uint8, device="cuda"), dy=K.Slice(gy, 32, C),
                           dx=torch.full((N, h, h, C), 7.0, device="cuda"), C=C, h=h, w=h))
        K.warp_forward_levels(lv, wr, N, 10, H0, H0, K.ACT_RELU)
        K.warp_backward_levels(lv, wr, N, 10, H0, H0, K.ACT_RELU, True)
        for d, (y, dx) in zip(lv, single):
            assert torch.equal(d["y"].t, y)
            assert rel_l2(d["dx"], dx) <= 1e-5      # fp32 atomics: order-dependent in the last bits


# ----------------------------------------------------------------------------- losses
def test_adv_loss(K):
    from oracle import restate
    g = gen(3)
    z = torch.randn(6, 49, generator=g) * 3
    zr = z.clone().requires_grad_(True)
    p = torch.sigmoid(zr)
    lt = restate.adv_true(p[:4]) * 0.25
    lf = restate.adv_fake(p[4:]) * 0.25
    (lt + lf).backward()
    loss = torch.zeros(2, device="cuda")
    d4 = torch.zeros(6 * 49, 4, device="cuda")
    K.adv_loss(z.cuda(), 6, 49, 4, 0.25, loss, d4, 4)
    assert abs(float(loss[0]) - float(lt)) < 1e-5 * abs(float(lt)) + 1e-6
    assert abs(float(loss[1]) - float(lf)) < 1e-5 * abs(float(lf)) + 1e-6
    assert rel_l2(d4[:, 0].reshape(6, 49), zr.grad) < 1e-5
    assert float(d4[:, 1:].abs().max()) == 0


def test_l1_loss(K):
    g = gen(4)
    a, b = torch.randn(2, 3, 9, 11, generator=g), torch.randn(2, 3, 9, 11, generator=g)
    ar = a.clone().requires_grad_(True)
    ref = (ar - b).abs().mean() * 100
    ref.backward()
    loss = torch.zeros(1, device="cuda")
    grad = torch.empty_like(a, device="cuda")
    K.l1_loss(a.cuda(), b.cuda(), 100.0, loss, grad)
    assert abs(float(loss) - float(ref)) < 1e-5 * float(ref)
    assert max_abs(grad, ar.grad) < 1e-7


@pytest.mark.parametrize("H,W,area", [(32, 48, 5), (20, 20, 3), (16, 16, 1), (37, 21, 5)])
def test_nnloss_fused_vgg(K, H, W, area):
    from oracle import restate, synth
    g = gen(H * 7 + area)
    pred = (torch.rand(2, 3, H, W, generator=g) * 2 - 1)
    tgt = (torch.rand(2, 3, H, W, generator=g) * 2 - 1)
    vw, vb = synth.vgg_conv1_1(1)
    pr = pred.clone().requires_grad_(True)
    ref = restate.nn_loss(restate.feature_extractor(vw, vb, pr), restate.feature_extractor(vw, vb, tgt), area, area) * 0.01
    ref.backward()
    loss = torch.zeros(1, device="cuda")
    argmin = torch.empty(2, H, W, dtype=torch.uint8, device="cuda")
    K.nnloss_forward(pred.cuda(), tgt.cuda(), vw.cuda(), vb.cuda(), area, 0.01, loss, argmin)
    assert abs(float(loss) - float(ref)) < 2e-5 * abs(float(ref))
    dpred = torch.empty(2, 3, H, W, device="cuda")
    K.nnloss_backward(pred.cuda(), tgt.cuda(), vw.cuda(), vb.cuda(), argmin, area, 0.01, dpred)
    # sign(gt - pred) flips on exact ties only; relu'(0) measure-zero
    assert rel_l2(dpred, pr.grad) < 1e-3


def test_tanh_bwd_combine(K):
    g = gen(8)
    out = torch.tanh(torch.randn(2, 3, 6, 5, generator=g))
    g1 = torch.randn(2, 3, 6, 5, generator=g)
    g2 = torch.randn(2, 6, 5, 44, generator=g)
    dz = torch.zeros(2, 6, 5, 4, device="cuda")
    K.tanh_bwd_combine(g1.cuda(), K.Slice(g2.cuda(), 21, 3), out.cuda(), dz, 4, 2, 3, 6, 5)
    ref = (g1 + nchw(g2[..., 21:24])) * (1 - out ** 2)
    assert max_abs(nchw(dz[..., :3]), ref) < 1e-6
    assert float(dz[..., 3].abs().max()) == 0


# ----------------------------------------------------------------------------- optimiser
def test_adam_matches_torch(K):
    g = gen(9)
    p0 = torch.randn(1001, generator=g)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=2e-4, betas=(0.5, 0.999))
    p = p0.clone().cuda()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        gr = torch.randn(1001, generator=g)
        ref.grad = gr.clone()
        opt.step()
        K.adam_step(p, gr.cuda(), m, v, 2e-4, 0.5, 0.999, 1e-8, step)
        assert max_abs(p, ref.detach()) < 1e-7
