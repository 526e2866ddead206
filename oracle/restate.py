"""TEST INFRASTRUCTURE ONLY.  Pure-torch fp32 restatement of the reference deformable-GAN step.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function here against
(a) ``tests/golden/*.npz`` produced by the unmodified reference modules (``oracle/make_golden.py``)
and (b) the reference itself when ``/root/reference`` is mounted, plus the reference's known answers
(parameter counts in src_deformable/logs/gen_full_fasion:158,193 and gen_full_h36m:136,171).

All citations are relative to /root/reference/src_deformable/.  Arithmetic the reference delegates
to torch (conv2d, conv_transpose2d, instance_norm, grid_sample ...) is restated with the closed form
where the reference composes several library calls (norm, warp, mask resize, losses) and with the
same torch CPU primitive where the reference calls exactly one (conv2d / conv_transpose2d).
"""
import torch
import torch.nn.functional as F

EPS_NORM = 1e-3  # networks.py:159


# ----------------------------------------------------------------------------- a1
def get_imgpose(inp, use_input_pose, pose_dim):
    """utils/pose_utils.py:227-233."""
    img = inp[:, :3]
    src = inp[:, 3:3 + pose_dim] if use_input_pose else None
    tgt = inp[:, (3 + pose_dim if use_input_pose else 6):]
    return img, src, tgt


# ----------------------------------------------------------------------------- a3
def block_norm(z, gamma, beta):
    """InstanceNorm3d(1, eps=1e-3, affine) on z.unsqueeze(1) (networks.py:159,164-169): per-sample
    mean / biased variance over C*H*W, scalar weight and bias."""
    n = z.shape[0]
    flat = z.reshape(n, -1)
    mean = flat.mean(dim=1).view(n, 1, 1, 1)
    var = flat.var(dim=1, unbiased=False).view(n, 1, 1, 1)
    return (z - mean) / torch.sqrt(var + EPS_NORM) * gamma + beta


def block_down(x, w, gamma=None, beta=None, pad=1):
    """Block(down=True): LeakyReLU(0.2) -> Conv k4 s2 p1 (no bias) -> norm (networks.py:148-160)."""
    z = F.conv2d(F.leaky_relu(x, 0.2), w, None, stride=2, padding=pad)
    return block_norm(z, gamma, beta) if gamma is not None else z


def block_up(x, w, gamma, beta, drop=None):
    """Block(down=False, leaky=False): ReLU -> ConvT k4 s2 p0 -> crop 1 -> norm -> Dropout2d
    (networks.py:152,156-161).  ConvT(p=0)+Cropping2D(1) == ConvT(p=1)."""
    z = F.conv_transpose2d(F.relu(x), w, None, stride=2, padding=1)
    y = block_norm(z, gamma, beta)
    if drop is not None:
        y = y * drop
    return y


# ----------------------------------------------------------------------------- a2
def encoder_forward(sd, prefix, x, levels):
    """encoder.forward (networks.py:193-202): returns every level's (pre-activation) output."""
    outs = [F.conv2d(x, sd[prefix + ".net.0.weight"], sd[prefix + ".net.0.bias"], padding=1)]
    for i in range(1, levels):
        w = sd["%s.net.%d.net.1.weight" % (prefix, i)]
        gk = "%s.net.%d.net.2.weight" % (prefix, i)
        if gk in sd:
            outs.append(block_down(outs[-1], w, sd[gk], sd["%s.net.%d.net.2.bias" % (prefix, i)]))
        else:
            outs.append(block_down(outs[-1], w))
    return outs


# ----------------------------------------------------------------------------- a6
def mask_pyramid_level(masks, h, w):
    """cv2.resize(INTER_LINEAR) of the [N,K,H0,W0] masks to (h,w) (utils/pose_transform.py:84)
    == half-pixel-centre bilinear (align_corners=False), then .float() (:87)."""
    if masks.shape[-2:] == (h, w):
        return masks.float()
    return F.interpolate(masks, size=(h, w), mode="bilinear", align_corners=False).float()


def normalized_theta(warps, h, w, H0, W0):
    """AffineTransformLayer.forward :72-76 then AffineLayer.normalize_transforms :48-58.
    The in-place assignments are sequential, so tx' sees the already-updated b' (and ty' sees c')."""
    a, b, tx, c, d, ty = [warps[..., i] for i in range(6)]
    tx = tx / (H0 / h)
    ty = ty / (W0 / w)
    b2 = b * w / h
    tx2 = tx * 2 / h + a + b2 - 1
    c2 = c * h / w
    ty2 = ty * 2 / w + c2 + d - 1
    return a, b2, tx2, c2, d, ty2


def affine_warp(x, warps, masks, init_size, align_corners=False):
    """AffineTransformLayer(K, init_size, 'mask')(x, warps, masks) (utils/pose_transform.py:16-92).
    Closed form of affine_grid + grid_sample(bilinear, zeros) for every part, mask multiply, max_k."""
    N, C, h, w = x.shape
    K = warps.shape[1]
    H0, W0 = init_size
    a, b2, tx2, c2, d, ty2 = [t.view(N, K, 1, 1).float() for t in normalized_theta(warps.float(), h, w, H0, W0)]
    jj = torch.arange(w, dtype=torch.float32).view(1, 1, 1, w)
    ii = torch.arange(h, dtype=torch.float32).view(1, 1, h, 1)
    if align_corners:
        gx = jj * 2 / max(w - 1, 1) - 1
        gy = ii * 2 / max(h - 1, 1) - 1
    else:
        gx = (2 * jj + 1) / w - 1
        gy = (2 * ii + 1) / h - 1
    sx = a * gx + b2 * gy + tx2
    sy = c2 * gx + d * gy + ty2
    if align_corners:
        px = (sx + 1) / 2 * (w - 1)
        py = (sy + 1) / 2 * (h - 1)
    else:
        px = ((sx + 1) * w - 1) / 2
        py = ((sy + 1) * h - 1) / 2
    x0 = torch.floor(px)
    y0 = torch.floor(py)
    fx = px - x0
    fy = py - y0
    m = mask_pyramid_level(masks, h, w)  # [N,K,h,w]
    out = None
    xf = x.reshape(N, C, h * w)
    for k in range(K):
        acc = torch.zeros_like(x)
        for dy, dx, wt in ((0, 0, (1 - fy[:, k]) * (1 - fx[:, k])), (0, 1, (1 - fy[:, k]) * fx[:, k]),
                           (1, 0, fy[:, k] * (1 - fx[:, k])), (1, 1, fy[:, k] * fx[:, k])):
            xi = x0[:, k] + dx
            yi = y0[:, k] + dy
            ok = (xi >= 0) & (xi <= w - 1) & (yi >= 0) & (yi <= h - 1)
            idx = (yi.clamp(0, h - 1) * w + xi.clamp(0, w - 1)).long().view(N, 1, h * w).expand(N, C, h * w)
            v = torch.gather(xf, 2, idx).view(N, C, h, w)
            acc = acc + v * (wt * ok.float()).view(N, 1, h, w)
        acc = acc * m[:, k].view(N, 1, h, w)
        out = acc if out is None else torch.maximum(out, acc)
    return out


# ----------------------------------------------------------------------------- a4 / a5
def generator_forward(sd, inp, warps, masks, image_size, pose_dim, drop=None):
    """Deformable_Generator.forward (networks.py:269-288) + decoder.forward (:236-250).
    `drop` = list of three [N,512,1,1] Dropout2d noise tensors (values {0,2}) or None for no dropout."""
    levels = 7 if max(image_size) >= 256 else 6
    img, src, tgt = get_imgpose(inp, True, pose_dim)
    skips_app = encoder_forward(sd, "encoder_app", torch.cat([img, src], 1), levels)
    skips_pose = encoder_forward(sd, "encoder_pose", tgt, levels)
    skips = []
    for i, (sa, sp) in enumerate(zip(skips_app, skips_pose)):
        if i < 4:
            sa = affine_warp(sa, warps, masks, image_size)
        skips.append(torch.cat([sa, sp], 1))
    out = None
    for j in range(levels - 1):
        x = skips[-1] if j == 0 else torch.cat([out, skips[-(j + 1)]], 1)
        out = block_up(x, sd["decoder.net.%d.net.1.weight" % j], sd["decoder.net.%d.net.3.weight" % j],
                       sd["decoder.net.%d.net.3.bias" % j], drop[j] if (drop is not None and j < 3) else None)
    x = torch.cat([out, skips[0]], 1)
    k = levels
    return torch.tanh(F.conv2d(F.relu(x), sd["decoder.net.%d.weight" % k], sd["decoder.net.%d.bias" % k], padding=1))


# ----------------------------------------------------------------------------- a7
def discriminator_forward(sd, x):
    """Discriminator.forward (networks.py:338-357): Conv k4 s2 p0 + bias, 3 normed Blocks, Block(512,1,
    bn=False), Sigmoid, Flatten."""
    y = F.conv2d(x, sd["net.0.weight"], sd["net.0.bias"], stride=2)
    for i in (1, 2, 3):
        y = block_down(y, sd["net.%d.net.1.weight" % i], sd["net.%d.net.2.weight" % i], sd["net.%d.net.2.bias" % i])
    y = block_down(y, sd["net.4.net.1.weight"])
    return torch.sigmoid(y).reshape(y.shape[0], -1)


# ----------------------------------------------------------------------------- a10
VGG_MEAN = (0.485, 0.456, 0.406)
VGG_STD = (0.229, 0.224, 0.225)


def get_layer_ind(layer_name):
    """utils/pose_utils.py:312-317 ('block1_conv2' -> 1)."""
    block, conv = layer_name.split("_")
    return int(["0", "5", "10", "19", "28"][int(block[-1]) - 1]) + int(conv[-1]) - 1


def vgg_preprocess(x):
    """utils/pose_utils.py:324-331: NCHW memory re-viewed as NHWC (no permute) => element with flat
    per-sample index i is normalised with mean[i % 3], std[i % 3]."""
    N, C, H, W = x.shape
    i = torch.arange(C * H * W) % 3
    mean = torch.tensor(VGG_MEAN)[i].view(1, C, H, W)
    std = torch.tensor(VGG_STD)[i].view(1, C, H, W)
    return (x - mean) / std


def feature_extractor(vgg_w, vgg_b, x):
    """Feature_Extractor(vgg, x, 'block1_conv2') == relu(conv1_1(preprocess(x))) (pose_utils.py:320-338)."""
    return F.relu(F.conv2d(vgg_preprocess(x), vgg_w, vgg_b, padding=1))


VGG19_CFG = (64, 64, 'M', 128, 128, 'M', 256, 256, 256, 256, 'M', 512, 512, 512, 512, 'M', 512, 512, 512, 512, 'M')


def feature_extractor_prefix(convs, x, layer_ind):
    """Feature_Extractor for ANY content_loss_layer (pose_utils.py:320-338): torchvision vgg19.features[0..layer_ind]
    (configuration 'E': 3x3/pad-1 convs each followed by ReLU, 'M' = MaxPool2d(2, 2)) on the view-normalised image.
    convs: [(weight, bias)] of the VGG convolutions in order (as many as the prefix needs)."""
    x = vgg_preprocess(x)
    it, ci = 0, 0
    for v in VGG19_CFG:
        if it > layer_ind:
            break
        if v == 'M':
            x = F.max_pool2d(x, 2, 2)
            it += 1
            continue
        w, b = convs[ci]
        ci += 1
        x = F.conv2d(x, w, b, padding=1)
        it += 1
        if it > layer_ind:
            break
        x = F.relu(x)
        it += 1
    return x


# ----------------------------------------------------------------------------- a11
def nn_loss(pred, gt, nh, nw):
    """DeformablePose_GAN.nn_loss (models/pose_gan.py:173-199), without materialising the 25x stack."""
    assert nh == nw, "reference only works for square windows (ConstantPad2d argument order)"
    p = nh // 2
    padded = F.pad(gt, (p, p, p, p), value=-10000.0)
    H, W = pred.shape[2:]
    best = None
    for i in range(nh):
        for j in range(nw):
            d = (padded[:, :, i:i + H, j:j + W] - pred).abs().sum(dim=1)
            best = d if best is None else torch.minimum(best, d)
    return best.mean()


# ----------------------------------------------------------------------------- a8 / a9
def adv_true(out):
    """sum_n -mean_j log(out[n,j] + 1e-7) (pose_gan.py:90-98,140-151)."""
    return -(torch.log(out + 1e-7).mean(dim=1)).sum()


def adv_fake(out):
    """sum_n -mean_j log(1 - out[n,j] + 1e-7) (pose_gan.py:152-160)."""
    return -(torch.log(1 - out + 1e-7).mean(dim=1)).sum()


class OracleGAN:
    """Restatement of DeformablePose_GAN (models/pose_gan.py:11-171) for gen_type='baseline'.

    Holds leaf tensors for G and D (keys == reference state_dict keys) and two torch Adam optimisers
    (lr, betas=(0.5,0.999), pose_gan.py:49-51).  Like the reference it does NOT detach out_gen in
    dis_update (pose_gan.py:129,166) -- the wasted G backward is part of the reference's cost."""

    def __init__(self, gen_sd, disc_sd, vgg_w, vgg_b, image_size, pose_dim, batch_size, lr=2e-4,
                 content_loss_layer="block1_conv2", nn_loss_area_size=5, faithful_waste=True):
        self.gen = {k: v.clone().requires_grad_(True) for k, v in gen_sd.items()}
        self.disc = {k: v.clone().requires_grad_(True) for k, v in disc_sd.items()}
        self.vgg_w, self.vgg_b = vgg_w, vgg_b
        self.image_size, self.pose_dim, self.batch_size = tuple(image_size), pose_dim, batch_size
        self.content_loss_layer, self.area = content_loss_layer, nn_loss_area_size
        self.gen_opt = torch.optim.Adam(list(self.gen.values()), lr=lr, betas=(0.5, 0.999))
        self.disc_opt = torch.optim.Adam(list(self.disc.values()), lr=lr, betas=(0.5, 0.999))
        self.faithful_waste = faithful_waste

    def gen_forward(self, inp, warps, masks, drop=None):
        return generator_forward(self.gen, inp, warps, masks, self.image_size, self.pose_dim, drop)

    def dis_update(self, inp, target, warps, masks, real_inp, real_target, gan_w=1.0, drop=None):
        for p in self.disc.values():
            p.grad = None
        out_gen = self.gen_forward(inp, warps, masks, drop)
        if not self.faithful_waste:
            out_gen = out_gen.detach()
        img, src, tgt = get_imgpose(inp, True, self.pose_dim)
        fake = torch.cat([img, src, out_gen, tgt], 1)
        rimg, rsrc, rtgt = get_imgpose(real_inp, True, self.pose_dim)
        real = torch.cat([rimg, rsrc, real_target, rtgt], 1)
        res = discriminator_forward(self.disc, torch.cat([real, fake], 0))
        B = self.batch_size
        true_l = adv_true(res[:B]) * gan_w / B
        fake_l = adv_fake(res[B:]) * gan_w / B
        loss = true_l + fake_l
        loss.backward()
        self.disc_opt.step()
        return [loss.item(), true_l.item(), fake_l.item()]

    def gen_update(self, inp, target, warps, masks, gan_w=1.0, l1_w=100.0, drop=None):
        for p in self.gen.values():
            p.grad = None
        out_gen = self.gen_forward(inp, warps, masks, drop)
        img, src, tgt = get_imgpose(inp, True, self.pose_dim)
        out_dis = discriminator_forward(self.disc, torch.cat([img, src, out_gen, tgt], 1))
        ad = adv_true(out_dis)
        if self.content_loss_layer != "none":
            ll = nn_loss(feature_extractor(self.vgg_w, self.vgg_b, out_gen),
                         feature_extractor(self.vgg_w, self.vgg_b, target), self.area, self.area)
        else:
            ll = (out_gen - target).abs().mean()
        ad = ad * gan_w / self.batch_size
        ll = ll * l1_w
        total = ad + ll
        total.backward()
        self.gen_opt.step()
        return out_gen.detach(), [total.item(), ll.item(), ad.item()]
