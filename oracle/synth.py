"""TEST INFRASTRUCTURE ONLY.  Seeded synthetic inputs and weights.

The same recipe feeds the reference (when generating ``tests/golden``), the CPU restatement, the CUDA
product path in the parity tests, and ``bench.py``.  Everything is produced with a private CPU
``torch.Generator`` so the bytes do not depend on the device, on global RNG state or on module
construction order.

Value distributions follow what the reference's data pipeline produces:
  * images in [-1, 1]                       (utils/pose_utils.py:216-217  _preprocess_image)
  * pose heat-maps exp(-d^2 / (2*6^2))      (utils/pose_utils.py:79-86    cords_to_map, sigma=6)
  * warps[N,10,8] = first 8 entries of a row-major 3x3 affine; missing parts carry ``no_point_tr``
    = [[1,0,1000],[0,1,1000],[0,0,1]]       (utils/pose_transform.py:221,289)
  * masks[N,10,H,W] float64, part 0 all ones (utils/pose_transform.py:149-150), other parts binary
"""
import math

import torch


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def make_batch(N, H, W, P, seed=0, K=10):
    """Returns dict(input[N,3+2P,H,W] f32, target[N,3,H,W] f32, warps[N,K,8] f32, masks[N,K,H,W] f64)."""
    g = _gen(1000003 * (seed + 1) + 17 * N + H + 3 * W + 7 * P)
    img = torch.rand(N, 3, H, W, generator=g) * 2 - 1
    target = torch.rand(N, 3, H, W, generator=g) * 2 - 1
    yy = torch.arange(H, dtype=torch.float32).view(1, 1, H, 1)
    xx = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W)
    ky = torch.rand(N, 2 * P, 1, 1, generator=g) * H
    kx = torch.rand(N, 2 * P, 1, 1, generator=g) * W
    pose = torch.exp(-((yy - ky) ** 2 + (xx - kx) ** 2) / (2 * 6.0 ** 2))
    missing = torch.rand(N, 2 * P, 1, 1, generator=g) < 0.1
    pose = torch.where(missing, torch.zeros_like(pose), pose)
    inp = torch.cat([img, pose], dim=1).contiguous()

    A = torch.eye(2).view(1, 1, 2, 2) + 0.2 * torch.randn(N, K, 2, 2, generator=g)
    t = 0.12 * torch.randn(N, K, 2, generator=g) * torch.tensor([float(H), float(W)])
    warps = torch.zeros(N, K, 8)
    warps[..., 0] = A[..., 0, 0]
    warps[..., 1] = A[..., 0, 1]
    warps[..., 2] = t[..., 0]
    warps[..., 3] = A[..., 1, 0]
    warps[..., 4] = A[..., 1, 1]
    warps[..., 5] = t[..., 1]
    absent = torch.rand(N, K, generator=g) < 0.2
    absent[:, 0] = False
    no_point = torch.tensor([1.0, 0.0, 1000.0, 0.0, 1.0, 1000.0, 0.0, 0.0])
    warps = torch.where(absent.unsqueeze(-1), no_point.view(1, 1, 8), warps)

    masks = torch.zeros(N, K, H, W, dtype=torch.float64)
    masks[:, 0] = 1.0
    yyd = torch.arange(H, dtype=torch.float64).view(H, 1)
    xxd = torch.arange(W, dtype=torch.float64).view(1, W)
    for n in range(N):
        for k in range(1, K):
            cy = float(torch.rand(1, generator=g)) * H
            cx = float(torch.rand(1, generator=g)) * W
            area = (0.02 + 0.13 * float(torch.rand(1, generator=g))) * H * W
            aspect = 0.4 + 1.6 * float(torch.rand(1, generator=g))
            hh = math.sqrt(area * aspect) / 2
            ww = math.sqrt(area / aspect) / 2
            th = float(torch.rand(1, generator=g)) * math.pi
            if bool(absent[n, k]):
                continue
            u = (yyd - cy) * math.cos(th) + (xxd - cx) * math.sin(th)
            v = -(yyd - cy) * math.sin(th) + (xxd - cx) * math.cos(th)
            masks[n, k] = ((u.abs() <= hh) & (v.abs() <= ww)).to(torch.float64)
    return {"input": inp, "target": target, "warps": warps, "masks": masks}


def make_keypoints(N, H, W, P, seed=0, missing=0.1):
    """[N, P, 2] int64 (y, x; -1 = missing): a plausible upright skeleton per sample for the 18-joint (OpenPose / PAF) or the
    16-joint (stacked-hourglass) label set of the reference (utils/pose_utils.py:27,36-37).  The four torso joints are
    always present (the reference's pose geometry dereferences them unconditionally)."""
    g = _gen(2000003 * (seed + 1) + 31 * N + H + 5 * W + 11 * P)
    names18 = ['nose', 'neck', 'Rsho', 'Relb', 'Rwri', 'Lsho', 'Lelb', 'Lwri', 'Rhip', 'Rkne', 'Rank', 'Lhip', 'Lkne', 'Lank',
               'Leye', 'Reye', 'Lear', 'Rear']
    names16 = ['Rank', 'Rknee', 'Rhip', 'Lhip', 'Lknee', 'Lank', 'pelv', 'spine', 'neck', 'head', 'Rwri', 'Relb', 'Rsho', 'Lsho',
               'Lelb', 'Lwri']
    names = names18 if P == 18 else names16
    out = torch.full((N, P, 2), -1, dtype=torch.int64)

    def rnd(a=1.0):
        return float(torch.rand(1, generator=g)) * a

    for n in range(N):
        s = H * (0.16 + 0.08 * rnd())                        # shoulder-to-hip length
        cy, cx = H * (0.30 + 0.1 * rnd()), W * (0.35 + 0.3 * rnd())
        lean = (rnd() - 0.5) * 0.5
        pts = {}
        pts['neck'] = (cy - 0.15 * s, cx)
        pts['Rsho'] = (cy, cx - 0.45 * s)
        pts['Lsho'] = (cy, cx + 0.45 * s)
        pts['Rhip'] = (cy + s, cx - 0.3 * s + lean * s)
        pts['Lhip'] = (cy + s, cx + 0.3 * s + lean * s)
        pts['pelv'] = (cy + s, cx + lean * s)
        pts['spine'] = (cy + 0.5 * s, cx + 0.5 * lean * s)
        pts['nose'] = pts['head'] = (cy - 0.55 * s, cx + (rnd() - 0.5) * 0.2 * s)
        pts['Reye'] = (cy - 0.62 * s, cx - 0.1 * s)
        pts['Leye'] = (cy - 0.62 * s, cx + 0.1 * s)
        pts['Rear'] = (cy - 0.58 * s, cx - 0.22 * s)
        pts['Lear'] = (cy - 0.58 * s, cx + 0.22 * s)
        for side, sx in (('R', -1), ('L', 1)):
            a1, a2 = (rnd() - 0.5) * 1.6, (rnd() - 0.5) * 1.6
            sy, sxx = pts[side + 'sho']
            ey, ex = sy + 0.6 * s * math.cos(a1), sxx + sx * 0.15 * s + 0.6 * s * math.sin(a1)
            wy, wx = ey + 0.55 * s * math.cos(a1 + a2), ex + 0.55 * s * math.sin(a1 + a2)
            pts[side + 'elb'], pts[side + 'wri'] = (ey, ex), (wy, wx)
            b1, b2 = (rnd() - 0.5) * 0.8, (rnd() - 0.5) * 0.6
            hy, hx = pts[side + 'hip']
            ky, kx = hy + 0.85 * s * math.cos(b1), hx + 0.85 * s * math.sin(b1)
            ay, ax = ky + 0.8 * s * math.cos(b1 + b2), kx + 0.8 * s * math.sin(b1 + b2)
            pts[side + 'kne'] = pts[side + 'knee'] = (ky, kx)
            pts[side + 'ank'] = (ay, ax)
        for i, name in enumerate(names):
            y, x = pts[name]
            keep = name in ('Rhip', 'Lhip', 'Rsho', 'Lsho') or rnd() >= missing
            if keep and 0 <= y < H and 0 <= x < W:
                out[n, i, 0], out[n, i, 1] = int(y), int(x)
            elif name in ('Rhip', 'Lhip', 'Rsho', 'Lsho'):
                out[n, i, 0], out[n, i, 1] = int(min(max(y, 0), H - 1)), int(min(max(x, 0), W - 1))
    return out


def fill_state_dict(shapes, seed=0):
    """Deterministic weights for a {key: shape} mapping (keys visited in sorted order).

    4-D tensors ~ U(-b, b) with b = sqrt(3 / (shape[1]*kh*kw)) (unit-gain fan-in scaling keeps the
    21-layer U-Net's activations O(1)); 1-element norm scalars: weight ~ U(0.5, 1.5),
    bias ~ U(-0.2, 0.2); other 1-D biases ~ U(-0.05, 0.05).
    """
    g = _gen(7919 * (seed + 1))
    out = {}
    for key in sorted(shapes):
        shape = tuple(shapes[key])
        if len(shape) == 4:
            b = math.sqrt(3.0 / (shape[1] * shape[2] * shape[3]))
            out[key] = (torch.rand(shape, generator=g) * 2 - 1) * b
        elif shape == (1,) and key.endswith("weight"):
            out[key] = torch.rand(shape, generator=g) + 0.5
        elif shape == (1,):
            out[key] = torch.rand(shape, generator=g) * 0.4 - 0.2
        else:
            out[key] = torch.rand(shape, generator=g) * 0.1 - 0.05
    return out


def vgg_conv1_1(seed=0):
    """Seeded stand-in for torchvision vgg19.features[0] (no ImageNet weights offline)."""
    g = _gen(104729 * (seed + 1))
    w = (torch.rand(64, 3, 3, 3, generator=g) * 2 - 1) * math.sqrt(3.0 / 27)
    b = torch.rand(64, generator=g) * 0.2 - 0.1
    return w, b


def fill_vgg(vgg, seed=0):
    """Seeded weights for EVERY conv of a torchvision vgg19 (deeper content_loss_layer fixtures): features[0] keeps
    vgg_conv1_1(seed); the others get uniform(+-sqrt(6 / fan_in)) so that activations keep their scale through ReLUs."""
    import torch.nn as nn
    w0, b0 = vgg_conv1_1(seed)
    k = 0
    with torch.no_grad():
        for m in vgg.features:
            if not isinstance(m, nn.Conv2d):
                continue
            if k == 0:
                m.weight.copy_(w0)
                m.bias.copy_(b0)
            else:
                g = _gen(7919 * (seed + 1) + 31 * k)
                fan_in = m.weight.shape[1] * 9
                m.weight.copy_((torch.rand(m.weight.shape, generator=g) * 2 - 1) * math.sqrt(6.0 / fan_in))
                m.bias.copy_(torch.rand(m.bias.shape, generator=g) * 0.2 - 0.1)
            k += 1
    return vgg


def dropout_masks(N, C, count, seed=0):
    """`count` Dropout2d(0.5) noise tensors [N,C,1,1] with values in {0, 2} (networks.py:161)."""
    g = _gen(15485863 * (seed + 1))
    return [(torch.rand(N, C, 1, 1, generator=g) < 0.5).float() * 2.0 for _ in range(count)]


def generator_shapes(P, image_size, use_input_pose=True):
    """state_dict key -> shape of the reference Deformable_Generator (networks.py:175-266)."""
    big = max(image_size) >= 256
    enc = (64, 128, 256, 512, 512, 512, 512) if big else (64, 128, 256, 512, 512, 512)
    dec = (512, 512, 512, 512, 256, 128, 3) if big else (512, 512, 512, 256, 128, 3)
    input_nc = 3 + 2 * P if use_input_pose else 3 + P
    shapes = {}
    for name, cin in (("encoder_app", input_nc - P), ("encoder_pose", P)):
        shapes["%s.net.0.weight" % name] = (enc[0], cin, 3, 3)
        shapes["%s.net.0.bias" % name] = (enc[0],)
        for i in range(1, len(enc)):
            shapes["%s.net.%d.net.1.weight" % (name, i)] = (enc[i], enc[i - 1], 4, 4)
            if i != len(enc) - 1:
                shapes["%s.net.%d.net.2.weight" % (name, i)] = (1,)
                shapes["%s.net.%d.net.2.bias" % (name, i)] = (1,)
    for i, nf in enumerate(dec):
        cin = 2 * enc[-1] if i == 0 else 2 * enc[-(i + 1)] + dec[i - 1]
        if i == len(dec) - 1:
            shapes["decoder.net.%d.weight" % (i + 1)] = (nf, cin, 3, 3)
            shapes["decoder.net.%d.bias" % (i + 1)] = (nf,)
        else:
            shapes["decoder.net.%d.net.1.weight" % i] = (cin, nf, 4, 4)
            shapes["decoder.net.%d.net.3.weight" % i] = (1,)
            shapes["decoder.net.%d.net.3.bias" % i] = (1,)
    return shapes


def discriminator_shapes(input_nc):
    """state_dict key -> shape of the reference Discriminator (networks.py:338-353)."""
    shapes = {"net.0.weight": (64, input_nc, 4, 4), "net.0.bias": (64,)}
    chans = (64, 128, 256, 512)
    for i in range(1, 4):
        shapes["net.%d.net.1.weight" % i] = (chans[i], chans[i - 1], 4, 4)
        shapes["net.%d.net.2.weight" % i] = (1,)
        shapes["net.%d.net.2.bias" % i] = (1,)
    shapes["net.4.net.1.weight"] = (1, 512, 4, 4)
    return shapes
