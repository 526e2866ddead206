"""TEST INFRASTRUCTURE ONLY.  Generate tests/golden/*.npz from the UNMODIFIED reference modules.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
Inputs and weights are never stored: they are regenerated from ``oracle/synth.py`` seeds by the
tests; only reference OUTPUTS (and compact summaries of big gradient sets) are committed.
"""
import argparse
import os

import numpy as np
import torch

from . import ref_import, synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

WARP_CASES = [
    # name, N, C, h, w, H0, W0, seed
    ("warp_sq_full", 2, 8, 32, 32, 32, 32, 1),
    ("warp_sq_div4", 2, 16, 16, 16, 64, 64, 2),
    ("warp_nonsq_div2", 3, 8, 64, 32, 128, 64, 3),   # H != W quirk (pose_transform.py:52,56,72-73)
    ("warp_224_div8", 2, 4, 28, 28, 224, 224, 4),
]


def sample_idx(numel, count=24):
    return (np.arange(count, dtype=np.int64) * 2654435761 % max(numel, 1)).astype(np.int64)


def summarize(t, count=24):
    f = t.detach().reshape(-1).double()
    idx = sample_idx(f.numel(), count)
    return np.concatenate([[f.norm().item(), f.sum().item()], f[torch.from_numpy(idx)].numpy()])


def warp_inputs(N, C, h, w, H0, W0, seed):
    b = synth.make_batch(N, H0, W0, 2, seed=seed)
    g = torch.Generator().manual_seed(4242 + seed)
    x = torch.randn(N, C, h, w, generator=g)
    gy = torch.randn(N, C, h, w, generator=g)
    return x, b["warps"], b["masks"], gy


def gen_warp(ns):
    for name, N, C, h, w, H0, W0, seed in WARP_CASES:
        x, warps, masks, gy = warp_inputs(N, C, h, w, H0, W0, seed)
        x = x.clone().requires_grad_(True)
        layer = ns.pose_transform.AffineTransformLayer(10, (H0, W0), "mask")
        y = layer(x, warps.clone(), masks.clone())
        y.backward(gy)
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), y=y.detach().numpy(), dx=x.grad.numpy())
        print(name, tuple(y.shape), float(y.abs().mean()))


class _DropPatch:
    """Feed explicit Dropout2d noise to the reference (it draws from global RNG, networks.py:161)."""

    def __init__(self, masks):
        self.masks = list(masks)
        self.i = 0

    def __enter__(self):
        self.old = torch.nn.Dropout2d.forward
        patch = self

        def fwd(mod, x):
            m = patch.masks[patch.i % len(patch.masks)]
            patch.i += 1
            return x * m
        torch.nn.Dropout2d.forward = fwd
        return self

    def __exit__(self, *a):
        torch.nn.Dropout2d.forward = self.old


def gen_networks(ns, H, W, P, N, seed, tag):
    image_size = (H, W)
    big = max(image_size) >= 256
    enc = (64, 128, 256, 512, 512, 512, 512) if big else (64, 128, 256, 512, 512, 512)
    dec = (512, 512, 512, 512, 256, 128, 3) if big else (512, 512, 512, 256, 128, 3)
    G = ns.networks.Deformable_Generator(3 + 2 * P, P, image_size, enc, dec, "mask")
    D = ns.networks.Discriminator(3 + 2 * P + 3)
    gshapes = {k: tuple(v.shape) for k, v in G.state_dict().items()}
    dshapes = {k: tuple(v.shape) for k, v in D.state_dict().items()}
    assert gshapes == synth.generator_shapes(P, image_size), "generator_shapes() drifted from reference"
    assert dshapes == synth.discriminator_shapes(3 + 2 * P + 3)
    G.load_state_dict(synth.fill_state_dict(gshapes, seed))
    D.load_state_dict(synth.fill_state_dict(dshapes, seed + 1))
    b = synth.make_batch(N, H, W, P, seed=seed)
    drop = synth.dropout_masks(N, 512, 3, seed=seed)
    with _DropPatch(drop):
        out = G(b["input"], b["warps"].clone(), b["masks"].clone())
    img, src, tgt = ns.pose_utils.get_imgpose(b["input"], True, P)
    d_in = torch.cat([img, src, out.detach(), tgt], 1)
    d_out = D(d_in)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "net_%s.npz" % tag), out_gen=out.detach().numpy(),
                        d_out=d_out.detach().numpy(),
                        n_params_g=np.int64(sum(p.numel() for p in G.parameters())),
                        n_params_d=np.int64(sum(p.numel() for p in D.parameters())))
    print("net", tag, tuple(out.shape), tuple(d_out.shape), float(out.abs().mean()), float(d_out.mean()))


def gen_step(ns, H, W, P, N, seed, tag, content="block1_conv2", area=5, l1_w=0.01, steps=2):
    import argparse as ap
    import torchvision
    image_size = (H, W)
    opt = ap.Namespace(image_size=image_size, use_input_pose=True, pose_dim=P, batch_size=N, num_stacks=4,
                       gen_type="baseline", warp_skip="mask", dataset="fasion", learning_rate=2e-4,
                       content_loss_layer=content, nn_loss_area_size=area, gan_penalty_weight=1.0,
                       l1_penalty_weight=l1_w)
    dsd = synth.fill_state_dict(synth.discriminator_shapes(3 + 2 * P + 3), seed + 1)
    vgg = torchvision.models.vgg19(weights=None)
    if content in ("none", "block1_conv2"):
        vw, vb = synth.vgg_conv1_1(seed)
        with torch.no_grad():
            vgg.features[0].weight.copy_(vw)
            vgg.features[0].bias.copy_(vb)
    else:
        synth.fill_vgg(vgg, seed)
    model = ref_import.make_reference_gan(opt, dsd, vgg)
    model.gen.load_state_dict(synth.fill_state_dict(synth.generator_shapes(P, image_size), seed))
    od = vars(opt)
    rec = {}
    for s in range(steps):
        b = synth.make_batch(N, H, W, P, seed=seed + 10 * s)
        r = synth.make_batch(N, H, W, P, seed=seed + 10 * s + 1)
        b2 = synth.make_batch(N, H, W, P, seed=seed + 10 * s + 2)
        drop_d = synth.dropout_masks(N, 512, 3, seed=seed + 10 * s)
        drop_g = synth.dropout_masks(N, 512, 3, seed=seed + 10 * s + 2)
        with _DropPatch(drop_d):
            dl = model.dis_update(b["input"], b["target"], {"warps": b["warps"].clone(), "masks": b["masks"].clone()},
                                  r["input"], r["target"], od)
        rec["d_loss_%d" % s] = np.array(dl)
        rec["d_grad_%d" % s] = np.stack([summarize(p.grad) for _, p in sorted(model.disc.named_parameters())])
        with _DropPatch(drop_g):
            out, _, gl = model.gen_update(b2["input"], b2["target"], {"warps": b2["warps"].clone(), "masks": b2["masks"].clone()}, od)
        rec["g_loss_%d" % s] = np.array(gl)
        rec["out_gen_%d" % s] = out.detach().numpy()
        rec["g_grad_%d" % s] = np.stack([summarize(p.grad) for _, p in sorted(model.gen.named_parameters())])
        rec["g_param_%d" % s] = np.stack([summarize(p) for _, p in sorted(model.gen.named_parameters())])
        rec["d_param_%d" % s] = np.stack([summarize(p) for _, p in sorted(model.disc.named_parameters())])
        print("step", tag, s, dl, gl)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "step_%s.npz" % tag), **rec)


BIG_SAMPLES = 16384     # elements of every gradient tensor kept by the benchmark-size step fixtures
BIG_CASES = [
    # tag, H, W, P, N, seed, out_gen pixel stride   (BASELINE.json configs[1], [2], [4] geometry at batch 2)
    ("256x256_p18_n2", 256, 256, 18, 2, 21, 2),
    ("224x224_p16_n2", 224, 224, 16, 2, 22, 2),
    ("512x512_p18_n2", 512, 512, 18, 2, 23, 4),
]


def big_sample_idx(numel, count=BIG_SAMPLES):
    """Fixed pseudo-random element subset of a tensor (all of it when it is small)."""
    if numel <= count:
        return np.arange(numel, dtype=np.int64)
    return ((np.arange(count, dtype=np.int64) * 2654435761 + 12345) % numel).astype(np.int64)


def big_summary(t):
    """(norm, sampled values) of a tensor: enough to check direction (cosine on the sample) and magnitude."""
    f = t.detach().reshape(-1).float().cpu()
    idx = torch.from_numpy(big_sample_idx(f.numel()))
    return np.float64(f.double().norm().item()), f[idx].numpy().astype(np.float32)


def gen_step_big(ns, tag, H, W, P, N, seed, stride):
    """One dis_update + gen_update of the UNMODIFIED reference at a benchmarked geometry (NN loss 5x5 on block1_conv2,
    the bench.py objective).  Records losses, out_gen on a pixel lattice, and for every parameter the gradient norm plus
    a fixed 16k-element sample of the gradient and of the Adam-updated parameter."""
    import argparse as ap
    import torchvision
    opt = ap.Namespace(image_size=(H, W), use_input_pose=True, pose_dim=P, batch_size=N, num_stacks=4,
                       gen_type="baseline", warp_skip="mask", dataset="fasion", learning_rate=2e-4,
                       content_loss_layer="block1_conv2", nn_loss_area_size=5, gan_penalty_weight=1.0,
                       l1_penalty_weight=0.01)
    dsd = synth.fill_state_dict(synth.discriminator_shapes(3 + 2 * P + 3), seed + 1)
    vgg = torchvision.models.vgg19(weights=None)
    vw, vb = synth.vgg_conv1_1(seed)
    with torch.no_grad():
        vgg.features[0].weight.copy_(vw)
        vgg.features[0].bias.copy_(vb)
    model = ref_import.make_reference_gan(opt, dsd, vgg)
    model.gen.load_state_dict(synth.fill_state_dict(synth.generator_shapes(P, (H, W)), seed))
    od = vars(opt)
    b = synth.make_batch(N, H, W, P, seed=seed)
    r = synth.make_batch(N, H, W, P, seed=seed + 1)
    b2 = synth.make_batch(N, H, W, P, seed=seed + 2)
    rec = {"seed": np.int64(seed), "stride": np.int64(stride)}
    with _DropPatch(synth.dropout_masks(N, 512, 3, seed=seed)):
        dl = model.dis_update(b["input"], b["target"], {"warps": b["warps"].clone(), "masks": b["masks"].clone()},
                              r["input"], r["target"], od)
    rec["d_loss"] = np.array(dl)
    dn, ds = zip(*[big_summary(p.grad) for _, p in sorted(model.disc.named_parameters())])
    rec["d_grad_norm"] = np.array(dn)
    for i, v in enumerate(ds):
        rec["d_grad_%02d" % i] = v
    with _DropPatch(synth.dropout_masks(N, 512, 3, seed=seed + 2)):
        out, _, gl = model.gen_update(b2["input"], b2["target"], {"warps": b2["warps"].clone(), "masks": b2["masks"].clone()}, od)
    rec["g_loss"] = np.array(gl)
    rec["out_gen"] = out.detach()[:, :, ::stride, ::stride].contiguous().numpy()
    rec["out_gen_norm"] = np.float64(out.detach().double().norm().item())
    gn, gs = zip(*[big_summary(p.grad) for _, p in sorted(model.gen.named_parameters())])
    rec["g_grad_norm"] = np.array(gn)
    for i, v in enumerate(gs):
        rec["g_grad_%02d" % i] = v
    rec["g_param"] = np.stack([summarize(p) for _, p in sorted(model.gen.named_parameters())])
    rec["d_param"] = np.stack([summarize(p) for _, p in sorted(model.disc.named_parameters())])
    print("big step", tag, dl, gl)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "step_%s.npz" % tag), **rec)


STACKED_CASE = ("stacked_64x64_p18_s2", 64, 64, 18, 2, 2, 41)     # tag, H, W, P, N, num_stacks, seed


def stacked_inputs(H, W, P, N, S, seed):
    """(input, target, interpol_pose [N,S*P,H,W], interpol_warps [N,S,10,8], interpol_masks [N,S,10,H,W]) from S batches."""
    bs = [synth.make_batch(N, H, W, P, seed=seed + i) for i in range(S)]
    return (bs[0]["input"], bs[0]["target"], torch.cat([b["input"][:, 3 + P:] for b in bs], 1),
            torch.stack([b["warps"] for b in bs], 1), torch.stack([b["masks"] for b in bs], 1))


def gen_step_stacked(tag, H, W, P, N, S, seed):
    """gen_type='stacked' (SURVEY 8f-3): one dis_update + gen_update of the unmodified reference trainer with num_stacks = S
    (L1 content loss), incl. the gradient that flows from stack i back into stack i-1 through the generated image."""
    import argparse as ap
    opt = ap.Namespace(image_size=(H, W), use_input_pose=True, pose_dim=P, batch_size=N, num_stacks=S,
                       gen_type="stacked", warp_skip="mask", dataset="fasion", learning_rate=2e-4,
                       content_loss_layer="none", nn_loss_area_size=1, gan_penalty_weight=1.0, l1_penalty_weight=100.0)
    dsd = synth.fill_state_dict(synth.discriminator_shapes(3 + 2 * P + 3), seed + 1)
    gsd = synth.fill_state_dict(synth.generator_shapes(P, (H, W)), seed)
    model = ref_import.make_reference_gan(opt, dsd, None, gen_state=gsd)
    od = vars(opt)
    rec = {"seed": np.int64(seed)}
    inp, tgt, ipose, iwarps, imasks = stacked_inputs(H, W, P, N, S, seed)
    r = synth.make_batch(N, H, W, P, seed=seed + 50)
    inp2, tgt2, ipose2, iwarps2, imasks2 = stacked_inputs(H, W, P, N, S, seed + 100)
    drop_d = [m for i in range(S) for m in synth.dropout_masks(N, 512, 3, seed=seed + i)]
    drop_g = [m for i in range(S) for m in synth.dropout_masks(N, 512, 3, seed=seed + 10 + i)]
    with _DropPatch(drop_d):
        dl = model.dis_update(inp, tgt, {"interpol_pose": ipose, "interpol_warps": iwarps.clone(), "interpol_masks": imasks.clone()},
                              r["input"], r["target"], od)
    rec["d_loss"] = np.array(dl)
    rec["d_grad"] = np.stack([summarize(p.grad) for _, p in sorted(model.disc.named_parameters())])
    with _DropPatch(drop_g):
        out, outs, gl = model.gen_update(inp2, tgt2, {"interpol_pose": ipose2, "interpol_warps": iwarps2.clone(),
                                                      "interpol_masks": imasks2.clone()}, od)
    rec["g_loss"] = np.array(gl)
    rec["out_gen"] = out.detach().numpy()
    rec["out_first"] = outs[0].detach().numpy()
    rec["g_grad"] = np.stack([summarize(p.grad) for _, p in sorted(model.gen.named_parameters())])
    rec["g_param"] = np.stack([summarize(p) for _, p in sorted(model.gen.named_parameters())])
    print("stacked step", tag, dl, gl)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "step_%s.npz" % tag), **rec)


def gen_step_baseline(H, W, P, N, seed, tag, l1_w=100.0):
    """BASELINE configs[0]: src_baseline Pose_GAN (single-encoder Generator, L1 + adversarial), one iteration.  Weights:
    the reference's own xavier initialisation under torch.manual_seed(seed); they are stored so the product starts from
    the same point without the reference tree."""
    import argparse as ap
    import contextlib
    import io
    nsb = ref_import.load_baseline()
    opt = ap.Namespace(image_size=(H, W), use_input_pose=True, pose_dim=P, batch_size=N, num_stacks=4, checkMode=0,
                       gen_type="baseline", dataset="fasion128", learning_rate=2e-4, gan_penalty_weight=1.0,
                       l1_penalty_weight=l1_w)
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        model = nsb.pose_gan.Pose_GAN(opt)
    # deterministic weights that travel: seeded fills instead of the xavier draw
    g = torch.Generator().manual_seed(seed + 100)
    with torch.no_grad():
        for net in (model.gen, model.disc):
            for k, p in sorted(net.named_parameters()):
                if p.dim() == 4:
                    fan = p.shape[1] * p.shape[2] * p.shape[3]
                    p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * (3.0 / fan) ** 0.5)
                elif p.numel() == 1:
                    p.fill_(1.0 if k.endswith("weight") else 0.0)
                else:
                    p.copy_((torch.rand(p.shape, generator=g) - 0.5) * 0.1)
    od = vars(opt)
    rec = {"seed": np.int64(seed)}
    b = synth.make_batch(N, H, W, P, seed=seed)
    r = synth.make_batch(N, H, W, P, seed=seed + 1)
    b2 = synth.make_batch(N, H, W, P, seed=seed + 2)
    with _DropPatch(synth.dropout_masks(N, 512, 3, seed=seed)):
        dl = model.dis_update(b["input"], b["target"], None, r["input"], r["target"], od)
    rec["d_loss"] = np.array(dl)
    rec["d_grad"] = np.stack([summarize(p.grad) for _, p in sorted(model.disc.named_parameters())])
    with _DropPatch(synth.dropout_masks(N, 512, 3, seed=seed + 2)):
        out, _, gl = model.gen_update(b2["input"], b2["target"], None, od)
    rec["g_loss"] = np.array(gl)
    rec["out_gen"] = out.detach().numpy()
    rec["g_grad"] = np.stack([summarize(p.grad) for _, p in sorted(model.gen.named_parameters())])
    rec["g_param"] = np.stack([summarize(p) for _, p in sorted(model.gen.named_parameters())])
    rec["d_param"] = np.stack([summarize(p) for _, p in sorted(model.disc.named_parameters())])
    print("baseline step", tag, dl, gl)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "step_%s.npz" % tag), **rec)


def baseline_initial_weights(model_gen, model_disc, seed):
    """The seeded fills of gen_step_baseline, applied to any pair of modules with the same parameter names (shared with
    the tests so the product starts from the fixture's weights)."""
    g = torch.Generator().manual_seed(seed + 100)
    with torch.no_grad():
        for net in (model_gen, model_disc):
            for k, p in sorted(net.named_parameters()):
                if p.dim() == 4:
                    fan = p.shape[1] * p.shape[2] * p.shape[3]
                    p.copy_(((torch.rand(p.shape, generator=g) * 2 - 1) * (3.0 / fan) ** 0.5).to(p.device))
                elif p.numel() == 1:
                    p.fill_(1.0 if k.endswith("weight") else 0.0)
                else:
                    p.copy_(((torch.rand(p.shape, generator=g) - 0.5) * 0.1).to(p.device))


def main():
    apx = argparse.ArgumentParser()
    apx.add_argument("--only", default="")
    a = apx.parse_args()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.manual_seed(0)
    ns = ref_import.load()
    if a.only in ("", "warp"):
        gen_warp(ns)
    if a.only in ("", "net"):
        gen_networks(ns, 64, 64, 18, 2, 0, "64x64_p18")
        gen_networks(ns, 128, 64, 16, 3, 1, "128x64_p16")
    if a.only in ("", "step"):
        gen_step(ns, 64, 64, 18, 2, 0, "64x64_p18_nn5")
        gen_step(ns, 64, 64, 18, 2, 3, "64x64_p18_l1", content="none", area=1, l1_w=100.0, steps=1)
    if a.only in ("", "step", "vggdeep"):
        # content_loss_layer deeper than block1_conv2: conv2_1 without its ReLU (features[0..5]) and relu(conv3_2) after
        # two max-pools (features[0..13]); seeded weights in every VGG conv
        gen_step(ns, 64, 64, 18, 2, 5, "64x64_p18_b2c1", content="block2_conv1", area=3, l1_w=0.01, steps=1)
        gen_step(ns, 64, 64, 18, 2, 6, "64x64_p18_b3c4", content="block3_conv4", area=3, l1_w=0.01, steps=1)
    if a.only in ("", "big") or a.only.startswith("big:"):
        for case in BIG_CASES:
            if a.only.startswith("big:") and a.only[4:] != case[0]:
                continue
            gen_step_big(ns, *case)
    if a.only in ("", "stacked"):
        gen_step_stacked(*STACKED_CASE)
    if a.only in ("", "baseline"):
        gen_step_baseline(128, 64, 18, 4, 11, "baseline_128x64_p18")
    if a.only in ("", "params"):
        # known answers from the reference logs (gen_full_fasion:158,193 ; gen_full_h36m:136,171)
        for (H, P, ng, nd) in ((256, 18, 82080611, 2803782), (224, 16, 61106781, 2799686)):
            g = sum(int(np.prod(s)) for s in synth.generator_shapes(P, (H, H)).values())
            d = sum(int(np.prod(s)) for s in synth.discriminator_shapes(3 + 2 * P + 3).values())
            assert (g, d) == (ng, nd), (g, d)
        print("param counts OK")


if __name__ == "__main__":
    main()
