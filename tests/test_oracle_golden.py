"""CPU: pin oracle/restate.py against the golden outputs of the UNMODIFIED reference (tests/golden,
made by oracle/make_golden.py) and against the reference's known answers."""
import numpy as np
import pytest
import torch

from oracle import restate, synth
from oracle.make_golden import WARP_CASES, warp_inputs
from helpers import golden, summarize, max_abs, assert_summary_close


@pytest.mark.parametrize("case", WARP_CASES, ids=[c[0] for c in WARP_CASES])
def test_warp_restatement_matches_reference(case):
    name, N, C, h, w, H0, W0, seed = case
    x, warps, masks, gy = warp_inputs(N, C, h, w, H0, W0, seed)
    x = x.clone().requires_grad_(True)
    y = restate.affine_warp(x, warps, masks, (H0, W0))
    y.backward(gy)
    g = golden(name)
    # fp32 coordinate rounding with +-30 px translations: <= 2e-4 abs
    assert max_abs(y, g["y"]) <= 2e-4
    assert max_abs(x.grad, g["dx"]) <= 2e-4


@pytest.mark.parametrize("tag,H,W,P,N,seed", [("64x64_p18", 64, 64, 18, 2, 0), ("128x64_p16", 128, 64, 16, 3, 1)])
def test_networks_restatement_matches_reference(tag, H, W, P, N, seed):
    g = golden("net_" + tag)
    gsd = synth.fill_state_dict(synth.generator_shapes(P, (H, W)), seed)
    dsd = synth.fill_state_dict(synth.discriminator_shapes(3 + 2 * P + 3), seed + 1)
    b = synth.make_batch(N, H, W, P, seed=seed)
    drop = synth.dropout_masks(N, 512, 3, seed=seed)
    with torch.no_grad():
        out = restate.generator_forward(gsd, b["input"], b["warps"], b["masks"], (H, W), P, drop)
        img, src, tgt = restate.get_imgpose(b["input"], True, P)
        d_out = restate.discriminator_forward(dsd, torch.cat([img, src, out, tgt], 1))
    assert max_abs(out, g["out_gen"]) <= 2e-5
    assert max_abs(d_out, g["d_out"]) <= 2e-6
    assert int(g["n_params_g"]) == sum(int(np.prod(s)) for s in synth.generator_shapes(P, (H, W)).values())


def test_known_answer_param_counts():
    # src_deformable/logs/gen_full_fasion:158,193 and gen_full_h36m:136,171
    for (H, P, ng, nd) in ((256, 18, 82080611, 2803782), (224, 16, 61106781, 2799686)):
        assert sum(int(np.prod(s)) for s in synth.generator_shapes(P, (H, H)).values()) == ng
        assert sum(int(np.prod(s)) for s in synth.discriminator_shapes(3 + 2 * P + 3).values()) == nd


def test_known_answer_layer_index_and_vgg_preprocess():
    assert restate.get_layer_ind("block1_conv2") == 1       # utils/pose_utils.py:312-317
    x = torch.arange(2 * 3 * 4 * 5, dtype=torch.float32).view(2, 3, 4, 5) / 100
    ref = x.view(2, 4, 5, 3)
    ref = ((ref - torch.tensor(restate.VGG_MEAN)) / torch.tensor(restate.VGG_STD)).view(2, 3, 4, 5)
    assert torch.equal(restate.vgg_preprocess(x), ref)


def test_adversarial_loss_recipe():
    # src_baseline/unitTests.py:158-182: -mean(log(x + 1e-7)) on a random 4x49 matrix
    x = torch.rand(4, 49, generator=torch.Generator().manual_seed(0))
    ref = sum(-torch.mean(torch.log(x[i] + 1e-7)) for i in range(4))
    assert abs(float(restate.adv_true(x)) - float(ref)) < 1e-6


def _run_steps(tag, content, area, l1_w, steps, seed):
    H = W = 64
    P, N = 18, 2
    g = golden("step_" + tag)
    vw, vb = synth.vgg_conv1_1(seed)
    model = restate.OracleGAN(synth.fill_state_dict(synth.generator_shapes(P, (H, W)), seed),
                              synth.fill_state_dict(synth.discriminator_shapes(3 + 2 * P + 3), seed + 1),
                              vw, vb, (H, W), P, N, content_loss_layer=content, nn_loss_area_size=area)
    for s in range(steps):
        b = synth.make_batch(N, H, W, P, seed=seed + 10 * s)
        r = synth.make_batch(N, H, W, P, seed=seed + 10 * s + 1)
        b2 = synth.make_batch(N, H, W, P, seed=seed + 10 * s + 2)
        dl = model.dis_update(b["input"], b["target"], b["warps"], b["masks"], r["input"], r["target"], 1.0,
                              synth.dropout_masks(N, 512, 3, seed=seed + 10 * s))
        # step 0 is a pure function of the inputs; later steps see Adam-updated weights (first Adam step is
        # +-lr*sign(g), chaotic in near-zero gradients) so the tolerance widens
        rt = 2e-5 if s == 0 else 5e-3
        np.testing.assert_allclose(dl, g["d_loss_%d" % s], rtol=rt)
        dg = np.stack([summarize(model.disc[k].grad) for k in sorted(model.disc)])
        loose = {} if s == 0 else dict(tol_norm=5e-2, tol_samp=0.25, tol_scalar=0.5)
        assert_summary_close(dg, g["d_grad_%d" % s], what="d_grad", **loose)
        out, gl = model.gen_update(b2["input"], b2["target"], b2["warps"], b2["masks"], 1.0, l1_w,
                                   synth.dropout_masks(N, 512, 3, seed=seed + 10 * s + 2))
        np.testing.assert_allclose(gl, g["g_loss_%d" % s], rtol=rt)
        assert max_abs(out, g["out_gen_%d" % s]) <= (5e-5 if s == 0 else 5e-3)
        gg = np.stack([summarize(model.gen[k].grad) for k in sorted(model.gen)])
        assert_summary_close(gg, g["g_grad_%d" % s], what="g_grad", **loose)
        gp = np.stack([summarize(model.gen[k]) for k in sorted(model.gen)])
        assert_summary_close(gp, g["g_param_%d" % s], tol_norm=1e-4 if s == 0 else 1e-3, tol_samp=2e-3 if s == 0 else 5e-2,
                             tol_scalar=2e-3 if s == 0 else 5e-2, what="g_param")


def test_train_step_nn_loss_matches_reference():
    _run_steps("64x64_p18_nn5", "block1_conv2", 5, 0.01, 2, 0)


def test_train_step_l1_matches_reference():
    _run_steps("64x64_p18_l1", "none", 1, 100.0, 1, 3)


def test_restatement_matches_live_reference_when_mounted():
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference tree not mounted (GPU box)")
    ns = ref_import.load()
    x, warps, masks, _ = warp_inputs(2, 4, 16, 8, 64, 32, 9)
    y_ref = ns.pose_transform.AffineTransformLayer(10, (64, 32), "mask")(x, warps.clone(), masks.clone())
    assert max_abs(restate.affine_warp(x, warps, masks, (64, 32)), y_ref) <= 2e-4


def test_benchmark_size_fixture_pins_restatement():
    """tests/golden/step_256x256_p18_n2.npz (one dis_update + gen_update of the unmodified reference at the benchmarked
    geometry, oracle/make_golden.py::gen_step_big) against the restatement: losses, out_gen lattice, gradient norm and
    sampled-gradient cosine of every parameter tensor."""
    from oracle.make_golden import BIG_CASES, big_sample_idx
    tag, H, W, P, N, seed, stride = BIG_CASES[0]
    g = golden("step_" + tag)
    gsd = synth.fill_state_dict(synth.generator_shapes(P, (H, W)), seed)
    dsd = synth.fill_state_dict(synth.discriminator_shapes(3 + 2 * P + 3), seed + 1)
    vw, vb = synth.vgg_conv1_1(seed)
    model = restate.OracleGAN(gsd, dsd, vw, vb, (H, W), P, N, faithful_waste=False)
    b, r, b2 = (synth.make_batch(N, H, W, P, seed=seed + i) for i in range(3))
    dl = model.dis_update(b["input"], b["target"], b["warps"], b["masks"], r["input"], r["target"], 1.0,
                          synth.dropout_masks(N, 512, 3, seed=seed))
    np.testing.assert_allclose(dl, g["d_loss"], rtol=2e-5)
    for i, k in enumerate(sorted(model.disc)):
        got = model.disc[k].grad.reshape(-1).double().numpy()
        tol = 5e-2 if got.size == 1 else 2e-3     # 1-element norm gains / biases: cancellation-heavy global sums
        assert abs(np.linalg.norm(got) - g["d_grad_norm"][i]) <= tol * g["d_grad_norm"][i] + 1e-9, k
    out, gl = model.gen_update(b2["input"], b2["target"], b2["warps"], b2["masks"], 1.0, 0.01,
                               synth.dropout_masks(N, 512, 3, seed=seed + 2))
    np.testing.assert_allclose(gl, g["g_loss"], rtol=2e-4)
    assert max_abs(out[:, :, ::stride, ::stride], g["out_gen"]) <= 5e-4
    worst = 1.0
    for i, k in enumerate(sorted(model.gen)):
        got = model.gen[k].grad.reshape(-1).double().numpy()
        if got.size < 64:
            continue
        s = got[big_sample_idx(got.size)]
        w = g["g_grad_%02d" % i].astype(np.float64)
        cos = float(s @ w / (np.linalg.norm(s) * np.linalg.norm(w)))
        worst = min(worst, cos)
        assert cos >= 0.9995, (k, cos)
        assert abs(np.linalg.norm(got) - g["g_grad_norm"][i]) <= 2e-2 * g["g_grad_norm"][i], k
    print("worst sampled-gradient cosine restatement vs reference fixture:", worst)
