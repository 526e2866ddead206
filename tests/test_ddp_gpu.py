"""GPU, world_size 2 over NCCL (needs two visible GPUs; skipped otherwise -- run with `gpurun --gpus 2`): the
data-parallel trainer on the real kernels.  (a) Replicas stay BIT-IDENTICAL over 3 full iterations (same broadcast
weights + identical all-reduced gradients + the same Adam kernel => same bits on every rank).  (b) 2 ranks x per-rank
batch 4 take the same optimiser step as 1 process x batch 8 on the concatenated samples: gradients agree to fp32
summation order (exact-fp32 conv mode), SURVEY Appendix D."""
import argparse
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H = W = 64
P = 18


def _opt(N):
    return argparse.Namespace(image_size=(H, W), use_input_pose=True, pose_dim=P, batch_size=N, num_stacks=4,
                              gen_type="baseline", warp_skip="mask", dataset="fasion", learning_rate=2e-4,
                              content_loss_layer="block1_conv2", nn_loss_area_size=5, gan_penalty_weight=1.0,
                              l1_penalty_weight=0.01)


def _model(N):
    import contextlib
    import io
    import pose_transfer_b200  # noqa: F401
    from pose_transfer_b200.models import pose_gan
    from oracle import synth
    with contextlib.redirect_stdout(io.StringIO()):
        model = pose_gan.DeformablePose_GAN(_opt(N)).cuda()
    model.gen.load_state_dict(synth.fill_state_dict(synth.generator_shapes(P, (H, W)), 0))
    model.disc.load_state_dict(synth.fill_state_dict(synth.discriminator_shapes(3 + 2 * P + 3), 1))
    vw, vb = synth.vgg_conv1_1(0)
    with torch.no_grad():
        model.content_model.features[0].weight.copy_(vw)
        model.content_model.features[0].bias.copy_(vb)
    return model


def _cat(bs):
    return {k: torch.cat([b[k] for b in bs], 0) for k in bs[0]}


def _step(model, b, r, drop, N):
    od = vars(_opt(N))
    io = {"warps": b["warps"].cuda(), "masks": b["masks"].cuda()}
    model.dis_update(b["input"].cuda(), b["target"].cuda(), io, r["input"].cuda(), r["target"].cuda(), od, drop=drop)
    dg = model.disc_arena.grad.clone()
    model.gen_update(b["input"].cuda(), b["target"].cuda(), io, od, drop=drop)
    return dg, model.gen_arena.grad.clone()


def _worker(rank, world, port, out_dir, mode):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), PTK_CONV_IMPL=mode)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from oracle import synth
    model = _model(4)
    res = {}
    for it in range(3):
        b = synth.make_batch(4, H, W, P, seed=100 * it + rank)
        r = synth.make_batch(4, H, W, P, seed=100 * it + 10 + rank)
        drop = [d[4 * rank:4 * rank + 4] for d in synth.dropout_masks(8, 512, 3, seed=it)]
        dg, gg = _step(model, b, r, drop, 4)
        if it == 0:
            res["d_grad"], res["g_grad"] = dg.cpu(), gg.cpu()        # all-reduced SUM over ranks (1/world is folded into Adam)
    res["gen"], res["disc"] = model.gen_arena.flat.cpu(), model.disc_arena.flat.cpu()
    torch.save(res, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["simt", "auto"])
def test_two_rank_nccl_training(tmp_path, mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from oracle import synth
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path), mode), nprocs=2, join=True)
    r0, r1 = (torch.load(os.path.join(str(tmp_path), "rank%d.pt" % r)) for r in (0, 1))
    # (a) bit-identical replicas after 3 iterations, and identical all-reduced gradients
    for k in ("gen", "disc", "d_grad", "g_grad"):
        assert torch.equal(r0[k], r1[k]), "rank 0 and rank 1 differ in " + k
    if mode != "simt":
        return
    # (b) 2 x 4 == 1 x 8 on the concatenated samples (first iteration, same weights)
    os.environ["PTK_CONV_IMPL"] = "simt"
    try:
        model = _model(8)
        b = _cat([synth.make_batch(4, H, W, P, seed=rk) for rk in (0, 1)])
        r = _cat([synth.make_batch(4, H, W, P, seed=10 + rk) for rk in (0, 1)])
        dg, gg = _step(model, b, r, synth.dropout_masks(8, 512, 3, seed=0), 8)
    finally:
        os.environ.pop("PTK_CONV_IMPL", None)
    # per-rank losses are means over 4 samples scaled by 1 / batch_size(4): the rank-sum equals 2x the batch-8 gradient
    for name, got, want in (("disc", r0["d_grad"], dg.cpu()), ("gen", r0["g_grad"], gg.cpu())):
        rel = float((got.double() / 2 - want.double()).norm() / want.double().norm())
        print("2x4 vs 1x8 %s gradient: rel-L2 %.3g" % (name, rel))
        assert rel <= 2e-4, (name, rel)
