"""CPU, world_size 2 over gloo: the data-parallel path of the trainer (gradient all-reduce on the flat arena,
1/world folded into Adam, rank-0 weight broadcast).  Kernels are the torch emulation (tests/emul_kernels.py);
the host logic under test is the product's.  Property checked: 2 ranks x per-rank batch 2 produce exactly the
update of 1 rank x batch 4 on the concatenated samples, and replicas stay bit-identical."""
import argparse
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H = W = 64
P = 18


def _opt(N):
    return argparse.Namespace(image_size=(H, W), use_input_pose=True, pose_dim=P, batch_size=N, num_stacks=4,
                              gen_type="baseline", warp_skip="mask", dataset="fasion", learning_rate=2e-4,
                              content_loss_layer="none", nn_loss_area_size=1, gan_penalty_weight=1.0,
                              l1_penalty_weight=100.0)


def _make_model(N):
    import pose_transfer_b200  # noqa: F401
    from pose_transfer_b200.models import pose_gan
    from oracle import synth
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    model = pose_gan.DeformablePose_GAN(_opt(N))
    model.gen.load_state_dict(synth.fill_state_dict(synth.generator_shapes(P, (H, W)), 0))
    model.disc.load_state_dict(synth.fill_state_dict(synth.discriminator_shapes(3 + 2 * P + 3), 1))
    return model


def _step(model, b, r, drop, N):
    od = vars(_opt(N))
    io = {"warps": b["warps"], "masks": b["masks"]}
    d = model.dis_update(b["input"], b["target"], io, r["input"], r["target"], od, drop=drop)
    _, _, g = model.gen_update(b["input"], b["target"], io, od, drop=drop)
    return d, g


def _cat(bs):
    return {k: torch.cat([b[k] for b in bs], 0) for k in bs[0]}


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import emul_kernels
    from pose_transfer_b200 import kernels as K
    from oracle import synth
    with emul_kernels.install(K):
        model = _make_model(2)
        assert model.world == world
        b = synth.make_batch(2, H, W, P, seed=10 + rank)
        r = synth.make_batch(2, H, W, P, seed=20 + rank)
        drop = synth.dropout_masks(2, 512, 3, seed=30 + rank)
        _step(model, b, r, drop, 2)
    torch.save({"gen": model.gen_arena.flat.clone(), "disc": model.disc_arena.flat.clone()},
               os.path.join(out_dir, "rank%d.pt" % rank))
    dist.destroy_process_group()


def test_two_ranks_equal_one_rank_with_double_batch(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(os.path.join(tmp_path, "rank0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "rank1.pt"))
    assert torch.equal(r0["gen"], r1["gen"]) and torch.equal(r0["disc"], r1["disc"]), "replicas diverged"

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import emul_kernels
    from pose_transfer_b200 import kernels as K
    from oracle import synth
    old = (torch.Tensor.cuda, torch.nn.Module.cuda)
    try:
        with emul_kernels.install(K):
            model = _make_model(4)
            b = _cat([synth.make_batch(2, H, W, P, seed=10 + k) for k in range(2)])
            r = _cat([synth.make_batch(2, H, W, P, seed=20 + k) for k in range(2)])
            drops = [synth.dropout_masks(2, 512, 3, seed=30 + k) for k in range(2)]
            drop = [torch.cat([drops[0][j], drops[1][j]], 0) for j in range(3)]
            _step(model, b, r, drop, 4)
    finally:
        torch.Tensor.cuda, torch.nn.Module.cuda = old
    # first Adam step moves every weight by ~lr*sign(g): compare the update directions
    for name, single, multi in (("gen", model.gen_arena.flat, r0["gen"]), ("disc", model.disc_arena.flat, r0["disc"])):
        diff = (single - multi).abs()
        frac_bad = float((diff > 1e-5).float().mean())
        assert frac_bad < 2e-3, "%s: %.4f of the weights disagree between 2x2 and 1x4" % (name, frac_bad)
