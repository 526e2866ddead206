"""CPU: the UNMODIFIED reference training script (src_deformable/main.py) driven end-to-end through the product's
module surface (tests/dropin_launcher.py = the launcher of INTEGRATION.md) on the torch emulation of the kernels.
Proves the drop-in boundary: constructor/opts contract, dis_update / gen_update / gen(...) call signatures and return
values, checkpoint file names and the state_dict ABI (the files written by our trainer load strict=True into the
REFERENCE classes).  The same script runs on the real CUDA library in tests/test_modules_gpu.py.
Skipped where no copy of the reference is available."""
import pytest
import torch

from oracle import ref_import

import dropin_launcher


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not available")
def test_reference_main_py_runs_on_the_product(tmp_path, monkeypatch):
    work = tmp_path / "work" / "src"
    work.mkdir(parents=True)
    monkeypatch.chdir(work)
    dropin_launcher.run_reference_main(
        ["main.py", "--dataset", "market", "--batch_size", "2", "--pose_dim", "18", "--number_of_epochs", "1",
         "--iters_per_epoch", "2", "--checkpoint_ratio", "1", "--display_ratio", "1", "--content_loss_layer", "none",
         "--l1_penalty_weight", "100", "--expID", "dropin"], emulate_kernels=True, monkeypatch=monkeypatch)
    ckpt = tmp_path / "work" / "exp" / "dropin" / "models"
    assert (ckpt / "gen_001.pkl").is_file() and (ckpt / "disc_001.pkl").is_file()
    assert (tmp_path / "work" / "exp" / "dropin" / "opt.txt").is_file()
    # checkpoint ABI: our files load strict=True into the REFERENCE classes
    ns = ref_import.load()
    enc, dec = (64, 128, 256, 512, 512, 512), (512, 512, 512, 256, 128, 3)
    G = ns.networks.Deformable_Generator(39, 18, (128, 64), enc, dec, "mask")
    D = ns.networks.Discriminator(42)
    G.load_state_dict(torch.load(ckpt / "gen_001.pkl"), strict=True)
    D.load_state_dict(torch.load(ckpt / "disc_001.pkl"), strict=True)
    assert all(torch.isfinite(p).all() for p in G.parameters())
