"""CPU: the UNMODIFIED reference training script (src_deformable/main.py) driven end-to-end through the product's
module surface (models.networks / models.pose_gan / utils.pose_transform re-exported under the reference's top-level
names, exactly as INTEGRATION.md describes), with the reference's own opts.py, a synthetic stand-in for the dataset
and the torch emulation of the kernels.  Proves the drop-in boundary: constructor/opts contract, dis_update /
gen_update / gen(...) call signatures and return values, checkpoint file names and the state_dict ABI (the files
written by our trainer load strict=True into the REFERENCE classes).  Skipped where /root/reference is absent."""
import os
import runpy
import sys
import types

import numpy as np
import pytest
import torch

from oracle import ref_import, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = ref_import.REFERENCE_ROOT


class _SyntheticPoseDataset(torch.utils.data.Dataset):
    """Stand-in for datasets/PoseTransfer_Dataset.py:163-189: (input, target, warps[10,8], masks[10,H,W] f64)."""

    def __init__(self, opt, split):
        H, W = opt["image_size"]
        self.b = synth.make_batch(8, H, W, opt["pose_dim"], seed=0 if split == "train" else 1)

    def __len__(self):
        return 8

    def __getitem__(self, i):
        b = self.b
        return b["input"][i], b["target"][i], b["warps"][i].double(), b["masks"][i]


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not mounted")
def test_reference_main_py_runs_on_the_product(tmp_path, monkeypatch):
    import pose_transfer_b200  # noqa: F401
    from pose_transfer_b200 import kernels as K
    from pose_transfer_b200.models import networks, pose_gan
    from pose_transfer_b200.utils import pose_transform, pose_utils as our_pu
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import emul_kernels

    work = tmp_path / "work" / "src"
    work.mkdir(parents=True)
    monkeypatch.chdir(work)
    monkeypatch.setattr(sys, "argv", ["main.py", "--dataset", "market", "--batch_size", "2", "--pose_dim", "18",
                                      "--number_of_epochs", "1", "--iters_per_epoch", "2", "--checkpoint_ratio", "1",
                                      "--display_ratio", "1", "--content_loss_layer", "none", "--l1_penalty_weight", "100",
                                      "--expID", "dropin"])
    saved_modules = dict(sys.modules)
    old_cuda = (torch.Tensor.cuda, torch.nn.Module.cuda)
    import torch.utils.data.dataloader as dl
    had_next = hasattr(dl._BaseDataLoaderIter, "next")
    try:
        # --- the launcher of INTEGRATION.md, plus stubs for what this image lacks
        models = types.ModuleType("models")
        models.networks, models.pose_gan = networks, pose_gan
        utils = types.ModuleType("utils")
        pu = types.ModuleType("utils.pose_utils")
        pu.get_imgpose, pu.Feature_Extractor, pu.get_model_list = our_pu.get_imgpose, our_pu.Feature_Extractor, our_pu.get_model_list
        pu.display = lambda *a, **k: np.zeros((8, 8, 3), dtype=np.float32)       # visualisation is out of scope
        utils.pose_utils, utils.pose_transform = pu, pose_transform
        datasets = types.ModuleType("datasets")
        dsm = types.ModuleType("datasets.PoseTransfer_Dataset")
        dsm.PoseTransfer_Dataset = _SyntheticPoseDataset
        datasets.PoseTransfer_Dataset = dsm
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        pylab = types.ModuleType("pylab")
        pylab.imsave = lambda *a, **k: None
        pylab.cm = types.SimpleNamespace(gray=None)
        sys.modules.update({"models": models, "models.networks": networks, "models.pose_gan": pose_gan, "utils": utils,
                            "utils.pose_utils": pu, "utils.pose_transform": pose_transform, "datasets": datasets,
                            "datasets.PoseTransfer_Dataset": dsm, "matplotlib": mpl, "pylab": pylab})
        sys.modules.pop("opts", None)
        sys.path.insert(0, REF)                                                     # for the reference's own opts.py
        dl._BaseDataLoaderIter.next = dl._BaseDataLoaderIter.__next__               # main.py:27 uses iter.next()
        if not torch.cuda.is_available():
            torch.Tensor.cuda = lambda self, *a, **k: self
            torch.nn.Module.cuda = lambda self, *a, **k: self
        monkeypatch.setattr(networks, "_require_cuda", lambda t, who: None)     # kernels are emulated on CPU here
        with emul_kernels.install(K):
            runpy.run_path(os.path.join(REF, "main.py"), run_name="reference_main")
    finally:
        torch.Tensor.cuda, torch.nn.Module.cuda = old_cuda
        if not had_next:
            del dl._BaseDataLoaderIter.next
        if REF in sys.path:
            sys.path.remove(REF)
        # drop only the names the launcher injected (native extension modules such as cv2 cannot be re-imported
        # once removed from sys.modules), then restore whatever was there before
        for k in list(sys.modules):
            if k not in saved_modules and k.split(".")[0] in ("models", "utils", "datasets", "matplotlib", "pylab", "opts"):
                del sys.modules[k]
        sys.modules.update(saved_modules)

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
