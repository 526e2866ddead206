"""Shared launcher of the drop-in tests (test infrastructure): execs the UNMODIFIED reference ``src_deformable/main.py``
on top of the product's module surface exactly as INTEGRATION.md describes -- the product's ``models.networks`` /
``models.pose_gan`` / ``utils.pose_transform`` (+ the three hot-path functions of ``utils.pose_utils``) under the
reference's top-level names, the reference's own ``opts.py``, a synthetic stand-in for the dataset (no dataset files
offline) and stubs for the visualisation modules this image lacks."""
import os
import runpy
import sys
import types

import numpy as np
import torch

from oracle import ref_import, synth


class SyntheticPoseDataset(torch.utils.data.Dataset):
    """Stand-in for datasets/PoseTransfer_Dataset.py:163-189: (input, target, warps[10,8] f64, masks[10,H,W] f64)."""

    def __init__(self, opt, split):
        H, W = opt["image_size"]
        self.b = synth.make_batch(8, H, W, opt["pose_dim"], seed=0 if split == "train" else 1)

    def __len__(self):
        return 8

    def __getitem__(self, i):
        b = self.b
        return b["input"][i], b["target"][i], b["warps"][i].double(), b["masks"][i]


def run_reference_main(argv, emulate_kernels, monkeypatch):
    """Exec main.py (module-level ``main()`` call) in the current working directory with sys.argv = argv.
    emulate_kernels=True: CPU run on tests/emul_kernels.py (host logic only); False: the real CUDA library."""
    import pose_transfer_b200  # noqa: F401
    from pose_transfer_b200 import kernels as K
    from pose_transfer_b200.models import networks, pose_gan
    from pose_transfer_b200.utils import pose_transform, pose_utils as our_pu
    REF = ref_import.REFERENCE_ROOT
    monkeypatch.setattr(sys, "argv", argv)
    saved_modules = dict(sys.modules)
    old_cuda = (torch.Tensor.cuda, torch.nn.Module.cuda)
    import torch.utils.data.dataloader as dl
    had_next = hasattr(dl._BaseDataLoaderIter, "next")
    try:
        models = types.ModuleType("models")
        models.networks, models.pose_gan = networks, pose_gan
        utils = types.ModuleType("utils")
        pu = types.ModuleType("utils.pose_utils")
        pu.get_imgpose, pu.Feature_Extractor, pu.get_model_list = our_pu.get_imgpose, our_pu.Feature_Extractor, our_pu.get_model_list
        pu.display = lambda *a, **k: np.zeros((8, 8, 3), dtype=np.float32)       # visualisation is out of scope
        utils.pose_utils, utils.pose_transform = pu, pose_transform
        datasets = types.ModuleType("datasets")
        dsm = types.ModuleType("datasets.PoseTransfer_Dataset")
        dsm.PoseTransfer_Dataset = SyntheticPoseDataset
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
        if emulate_kernels:
            sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
            import emul_kernels
            if not torch.cuda.is_available():
                torch.Tensor.cuda = lambda self, *a, **k: self
                torch.nn.Module.cuda = lambda self, *a, **k: self
            monkeypatch.setattr(networks, "_require_cuda", lambda t, who: None)     # kernels are emulated on CPU here
            with emul_kernels.install(K):
                runpy.run_path(os.path.join(REF, "main.py"), run_name="reference_main")
        else:
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
