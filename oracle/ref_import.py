"""TEST INFRASTRUCTURE ONLY.  Import the unmodified reference modules.

The reference tree (``/root/reference/src_deformable``) only exists in the build container, never on
the GPU box, so everything here is used to *generate* fixtures (``oracle/make_golden.py``) and to pin
``oracle/restate.py`` in the CPU test-suite (tests skip when the tree is absent).

Shims (none of them touches arithmetic):
  * stub modules for dead/unavailable imports: keras (networks.py:9), skimage (pose_transform.py:3-6,
    pose_utils.py:5,13), matplotlib/pylab (pose_utils.py:8-11, pose_transform.py:1);
  * ``.cuda()`` is hard-coded (pose_transform.py:74,84; pose_utils.py:321,327; pose_gan.py:59-60) so
    on a CPU-only host ``Tensor.cuda``/``Module.cuda`` become identity;
  * ``DeformablePose_GAN.__init__`` unconditionally loads ``disc_090.pkl`` (pose_gan.py:40-42) and
    ``vgg19(pretrained=True)`` (pose_gan.py:56): both are patched to seeded in-memory objects.
"""
import os
import sys
import types

import torch

from . import fetch_ref

# the mounted reference in the build container, else the verbatim copy that travels to the GPU box (oracle/fetch_ref.py)
REFERENCE_ROOT = os.environ.get("PTK_REFERENCE_ROOT") or fetch_ref.root("src_deformable") or "/root/reference/src_deformable"


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "networks.py"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_loaded = None
_loaded_by_root = {}
BASELINE_ROOT = os.path.join(os.path.dirname(REFERENCE_ROOT), "src_baseline")


def load_baseline():
    """The src_baseline tree (Generator / Pose_GAN), imported the same way under private module names."""
    return load(BASELINE_ROOT)


def load(root=None):
    """Returns a namespace with the reference modules: networks, pose_gan, pose_transform, pose_utils."""
    global _loaded
    if root is not None and root != REFERENCE_ROOT:
        if root not in _loaded_by_root:
            _loaded_by_root[root] = _load_tree(root, "_ptk_reference_" + os.path.basename(root) + ".")
        return _loaded_by_root[root]
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _loaded = _load_tree(REFERENCE_ROOT, "_ptk_reference.")
    return _loaded


def _load_tree(REFERENCE_ROOT, private_prefix):
    if not os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "networks.py")):
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    saved = {k: sys.modules.get(k) for k in ("models", "utils", "models.networks", "models.pose_gan",
                                             "utils.pose_transform", "utils.pose_utils")}
    if "keras" not in sys.modules:
        _stub("keras")
        _stub("keras.optimizers", Adam=object)
    if "skimage" not in sys.modules:
        sk = _stub("skimage")
        sk.io = _stub("skimage.io", imread=None)
        sk.transform = _stub("skimage.transform", warp_coords=None, estimate_transform=None)
        sk.measure = _stub("skimage.measure")
        sk.draw = _stub("skimage.draw", circle=None, line_aa=None, polygon=None)
    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib", use=lambda *a, **k: None)
        mpl.pyplot = _stub("matplotlib.pyplot")
        mpl.patches = _stub("matplotlib.patches")
    if "pylab" not in sys.modules:
        _stub("pylab")
    if not torch.cuda.is_available():
        _patch_cuda_identity()
    for k in list(saved):
        sys.modules.pop(k, None)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        from models import networks, pose_gan          # noqa
        from utils import pose_transform, pose_utils   # noqa
    finally:
        sys.path.remove(REFERENCE_ROOT)
    ns = types.SimpleNamespace(networks=networks, pose_gan=pose_gan,
                               pose_transform=pose_transform, pose_utils=pose_utils)
    # keep the reference modules reachable under private names only, so that they never shadow the
    # product's own `models` / `utils` drop-in packages.
    for k in ("models", "utils", "models.networks", "models.pose_gan", "utils.pose_transform",
              "utils.pose_utils"):
        mod = sys.modules.pop(k, None)
        if mod is not None:
            sys.modules[private_prefix + k] = mod
        if saved[k] is not None:
            sys.modules[k] = saved[k]
    return ns


_orig_cuda = None


def _patch_cuda_identity():
    global _orig_cuda
    if _orig_cuda is None:
        _orig_cuda = (torch.Tensor.cuda, torch.nn.Module.cuda)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self


class cpu_only:
    """Context manager: the reference hard-codes ``.cuda()``; inside this block those calls are the identity even on a
    box that has a GPU, so the reference's step runs on the host cores (bench.py's cpu_baseline / reference arm)."""

    def __enter__(self):
        self.saved = (torch.Tensor.cuda, torch.nn.Module.cuda)
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        return self

    def __exit__(self, *a):
        torch.Tensor.cuda, torch.nn.Module.cuda = self.saved


def make_reference_gan(opt, disc_state, vgg, gen_state=None):
    """Instantiate the reference DeformablePose_GAN with its file loads patched (disc_090.pkl, pose_gan.py:40-42; for
    gen_type='stacked' also gen_090.pkl, :31-32; vgg19(pretrained=True), :56)."""
    ns = load()
    pg = ns.pose_gan
    old_load, old_vgg = pg.torch.load, pg.vgg19
    pg.torch.load = lambda path, *a, **k: gen_state if "gen_" in os.path.basename(str(path)) else disc_state
    pg.vgg19 = lambda pretrained=True: vgg
    try:
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            model = pg.DeformablePose_GAN(opt)
    finally:
        pg.torch.load, pg.vgg19 = old_load, old_vgg
    return model
