"""TEST / BASELINE INFRASTRUCTURE -- runs the reference's own modules (staged in oracle/_ref/ by oracle/build_ref.py, or
imported straight from /root/reference when that tree is present) on the CPU.  Only tests/, smoke() and bench.py's
cpu_baseline / --impl reference legs may import this; the product package never does.

What is patched around the unmodified reference code, and why:
  * ``pytorch3d.loss.chamfer_distance`` (third party, not installable here) := oracle.adabins_oracle.chamfer_distance, so
    BinsChamferLoss (loss.py:33-46) runs unmodified above the call boundary;
  * the loaders call ``.cuda()`` unconditionally (SemanticsLoader.py:122,130,142; InstanceSegmentationLoader.py:109-118):
    ``Tensor.cuda`` is the identity while a loader runs, so the reference's CPU gather is what is executed and timed;
  * ``UnetAdaptiveBins.build`` needs torch.hub (network): the class is constructed directly with the geffnet-shaped
    random-init backbone of mde_biological_vision_systems_b200.models.efficientnet.
"""
import contextlib
import importlib
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
_STATE = {}


def locate():
    """Directory holding the reference modules: oracle/_ref (staged) or /root/reference; None if neither exists."""
    staged = os.path.join(HERE, "_ref")
    if os.path.exists(os.path.join(staged, "MANIFEST.json")):
        return staged
    ref = os.environ.get("MDE_REFERENCE_ROOT", "/root/reference")
    return ref if os.path.isdir(os.path.join(ref, "models")) else None


def available():
    return locate() is not None


def load():
    """-> namespace with the reference's UnetAdaptiveBins, SILogLoss, BinsChamferLoss, SemanticsLoader,
    InstanceSegmentationLoader classes (imported once under private module names)."""
    if "ns" in _STATE:
        return _STATE["ns"]
    base = locate()
    if base is None:
        raise RuntimeError("reference modules unavailable: neither oracle/_ref nor /root/reference exists")
    from oracle import adabins_oracle as oracle
    stub, stub_loss = types.ModuleType("pytorch3d"), types.ModuleType("pytorch3d.loss")
    stub_loss.chamfer_distance = oracle.chamfer_distance
    stub.loss = stub_loss
    sys.modules.setdefault("pytorch3d", stub)
    sys.modules.setdefault("pytorch3d.loss", stub_loss)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "models" or k.startswith("models.") or k == "loss"
             or k == "ExternalInfoLoaders" or k.startswith("ExternalInfoLoaders.")}
    sys.path.insert(0, base)
    try:
        ns = types.SimpleNamespace(
            base=base,
            UnetAdaptiveBins=importlib.import_module("models").UnetAdaptiveBins,
            SILogLoss=importlib.import_module("loss").SILogLoss,
            BinsChamferLoss=importlib.import_module("loss").BinsChamferLoss,
            SemanticsLoader=importlib.import_module("ExternalInfoLoaders.SemanticsLoader").SemanticsLoader,
            InstanceSegmentationLoader=importlib.import_module("ExternalInfoLoaders.InstanceSegmentationLoader").InstanceSegmentationLoader)
    finally:
        sys.path.remove(base)
        for k in list(sys.modules):  # keep the reference's top-level names out of the product's import namespace
            if k == "models" or k.startswith("models.") or k == "loss" or k == "ExternalInfoLoaders" or k.startswith("ExternalInfoLoaders."):
                sys.modules["_mde_ref_" + k] = sys.modules.pop(k)
        sys.modules.update(saved)
    _STATE["ns"] = ns
    return ns


@contextlib.contextmanager
def cpu_loaders():
    """The reference loaders open ``data/*.npy`` relative to the cwd and end with ``.cuda()``."""
    real_cuda, cwd = torch.Tensor.cuda, os.getcwd()
    torch.Tensor.cuda = lambda self, *a, **k: self
    os.chdir(ROOT)
    try:
        yield
    finally:
        torch.Tensor.cuda = real_cuda
        os.chdir(cwd)


def build_model(encoder_name="efficientnet-b1", state_dict=None, **kw):
    """The reference's UnetAdaptiveBins on the product's geffnet-shaped backbone (same construction as tests/helpers.py
    make_model / tests/golden/make_golden.py), optionally loaded with ``state_dict``; eval mode, CPU."""
    from mde_biological_vision_systems_b200.models.efficientnet import SamePadConv2d, build_backbone
    ns = load()
    bb = "tf_efficientnet_b5_ap" if "b5" in encoder_name else "tf_efficientnet_b1_ap"
    backbone = build_backbone(bb, seed=0)
    backbone.global_pool = torch.nn.Identity()
    backbone.classifier = torch.nn.Identity()
    add = ns.UnetAdaptiveBins.get_num_channels_to_add(encoder_name, kw.get("semantics_mode"),
                                                      kw.get("instance_segmentation_mode"), kw.get("image", "rgb"))
    if kw.get("insertion_point") == "input" and add:
        backbone.conv_stem = SamePadConv2d(3 + add, 32, 3, 2)
    m = ns.UnetAdaptiveBins(backbone, n_bins=256, min_val=1e-3, max_val=10, norm="linear", encoder_name=encoder_name, **kw)
    if state_dict is not None:
        m.load_state_dict(state_dict)
    return m.eval()
