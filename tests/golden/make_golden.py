"""Generate tests/golden/*.npz by running the REFERENCE modules themselves (imported from /root/reference)
on the seeded synthetic inputs of mde_biological_vision_systems_b200.synthetic.

Run once in the build container (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py

What is patched, and why (nothing in /root/reference is modified):
  * ``pytorch3d.loss.chamfer_distance`` is not installable -> oracle.adabins_oracle.chamfer_distance is
    injected as that module, so BinsChamferLoss (loss.py:33-46) runs unmodified above the call boundary.
  * the loaders call ``.cuda()`` unconditionally (SemanticsLoader.py:122,130,142) -> ``Tensor.cuda`` is
    replaced by the identity for the duration of this script so they run on CPU.
  * ``UnetAdaptiveBins.build`` needs torch.hub (network) -> the class is constructed directly with the
    geffnet-shaped random-init backbone of models/efficientnet.py.
"""
import hashlib
import os
import sys
import types
from argparse import Namespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import adabins_oracle as oracle  # noqa: E402
from mde_biological_vision_systems_b200 import synthetic  # noqa: E402
from mde_biological_vision_systems_b200.models.efficientnet import build_backbone  # noqa: E402


def _import_reference():
    stub = types.ModuleType("pytorch3d")
    stub_loss = types.ModuleType("pytorch3d.loss")
    stub_loss.chamfer_distance = oracle.chamfer_distance
    stub.loss = stub_loss
    sys.modules["pytorch3d"] = stub
    sys.modules["pytorch3d.loss"] = stub_loss
    sys.path.insert(0, REF)
    import importlib
    ref_models = importlib.import_module("models")
    ref_loss = importlib.import_module("loss")
    ref_sem = importlib.import_module("ExternalInfoLoaders.SemanticsLoader")
    ref_inst = importlib.import_module("ExternalInfoLoaders.InstanceSegmentationLoader")
    return ref_models, ref_loss, ref_sem, ref_inst


def digest(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest() + f":{a.dtype}:{'x'.join(map(str, a.shape))}"


SEM_MODES = ["glove-25d-ade20k-places", "glove-25d-ade20k-places-random", "glove-25d-ade20k-places-human-sizes",
             "glove-25d-ade20k-places-size_shuffled", "glove-25d", "glove-25d-inst-areas", "glove", "raw"]
INST_MODES = ["ade20k_swin", "ade20k_swin_human_sizes", "ade20k_swin_bbox_human_sizes_shuffled", "coco"]


def golden_loaders(ref_sem, ref_inst, out):
    """A1/A2: digests (bit-exact contract) of every loader mode on a 2x1x48x64 label map."""
    os.chdir(REF)  # loaders open "data/*.npy" relative to the cwd
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        for mode in SEM_MODES:
            places = "ade20k-places" in mode
            lab, _ = synthetic.label_maps(2, 48, 64, seed=11, lo=-1 if places else 0, hi=100 if places else 149,
                                          inject=(-7, 101, 255, 1000) if places else ())
            loader = ref_sem.SemanticsLoader(Namespace(use_semantics=mode))
            raw, sem = loader.get_semantics({"semantics": lab.clone()})
            out[f"sem/{mode}/raw"] = digest(raw.numpy())
            out[f"sem/{mode}/out"] = digest(sem.numpy())
        for mode in INST_MODES:
            hi = 80 if mode == "coco" else 100
            lab, areas = synthetic.label_maps(2, 48, 64, seed=12, lo=-1, hi=hi, inject=(-7, 101, 255, 1000))
            loader = ref_inst.InstanceSegmentationLoader(Namespace(use_instance_segmentation=mode))
            raw, emb, ar = loader.get_instance_segmentation({"instance_labels": lab.clone(), "instance_areas": areas.clone()})
            out[f"inst/{mode}/raw"] = digest(raw.numpy())
            out[f"inst/{mode}/emb"] = digest(emb.numpy())
            out[f"inst/{mode}/areas"] = digest(ar.numpy())
    finally:
        torch.Tensor.cuda = real_cuda
        os.chdir(ROOT)


def _ref_model(ref_models, encoder_name="efficientnet-b1", **kw):
    bb = "tf_efficientnet_b5_ap" if "b5" in encoder_name else "tf_efficientnet_b1_ap"
    backbone = build_backbone(bb, seed=0)
    backbone.global_pool = torch.nn.Identity()
    backbone.classifier = torch.nn.Identity()
    add = ref_models.UnetAdaptiveBins.get_num_channels_to_add(
        encoder_name, kw.get("semantics_mode"), kw.get("instance_segmentation_mode"), kw.get("image", "rgb"))
    if kw.get("insertion_point") == "input" and add:
        from mde_biological_vision_systems_b200.models.efficientnet import SamePadConv2d
        backbone.conv_stem = SamePadConv2d(3 + add, 32, 3, 2)
    m = ref_models.UnetAdaptiveBins(backbone, n_bins=256, min_val=1e-3, max_val=10, norm="linear",
                                    encoder_name=encoder_name, **kw)
    synthetic.fill_state_dict(m, seed=5)
    return m.eval()


def golden_head(ref_models, out):
    """A4-A8: the reference mViT + conv_out + bin pipeline on a synthetic unet_out [2,128,176,192]."""
    m = _ref_model(ref_models, insertion_point="input", semantics_mode=None, instance_segmentation_mode=None)
    x = synthetic.decoder_features(2, 128, 176, 192, seed=21)
    with torch.no_grad():
        tgt = m.adaptive_bins_layer.patch_transformer(x.clone())
        widths, ram = m.adaptive_bins_layer(x)
        sm = m.conv_out(ram)
        w = (m.max_val - m.min_val) * widths
        w = torch.nn.functional.pad(w, (1, 0), mode="constant", value=m.min_val)
        edges = torch.cumsum(w, dim=1)
        centers = 0.5 * (edges[:, :-1] + edges[:, 1:])
        pred = torch.sum(sm * centers.view(2, 256, 1, 1), dim=1, keepdim=True)
    out["head/tgt"] = tgt.numpy()
    out["head/widths"] = widths.numpy()
    out["head/edges"] = edges.numpy()
    out["head/pred"] = pred.numpy()
    out["head/ram_sub"] = ram[:, :, ::16, ::16].contiguous().numpy()  # subsampled range-attention maps


def golden_full(ref_models, out):
    """Whole model (config 1 shape family, reduced to 352x384 so 11x12 = 132 >= 129 tokens) incl. backbone."""
    m = _ref_model(ref_models, insertion_point="input", semantics_mode=None, instance_segmentation_mode=None)
    x = synthetic.image(1, 352, 384, seed=31)
    with torch.no_grad():
        edges, pred = m(x)
    out["full/edges"] = edges.numpy()
    out["full/pred"] = pred.numpy()


def golden_insertion(ref_models, ref_sem, ref_inst, out):
    """A3: tensor entering the encoder for configs 2 and 3 (+ inst-areas / human-sizes semantics)."""
    cases = {
        "cfg2": dict(semantics_mode="glove-25d-ade20k-places", instance_segmentation_mode=None),
        "cfg3": dict(semantics_mode="glove-25d", instance_segmentation_mode="ade20k_swin_human_sizes"),
        "areas": dict(semantics_mode="glove-25d-inst-areas", instance_segmentation_mode="coco"),
        "hsizes": dict(semantics_mode="glove-25d-ade20k-places-human-sizes", instance_segmentation_mode="ade20k_swin"),
    }
    os.chdir(REF)
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        for name, kw in cases.items():
            m = _ref_model(ref_models, insertion_point="input", image="rgb", **kw)
            captured = {}
            m.encoder.register_forward_pre_hook(lambda mod, inp: captured.__setitem__("x", inp[0].detach().clone()))
            h, w = 32, 32
            img = synthetic.image(2, h, w, seed=41)
            places = "ade20k-places" in kw["semantics_mode"]
            slab, _ = synthetic.label_maps(2, h, w, seed=42, lo=-1 if places else 0, hi=100 if places else 149,
                                           inject=(-7, 101, 255, 1000) if places else (), n_rect=(5, 10))
            sem_loader = ref_sem.SemanticsLoader(Namespace(use_semantics=kw["semantics_mode"]))
            _, sem = sem_loader.get_semantics({"semantics": slab.clone()})
            args = dict(x=img, semantics=sem)
            if kw["instance_segmentation_mode"]:
                hi = 80 if kw["instance_segmentation_mode"] == "coco" else 100
                ilab, iar = synthetic.label_maps(2, h, w, seed=43, lo=-1, hi=hi, n_rect=(5, 10))
                il = ref_inst.InstanceSegmentationLoader(Namespace(use_instance_segmentation=kw["instance_segmentation_mode"]))
                _, emb, ar = il.get_instance_segmentation({"instance_labels": ilab.clone(), "instance_areas": iar.clone()})
                args.update(instance_labels=emb, instance_areas=ar)
            try:
                with torch.no_grad():
                    m(**args)
            except Exception:  # 32x32 is too small for the head; the pre-hook has already fired
                pass
            out[f"insert/{name}"] = captured["x"].numpy()
    finally:
        torch.Tensor.cuda = real_cuda
        os.chdir(ROOT)


def golden_losses(ref_loss, out):
    """A9/A10: SILogLoss and BinsChamferLoss (reference wrapper; pytorch3d call restated)."""
    silog, chamfer = ref_loss.SILogLoss(), ref_loss.BinsChamferLoss()
    rng = np.random.default_rng(51)
    for name, (b, h, w) in {"a": (2, 104, 136), "b": (3, 64, 96)}.items():
        depth = synthetic.depth(b, h, w, seed=52)
        pred = torch.from_numpy((0.3 + 9 * rng.random((b, 1, h // 2, w // 2), dtype=np.float32)).astype(np.float32))
        mask = depth > 1e-3
        widths = torch.from_numpy(rng.random((b, 256), dtype=np.float32) + 0.1)
        widths = widths / widths.sum(1, keepdim=True) * (10 - 1e-3)
        edges = torch.cumsum(torch.nn.functional.pad(widths, (1, 0), value=1e-3), dim=1)
        out[f"loss/{name}/pred"] = pred.numpy()
        out[f"loss/{name}/edges"] = edges.numpy()
        out[f"loss/{name}/silog"] = silog(pred, depth, mask=mask, interpolate=True).numpy()
        out[f"loss/{name}/silog_nomask_noint"] = silog(
            torch.nn.functional.interpolate(pred, depth.shape[-2:], mode="nearest"), depth.clamp_min(0.2),
            mask=None, interpolate=False).numpy()
        out[f"loss/{name}/chamfer"] = chamfer(edges, depth).numpy()
    # all-valid and nearly-all-invalid edge cases
    depth = synthetic.depth(2, 64, 96, seed=53, all_valid=True)
    depth[1, :, 1:, :] = 0.0
    depth[1, :, 0, 5:] = 0.0  # 5 valid pixels in image 1
    pred = torch.from_numpy((0.3 + 9 * rng.random((2, 1, 32, 48), dtype=np.float32)).astype(np.float32))
    edges = torch.linspace(1e-3, 10, 257).repeat(2, 1).contiguous()
    out["loss/edge/pred"] = pred.numpy()
    out["loss/edge/silog"] = silog(pred, depth, mask=depth > 1e-3, interpolate=True).numpy()
    out["loss/edge/chamfer"] = chamfer(edges, depth).numpy()


def golden_state_keys(ref_models):
    """state_dict key/shape contract of the reference (model_io.py:36-72 loads by key)."""
    cases = {
        "b1_plain": dict(encoder_name="efficientnet-b1", insertion_point="input", semantics_mode=None,
                         instance_segmentation_mode=None),
        "b1_cfg3": dict(encoder_name="efficientnet-b1", insertion_point="input", semantics_mode="glove-25d",
                        instance_segmentation_mode="ade20k_swin_human_sizes"),
        "b1_areas_before_attn": dict(encoder_name="efficientnet-b1", insertion_point="before-attn",
                                     semantics_mode="glove-25d-inst-areas", instance_segmentation_mode="coco"),
        "b5_plain": dict(encoder_name="efficientnet-b5", insertion_point="before-attn", semantics_mode=None,
                         instance_segmentation_mode=None),
        "b1_noadabins": dict(encoder_name="efficientnet-b1-noAdaBins", insertion_point="input", semantics_mode=None,
                             instance_segmentation_mode=None),
    }
    with open(os.path.join(HERE, "golden_state_keys.txt"), "w") as f:
        for name, kw in cases.items():
            m = _ref_model(ref_models, **kw)
            for k, v in m.state_dict().items():
                if k.startswith("encoder."):
                    continue  # backbone keys are geffnet's, not the reference's
                f.write(f"{name} {k} {'x'.join(map(str, v.shape)) or 'scalar'}\n")


def main():
    torch.set_grad_enabled(False)
    ref_models, ref_loss, ref_sem, ref_inst = _import_reference()
    digests, arrays = {}, {}
    golden_state_keys(ref_models)
    golden_loaders(ref_sem, ref_inst, digests)
    golden_head(ref_models, arrays)
    golden_full(ref_models, arrays)
    golden_insertion(ref_models, ref_sem, ref_inst, arrays)
    golden_losses(ref_loss, arrays)
    np.savez_compressed(os.path.join(HERE, "golden_arrays.npz"), **{k.replace("/", "__"): v for k, v in arrays.items()})
    with open(os.path.join(HERE, "golden_digests.txt"), "w") as f:
        for k in sorted(digests):
            f.write(f"{k} {digests[k]}\n")
    for k, v in arrays.items():
        print(k, v.dtype, v.shape)
    print(len(digests), "digests")


if __name__ == "__main__":
    main()
