"""Shared builders for the parity tests (weights and inputs come from the seeded numpy streams of
mde_biological_vision_systems_b200.synthetic, exactly as tests/golden/make_golden.py used them)."""
import hashlib
import os

import numpy as np
import torch

from mde_biological_vision_systems_b200 import synthetic
from mde_biological_vision_systems_b200.models import UnetAdaptiveBins
from mde_biological_vision_systems_b200.models.efficientnet import SamePadConv2d, build_backbone

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SEM_MODES = ["glove-25d-ade20k-places", "glove-25d-ade20k-places-random", "glove-25d-ade20k-places-human-sizes",
             "glove-25d-ade20k-places-size_shuffled", "glove-25d", "glove-25d-inst-areas", "glove", "raw"]
INST_MODES = ["ade20k_swin", "ade20k_swin_human_sizes", "ade20k_swin_bbox_human_sizes_shuffled", "coco"]


def digest(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest() + f":{a.dtype}:{'x'.join(map(str, a.shape))}"


def make_model(encoder_name="efficientnet-b1", seed=5, **kw):
    """Product model built exactly like make_golden._ref_model builds the reference one (same backbone seed, same
    state_dict fill), so both hold identical weights."""
    bb = "tf_efficientnet_b5_ap" if "b5" in encoder_name else "tf_efficientnet_b1_ap"
    backbone = build_backbone(bb, seed=0)
    backbone.global_pool = torch.nn.Identity()
    backbone.classifier = torch.nn.Identity()
    add = UnetAdaptiveBins.get_num_channels_to_add(encoder_name, kw.get("semantics_mode"),
                                                   kw.get("instance_segmentation_mode"), kw.get("image", "rgb"))
    if kw.get("insertion_point") == "input" and add:
        backbone.conv_stem = SamePadConv2d(3 + add, 32, 3, 2)
    m = UnetAdaptiveBins(backbone, n_bins=256, min_val=1e-3, max_val=10, norm="linear", encoder_name=encoder_name, **kw)
    synthetic.fill_state_dict(m, seed=seed)
    return m.eval()


def sem_labels(mode, b=2, h=48, w=64, seed=11, **kw):
    places = "ade20k-places" in mode
    return synthetic.label_maps(b, h, w, seed=seed, lo=-1 if places else 0, hi=100 if places else 149,
                                inject=(-7, 101, 255, 1000) if places else (), **kw)


def inst_labels(mode, b=2, h=48, w=64, seed=12, **kw):
    return synthetic.label_maps(b, h, w, seed=seed, lo=-1, hi=80 if mode == "coco" else 100,
                                inject=kw.pop("inject", (-7, 101, 255, 1000)), **kw)


def load_table(name):
    return np.load(os.path.join(ROOT, "data", name))


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-12)))


def rel_stats(a, b):
    """(max, 99.9th percentile) of the element-wise relative error."""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    r = np.abs(a - b) / np.maximum(np.abs(b), 1e-12)
    return float(r.max()), float(np.quantile(r, 0.999))
