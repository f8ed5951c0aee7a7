"""Hot-path-only workload for ncu: gather -> mViT head -> SILog + chamfer on a fixed synthetic unet_out (config 2
shapes, B=16, 416x544).  Usage: python scripts/profile_head.py [iters]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mde_biological_vision_systems_b200 import synthetic  # noqa: E402
from mde_biological_vision_systems_b200.ExternalInfoLoaders.SemanticsLoader import SemanticsLoader  # noqa: E402
from mde_biological_vision_systems_b200.loss import DepthLosses  # noqa: E402
from mde_biological_vision_systems_b200.models import UnetAdaptiveBins  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B, H, W = 16, 416, 544
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = UnetAdaptiveBins.build(n_bins=256, min_val=1e-3, max_val=10.0, norm="linear", encoder_name="efficientnet-b1",
                               semantics_mode="glove-25d-ade20k-places", instance_segmentation_mode=None,
                               insertion_point="input", image="rgb").to(dev).eval()
loader = SemanticsLoader(argparse.Namespace(use_semantics="glove-25d-ade20k-places"), device=dev)
loader.bind_encoder_input(model)
both = DepthLosses(1e-3)
batch = {"semantics": synthetic.label_maps(B, H, W, seed=2)[0].to(dev)}
depth = synthetic.depth(B, H, W, seed=1).to(dev)
unet_out = synthetic.decoder_features(B, 128, H // 2, W // 2, seed=3).to(dev)
with torch.no_grad():
    for _ in range(iters):
        _, sem = loader.get_semantics(batch)
        edges, pred = model._head(unet_out)
        l1, l2 = both(pred, edges, depth, interpolate=True)
torch.cuda.synchronize()
print("ok", float(l1), float(l2))
