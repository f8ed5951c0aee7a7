"""Kernel-time table (torch.profiler / CUPTI) of one inference step and one training step of BASELINE config 2
(B = 16, 416x544).  Usage: python scripts/profile_step.py [infer|train|both] [batch]
Shows where the time of the whole public-API step goes (backbone passthrough vs the hand-written hot path)."""
import argparse
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mde_biological_vision_systems_b200 import synthetic  # noqa: E402
from mde_biological_vision_systems_b200.ExternalInfoLoaders.SemanticsLoader import SemanticsLoader  # noqa: E402
from mde_biological_vision_systems_b200.loss import BinsChamferLoss, SILogLoss  # noqa: E402
from mde_biological_vision_systems_b200.models import UnetAdaptiveBins  # noqa: E402
from mde_biological_vision_systems_b200.training import TrainStep  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "both"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
H, W = 416, 544
MODE = "glove-25d-ade20k-places"
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = UnetAdaptiveBins.build(n_bins=256, min_val=1e-3, max_val=10.0, norm="linear", encoder_name="efficientnet-b1",
                               semantics_mode=MODE, instance_segmentation_mode=None, insertion_point="input",
                               image="rgb").to(dev)
loader = SemanticsLoader(argparse.Namespace(use_semantics=MODE), device=dev)
silog, chamfer = SILogLoss(), BinsChamferLoss()
batch = {"image": synthetic.image(B, H, W, seed=0).to(dev), "depth": synthetic.depth(B, H, W, seed=1).to(dev),
         "semantics": synthetic.label_maps(B, H, W, seed=2)[0].to(dev)}


def table(prof, title, rows=45):
    evs = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
    total = sum(e.device_time_total for e in evs)
    print(f"== {title}: {total / 1e3:.2f} ms of kernels, {sum(e.count for e in evs)} launches")
    for e in sorted(evs, key=lambda e: -e.device_time_total)[:rows]:
        print(f"{e.device_time_total / 1e3:9.3f} ms {100 * e.device_time_total / total:5.1f}%  x{e.count:<4d} {e.key[:110]}")


if what in ("infer", "both"):
    # the benchmarked mode: loader bound to the encoder input, fused losses (bench.build_gpu / bench.infer_step_fn)
    import bench
    cfg = dict(bench.CONFIGS[2], batch=B)
    ctx = bench.Ctx()
    bmodel, sem_loader, inst_loader = bench.build_gpu(cfg, ctx)
    bmodel.eval()
    step = bench.infer_step_fn(cfg, bmodel, sem_loader, inst_loader, ctx.dev)
    hb = bench.host_batch(cfg, B)
    dbatch = {k: v.to(dev) for k, v in hb.items()}
    for _ in range(3):
        step(dbatch)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step(dbatch)
        torch.cuda.synchronize()
    table(prof, f"inference step B={B} (bench mode)", rows=60)

if what in ("train", "both"):
    model.train()
    if os.environ.get("MDE_EXPERIMENT_CL") == "1":
        # experiment: whole model in channels_last with the stock resize + cat (how much do cuDNN's NHWC kernels buy?)
        import torch.nn.functional as F
        from mde_biological_vision_systems_b200.models import unet_adaptive_bins as uab

        def stock_forward(self, x, concat_with):
            x = F.interpolate(x, size=concat_with.shape[-2:], mode='bilinear', align_corners=True)
            return self._net(torch.cat((x, concat_with), dim=1))

        uab.UpSampleBN.forward = stock_forward
        model = model.to(memory_format=torch.channels_last)
        batch["image"] = batch["image"].contiguous(memory_format=torch.channels_last)
    stepper = TrainStep(model, semantics_loader=loader, total_steps=1000)
    for _ in range(3):
        stepper(batch, dev)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        stepper(batch, dev)
        torch.cuda.synchronize()
    table(prof, f"training step B={B}", rows=60)
