"""Kernel-time table of one DDP training step (rank 0) -- launched with torchrun, one rank per GPU.
Usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/profile_train_ddp.py [stock|ours]"""
import argparse
import os
import sys

import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mde_biological_vision_systems_b200 import synthetic  # noqa: E402
from mde_biological_vision_systems_b200.ExternalInfoLoaders.SemanticsLoader import SemanticsLoader  # noqa: E402
from mde_biological_vision_systems_b200.models import UnetAdaptiveBins  # noqa: E402
from mde_biological_vision_systems_b200.training import TrainStep  # noqa: E402

impl = sys.argv[1] if len(sys.argv) > 1 else "stock"
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, H, W, MODE = 16, 416, 544, "glove-25d-ade20k-places"
torch.manual_seed(0)
model = UnetAdaptiveBins.build(n_bins=256, min_val=1e-3, max_val=10.0, norm="linear", encoder_name="efficientnet-b1",
                               semantics_mode=MODE, instance_segmentation_mode=None, insertion_point="input", image="rgb").to(dev)
if impl == "stock":
    model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
elif impl in ("ours", "p2p"):
    from mde_biological_vision_systems_b200 import parallel
    model = parallel.convert_sync_batchnorm(model, p2p=(impl == "p2p"))
model.train()
loader = SemanticsLoader(argparse.Namespace(use_semantics=MODE), device=dev)
batch = {"image": synthetic.image(B, H, W, seed=10 * rank).to(dev), "depth": synthetic.depth(B, H, W, seed=10 * rank + 1).to(dev),
         "semantics": synthetic.label_maps(B, H, W, seed=10 * rank + 2)[0].to(dev)}
stepper = TrainStep(model, semantics_loader=loader, total_steps=1000)
for _ in range(3):
    stepper(batch, dev)
dist.barrier()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    loss = stepper(batch, dev)
e.record()
torch.cuda.synchronize()
wall = s.elapsed_time(e) / 5
if len(sys.argv) > 2 and sys.argv[2] == "quick":
    wall_all = torch.tensor([wall], device=dev)
    dist.all_reduce(wall_all, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"== {impl} SyncBN, world {world}: {float(wall_all):.2f} ms/step (max over ranks), {world * B / float(wall_all) * 1e3:.1f} img/s, loss {float(loss):.4f}")
    dist.destroy_process_group()
    sys.exit(0)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    stepper(batch, dev)
    torch.cuda.synchronize()
if rank == 0:
    evs = [ev for ev in prof.key_averages() if ev.device_time_total > 0 and ev.device_type.name == "CUDA"]
    total = sum(ev.device_time_total for ev in evs)
    print(f"== {impl} SyncBN, world {world}: {wall:.2f} ms/step wall, {total / 1e3:.2f} ms of kernels, {sum(ev.count for ev in evs)} launches, loss {float(loss):.4f}")
    for ev in sorted(evs, key=lambda ev: -ev.device_time_total)[:28]:
        print(f"{ev.device_time_total / 1e3:9.3f} ms {100 * ev.device_time_total / total:5.1f}%  x{ev.count:<4d} {ev.key[:120]}")
dist.destroy_process_group()
