"""Kernel-level profile of one single-GPU training step (config 2 shapes): python scripts/profile_train.py [batch]"""
import argparse, os, sys, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mde_biological_vision_systems_b200.training import TrainStep

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cfg = bench.CONFIGS[2]
ctx = bench.Ctx()
model, sem_loader, inst_loader = bench.build_gpu(cfg, ctx)
model.train()
impl = os.environ.get("MDE_TRAIN_CONV", "tc")
for m in model.modules():
    if hasattr(m, "train_conv_impl"):
        m.train_conv_impl = impl
model.adaptive_bins_layer.train_conv_impl = impl
stepper = TrainStep(model, semantics_loader=sem_loader, total_steps=100)
host = bench.host_batch(cfg, batch, 0, pin=True)
for _ in range(3):
    stepper(host, ctx.dev)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    stepper(host, ctx.dev)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        agg[e.name[:120]][0] += 1
        agg[e.name[:120]][1] += e.device_time
tot = sum(v for _, v in agg.values())
ours = sum(v for k, (_, v) in agg.items() if "mde::" in k or "tc::" in k)
print(f"== train step ({impl} convs), B={batch}: {tot/1000:.2f} ms of kernels, {sum(c for c,_ in agg.values())} launches, mde kernels {100*ours/tot:.1f} %")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{v/1000:9.3f} ms {100*v/tot:5.1f}% x{c:<4d} {k}")
