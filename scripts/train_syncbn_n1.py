"""Where does the N > 1 training step lose time against N = 1?  One GPU, no communication: the config-2 training step with stock
BatchNorm (what N = 1 runs) vs the SyncBatchNorm2d kernels forced on at world size 1 (same kernels and launch path as N > 1, the
all-reduce / peer exchange being the identity).  The difference is the cost of the kernel path itself; what remains of the N > 1
gap is communication and rank skew (bench.py: exposed_allreduce_ms, syncbn_wait_ms).
usage: python scripts/train_syncbn_n1.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mde_biological_vision_systems_b200 import parallel  # noqa: E402
from mde_biological_vision_systems_b200.training import TrainStep  # noqa: E402

cfg = bench.CONFIGS[2]
ctx = bench.Ctx()
for mode in ("stock", "kernels"):
    model, sem_loader, inst_loader = bench.build_gpu(cfg, ctx)
    if mode == "kernels":
        model = parallel._convert_sync_batchnorm(model)
        n = 0
        for m in model.modules():
            if isinstance(m, parallel.SyncBatchNorm2d):
                m.force_kernels = True
                n += 1
    model.train()
    stepper = TrainStep(model, semantics_loader=sem_loader, instance_loader=inst_loader, total_steps=1000)
    host = bench.host_batch(cfg, cfg["batch"], 0, pin=True)
    for _ in range(3):
        stepper(host, ctx.dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(5):
        loss = stepper(host, ctx.dev)
    e1.record()
    t_cpu = (time.perf_counter() - t0) / 5 * 1e3  # host time to ENQUEUE a step (no sync inside): > GPU time means launch-bound
    torch.cuda.synchronize()
    print(f"{mode:8s}: {e0.elapsed_time(e1) / 5:.2f} ms per step on the GPU, host enqueue {t_cpu:.2f} ms per step, loss {float(loss):.4f}")
    del stepper, model
    torch.cuda.empty_cache()
