"""One inference step of a BASELINE config in the benchmarked mode (bound loader, fused losses, default flags), bracketed by
cudaProfilerStart/Stop so that `ncu --profile-from-start off` captures exactly that step:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python scripts/ncu_step.py [config]
The same script without ncu prints the event-timed duration of the step (the number the launch list's total is compared with)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

cfg = bench.CONFIGS[int(sys.argv[1]) if len(sys.argv) > 1 else 2]
ctx = bench.Ctx()
model, sem_loader, inst_loader = bench.build_gpu(cfg, ctx)
model.eval()
step = bench.infer_step_fn(cfg, model, sem_loader, inst_loader, ctx.dev)
hb = bench.host_batch(cfg, cfg["batch"])
batch = {k: v.to(ctx.dev) for k, v in hb.items()}
for _ in range(3):
    step(batch)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.cudart().cudaProfilerStart()
e0.record()
loss = step(batch)
e1.record()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(f"{cfg['name'][:18]}: one eager step {e0.elapsed_time(e1):.3f} ms, loss {float(loss):.6f}")
