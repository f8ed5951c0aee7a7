"""One B200: the BASELINE config-2 inference step (graph replay and eager) under every programmatic-dependent-launch mask
(mde_set_pdl: 1 = chains of small kernels, 2 = persistent tcgen05 kernels, 4 = streaming kernels).  Prints one JSON line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mde_biological_vision_systems_b200 import _lib  # noqa: E402
from mde_biological_vision_systems_b200.graphs import GraphedStep  # noqa: E402


def main():
    ctx = bench.Ctx()
    cfg = bench.CONFIGS[2]
    batch = cfg["batch"]
    model, sem_loader, inst_loader = bench.build_gpu(cfg, ctx)
    model.eval()
    step = bench.infer_step_fn(cfg, model, sem_loader, inst_loader, ctx.dev)
    host = bench.host_batch(cfg, batch, 0, pin=True)
    resident = {k: v.to(ctx.dev) for k, v in host.items()}
    lib = _lib.load()
    for _ in range(3):
        step(resident)
    ref = float(step(resident))
    out = {}
    masks = [int(m) for m in os.environ.get("MASKS", "0,1,2,4,5,3,7,0").split(",")]
    for i, mask in enumerate(masks):
        lib.mde_set_pdl(mask)
        g = GraphedStep(lambda **kw: step(kw), resident, warmup=1)
        got = float(g(**resident))
        ms_g = ctx.timed(lambda: g(**resident), 20) / 20
        ms_e = ctx.timed(lambda: step(resident), 10) / 10
        out[f"{i}:mask{mask}"] = {"graph_ms": round(ms_g, 4), "eager_ms": round(ms_e, 4), "loss_ok": abs(got - ref) <= 1e-6 * abs(ref)}
        del g
    print(json.dumps(out))


if __name__ == "__main__":
    main()
