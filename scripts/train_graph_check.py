"""torchrun -- N-rank check of training.GraphedTrainStep against the eager TrainStep (SyncBatchNorm over peer memory with the
device-resident epoch counter, one all-reduce on the gradient arena after each replay): same loss trajectory, identical
parameters on every rank afterwards, and the step times of both launch modes.  Prints one JSON line on rank 0."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mde_biological_vision_systems_b200 import parallel  # noqa: E402
from mde_biological_vision_systems_b200.training import GraphedTrainStep, TrainStep  # noqa: E402


def main():
    ctx = bench.Ctx()
    cfg = bench.CONFIGS[int(os.environ.get("CFG", "2"))]
    batch = int(os.environ.get("BATCH", "8"))
    steps = int(os.environ.get("STEPS", "4"))
    host = bench.host_batch(cfg, batch, ctx.rank, pin=True)
    if ctx.world > 1:  # a protocol bug must end this check in seconds, not after the production bound of 600 s
        from mde_biological_vision_systems_b200 import _lib
        _lib.load().mde_bn_set_peer_timeout_seconds(20.0)
    out = {"world": ctx.world, "batch_per_gpu": batch}
    for kind in ("eager", "graph"):
        torch.manual_seed(0)
        model, sem_loader, inst_loader = bench.build_gpu(cfg, ctx)
        if ctx.world > 1:
            model = parallel.convert_sync_batchnorm(model, p2p=True)
        model.train()
        for mod in model.modules():  # the two launch modes draw different dropout masks
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
            if isinstance(mod, torch.nn.MultiheadAttention):
                mod.dropout = 0.0
        kw = dict(semantics_loader=sem_loader, instance_loader=inst_loader, total_steps=1000)
        if kind == "eager":
            st = TrainStep(model, **kw)
            losses = [float(st(host, ctx.dev)) for _ in range(3 + steps)]
        else:
            st = GraphedTrainStep(model, host, ctx.dev, warmup=3, **kw)
            losses = [None] * 3 + [float(st(host)) for _ in range(steps)]
        if kind == "graph":  # with the next batch's H2D copy overlapped (GraphedTrainStep.prefetch), as bench.py runs it
            def one():
                st() if st._staged else st(host)
                st.prefetch(host)
            ms = ctx.timed(one, 5) / 5
            out["graph_no_prefetch_ms"] = round(ctx.timed(lambda: st(host), 5) / 5, 2)
        else:
            ms = ctx.timed(lambda: st(host, ctx.dev), 5) / 5
        # replicas must stay identical: compare a parameter checksum across ranks
        chk = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        if ctx.world > 1:
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        out[kind] = {"losses": [None if v is None else round(v, 5) for v in losses], "ms": round(ms, 2),
                     "imgs_s": round(ctx.world * batch / (ms * 1e-3), 1), "replicas_equal": bool(float(hi - lo) == 0.0),
                     "exposed_allreduce_ms": st.averager.last_exposed_wait_ms() if ctx.world > 1 else None}
        del st, model
        torch.cuda.empty_cache()
    e, g = out["eager"]["losses"], out["graph"]["losses"]
    out["max_rel_loss_diff"] = max(abs(a - b) / abs(a) for a, b in zip(e[3:], g[3:]))
    if ctx.rank == 0:
        print(json.dumps(out))
    if ctx.world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
