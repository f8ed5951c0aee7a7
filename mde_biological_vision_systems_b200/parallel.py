"""Multi-GPU plumbing for the hot path: one process per GPU, batch sharded across ranks.

The forward path (loaders -> model -> SILog/chamfer) needs no collective: images are independent and the reference
computes its loss per rank on the local shard (train.py:414-426).  The only exchanges of the *training* path are the
gradient all-reduce that DistributedDataParallel performs (train.py:298-299) and SyncBatchNorm's statistics
(train.py:296, stock torch module).  ``GradientAverager`` is that all-reduce: gradients are packed into flat buckets,
summed with ``torch.distributed.all_reduce`` (NCCL over NVLink on the GPU box, gloo in the CPU tests) and divided by
the world size -- DDP's mean semantics -- bucket by bucket so the collective of one bucket overlaps the packing of the
next.
"""
import os

import torch
import torch.distributed as dist


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment (1 process if unset)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_indices(n_items, rank, world, pad=True):
    """DistributedSampler-style disjoint shard (dataloader.py:33-34): item i goes to rank i % world; with ``pad`` the
    list is wrapped around so every rank gets ceil(n/world) items (what DistributedSampler does without drop_last)."""
    per = -(-n_items // world)
    if not pad:
        return list(range(rank, n_items, world))
    idx = list(range(n_items))
    idx += idx[: per * world - n_items]
    return idx[rank: per * world: world]


def per_rank_batch(global_batch, world, use_new_batching=False):
    """train.py:286-291: by default the global --bs is divided by the GPUs; --use_new_batching keeps bs per GPU."""
    return global_batch if use_new_batching else max(1, global_batch // world)


def max_over_ranks(value, device=None):
    """Max of a python float over all ranks (used for the device-timed step duration)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class GradientAverager:
    """Bucketed mean all-reduce of ``.grad`` over all ranks (the collective of the DDP training path)."""

    def __init__(self, params, bucket_mb=25.0):
        self.params = [p for p in params if p.requires_grad]
        self.bucket_elems = max(1, int(bucket_mb * 1024 * 1024 / 4))

    def _buckets(self):
        cur, n = [], 0
        for p in reversed(self.params):  # gradients become ready roughly in reverse registration order
            if p.grad is None:
                continue
            cur.append(p)
            n += p.grad.numel()
            if n >= self.bucket_elems:
                yield cur
                cur, n = [], 0
        if cur:
            yield cur

    def reduce(self):
        if not (dist.is_available() and dist.is_initialized()):
            return 0
        world = dist.get_world_size()
        if world == 1:
            return 0
        pending = []
        for bucket in self._buckets():
            flat = torch.cat([p.grad.reshape(-1) for p in bucket])
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)
            pending.append((bucket, flat, work))
        total = 0
        for bucket, flat, work in pending:
            work.wait()
            flat.div_(world)
            off = 0
            for p in bucket:
                n = p.grad.numel()
                p.grad.copy_(flat[off:off + n].view_as(p.grad))
                off += n
            total += off
        return total  # elements exchanged
