"""Multi-GPU plumbing for the hot path: one process per GPU, batch sharded across ranks.

The forward path (loaders -> model -> SILog/chamfer) needs no collective: images are independent and the reference
computes its loss per rank on the local shard (train.py:414-426).  The only exchanges of the *training* path are the
gradient all-reduce that DistributedDataParallel performs (train.py:298-299) and SyncBatchNorm's statistics
(train.py:296).  ``GradientAverager`` is that all-reduce: every gradient lives in a flat per-bucket arena, each bucket is
summed in place with ``torch.distributed.all_reduce`` (NCCL over NVLink on the GPU box, gloo in the CPU tests) as soon as
autograd has filled it -- overlapped with the rest of the backward pass -- and divided by the world size (DDP's mean
semantics); ``broadcast_module_state`` is DDP's construction-time parameter / buffer broadcast.
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


def broadcast_module_state(module, src=0, group=None):
    """What DistributedDataParallel does at construction (train.py:298-299): rank ``src``'s parameters and buffers overwrite
    every other rank's, so replicas cannot silently start from different weights / BatchNorm statistics (different seeds,
    per-rank checkpoints).  One coalesced broadcast per dtype."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    tensors = [p.data for p in module.parameters()] + [b for b in module.buffers()]
    by_dtype = {}
    for t in tensors:
        by_dtype.setdefault((t.dtype, t.device), []).append(t)
    total = 0
    for ts in by_dtype.values():
        flat = torch.cat([t.reshape(-1) for t in ts])
        dist.broadcast(flat, src=src, group=group)
        off = 0
        for t in ts:
            n = t.numel()
            t.copy_(flat[off:off + n].view(t.shape))
            off += n
        total += off
    return total


class GradientAverager:
    """Bucketed mean all-reduce of the gradients over all ranks -- the collective DistributedDataParallel adds to the
    training path (train.py:298-299) -- overlapped with the backward pass:

    * fixed layout: ALL trainable parameters, in reverse registration order (the order their gradients become ready: head ->
      decoder -> encoder), are packed into buckets of ``bucket_mb``; every rank therefore issues the same collectives in the
      same order whatever subset of parameters received a gradient (a parameter that got none contributes zeros, which is
      what DDP's find_unused_parameters=True does in the reference);
    * gradient arena: each bucket is ONE flat tensor and every ``p.grad`` is a view into it, so autograd accumulates straight
      into the communication buffer -- no packing, no copy-back, the all-reduce is in place;
    * overlap: a post-accumulate hook per parameter counts the bucket down and launches its asynchronous all-reduce the
      moment the bucket is complete (in bucket order), while autograd is still producing the gradients of the layers below;
      ``reduce()`` after backward only launches what is left, waits and scales by 1 / world.

    Use ``zero_grad()`` of this object (one fill per bucket) instead of ``optimizer.zero_grad(set_to_none=True)``, which
    would detach the views."""

    def __init__(self, params, bucket_mb=25.0, group=None, overlap=True):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.overlap = overlap
        self.bucket_elems = max(1, int(bucket_mb * 1024 * 1024 / 4))
        self.buckets = []   # dicts: params, flat, views, pending (per step), work
        cur, n = [], 0
        for p in reversed(self.params):
            cur.append(p)
            n += p.numel()
            if n >= self.bucket_elems:
                self.buckets.append(self._make_bucket(cur))
                cur, n = [], 0
        if cur:
            self.buckets.append(self._make_bucket(cur))
        self._index = {}
        for bi, bk in enumerate(self.buckets):
            for p in bk["params"]:
                self._index[p] = bi
        self._next_launch = 0
        self._hooks = []
        self.exposed_wait_ms = None  # device-side time reduce() spent waiting for collectives (last call; needs CUDA)
        if self._active() and overlap:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self.attach()

    def _active(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    @staticmethod
    def _make_bucket(params):
        dtype, device = params[0].dtype, params[0].device
        total = sum(p.numel() for p in params)
        return {"params": list(params), "flat": torch.zeros(total, dtype=dtype, device=device), "views": None, "pending": 0,
                "work": None}

    def attach(self):
        """(Re-)point every p.grad at its slice of the bucket arena (memory order = the parameter's own: channels_last conv
        weights keep channels_last gradients)."""
        for bk in self.buckets:
            off = 0
            views = []
            for p in bk["params"]:
                n = p.numel()
                view = bk["flat"][off:off + n]
                if p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last) and not p.is_contiguous():
                    g = view.view(p.shape[0], p.shape[2], p.shape[3], p.shape[1]).permute(0, 3, 1, 2)
                else:
                    g = view.view(p.shape)
                if p.grad is None or p.grad.data_ptr() != g.data_ptr():
                    if p.grad is not None:  # constructed after a backward pass: keep what autograd already produced
                        g.copy_(p.grad)
                    p.grad = g
                views.append(p.grad)
                off += n
            bk["views"] = views
            bk["pending"] = len(bk["params"])
        self._next_launch = 0

    def zero_grad(self):
        """One fill per bucket.  The p.grad views are re-created only if something replaced them (e.g. an
        optimizer.zero_grad(set_to_none=True)): rebuilding 300+ views every step cost milliseconds of host time in a step that
        is host-launch-bound."""
        intact = True
        for bk in self.buckets:
            bk["flat"].zero_()
            bk["work"] = None
            bk["pending"] = len(bk["params"])
            views = bk["views"]
            if views is None or any(p.grad is not v for p, v in zip(bk["params"], views)):
                intact = False
        self._next_launch = 0
        if not intact:
            self.attach()

    def clip_grad_norm_(self, max_norm, eps=1e-6):
        """torch.nn.utils.clip_grad_norm_ (2-norm) over all parameters, on the bucket arena: a handful of launches instead of a
        multi-tensor pass over every parameter.  Same formula: coef = max_norm / (total + 1e-6), clamped to 1.  Returns the
        total norm (a 0-dim tensor; no host synchronisation)."""
        norms = torch.stack([torch.linalg.vector_norm(bk["flat"].float()) for bk in self.buckets])
        total = torch.linalg.vector_norm(norms)
        coef = torch.clamp(max_norm / (total + eps), max=1.0)
        for bk in self.buckets:
            bk["flat"].mul_(coef.to(bk["flat"].dtype))
        return total

    def _launch_ready(self, force=False):
        while self._next_launch < len(self.buckets):
            bk = self.buckets[self._next_launch]
            if bk["pending"] > 0 and not force:
                break
            bk["work"] = dist.all_reduce(bk["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._next_launch += 1

    def _on_grad(self, p):
        bk = self.buckets[self._index[p]]
        bk["pending"] -= 1
        if bk["pending"] == 0:
            self._launch_ready()

    def reduce(self):
        """Call after backward: launches the buckets that are still pending (parameters without a gradient this step), waits
        for all collectives and turns the sums into means.  Returns the number of elements exchanged."""
        if not self._active():
            return 0
        world = dist.get_world_size(self.group)
        cuda = self.buckets and self.buckets[0]["flat"].is_cuda
        if cuda:
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
        self._launch_ready(force=True)
        total = 0
        for bk in self.buckets:
            bk["work"].wait()
            bk["flat"].div_(world)
            bk["work"] = None
            bk["pending"] = len(bk["params"])
            total += bk["flat"].numel()
        # ready for the next backward pass even if zero_grad() is not called in between on the host (a replayed CUDA graph
        # performs the fill on the device: training.GraphedTrainStep)
        self._next_launch = 0
        if cuda:
            t1.record()
            self._wait_events = (t0, t1)
        return total

    def last_exposed_wait_ms(self):
        """Device time between the end of backward and the last bucket's mean being ready (call after a synchronize)."""
        ev = getattr(self, "_wait_events", None)
        return None if ev is None else ev[0].elapsed_time(ev[1])


# ---------------------------------------------------------------------------------------------------------------
# SyncBatchNorm on the B200 kernels (csrc/bn_sync.cu)
# ---------------------------------------------------------------------------------------------------------------
class _SyncBatchNormFn(torch.autograd.Function):
    """Training-mode batch norm with statistics over the global batch: two streaming kernels and ONE all-reduce of the raw
    per-channel moments per direction (nn.SyncBatchNorm: three kernels + an all_gather forward, two + an all_reduce
    backward, plus bookkeeping copies).  x must be a channels_last CUDA tensor with C % 4 == 0."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, group, bufs, p2p=None):
        from . import _lib, ops
        lib = _lib.load()
        b, c, h, w = x.shape
        n = b * h * w
        # bufs: the module's two pairs of float64 [2C+1] scratch rows (forward / backward), used alternately: the kernel
        # of call k zeroes the row of call k+1, so the loop issues no memsets
        stats, nxt = bufs.next_forward(c, x.device)
        if p2p is not None:
            # fused exchange over NVLink peer memory: the statistics kernel publishes into every peer's arena, the
            # normalise kernel waits for the world's partial sums -- no collective call between the two kernels
            arena, epoch, slot_off, flag_off = p2p.next(0)
            y = torch.empty_like(x, memory_format=torch.channels_last)
            save_mean = torch.empty(c, dtype=torch.float32, device=x.device)
            save_invstd = torch.empty(c, dtype=torch.float32, device=x.device)
            rc = lib.mde_bn_stats_p2p_nhwc(ops._p(x), n, c, ops._p(stats), ops._p(nxt), arena.ptrs, arena.world, arena.rank,
                                           slot_off, flag_off, epoch, ops._s())
            _lib.check(rc, "mde_bn_stats_p2p_nhwc")
            count = float(n * arena.world)
            rc = lib.mde_bn_apply_p2p_nhwc(ops._p(x), ops._p(y), n, c, arena.my_base, slot_off, flag_off, arena.world, epoch,
                                           count, ops._p(weight), ops._p(bias), float(eps), ops._p(save_mean),
                                           ops._p(save_invstd), ops._p(running_mean), ops._p(running_var), float(momentum),
                                           ops._s())
            _lib.check(rc, "mde_bn_apply_p2p_nhwc")
            ctx.save_for_backward(x, weight, save_mean, save_invstd)
            ctx.count, ctx.group, ctx.world, ctx.bufs, ctx.p2p = count, group, arena.world, bufs, p2p
            return y
        _lib.check(lib.mde_bn_stats_nhwc(ops._p(x), n, c, ops._p(stats), ops._p(nxt), ops._s()), "mde_bn_stats_nhwc")
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        if world > 1:
            if _COUNT_FROM_ALLREDUCE:
                stats[2 * c] = float(n)
            dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
            count = float(stats[2 * c].item()) if _COUNT_FROM_ALLREDUCE else float(n * world)
        else:
            count = float(n)
        y = torch.empty_like(x, memory_format=torch.channels_last)
        save_mean = torch.empty(c, dtype=torch.float32, device=x.device)
        save_invstd = torch.empty(c, dtype=torch.float32, device=x.device)
        rc = lib.mde_bn_apply_nhwc(ops._p(x), ops._p(y), n, c, ops._p(stats), count, ops._p(weight), ops._p(bias), float(eps),
                                   ops._p(save_mean), ops._p(save_invstd), ops._p(running_mean), ops._p(running_var),
                                   float(momentum), ops._s())
        _lib.check(rc, "mde_bn_apply_nhwc")
        ctx.save_for_backward(x, weight, save_mean, save_invstd)
        ctx.count, ctx.group, ctx.world, ctx.bufs, ctx.p2p = count, group, world, bufs, None
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import _lib, ops
        lib = _lib.load()
        x, weight, mean, invstd = ctx.saved_tensors
        b, c, h, w = x.shape
        n = b * h * w
        if not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.contiguous(memory_format=torch.channels_last)
        sums, nxt = ctx.bufs.next_backward(c, x.device)
        if ctx.p2p is not None:
            arena, epoch, slot_off, flag_off = ctx.p2p.next(1)
            rc = lib.mde_bn_bwd_reduce_p2p_nhwc(ops._p(x), ops._p(dy), n, c, ops._p(mean), ops._p(invstd), ops._p(sums),
                                                ops._p(nxt), arena.ptrs, arena.world, arena.rank, slot_off, flag_off, epoch,
                                                ops._s())
            _lib.check(rc, "mde_bn_bwd_reduce_p2p_nhwc")
            local = sums[:2 * c].float()
            dx = torch.empty_like(x, memory_format=torch.channels_last)
            rc = lib.mde_bn_bwd_apply_p2p_nhwc(ops._p(x), ops._p(dy), ops._p(dx), n, c, ops._p(mean), ops._p(invstd),
                                               ops._p(weight), arena.my_base, slot_off, flag_off, arena.world, epoch,
                                               ctx.count, ops._s())
            _lib.check(rc, "mde_bn_bwd_apply_p2p_nhwc")
            return dx, (local[c:] if weight is not None else None), (local[:c] if weight is not None else None), \
                None, None, None, None, None, None, None
        rc = lib.mde_bn_bwd_reduce_nhwc(ops._p(x), ops._p(dy), n, c, ops._p(mean), ops._p(invstd), ops._p(sums), ops._p(nxt),
                                        ops._s())
        _lib.check(rc, "mde_bn_bwd_reduce_nhwc")
        local = sums[:2 * c].float()  # d bias, d weight are the LOCAL sums (DDP averages parameter gradients afterwards)
        if ctx.world > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=ctx.group)
        dx = torch.empty_like(x, memory_format=torch.channels_last)
        rc = lib.mde_bn_bwd_apply_nhwc(ops._p(x), ops._p(dy), ops._p(dx), n, c, ops._p(mean), ops._p(invstd), ops._p(weight),
                                       ops._p(sums), ctx.count, ops._s())
        _lib.check(rc, "mde_bn_bwd_apply_nhwc")
        gw = local[c:] if weight is not None else None
        gb = local[:c] if weight is not None else None
        return dx, gw, gb, None, None, None, None, None, None, None


class P2PArena:
    """Symmetric-memory arena shared by the ranks of a process group (torch.distributed._symmetric_memory): every rank's
    buffer is mapped into every process, so kernels exchange data with plain loads / stores over NVLink."""

    def __init__(self, nbytes, group=None):
        import ctypes
        import torch.distributed._symmetric_memory as symm_mem
        group = dist.group.WORLD if group is None else group
        enable = getattr(symm_mem, "enable_symm_mem_for_group", None)
        if enable is not None:  # needed by older torch releases, a deprecated no-op on current ones
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", FutureWarning)
                try:
                    enable(group.group_name)
                except Exception:
                    pass
        self.buf = symm_mem.empty(int(nbytes), dtype=torch.uint8, device=torch.device("cuda", torch.cuda.current_device()))
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, group)
        self.graph_owned = False
        self.rank, self.world = int(self.hdl.rank), int(self.hdl.world_size)
        if self.world > 8:
            raise RuntimeError("P2PArena supports at most 8 ranks (one NVSwitch domain)")
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.ptrs = (ctypes.c_uint64 * 8)(*(ptrs + [0] * (8 - len(ptrs))))
        self.my_base = ptrs[self.rank]
        # CUDA-graph replay (training.GraphedTrainStep): a device-resident step counter the graph's first node increments; the
        # SyncBatchNorm kernels add it to the epoch base frozen into the graph (csrc/bn_sync.cu: bn_resolve_epoch)
        self.replay_counter = torch.zeros(1, dtype=torch.int64, device=self.buf.device)
        self.capture_base = None  # counter value + 1 at capture (None: eager launches)
        torch.cuda.synchronize()
        dist.barrier(group=group)  # every arena is zeroed before anyone publishes into it

    def begin_capture(self):
        """Call right before capturing a training step (host-synchronising): from here on `_P2PRegion.next` hands the kernels
        epoch BASES and parity-0 offsets, and the launches read the step counter on the device."""
        from . import _lib
        if self.graph_owned:
            raise RuntimeError("P2PArena: a training graph was already captured over this arena")
        self.capture_base = int(self.replay_counter.item()) + 1
        _lib.check(_lib.load().mde_bn_p2p_set_epoch_counter(self.replay_counter.data_ptr()), "mde_bn_p2p_set_epoch_counter")

    def end_capture(self, captured=True):
        from . import _lib
        _lib.check(_lib.load().mde_bn_p2p_set_epoch_counter(None), "mde_bn_p2p_set_epoch_counter")
        self.capture_base = None
        if captured:
            self.graph_owned = True  # eager SyncBatchNorm calls would now reuse epochs the replays consume


class _P2PRegion:
    """One SyncBatchNorm2d module's slice of the arena: [direction][epoch parity] x (world slots of 2C doubles) and, behind
    them, [direction][parity] x world epoch flags."""

    def __init__(self, arena, offset, channels):
        self.arena, self.c = arena, channels
        self.slot_bytes = arena.world * 2 * channels * 8
        self.slots0 = offset
        self.flags0 = offset + 4 * self.slot_bytes
        self.epochs = [0, 0]

    @staticmethod
    def nbytes(world, channels):
        return (4 * world * 2 * channels * 8 + 4 * world * 8 + 255) // 256 * 256

    def next(self, direction):
        base = getattr(self.arena, "capture_base", None)
        if getattr(self.arena, "graph_owned", False) and base is None:
            raise RuntimeError("SyncBatchNorm2d: the peer-memory epochs of this model belong to a captured training graph; "
                               "eager training-mode calls after the capture are not supported")
        self.epochs[direction] += 1
        e = self.epochs[direction]
        if base is not None:
            # captured launch: epoch = base + device step counter (== e at the first replay), parity resolved by the kernel
            k = 2 * direction
            return self.arena, e - base, self.slots0 + k * self.slot_bytes, \
                self.flags0 + k * self.arena.world * 8
        k = 2 * direction + (e & 1)
        return self.arena, e, self.slots0 + k * self.slot_bytes, self.flags0 + k * self.arena.world * 8


def enable_p2p_statistics(module, process_group=None):
    """Give every SyncBatchNorm2d of ``module`` a region of one symmetric-memory arena: the statistics are then exchanged by
    the kernels themselves over NVLink peer memory instead of one NCCL all-reduce per layer and direction."""
    mods = [m for m in module.modules() if isinstance(m, SyncBatchNorm2d)]
    world = dist.get_world_size(process_group)
    total, offs = 0, []
    for m in mods:
        offs.append(total)
        total += _P2PRegion.nbytes(world, m.num_features)
    arena = P2PArena(max(total, 256), process_group)
    for m, off in zip(mods, offs):
        m._p2p = _P2PRegion(arena, off, m.num_features)
    return arena


class _BnScratch:
    """Per-module float64 scratch rows for the raw moments: [2][2C+1] for the forward and for the backward, allocated
    zeroed once; each kernel zeroes the row the next call will accumulate into."""

    def __init__(self):
        self.fwd = self.bwd = None
        self.fi = self.bi = 0

    def _rows(self, cur, c, device):
        if cur is None or cur.shape[1] != 2 * c + 1 or cur.device != device:
            cur = torch.zeros((2, 2 * c + 1), dtype=torch.float64, device=device)
        return cur

    def next_forward(self, c, device):
        self.fwd = self._rows(self.fwd, c, device)
        self.fi ^= 1
        return self._take(self.fwd, self.fi)

    def next_backward(self, c, device):
        self.bwd = self._rows(self.bwd, c, device)
        self.bi ^= 1
        return self._take(self.bwd, self.bi)

    @staticmethod
    def _take(rows, i):
        if rows.is_cuda and torch.cuda.is_current_stream_capturing():
            # a replayed graph accumulates into the SAME row every step (the alternation is frozen at capture): zero it inside
            # the graph instead of relying on the previous call
            rows[i].zero_()
        return rows[i], rows[i ^ 1]


# every rank of this path holds the same per-GPU batch (weak scaling, train.py:286-287), so the global pixel count is
# n * world and needs no device->host read; set True for ragged last batches
_COUNT_FROM_ALLREDUCE = False


class SyncBatchNorm2d(torch.nn.BatchNorm2d):
    """Drop-in for nn.SyncBatchNorm on 4-D inputs (same parameters, buffers and state_dict keys).  Training mode with a
    process group of more than one rank -- or ``force_kernels`` -- runs the B200 kernels; eval mode is plain batch norm
    with the running statistics, exactly as nn.SyncBatchNorm does."""

    def __init__(self, *args, process_group=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.process_group = process_group
        self.force_kernels = False
        self._scratch = _BnScratch()
        self._p2p = None  # set by enable_p2p_statistics

    def forward(self, x):
        world = dist.get_world_size(self.process_group) if (dist.is_available() and dist.is_initialized()) else 1
        if self.training and x.is_cuda and x.dtype in (torch.bfloat16, torch.float16) and (world > 1 or self.force_kernels):
            x = x.float()  # autocast activations: statistics and normalisation run in fp32, like torch's batch norm
        use = self.training and x.is_cuda and x.dim() == 4 and x.dtype == torch.float32 and x.shape[1] % 4 == 0 \
            and self.track_running_stats and (world > 1 or self.force_kernels)
        if not use:
            if self.training and world > 1:
                raise RuntimeError("SyncBatchNorm2d: unsupported input for the synchronised kernels "
                                   f"(shape {tuple(x.shape)}, dtype {x.dtype}, cuda {x.is_cuda})")
            return super().forward(x)
        if not x.is_contiguous(memory_format=torch.channels_last):
            x = x.contiguous(memory_format=torch.channels_last)
        momentum = self.momentum
        if self.num_batches_tracked is not None:
            self.num_batches_tracked.add_(1)
            if momentum is None:
                momentum = 1.0 / float(self.num_batches_tracked)
        return _SyncBatchNormFn.apply(x, self.weight, self.bias, self.running_mean, self.running_var, self.eps,
                                      0.0 if momentum is None else momentum, self.process_group, self._scratch,
                                      self._p2p if world > 1 else None)


def convert_sync_batchnorm(module, process_group=None, p2p=True):
    """train.py:296 counterpart: replace every nn.BatchNorm2d by SyncBatchNorm2d sharing its parameters and buffers
    (state_dict keys unchanged); other batch-norm flavours fall back to torch's own SyncBatchNorm conversion.
    With ``p2p`` (default) and an initialised process group of 2..8 ranks on CUDA, the statistics are exchanged by the
    kernels themselves over NVLink peer memory (enable_p2p_statistics); if the symmetric-memory arena cannot be set up,
    the modules keep the NCCL all-reduce path."""
    out = _convert_sync_batchnorm(module, process_group)
    if p2p and dist.is_available() and dist.is_initialized() and torch.cuda.is_available():
        world = dist.get_world_size(process_group)
        if 1 < world <= 8:
            try:
                enable_p2p_statistics(out, process_group)
            except Exception as exc:  # noqa: BLE001 -- any set-up failure leaves the (slower) all-reduce path in place
                import warnings
                warnings.warn(f"SyncBatchNorm2d: peer-memory statistics exchange unavailable ({exc!r}); using all-reduce")
                for m in out.modules():
                    if isinstance(m, SyncBatchNorm2d):
                        m._p2p = None
    return out


def _convert_sync_batchnorm(module, process_group=None):
    out = module
    if isinstance(module, torch.nn.BatchNorm2d) and not isinstance(module, SyncBatchNorm2d):
        if module.num_features % 4 == 0 and module.track_running_stats:
            out = SyncBatchNorm2d(module.num_features, module.eps, module.momentum, module.affine,
                                  module.track_running_stats, process_group=process_group)
            if module.affine:
                out.weight, out.bias = module.weight, module.bias
            out.running_mean, out.running_var = module.running_mean, module.running_var
            out.num_batches_tracked = module.num_batches_tracked
            out.training = module.training
        else:
            return torch.nn.SyncBatchNorm.convert_sync_batchnorm(module, process_group)
    for name, child in module.named_children():
        new = _convert_sync_batchnorm(child, process_group)
        if new is not child:
            out.add_module(name, new)
    return out
