"""One training iteration with the reference's semantics (train.py:338-455), on the B200 kernels.

    zero_grad -> loaders (GPU gather) -> model -> mask = depth > min_depth -> SILog + w_chamfer * chamfer ->
    backward -> [gradient mean all-reduce over ranks] -> clip_grad_norm_(0.1) -> AdamW.step -> OneCycleLR.step

Loss is computed per rank on the local shard (train.py:414-426); only gradients (and SyncBatchNorm statistics when
``sync_bn`` is on, train.py:296) cross GPUs.  This is the harness bench.py times for "train imgs/s"; it is host logic,
not a re-implementation of train.py's logging / validation / checkpoint code (out of scope).
"""
import torch
import torch.nn as nn

from . import ops
from .loss import BinsChamferLoss, DepthLosses, SILogLoss
from .parallel import GradientAverager, broadcast_module_state


class TrainStep:
    def __init__(self, model, semantics_loader=None, instance_loader=None, lr=0.000357, wd=0.1, w_chamfer=0.1,
                 min_depth=1e-3, total_steps=1000, div_factor=25, final_div_factor=100, same_lr=False, bucket_mb=25.0,
                 cudnn_benchmark=True, per_group_max_lr=False, autocast=None, overlap=True):
        # the stock torch bodies (encoder, decoder convolutions in training mode) run fixed shapes every step: let cuDNN
        # pick its kernels by measurement (64.2 -> 58.4 ms per step on B200; the reference leaves the flag off)
        if cudnn_benchmark:
            torch.backends.cudnn.benchmark = True
        self.model = model
        self.autocast = autocast  # e.g. torch.bfloat16 (BASELINE config 4): the stock torch bodies run under autocast
        self.semantics_loader = semantics_loader
        self.instance_loader = instance_loader
        self.w_chamfer = w_chamfer
        self.min_depth = min_depth
        self.criterion_ueff = SILogLoss()
        self.criterion_bins = BinsChamferLoss() if w_chamfer > 0 else None
        self.criterion_fused = DepthLosses(min_depth) if w_chamfer > 0 else None  # both criteria, one pass over the depth map
        if same_lr:
            params = model.parameters()
        else:  # train.py:351-352
            params = [{"params": model.get_1x_lr_params(), "lr": lr / 10}, {"params": model.get_10x_lr_params(), "lr": lr}]
        # DDP construction semantics (train.py:298-299): every replica starts from rank 0's parameters and buffers
        broadcast_module_state(model)
        # fused multi-tensor AdamW on CUDA (same update rule, a few launches instead of dozens: the step is host-launch-bound)
        on_cuda = all(p.is_cuda for p in model.parameters())
        self.optimizer = torch.optim.AdamW(params, weight_decay=wd, lr=lr, **({"fused": True} if on_cuda else {}))
        # train.py:364: OneCycleLR(optimizer, args.lr, ...) -- a SCALAR max_lr, which overrides the per-group learning rates:
        # encoder and decoder both peak at ``lr`` in the reference (the lr/10 of :351 only survives as initial value until the
        # scheduler's first step).  ``per_group_max_lr=True`` keeps the encoder at lr/10 instead (a deviation, off by default).
        max_lr = [g["lr"] for g in self.optimizer.param_groups] if per_group_max_lr else lr
        self.scheduler = torch.optim.lr_scheduler.OneCycleLR(
            self.optimizer, max_lr, total_steps=total_steps, cycle_momentum=True, base_momentum=0.85, max_momentum=0.95,
            div_factor=div_factor, final_div_factor=final_div_factor)
        self.averager = GradientAverager(model.parameters(), bucket_mb=bucket_mb, overlap=overlap)

    def __call__(self, batch, device):
        loss = self.forward_backward(batch, device)
        self.update()
        return loss

    def update(self):
        """[gradient mean all-reduce] -> clip 0.1 -> AdamW -> OneCycle (train.py:426-430)."""
        self.averager.reduce()
        self.averager.clip_grad_norm_(0.1)  # train.py:427 nn.utils.clip_grad_norm_(model.parameters(), 0.1), on the gradient arena
        self.optimizer.step()
        self.scheduler.step()

    def forward_backward(self, batch, device):
        self.averager.zero_grad()  # gradients live in the averager's bucket arena (views); one fill per bucket
        img = batch["image"].to(device, non_blocking=True)
        depth = batch["depth"].to(device, non_blocking=True)
        kwargs = {}
        if self.semantics_loader is not None:
            _, sem = self.semantics_loader.get_semantics(dict(batch, image=img))  # (a bound loader also writes the image planes)
            if sem is not None:
                kwargs["semantics"] = sem
        if self.instance_loader is not None:
            _, emb, areas = self.instance_loader.get_instance_segmentation(batch)
            if emb is not None:
                kwargs.update(instance_labels=emb, instance_areas=areas)
        with torch.autocast("cuda", dtype=self.autocast, enabled=self.autocast is not None):
            bin_edges, pred = self.model(img, **kwargs)
            if self.criterion_fused is not None and bin_edges is not None:
                l_dense, l_chamfer = self.criterion_fused(pred, bin_edges, depth, interpolate=True)  # train.py:414-419
                loss = l_dense + self.w_chamfer * l_chamfer
            else:
                mask = depth > self.min_depth
                loss = self.criterion_ueff(pred, depth, mask=mask.to(torch.bool), interpolate=True)
        loss.backward()
        return loss.detach()


class GraphedTrainStep(TrainStep):
    """TrainStep with zero_grad + loaders + forward + losses + backward replayed as ONE CUDA graph per iteration.

    The eager iteration is ~2400 kernel launches: 56 ms of host time to enqueue 58 ms of GPU work on one B200, so at N GPUs the
    step pays for every microsecond of host contention and rank skew (8 processes + NCCL proxy threads on the box's cores:
    weak-scaling efficiency 0.86 at N = 8).  Replayed as a graph the host issues one launch; what stays
    eager is what must: the gradient all-reduce (NCCL, ONE in-place collective on the gradient arena -- nothing inside the
    graph talks to NCCL), the clip on the arena, fused AdamW and the OneCycle schedule (its learning rate and momentum are
    host scalars that change every step, train.py:364-368).

    What makes the iteration capturable:
      * every kernel of this package launches on the current stream without host synchronisation (include/mde_b200.h);
      * gradients accumulate in place into the averager's arena (p.grad are views; the arena fill is the graph's first node);
      * inputs are copied into static device buffers before each replay;
      * SyncBatchNorm's peer-memory exchange takes its epoch from a device-resident step counter that the graph's first node
        increments (parallel.P2PArena.begin_capture; csrc/bn_sync.cu bn_resolve_epoch), and its float64 scratch rows are zeroed
        inside the graph.  The NCCL all-reduce flavour of SyncBatchNorm2d cannot be captured: use the peer-memory one.
    Shapes, modes and the set of inputs must not change after construction; dropout keeps drawing fresh masks (torch registers
    the generator with the graph).  After the capture the model's SyncBatchNorm epochs belong to the graph."""

    def __init__(self, model, example_batch, device, warmup=3, check_labels_every=1, **kwargs):
        kwargs.setdefault("bucket_mb", 1 << 20)  # one bucket: one collective after the replay
        super().__init__(model, overlap=False, **kwargs)
        device = torch.device(device)
        self.device = device
        self.check_labels_every = check_labels_every
        self.static = {k: torch.empty(v.shape, dtype=v.dtype, device=device) for k, v in example_batch.items()
                       if torch.is_tensor(v)}
        self.extra = {k: v for k, v in example_batch.items() if not torch.is_tensor(v)}
        self.arena = self._p2p_arena(model)
        self.replays = 0
        self._stage, self._staged = None, False
        side = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(warmup):  # cuDNN plans, cached operands, shared-memory attributes: everything lazy, outside the capture
                self._load(example_batch)
                self.forward_backward(dict(self.extra, **self.static), device)
                self.update()
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        self._load(example_batch)
        if self.arena is not None:
            self.arena.begin_capture()
        self.graph = torch.cuda.CUDAGraph()
        captured = False
        launches0 = ops.launch_count()
        ops.begin_deferred_checks()
        try:
            with torch.cuda.graph(self.graph):
                if self.arena is not None:
                    self.arena.replay_counter.add_(1)
                self.static_loss = self.forward_backward(dict(self.extra, **self.static), device)
            captured = True
        finally:
            self.deferred = ops.end_deferred_checks()
            self.captured_launches = ops.launch_count() - launches0  # C-ABI launches of this package inside one replay
            if self.arena is not None:
                self.arena.end_capture(captured)

    @staticmethod
    def _p2p_arena(model):
        import torch.distributed as dist
        from .parallel import SyncBatchNorm2d
        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        arena = None
        for m in model.modules():
            if isinstance(m, SyncBatchNorm2d):
                if m._p2p is not None:
                    arena = m._p2p.arena
                elif world > 1 and m.training:
                    raise RuntimeError("GraphedTrainStep: SyncBatchNorm2d without the peer-memory exchange issues NCCL calls "
                                       "inside forward / backward and cannot be captured (convert_sync_batchnorm(p2p=True))")
            elif world > 1 and isinstance(m, nn.SyncBatchNorm):
                raise RuntimeError("GraphedTrainStep: nn.SyncBatchNorm cannot be captured; use parallel.convert_sync_batchnorm")
        return arena

    def _load(self, batch):
        for k, buf in self.static.items():
            buf.copy_(batch[k], non_blocking=True)

    def prefetch(self, batch):
        """Start the NEXT iteration's host -> device copy on a side stream while the current one computes; the following
        ``step()`` call (no batch argument) moves it into the graph's static buffers with a device-to-device copy.  Without it the
        87 MB of a config-2 batch cross PCIe in front of every replay (1.6 ms of a 55 ms step)."""
        if self._stage is None:
            self._stage = {k: torch.empty_like(v) for k, v in self.static.items()}
            self._h2d = torch.cuda.Stream(self.device)
            self._stage_ready = torch.cuda.Event()
            self._stage_free = None
            # the staging buffers come from the consumer stream's allocator pool: whatever that stream still has in flight on
            # a recycled block must finish before the side stream's first copy lands in it
            self._h2d.wait_stream(torch.cuda.current_stream(self.device))
        if self._stage_free is not None:
            self._h2d.wait_event(self._stage_free)  # the previous copy OUT of the staging buffers
        with torch.cuda.stream(self._h2d):
            for k, buf in self._stage.items():
                buf.copy_(batch[k], non_blocking=True)
            self._stage_ready.record(self._h2d)
        self._staged = True

    def __call__(self, batch=None, device=None):
        if batch is None:
            if not self._staged:
                raise RuntimeError("GraphedTrainStep(): no batch given and none prefetched")
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(self._stage_ready)
            for k, buf in self.static.items():
                buf.copy_(self._stage[k], non_blocking=True)
            self._stage_free = torch.cuda.Event()
            self._stage_free.record(cur)
            self._staged = False
        else:
            self._load(batch)
        self.graph.replay()
        self.replays += 1
        self.update()
        if self.deferred and self.check_labels_every and self.replays % self.check_labels_every == 0:
            # the reference's index_select raises on an out-of-range label (150-class table, SemanticsLoader.py:125); a captured
            # gather cannot raise, so its flag is read here, after the whole iteration has been enqueued
            ops.raise_deferred_checks(self.deferred)
        return self.static_loss
