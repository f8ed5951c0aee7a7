"""One training iteration with the reference's semantics (train.py:338-455), on the B200 kernels.

    zero_grad -> loaders (GPU gather) -> model -> mask = depth > min_depth -> SILog + w_chamfer * chamfer ->
    backward -> [gradient mean all-reduce over ranks] -> clip_grad_norm_(0.1) -> AdamW.step -> OneCycleLR.step

Loss is computed per rank on the local shard (train.py:414-426); only gradients (and SyncBatchNorm statistics when
``sync_bn`` is on, train.py:296) cross GPUs.  This is the harness bench.py times for "train imgs/s"; it is host logic,
not a re-implementation of train.py's logging / validation / checkpoint code (out of scope).
"""
import torch
import torch.nn as nn

from .loss import BinsChamferLoss, DepthLosses, SILogLoss
from .parallel import GradientAverager, broadcast_module_state


class TrainStep:
    def __init__(self, model, semantics_loader=None, instance_loader=None, lr=0.000357, wd=0.1, w_chamfer=0.1,
                 min_depth=1e-3, total_steps=1000, div_factor=25, final_div_factor=100, same_lr=False, bucket_mb=25.0,
                 cudnn_benchmark=True, per_group_max_lr=False, autocast=None):
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
        self.averager = GradientAverager(model.parameters(), bucket_mb=bucket_mb)

    def __call__(self, batch, device):
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
        self.averager.reduce()
        self.averager.clip_grad_norm_(0.1)  # train.py:427 nn.utils.clip_grad_norm_(model.parameters(), 0.1), on the gradient arena
        self.optimizer.step()
        self.scheduler.step()
        return loss.detach()
