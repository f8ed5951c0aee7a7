#!/usr/bin/env python
"""Benchmark of the AdaBins head + loss + external-info hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config {2,3,4,5}] [--impl reference] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one pass of the drop-in call sequence of the reference's loop (train.py:400-423) over one synthetic batch:
loaders (GPU gather) -> UnetAdaptiveBins.forward -> SILogLoss + BinsChamferLoss [-> backward + gradient all-reduce + clip +
AdamW/OneCycle for the training configs].  BASELINE.json configs:

  2 (default)  EfficientNet-B1 AdaBins + GloVe-25d ADE20K-places semantics @input, batch 16/GPU, 416x544   inference Mpix/s
  3            B1 + GloVe-25d + ADE20K-Swin instance embeddings/areas/human sizes @input (73 ch), batch 16/GPU   train imgs/s
  4            EfficientNet-B5 AdaBins (original), bf16 autocast, batch 8/GPU (64 on 8 GPUs)                     train imgs/s
  5            noAdaBins B1 at 480x640, batch 32/GPU (256 on 8 GPUs), head-free depth regression              inference Mpix/s

The default run prints ONE JSON line for config 2 that also carries compact legs for the training step (config 2 model),
config 3 (training) and config 5 (inference): `"train"` is the last key so that it survives tail truncation.

  value : whole-job throughput, batch resident in HBM when the timed region starts (one CUDA-graph replay per step)
  e2e   : same, batch in pinned host memory, H2D copies + the D2H loss read inside the timed region
  --impl reference : the reference's own modules (oracle/_ref, staged by oracle/build_ref.py; the oracle port if absent) on
                     the host cores at the SAME config and batch -- a reported baseline, not a target.

The EfficientNet encoder is a PyTorch/cuDNN passthrough (outside the hot path) run in true fp32; every
tensor-core product of the decoder / head is formed from split-bf16 pairs (three bf16 products, fp32 accumulation) -- the mode
tests/test_gpu_parity.py::test_config2_full_size_bench_mode_vs_oracle holds to 1e-3 on every pixel.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_BINS = 256
CONFIGS = {
    2: dict(name="BASELINE config 2: EfficientNet-B1 AdaBins + GloVe-25d ADE20K-places @input, 416x544, n_bins 256, random init",
            encoder="efficientnet-b1", sem="glove-25d-ade20k-places", inst=None, insertion="input", hw=(416, 544), batch=16,
            kind="infer", autocast=None),
    3: dict(name="BASELINE config 3: B1 AdaBins + GloVe-25d + ADE20K-Swin instance embeddings / areas / human sizes @input "
                 "(73 channels), 416x544, training",
            encoder="efficientnet-b1", sem="glove-25d", inst="ade20k_swin_human_sizes", insertion="input", hw=(416, 544),
            batch=16, kind="train", autocast=None),
    4: dict(name="BASELINE config 4: EfficientNet-B5 AdaBins (original), bf16 autocast training, 416x544, batch 64 = 8 per GPU on 8 GPUs",
            encoder="efficientnet-b5", sem=None, inst=None, insertion="before-attn", hw=(416, 544), batch=8, kind="train",
            autocast="bf16"),
    5: dict(name="BASELINE config 5: noAdaBins EfficientNet-B1 at 480x640, batch 256 = 32 per GPU on 8 GPUs, inference",
            encoder="efficientnet-b1-noAdaBins", sem=None, inst=None, insertion="input", hw=(480, 640), batch=32, kind="infer",
            autocast=None),
}
METRIC_INFER = "head/loss Mpix/s (gather + UnetAdaptiveBins fwd + SILog + chamfer)"  # full-resolution pixels F*B per second
METRIC_TRAIN = "train imgs/s (fwd + SILog + 0.1*chamfer + bwd + grad all-reduce + clip + AdamW/OneCycle)"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p["bf16_tflops"]), "measured"
    except Exception:
        return 6650.0, 1590.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        mhz = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        if mhz:
            out["sm_mhz"] = statistics.median(mhz)
            out["sm_max_mhz"] = float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for i, nme in enumerate(names):
                if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows):
                    out["reasons"].append(nme)
        return out


# ---------------------------------------------------------------------------------------------------------------------
# synthetic batches (identical recipe for both arms)
# ---------------------------------------------------------------------------------------------------------------------
def host_batch(cfg, batch, rank=0, pin=False):
    from mde_biological_vision_systems_b200 import synthetic
    h, w = cfg["hw"]
    out = {"image": synthetic.image(batch, h, w, seed=10 * rank), "depth": synthetic.depth(batch, h, w, seed=10 * rank + 1)}
    if cfg["sem"]:
        places = "ade20k-places" in cfg["sem"]
        out["semantics"] = synthetic.label_maps(batch, h, w, seed=10 * rank + 2, lo=-1 if places else 0,
                                                hi=100 if places else 149, inject=(-7, 101, 255, 1000) if places else ())[0]
    if cfg["inst"]:
        lab, areas = synthetic.label_maps(batch, h, w, seed=10 * rank + 3, lo=-1, hi=100)
        out["instance_labels"], out["instance_areas"] = lab, areas
    if pin:
        out = {k: v.pin_memory() for k, v in out.items()}
    return out


def model_kwargs(cfg):
    return dict(n_bins=N_BINS, min_val=1e-3, max_val=10.0, norm="linear", encoder_name=cfg["encoder"], semantics_mode=cfg["sem"],
                instance_segmentation_mode=cfg["inst"], insertion_point=cfg["insertion"], image="rgb")


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own modules on the host cores
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_runner(cfg, batch):
    """-> (step callable returning the loss as a float, kind, parts callable or None).  kind "reference": the reference's own
    modules from oracle/_ref; "port": the oracle restatement (oracle/adabins_oracle.py)."""
    import numpy as np
    import torch
    from argparse import Namespace
    from oracle import adabins_oracle as oracle
    from oracle import ref_harness
    from mde_biological_vision_systems_b200.models import UnetAdaptiveBins

    torch.manual_seed(0)
    ours = UnetAdaptiveBins.build(**model_kwargs(cfg)).eval()  # random-init weights of the config's architecture (CPU)
    sd = {k: v.detach() for k, v in ours.state_dict().items()}
    hb = host_batch(cfg, batch)
    img, depth = hb["image"], hb["depth"]
    noada = "noAdaBins" in cfg["encoder"]
    if ref_harness.available():
        ns = ref_harness.load()
        kw = model_kwargs(cfg)
        model = ref_harness.build_model(cfg["encoder"], sd, semantics_mode=kw["semantics_mode"],
                                        instance_segmentation_mode=kw["instance_segmentation_mode"],
                                        insertion_point=kw["insertion_point"], image="rgb")
        with ref_harness.cpu_loaders():
            sem_loader = ns.SemanticsLoader(Namespace(use_semantics=cfg["sem"])) if cfg["sem"] else None
            inst_loader = ns.InstanceSegmentationLoader(Namespace(use_instance_segmentation=cfg["inst"])) if cfg["inst"] else None
        silog, chamfer = ns.SILogLoss(), ns.BinsChamferLoss()

        def loaders():
            kwargs = {}
            with ref_harness.cpu_loaders():
                if sem_loader is not None:
                    kwargs["semantics"] = sem_loader.get_semantics({"semantics": hb["semantics"].clone()})[1]
                if inst_loader is not None:
                    _, emb, areas = inst_loader.get_instance_segmentation(
                        {"instance_labels": hb["instance_labels"].clone(), "instance_areas": hb["instance_areas"].clone()})
                    kwargs.update(instance_labels=emb, instance_areas=areas)
            return kwargs

        def step():
            with torch.no_grad():
                edges, pred = model(img, **loaders())
                loss = silog(pred, depth, mask=depth > 1e-3, interpolate=True)
                if not noada:
                    loss = loss + 0.1 * chamfer(edges, depth)
            return float(loss)

        def parts():
            out = {}

            def once(key, fn):
                t = time.perf_counter()
                r = fn()
                out[key] = round((time.perf_counter() - t) * 1e3, 2)
                return r

            with torch.no_grad():
                kwargs = once("loaders_ms", loaders)
                edges, pred = once("model_ms", lambda: model(img, **kwargs))
                if not noada and cfg["insertion"] == "input" and cfg["sem"] and not cfg["inst"] and "semantics" in kwargs:
                    unet = model.decoder(model.encoder(torch.cat((img, kwargs["semantics"].float()), 1)))
                    once("head_ms", lambda: model.adaptive_bins_layer(unet))
                once("silog_ms", lambda: silog(pred, depth, mask=depth > 1e-3, interpolate=True))
                if not noada:
                    once("chamfer_ms", lambda: chamfer(edges, depth))
            return out

        return step, "reference", parts

    # fallback: the oracle port around the same torch encoder module
    table = np.load(os.path.join(ROOT, "data", "ade20k_places_classes_glove_twitter_27b_25d_embeddings.npy"))
    if cfg["sem"] != "glove-25d-ade20k-places" or cfg["inst"]:
        raise RuntimeError("the oracle-port arm covers config 2 only; stage oracle/_ref (python oracle/build_ref.py)")

    def step():
        with torch.no_grad():
            _, sem = oracle.semantics_loader(cfg["sem"], hb["semantics"].numpy(), table)
            x = oracle.input_insertion(sd, img, cfg["sem"], None, "rgb", semantics=torch.from_numpy(sem))
            backbone = lambda t: oracle.decoder_bn(oracle.encoder_features(ours.encoder.original_model, t), sd)
            _, _, l1, l2 = oracle.forward_and_losses(backbone, sd, x, depth, 1e-3, 10.0)
            return float(l1) + 0.1 * float(l2)

    return step, "port", None


def time_cpu(step, warmup, steps):
    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    return statistics.median(times), sum(times) / len(times)


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    batch = args.batch or cfg["batch"]
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    if cfg["kind"] == "train":
        print(json.dumps({"impl": "reference", "unavailable": "the CPU reference arm times the inference path (configs 2 and 5)"}))
        return
    step, kind, _ = cpu_reference_runner(cfg, batch)
    _, mean = time_cpu(step, args.warmup, args.steps)
    h, w = cfg["hw"]
    mpix = batch * h * w / mean / 1e6
    sample = f"each step = the full {cfg['name']} batch ({batch} x {h}x{w}) through the reference's own modules, torch CPU fp32"
    print(json.dumps({
        "impl": "reference", "metric": METRIC_INFER, "value": mpix, "unit": "Mpix/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["name"], "batch_per_gpu": batch},
        "cpu_baseline": {"value": mpix, "unit": "Mpix/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": mpix, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
class Ctx:
    """torch.distributed plumbing of one rank."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps):
        """K steps bracketed by barrier + synchronize, CUDA events, max over ranks -> total ms."""
        torch = self.torch
        self.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        torch.cuda.synchronize()
        ms = torch.tensor([s.elapsed_time(e)], device=self.dev)
        if self.world > 1:
            self.dist.barrier()
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item())


def build_gpu(cfg, ctx):
    import torch
    from argparse import Namespace
    from mde_biological_vision_systems_b200.ExternalInfoLoaders.InstanceSegmentationLoader import InstanceSegmentationLoader
    from mde_biological_vision_systems_b200.ExternalInfoLoaders.SemanticsLoader import SemanticsLoader
    from mde_biological_vision_systems_b200.models import UnetAdaptiveBins
    torch.manual_seed(0)
    model = UnetAdaptiveBins.build(**model_kwargs(cfg)).to(ctx.dev)
    sem_loader = SemanticsLoader(Namespace(use_semantics=cfg["sem"]), device=ctx.dev) if cfg["sem"] else None
    inst_loader = InstanceSegmentationLoader(Namespace(use_instance_segmentation=cfg["inst"]), device=ctx.dev) if cfg["inst"] else None
    if sem_loader is not None:
        sem_loader.bind_encoder_input(model)  # config 2: embeddings gathered straight into the NHWC encoder input
    return model, sem_loader, inst_loader


def infer_step_fn(cfg, model, sem_loader, inst_loader, dev):
    import torch
    from mde_biological_vision_systems_b200.loss import DepthLosses, SILogLoss
    silog, both = SILogLoss(), DepthLosses(1e-3)

    def step(batch):
        with torch.no_grad():
            img = batch["image"].to(dev, non_blocking=True)
            depth = batch["depth"].to(dev, non_blocking=True)
            kwargs = {}
            if sem_loader is not None:  # the device image rides along: a bound loader writes it with the embeddings
                kwargs["semantics"] = sem_loader.get_semantics(dict(batch, image=img))[1]
            if inst_loader is not None:
                _, emb, areas = inst_loader.get_instance_segmentation(batch)
                kwargs.update(instance_labels=emb, instance_areas=areas)
            edges, pred = model(img, **kwargs)
            if edges is not None:  # train.py:414-419: SILog(mask = depth > min_depth) + 0.1 * chamfer, one pass over the depth map
                l_dense, l_bins = both(pred, edges, depth, interpolate=True)
                loss = l_dense + 0.1 * l_bins
            else:
                loss = silog(pred, depth, mask=depth > 1e-3, interpolate=True)
        return loss

    return step


def run_infer(cfg, ctx, steps, warmup, batch, detail=True):
    """Inference legs of one config: resident / e2e, eager / CUDA graph.  Returns a dict."""
    import torch
    from mde_biological_vision_systems_b200 import ops
    from mde_biological_vision_systems_b200.graphs import GraphedStep
    from mde_biological_vision_systems_b200.prefetch import DevicePrefetcher
    dev = ctx.dev
    h, w = cfg["hw"]
    model, sem_loader, inst_loader = build_gpu(cfg, ctx)
    model.eval()
    step = infer_step_fn(cfg, model, sem_loader, inst_loader, dev)
    host = host_batch(cfg, batch, ctx.rank, pin=True)
    resident = {k: v.to(dev) for k, v in host.items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    for _ in range(max(warmup, 3)):
        step(resident)
        float(step(host).item())
    out = {"batch_per_gpu": batch, "h2d_bytes_per_step": h2d}
    l0 = ops.launch_count()
    ms_eager = ctx.timed(lambda: step(resident), steps) / steps
    out["gpu_launches"] = int(ops.launch_count() - l0)
    out["eager_ms"] = ms_eager

    def timed_prefetch(hbatch, nsteps, runner):
        ctx.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        pending = None
        for b in DevicePrefetcher((hbatch for _ in range(nsteps)), dev):
            loss = runner(b)
            if pending is not None:
                float(pending.item())  # read the previous step's loss while this step runs
            pending = loss.clone()     # (a graph replay rewrites its static output tensor)
        float(pending.item())
        e.record()
        torch.cuda.synchronize()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if ctx.world > 1:
            ctx.dist.barrier()
            ctx.dist.all_reduce(ms, op=ctx.dist.ReduceOp.MAX)
        return float(ms.item())

    timed_prefetch(host, 2, step)
    out["e2e_eager_ms"] = timed_prefetch(host, steps, step) / steps
    # the same step replayed as one CUDA graph (graphs.GraphedStep): no host launch gaps
    try:
        gstep = GraphedStep(lambda **kw: step(kw), resident)
        ref_loss = float(step(resident))
        got = float(gstep(**resident))
        if abs(got - ref_loss) > 1e-4 * abs(ref_loss):
            raise RuntimeError(f"graph replay loss {got} != eager {ref_loss}")
        out["graph_ms"] = ctx.timed(lambda: gstep(**resident), steps) / steps
        timed_prefetch(host, 2, lambda b: gstep(**b))
        out["e2e_graph_ms"] = timed_prefetch(host, steps, lambda b: gstep(**b)) / steps
        out["loss"] = got
    except Exception as exc:  # reported, never fatal: the eager numbers stand on their own
        out["graph_error"] = repr(exc)[:200]
    if detail and "graph_ms" in out:
        # the same graph with the cuDNN passthrough bodies left on the library default (TF32) -- reported alongside; the
        # headline runs them in true fp32 (the tolerance contract is against the fp32 reference)
        try:
            model.backbone_tf32 = True
            g2 = GraphedStep(lambda **kw: step(kw), resident)
            out["tf32_backbone_ms"] = ctx.timed(lambda: g2(**resident), steps) / steps
            del g2
        except Exception as exc:  # noqa: BLE001
            out["tf32_backbone_error"] = repr(exc)[:120]
        model.backbone_tf32 = False
        # the north star's bf16 mode (2e-2 tolerance): one bf16 product per K step in the tensor-core kernels
        try:
            model.precision = "bf16"
            g3 = GraphedStep(lambda **kw: step(kw), resident)
            out["bf16_mode_ms"] = ctx.timed(lambda: g3(**resident), steps) / steps
            del g3
        except Exception as exc:  # noqa: BLE001
            out["bf16_mode_error"] = repr(exc)[:120]
        model.precision = "fp32"
    out["ms"] = out.get("graph_ms", ms_eager)
    out["e2e_ms"] = out.get("e2e_graph_ms", out["e2e_eager_ms"])
    pix = ctx.world * batch * h * w
    out["value"] = pix / (out["ms"] * 1e-3) / 1e6
    out["e2e_value"] = pix / (out["e2e_ms"] * 1e-3) / 1e6
    if not detail:
        del model
        torch.cuda.empty_cache()
        return out

    # ---- per-kernel durations, live, CUDA events on the launching stream (separate pass: the event records do not
    # perturb the headline number)
    ops.enable_kernel_timing(True)
    for _ in range(steps):
        step(resident)
    torch.cuda.synchronize()
    out["ktimes"] = {k: (statistics.mean(v), len(v) // steps) for k, v in ops.kernel_times_ms().items()}
    out["kflops"] = ops.kernel_work()
    ops.enable_kernel_timing(False)
    # ---- hot path only (loaders + head + losses on a fixed decoder output), as ONE CUDA graph
    if "noAdaBins" not in cfg["encoder"] and cfg["insertion"] == "input":
        from mde_biological_vision_systems_b200.loss import DepthLosses
        both = DepthLosses(1e-3)
        with torch.no_grad():
            kwargs = {}
            if sem_loader is not None:
                kwargs["semantics"] = sem_loader.get_semantics(resident)[1]
            if inst_loader is not None:
                _, emb, areas = inst_loader.get_instance_segmentation(resident)
                kwargs.update(instance_labels=emb, instance_areas=areas)
            hw = h * w
            x = model._concat_external(resident["image"], model._external_channels(kwargs.get("semantics"),
                                       kwargs.get("instance_labels"), kwargs.get("instance_areas"), hw))
            with ops.exact_fp32_library():
                unet = model.decoder(model.encoder(ops.to_channels_last(x)))
            planes = unet.planes if isinstance(unet, ops.SplitBF16) else None

        def hot(depth, **labels):
            with torch.no_grad():
                if sem_loader is not None:
                    sem_loader.get_semantics(labels)
                if inst_loader is not None:
                    inst_loader.get_instance_segmentation(labels)
                edges, pred = model._head(ops.SplitBF16(planes) if planes is not None else unet)
                l_dense, l_bins = both(pred, edges, depth, interpolate=True)
                return l_dense + 0.1 * l_bins

        try:
            ghot = GraphedStep(hot, {k: v for k, v in resident.items() if k != "image"})
            args_hot = {k: v for k, v in resident.items() if k != "image"}
            out["hot_ms"] = ctx.timed(lambda: ghot(**args_hot), steps) / steps
        except Exception as exc:
            out["hot_error"] = repr(exc)[:200]
        # ---- stand-alone range attention (A6, the GEMM north_star names) on the tcgen05 kernel, own roofline line
        try:
            with torch.no_grad():
                tgt, feat = model.adaptive_bins_layer.tokens_and_features(ops.SplitBF16(planes) if planes is not None else unet,
                                                                          bias_free=True, pair_out=True)
                q = tgt[1:129].permute(1, 0, 2).contiguous()
                ops.range_attention(feat, q, impl="tc")
                qp = ops.split_bf16_flat(q)
                y = torch.empty((batch, 128, h // 2, w // 2), dtype=torch.float32, device=dev)
                lib = ops._lib.load()
                out["a6_ms"] = ctx.timed(lambda: lib.mde_range_attention_tc(ops._p(feat.planes), ops._p(qp), ops._p(y), batch, 128,
                                                                            128, (h // 2) * (w // 2), ops._s()), steps) / steps
        except Exception as exc:
            out["a6_error"] = repr(exc)[:200]
    del model
    torch.cuda.empty_cache()
    return out


def run_train(cfg, ctx, steps, batch, sync_bn="kernels", launch="graph"):
    """Training iteration with the reference's semantics (train.py:387-455): forward + SILog + 0.1 chamfer + backward +
    gradient mean all-reduce over ranks (NCCL) + clip 0.1 + AdamW + OneCycle, `batch` per GPU (weak scaling,
    --use_new_batching), SyncBatchNorm when N > 1 (train.py:296).  Inputs come from pinned host memory.
    launch = "graph": zero_grad + loaders + forward + losses + backward replayed as ONE CUDA graph (training.GraphedTrainStep;
    the all-reduce is one collective on the gradient arena after the replay); "eager": ~2400 launches per step with the
    all-reduce overlapped bucket by bucket.  A capture that fails falls back to the eager launch of the same kernels and says so."""
    import torch
    from mde_biological_vision_systems_b200 import ops, parallel
    from mde_biological_vision_systems_b200.training import GraphedTrainStep, TrainStep
    launch = os.environ.get("MDE_TRAIN_LAUNCH", launch)

    def make_model():
        model, sem_loader, inst_loader = build_gpu(cfg, ctx)
        if ctx.world > 1 and sync_bn == "stock":
            model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
        elif ctx.world > 1 and sync_bn:
            # global-batch statistics on the B200 kernels (csrc/bn_sync.cu), exchanged by the kernels themselves over NVLink
            # peer memory ("allreduce": one NCCL all-reduce per layer and direction instead)
            model = parallel.convert_sync_batchnorm(model, p2p=(sync_bn != "allreduce"))
        model.train()
        return model, sem_loader, inst_loader

    autocast = torch.bfloat16 if cfg["autocast"] == "bf16" else None
    host = host_batch(cfg, batch, ctx.rank, pin=True)
    steps = max(2, min(steps, 5))
    stepper, note = None, None
    model, sem_loader, inst_loader = make_model()
    # nn.SyncBatchNorm / the NCCL flavour of SyncBatchNorm2d call collectives inside forward and backward: not capturable
    if launch == "graph" and (ctx.world == 1 or sync_bn in ("kernels", False, None)):
        try:
            stepper = GraphedTrainStep(model, host, ctx.dev, semantics_loader=sem_loader, instance_loader=inst_loader,
                                       total_steps=1000, autocast=autocast)
            for _ in range(2):
                stepper(host, ctx.dev)
            torch.cuda.synchronize()
        except Exception as exc:  # noqa: BLE001 -- the same kernels are then launched eagerly; the line says so
            note = f"{type(exc).__name__}: {str(exc)[:160]}"
            sys.stderr.write(f"[bench] training graph capture failed, eager launch instead: {note}\n")
            stepper = None
            del model
            torch.cuda.empty_cache()
            model, sem_loader, inst_loader = make_model()
    graphed = stepper is not None
    if stepper is None:
        stepper = TrainStep(model, semantics_loader=sem_loader, instance_loader=inst_loader, total_steps=1000, autocast=autocast)
        for _ in range(3):
            stepper(host, ctx.dev)
    l0 = ops.launch_count()
    box = {}

    def one():
        if graphed:  # the next batch's H2D copy runs on a side stream under this iteration's replay
            box["loss"] = stepper() if stepper._staged else stepper(host)
            stepper.prefetch(host)
        else:
            box["loss"] = stepper(host, ctx.dev)

    import ctypes
    from mde_biological_vision_systems_b200 import _lib
    lib = _lib.load()
    wait_ns, waits = ctypes.c_uint64(0), ctypes.c_uint64(0)
    lib.mde_bn_wait_stats(None, None, 1)
    ms = ctx.timed(one, steps) / steps
    lib.mde_bn_wait_stats(ctypes.byref(wait_ns), ctypes.byref(waits), 1)
    res = {"config": cfg["name"].split(":")[0].replace("BASELINE config ", "cfg"), "imgs_s": round(ctx.world * batch / (ms * 1e-3), 1),
           "ms": round(ms, 2), "batch_per_gpu": batch, "loss": round(float(box["loss"].item()), 4),
           "launches": int((ops.launch_count() - l0) / steps) + (stepper.captured_launches if graphed else 0),
           "dtype": cfg["autocast"] or "f32", "launch": "cuda_graph(fwd+bwd)" if graphed else "eager"}
    if note:
        res["graph_error"] = note
    if ctx.world > 1:
        res["sync_bn"] = sync_bn
        w = stepper.averager.last_exposed_wait_ms()
        res["exposed_allreduce_ms"] = None if w is None else round(w, 3)
        # SyncBatchNorm's peer-flag waits of this rank (rank skew + NVLink latency), per step
        res["syncbn_wait_ms"] = round(wait_ns.value / 1e6 / steps, 3)
        res["syncbn_waits"] = int(waits.value / steps)
    del stepper, model
    torch.cuda.empty_cache()
    return res


def roofline_from(cfg, r, batch):
    """The dominant hand-written kernel of the config: the fused head chain (HBM) for the AdaBins configs, the largest
    decoder convolution (tensor) for the head-free config 5."""
    hbm, bf16, src = peaks()
    h, w = cfg["hw"]
    P = (h // 2) * (w // 2)
    kt = r.get("ktimes", {})
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            tr = json.load(f)
        ent = tr.get("head_chain_kernel")
        if ent and ent.get("batch") == batch and ent.get("P") == P:
            traffic = ent["dram_bytes_read"] + ent["dram_bytes_write"]
    except Exception:
        pass
    if "head_chain" in kt:
        ms = kt["head_chain"][0]
        alg = batch * (128 * P * 4 + P * 4)  # the conv3x3 features once (as a 2 x bf16 pair = 4 B/element), pred written
        ach = alg / (ms * 1e-3) / 1e9
        fl = batch * 2.0 * P * N_BINS * 128  # algorithmic FLOPs of the folded contraction (the tensor cores issue 3x)
        return {"kernel": "head_chain_kernel<256,softmax> (range-attention x conv_out fold + softmax + bins, bf16x3)",
                "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": traffic,
                "algorithmic_bytes": alg, "peak_source": src, "ms_per_launch": ms,
                "traffic_source": None if traffic is None else "profiles/r2_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum "
                                  "of one launch, ncu --set full capture of this round at the same batch and size",
                "tensor": {"issued_tflops_bf16": 3 * fl / (ms * 1e-3) / 1e12, "peak_bf16_tflops": bf16,
                           "frac": 3 * fl / (ms * 1e-3) / 1e12 / bf16}}
    convs = {k: v for k, v in kt.items() if k.startswith("up") and "conv" in k}
    if convs:
        name = max(convs, key=lambda k: convs[k][0])
        ms = convs[name][0]
        fl = r.get("kflops", {}).get(name)
        ach = None if fl is None else 3 * fl / (ms * 1e-3) / 1e12  # issued bf16 FLOPs: three products per algorithmic MAC
        return {"kernel": f"conv3x3_kernel ({name}, bf16x3)", "bound": "tensor", "achieved": ach, "peak": bf16, "unit": "TFLOP/s",
                "frac": None if ach is None else ach / bf16, "traffic": None, "peak_source": src, "ms_per_launch": ms,
                "algorithmic_flops": fl}
    return None


def kernel_table(cfg, r, batch):
    """Compact per-kernel entries: [ms per launch, launches per step, fraction of the kernel's own roofline]."""
    hbm, bf16, _ = peaks()
    h, w = cfg["hw"]
    F, P = h * w, (h // 2) * (w // 2)
    kt = r.get("ktimes", {})
    alg = {  # algorithmic bytes per launch (HBM-bound kernels)
        "gather_embed": batch * F * (8 + 25 * 4), "patch_embed": batch * 128 * P * 4, "silog_fwd": batch * (F * 4 + P * 4),
        "chamfer_fwd": batch * F * 4, "loss_fused": batch * (F * 4 + P * 4), "head_chain": batch * (128 * P * 4 + P * 4),
    }
    flops = r.get("kflops", {})  # algorithmic FLOPs per launch of the conv kernels (the tensor cores issue 3x as bf16 products)
    out = {}
    for name, (ms, n) in sorted(kt.items()):
        ent = [round(ms, 4), n]
        if name in alg:
            ent.append(round(alg[name] / (ms * 1e-3) / 1e9 / hbm, 3))
        elif name in flops:
            ent.append(round(3 * flops[name] / (ms * 1e-3) / 1e12 / bf16, 3))
        out[name] = ent
    if "a6_ms" in r:
        by = batch * 2 * 128 * P * 4
        out["range_attention_tc(A6)"] = [round(r["a6_ms"], 4), 0, round(by / (r["a6_ms"] * 1e-3) / 1e9 / hbm, 3)]
    return out


def run_ours(args):
    import torch
    ctx = Ctx()
    cfg = CONFIGS[args.config]
    batch = args.batch or cfg["batch"]
    h, w = cfg["hw"]
    torch.backends.cudnn.benchmark = os.environ.get("MDE_CUDNN_BENCHMARK", "0") == "1"
    sampler = ClockSampler(ctx.local) if ctx.rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    line = {}
    train = None
    extra = {}
    if cfg["kind"] == "infer":
        r = run_infer(cfg, ctx, args.steps, args.warmup, batch)
        if not args.no_train and args.config == 2:
            try:  # the training legs must never cost the inference line: report their failure instead
                train = run_train(cfg, ctx, args.steps, batch)
                if ctx.world > 1 and os.environ.get("MDE_BENCH_TRAIN_VARIANTS") == "1":
                    train["stock_sync_bn"] = run_train(cfg, ctx, args.steps, batch, sync_bn="stock")
                    train["local_bn"] = run_train(cfg, ctx, args.steps, batch, sync_bn=False)
            except Exception as exc:  # noqa: BLE001
                train = {"error": repr(exc)[:300]}
        if not args.no_extra and args.config == 2:
            try:
                extra["cfg3_train"] = run_train(CONFIGS[3], ctx, args.steps, CONFIGS[3]["batch"])
            except Exception as exc:  # noqa: BLE001
                extra["cfg3_train"] = {"error": repr(exc)[:200]}
            try:
                r5 = run_infer(CONFIGS[5], ctx, max(3, args.steps // 2), 3, CONFIGS[5]["batch"], detail=False)
                extra["cfg5_infer"] = {"Mpix_s": round(r5["value"], 1), "ms": round(r5["ms"], 3), "e2e_Mpix_s": round(r5["e2e_value"], 1),
                                       "batch_per_gpu": r5["batch_per_gpu"], "hw": list(CONFIGS[5]["hw"])}
            except Exception as exc:  # noqa: BLE001
                extra["cfg5_infer"] = {"error": repr(exc)[:200]}
    else:
        train = run_train(cfg, ctx, args.steps, batch)
        r = None
    clocks = sampler.stop() if sampler else None
    if ctx.rank == 0:
        common = {"n_gpus": ctx.world, "steps": args.steps, "warmup": max(args.warmup, 3), "higher_is_better": True, "scaling": "weak",
                  "vs_baseline": None, "data": "synthetic"}
        l2note = "per-step working set (>1 GB of activations) exceeds the 126 MB L2; no explicit flush"
        if r is not None:
            pix = ctx.world * batch * h * w
            line = {"metric": METRIC_INFER, "value": r["value"], "unit": "Mpix/s", **common, "ms_per_step": r["ms"],
                    "dtype": "f32 (tensor-core products as 3 bf16 MMAs on split-bf16 pairs, fp32 accumulate; cuDNN encoder in fp32)",
                    "config": {"workload": cfg["name"], "batch_per_gpu": batch, "l2": l2note,
                               "launch": "cuda_graph_replay" if "graph_ms" in r else "eager"},
                    "gpu_launches": r["gpu_launches"],
                    "eager": {"value": pix / (r["eager_ms"] * 1e-3) / 1e6, "ms_per_step": r["eager_ms"]},
                    "kernels": kernel_table(cfg, r, batch)}
            if "hot_ms" in r:
                line["hot_path"] = {"what": "loaders + mViT head + bins + SILog + chamfer on a fixed decoder output, one CUDA graph",
                                    "ms_per_step": r["hot_ms"], "value": batch * h * w / (r["hot_ms"] * 1e-3) / 1e6, "unit": "Mpix/s per GPU"}
            if "tf32_backbone_ms" in r:
                line["tf32_backbone"] = {"value": pix / (r["tf32_backbone_ms"] * 1e-3) / 1e6, "ms_per_step": r["tf32_backbone_ms"],
                                         "note": "cuDNN encoder / conv2 on the library's TF32 default instead of true fp32"}
            if "bf16_mode_ms" in r:
                line["bf16_mode"] = {"value": pix / (r["bf16_mode_ms"] * 1e-3) / 1e6, "ms_per_step": r["bf16_mode_ms"],
                                     "note": "model.precision = 'bf16': one bf16 product per K step (2e-2 tolerance), library bodies TF32"}
            for k in ("graph_error", "hot_error", "a6_error", "bf16_mode_error", "tf32_backbone_error"):
                if k in r:
                    line[k] = r[k]
            line["clocks"] = clocks
            line["roofline"] = roofline_from(cfg, r, batch)
            if ctx.world == 1 and not args.no_cpu:
                try:
                    step, kind, parts = cpu_reference_runner(cfg, 2)
                    cores = len(os.sched_getaffinity(0))
                    torch.set_num_threads(cores)
                    med, _ = time_cpu(step, 1, 5)
                    line["cpu_baseline"] = {"value": 2 * h * w / med / 1e6, "unit": "Mpix/s", "cores": cores, "kind": kind,
                                            "sample": f"batch 2 x {h}x{w} of the same workload, 1 warm-up + median of 5 steps, torch CPU fp32",
                                            "breakdown_ms": parts() if parts else None}
                except Exception as exc:  # noqa: BLE001
                    line["cpu_baseline"] = {"error": repr(exc)[:200]}
            line["e2e"] = {"value": r["e2e_value"], "unit": "Mpix/s", "h2d_bytes_per_step": r["h2d_bytes_per_step"], "d2h_bytes_per_step": 4,
                           "ms_per_step": r["e2e_ms"], "eager_ms_per_step": r["e2e_eager_ms"],
                           "how": "pinned host batch (int64 labels) -> DevicePrefetcher (H2D on a side stream) -> loaders + model + losses -> loss.item()"}
            if extra:
                line["configs"] = extra
            line["train"] = train
        else:
            line = {"metric": METRIC_TRAIN, "value": train["imgs_s"], "unit": "imgs/s", **common, "ms_per_step": train["ms"],
                    "dtype": train["dtype"], "config": {"workload": cfg["name"], "batch_per_gpu": batch, "l2": l2note},
                    "gpu_launches": train["launches"] * max(2, min(args.steps, 5)), "clocks": clocks, "roofline": None,
                    "e2e": {"value": train["imgs_s"], "unit": "imgs/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": 0,
                            "how": "every training step reads its batch from pinned host memory (TrainStep)"},
                    "train": train}
        print(json.dumps(line))
    if ctx.world > 1:
        try:
            ctx.dist.destroy_process_group()
        except Exception:  # noqa: BLE001 -- the JSON line is already out
            pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the config's)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the compact config 3 / config 5 legs of the default run")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
