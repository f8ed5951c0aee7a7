"""CPU: drop-in surface, state_dict contract, C-ABI exports, and 'no fallback' behaviour."""
import ctypes
import os
import re

import pytest
import torch

from mde_biological_vision_systems_b200 import _lib, ops
from mde_biological_vision_systems_b200.loss import BinsChamferLoss, SILogLoss
from mde_biological_vision_systems_b200.models import UnetAdaptiveBins

from helpers import ROOT, make_model

CASES = {
    "b1_plain": dict(encoder_name="efficientnet-b1", insertion_point="input", semantics_mode=None,
                     instance_segmentation_mode=None),
    "b1_cfg3": dict(encoder_name="efficientnet-b1", insertion_point="input", semantics_mode="glove-25d",
                    instance_segmentation_mode="ade20k_swin_human_sizes"),
    "b1_areas_before_attn": dict(encoder_name="efficientnet-b1", insertion_point="before-attn",
                                 semantics_mode="glove-25d-inst-areas", instance_segmentation_mode="coco"),
    "b5_plain": dict(encoder_name="efficientnet-b5", insertion_point="before-attn", semantics_mode=None,
                     instance_segmentation_mode=None),
    "b1_noadabins": dict(encoder_name="efficientnet-b1-noAdaBins", insertion_point="input", semantics_mode=None,
                         instance_segmentation_mode=None),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_state_dict_keys_match_reference(case, golden_state_keys):
    m = make_model(**CASES[case])
    mine = {k: "x".join(map(str, v.shape)) or "scalar" for k, v in m.state_dict().items() if not k.startswith("encoder.")}
    assert mine == golden_state_keys[case]


def test_num_channels_to_add():
    f = UnetAdaptiveBins.get_num_channels_to_add
    assert f("efficientnet-b1", "glove-25d-ade20k-places", None, "rgb") == 25          # config 2 -> 28 inputs
    assert f("efficientnet-b1", "glove-25d", "ade20k_swin_human_sizes", "rgb") == 70   # config 3 -> 73 inputs
    assert f("efficientnet-b1", "glove-25d-inst-areas", "coco", "rgb") == 70
    assert f("efficientnet-b1", "glove", None, "rgb") == 300
    assert f("efficientnet-b1", "raw", None, "rgb") == 1
    assert f("efficientnet-b1", None, None, "rgb") == 0
    # (f)4 extension: the reference has a params file for this mode but exits on it (unet_adaptive_bins.py:377-378)
    assert f("efficientnet-b1", "one-hot-ade20k-places", None, "rgb") == 101
    with pytest.raises(SystemExit):
        f("efficientnet-b1", "not-a-mode", None, "rgb")


def test_build_surface():
    m = UnetAdaptiveBins.build(n_bins=256, min_val=1e-3, max_val=10, norm="linear", encoder_name="efficientnet-b1",
                               semantics_mode="glove-25d-ade20k-places", instance_segmentation_mode=None,
                               insertion_point="input", image="rgb")
    assert m.encoder.original_model.conv_stem.weight.shape == (32, 28, 3, 3)
    n1 = sum(p.numel() for p in m.get_1x_lr_params())
    n10 = sum(p.numel() for p in m.get_10x_lr_params())
    assert n1 + n10 == sum(p.numel() for p in m.parameters())
    assert SILogLoss().name == "SILog" and BinsChamferLoss().name == "ChamferLoss"
    # geffnet child order (the Encoder indexes features by position)
    names = list(m.encoder.original_model._modules.keys())
    assert names == ["conv_stem", "bn1", "act1", "blocks", "conv_head", "bn2", "act2", "global_pool", "classifier"]


def test_backbone_feature_shapes():
    """features [4],[5],[6],[8],[11] = 16/24/40/112/1280 channels at /2 /4 /8 /16 /32 for B1."""
    m = make_model(**CASES["b1_plain"])
    with torch.no_grad():
        feats = m.encoder(torch.zeros(1, 3, 64, 96))
    got = [(feats[i].shape[1], feats[i].shape[2], feats[i].shape[3]) for i in (4, 5, 6, 8, 11)]
    assert got == [(16, 32, 48), (24, 16, 24), (40, 8, 12), (112, 4, 6), (1280, 2, 3)]
    # the decoder itself has no CPU path in the product; its CPU restatement gives the shape contract
    from oracle import adabins_oracle as oracle
    with torch.no_grad():
        assert [f.shape for f in oracle.encoder_features(m.encoder.original_model, torch.zeros(1, 3, 64, 96))] == \
            [f.shape for f in feats]
        out = oracle.decoder_bn(feats, m.state_dict())
    assert out.shape == (1, 128, 32, 48)
    with pytest.raises(_lib.MdeError):
        m.decoder(feats)


def test_backbone_bn_folding_matches_unfolded():
    """Inference-time conv+BatchNorm folding inside the EfficientNet blocks gives the unfolded module's values."""
    from mde_biological_vision_systems_b200.models import efficientnet as E
    m = make_model(**CASES["b1_plain"])
    x = torch.randn(1, 3, 64, 96)
    with torch.no_grad():
        folded = m.encoder(x)
    ref = m.encoder(x)  # autograd enabled -> the BatchNorm modules run un-folded
    for a, b in zip(folded, ref):
        assert torch.allclose(a, b.detach(), rtol=1e-4, atol=1e-5)
    blk = m.encoder.original_model.blocks[1][0]
    w0 = E._folded_conv_bn(blk.conv_pw, blk.bn1)[0]
    with torch.no_grad():
        blk.bn1.running_var.mul_(2.0)  # in-place update bumps the version -> the cached fold is rebuilt
    assert not torch.equal(E._folded_conv_bn(blk.conv_pw, blk.bn1)[0], w0)


def test_abi_exports_every_declared_symbol():
    """Every function declared in include/mde_b200.h is exported by the built library and bound in _lib.py."""
    header = open(os.path.join(ROOT, "include", "mde_b200.h")).read()
    declared = set(re.findall(r"\b(mde_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.load(check_device=False).mde_version() >= 100


def test_no_cpu_fallback():
    """Operators refuse CPU tensors / a box without a B200 instead of silently computing elsewhere."""
    if torch.cuda.is_available():
        pytest.skip("only meaningful on a CPU-only box")
    with pytest.raises(_lib.MdeError):
        ops.silog(torch.ones(1, 1, 4, 4), torch.ones(1, 1, 8, 8))
    with pytest.raises(_lib.MdeError):
        SILogLoss()(torch.ones(1, 1, 4, 4), torch.ones(1, 1, 8, 8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mde_biological_vision_systems_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no CPU oracle", ""), f"{f} mentions the oracle"


def test_label_io_readers_follow_the_dataloader(tmp_path):
    """(f)3 host side: the on-disk label formats and their 'no prediction' sentinels (dataloader.py:98-161)."""
    import numpy as np
    from mde_biological_vision_systems_b200 import label_io
    hw = (6, 8)
    rng = np.random.default_rng(0)
    sem150 = rng.integers(0, 150, hw).astype(np.int64)
    np.save(tmp_path / "semantic_seg_0.npy", sem150)
    out = label_io.load_semantic_labels(str(tmp_path / "semantic_seg_0.npy"), "glove-25d", hw)
    assert out.dtype == np.uint8 and np.array_equal(out, sem150.astype(np.ubyte))
    inst = rng.integers(-1, 101, hw).astype(np.int32)
    areas = rng.integers(0, 5000, hw).astype(np.int32)
    np.savez(tmp_path / "instance_labels_ade20k_swin_0.npz", inst)
    np.savez(tmp_path / "instance_areas_ade20k_swin_0.npz", areas)
    np.savez(tmp_path / "instance_labels_ade20k_swin_1.npz", np.array(None, dtype=object))   # detector found nothing
    np.savez(tmp_path / "instance_areas_ade20k_swin_1.npz", np.array(None, dtype=object))
    # ade20k-places semantics are read from the Swin label file and cast to ubyte: -1 arrives as 255
    sem = label_io.load_semantic_labels(str(tmp_path / "instance_labels_ade20k_swin_0.npz"), "glove-25d-ade20k-places", hw)
    assert sem.dtype == np.uint8 and np.array_equal(sem, inst.astype(np.ubyte)) and (sem[inst == -1] == 255).all()
    empty = label_io.load_semantic_labels(str(tmp_path / "instance_labels_ade20k_swin_1.npz"), "glove-25d-ade20k-places", hw)
    assert empty.shape == hw and (empty == 255).all()
    l0, a0 = label_io.load_instance_maps(str(tmp_path / "instance_labels_ade20k_swin_0.npz"),
                                         str(tmp_path / "instance_areas_ade20k_swin_0.npz"), "ade20k_swin", hw)
    assert l0.dtype == np.int32 and np.array_equal(l0, inst) and np.array_equal(a0, areas)
    l1, a1 = label_io.load_instance_maps(str(tmp_path / "instance_labels_ade20k_swin_1.npz"),
                                         str(tmp_path / "instance_areas_ade20k_swin_1.npz"), "ade20k_swin_human_sizes", hw)
    assert (l1 == -1).all() and (a1 == 0).all() and l1.shape == hw
    batch = label_io.collate([{"semantics": sem, "instance_labels": l0, "instance_areas": a0},
                              {"semantics": empty, "instance_labels": l1, "instance_areas": a1}], pin=False)
    assert batch["semantics"].shape == (2, 1, 6, 8) and batch["semantics"].dtype == torch.uint8
    assert batch["instance_labels"].dtype == torch.int32 and batch["instance_areas"].dtype == torch.int32


def test_evaluation_helpers_cpu_side():
    """crop boxes and the running average of the evaluation module (host logic of (f)2)."""
    from oracle import adabins_oracle as oracle
    from mde_biological_vision_systems_b200 import evaluation
    for (h, w, garg, eigen, ds) in [(480, 640, False, True, "nyu"), (352, 1216, True, False, "kitti"),
                                    (352, 1216, False, True, "kitti"), (96, 128, False, False, "nyu")]:
        assert evaluation.crop_box(h, w, garg, eigen, ds) == oracle.eval_crop_box(h, w, garg, eigen, ds)
    avg = evaluation.RunningAverageDict()
    avg.update({"a1": 1.0, "rmse": 2.0})
    avg.update({"a1": 0.0, "rmse": 4.0})
    assert avg.get_value() == {"a1": 0.5, "rmse": 3.0}


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's own modules from oracle/_ref on the host cores; the oracle port where that
    directory was never staged) prints one JSON line with the contract's keys; it needs no GPU."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--batch", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mpix/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "BASELINE config 2" in line["config"]["workload"]


def test_torch_library_ops_registered_with_fake_impls():
    """torch.ops.mde.*: every hot-path operator is registered with torch.library (schema + fake implementation), so shapes
    propagate under FakeTensorMode without a GPU; a real CPU tensor is rejected (there is no CPU implementation)."""
    import pytest
    from torch._subclasses.fake_tensor import FakeTensorMode
    from mde_biological_vision_systems_b200 import _lib, torch_ops
    for name in torch_ops.OPS:
        assert hasattr(torch.ops.mde, name), name
    with FakeTensorMode():
        lab = torch.empty((2, 1, 32, 48), dtype=torch.int64, device="cuda")
        table = torch.empty((101, 25), dtype=torch.float32, device="cuda")
        assert torch.ops.mde.gather_embed(lab, table, 100).shape == (2, 25, 32, 48)
        x = torch.empty((2, 128, 32, 48), device="cuda")
        planes = torch.ops.mde.split_bf16(x)
        assert planes.shape == (2, 2, 32, 48, 128) and planes.dtype == torch.bfloat16
        w = torch.empty((2, 3, 3, 128, 128), dtype=torch.bfloat16, device="cuda")
        y = torch.ops.mde.conv3x3_x3(planes, w, None, None, 1.0, False)
        assert y.shape == (2, 128, 32, 48) and y.is_contiguous(memory_format=torch.channels_last)
        assert torch.ops.mde.conv3x3_x3(planes, w, None, None, 1.0, True).shape == (2, 2, 32, 48, 128)
        wq, bq = torch.ops.mde.fold_queries(torch.empty((256, 128, 1, 1), device="cuda"), torch.empty(256, device="cuda"),
                                            torch.empty((2, 128, 128), device="cuda"), None)
        assert wq.shape == (2, 2, 256, 128) and bq.shape == (2, 256)
        pred = torch.ops.mde.head_chain(planes, wq, bq, torch.empty((2, 256), device="cuda"))
        assert pred.shape == (2, 1, 32, 48)
        depth = torch.empty((2, 1, 64, 96), device="cuda")
        loss, ws = torch.ops.mde.silog_fwd(pred, depth, None, True)
        assert loss.shape == () and ws.dtype == torch.uint8
        s, c, _, _ = torch.ops.mde.depth_losses_fwd(pred, torch.empty((2, 257), device="cuda"), depth, 1e-3, 1e-3, True, False)
        assert s.shape == () and c.shape == ()
    with pytest.raises(_lib.MdeError):
        torch.ops.mde.split_bf16(torch.zeros(1, 8, 4, 4))


def test_gradient_arena_clip_matches_torch_clip():
    """GradientAverager.clip_grad_norm_ (norm of the bucket norms, one scale per bucket) == nn.utils.clip_grad_norm_ (train.py:427),
    and zero_grad keeps the p.grad views (re-attaching only when something replaced them)."""
    from mde_biological_vision_systems_b200.parallel import GradientAverager
    torch.manual_seed(11)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.ReLU(), torch.nn.Conv2d(8, 4, 1), torch.nn.Flatten(),
                              torch.nn.Linear(4 * 6 * 6, 5))
    ref = [p.detach().clone().requires_grad_(True) for p in net.parameters()]
    avg = GradientAverager(net.parameters(), bucket_mb=0.0005)  # several buckets
    assert len(avg.buckets) > 1
    x = torch.randn(2, 3, 8, 8)
    for mx in (0.1, 1e6):  # clipping active / inactive
        avg.zero_grad()
        views = [p.grad for p in net.parameters()]
        (net(x) ** 2).sum().backward()
        assert all(p.grad is v for p, v in zip(net.parameters(), views))
        for r, p in zip(ref, net.parameters()):
            r.grad = p.grad.detach().clone()
        t_ref = torch.nn.utils.clip_grad_norm_(ref, mx)
        t_ours = avg.clip_grad_norm_(mx)
        assert torch.allclose(t_ours, t_ref, rtol=1e-6)
        for r, p in zip(ref, net.parameters()):
            assert torch.allclose(p.grad, r.grad, rtol=1e-6, atol=1e-12)
    for p in net.parameters():  # something replaces the gradients: the next zero_grad re-attaches
        p.grad = None
    avg.zero_grad()
    assert all(p.grad is not None and float(p.grad.abs().max()) == 0.0 for p in net.parameters())


def test_train_step_host_protocol_and_deferred_checks():
    """Host-side pieces of training.GraphedTrainStep that need no GPU: TrainStep is split into forward_backward() (what the graph
    replays) and update() (what stays eager), GraphedTrainStep refuses a SyncBatchNorm flavour that would call NCCL inside the
    capture, and the deferred out-of-range check of a captured gather raises IndexError like the reference's index_select."""
    import inspect
    from mde_biological_vision_systems_b200 import ops, training
    assert {"forward_backward", "update", "__call__"} <= set(vars(training.TrainStep))
    src = inspect.getsource(training.TrainStep.__call__)
    assert "forward_backward" in src and "update" in src
    assert issubclass(training.GraphedTrainStep, training.TrainStep)
    # nothing to capture around: a plain module passes, its peer-memory arena is None
    assert training.GraphedTrainStep._p2p_arena(torch.nn.Sequential(torch.nn.Conv2d(3, 4, 1), torch.nn.BatchNorm2d(4))) is None
    # deferred label checks: flags collected between begin / end, read (and raised) afterwards
    ops.begin_deferred_checks()
    assert ops._DEFERRED == []
    ops._DEFERRED.append((torch.zeros(1, dtype=torch.int32), 150))
    flags = ops.end_deferred_checks()
    assert ops._DEFERRED is None and len(flags) == 1
    ops.raise_deferred_checks(flags)          # flag clear: passes
    flags[0][0].fill_(1)                      # a replay met an out-of-range label
    with pytest.raises(IndexError):
        ops.raise_deferred_checks(flags)
