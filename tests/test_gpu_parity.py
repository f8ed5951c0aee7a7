"""GPU parity tests: every kernel, called through the C ABI (ctypes -> libmde_b200.so), against the CPU oracle on
the same seeded inputs, against the committed golden vectors produced by the reference modules, and -- at
BASELINE.json's full sizes -- through size-independent properties.

Tolerances (BASELINE.json north_star): gathers / label indexing bit-exact; depth maps and bin edges <= 1e-3
relative (fp32 / TF32); SILog and chamfer <= 1e-4 relative.
"""
import numpy as np
import pytest
import torch

from oracle import adabins_oracle as oracle
from mde_biological_vision_systems_b200 import ops, synthetic
from mde_biological_vision_systems_b200.ExternalInfoLoaders.InstanceSegmentationLoader import InstanceSegmentationLoader
from mde_biological_vision_systems_b200.ExternalInfoLoaders.SemanticsLoader import SemanticsLoader
from mde_biological_vision_systems_b200.loss import BinsChamferLoss, DepthLosses, SILogLoss
from mde_biological_vision_systems_b200.models.layers import PixelWiseDotProduct

from helpers import INST_MODES, SEM_MODES, digest, inst_labels, load_table, make_model, rel_err, rel_stats, sem_labels

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
REL_DEPTH = 1e-3  # depth maps / bin edges, fp32-TF32
REL_LOSS = 1e-4   # SILog / chamfer
# Every tensor-core product of the inference path is formed from split-bf16 pairs (hi*hi + mid*hi + hi*mid, fp32
# accumulation; ops.SplitBF16) and the cuDNN passthrough bodies run in true fp32 inside the model, so depth maps are held to
# 1e-3 on EVERY pixel in the default mode -- the mode bench.py measures.  These tests run with PyTorch's default backend
# flags (cudnn.allow_tf32 = True): exactness is the model's job, not the test fixture's.


def assert_depth_close(pred, ref, tol=REL_DEPTH):
    mx, p999 = rel_stats(pred, ref)
    assert mx < tol, (mx, p999)


def _as_f32(t):
    return t.float() if isinstance(t, ops.SplitBF16) else t


class Args:
    def __init__(self, **kw):
        self.use_semantics = kw.get("use_semantics")
        self.use_instance_segmentation = kw.get("use_instance_segmentation")


# ------------------------------------------------------------------------------------------------------------
# K3 loaders: bit exact
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", SEM_MODES)
def test_semantics_loader_golden(mode, golden_digests):
    lab, _ = sem_labels(mode)
    batch = {"semantics": lab.clone()}
    loader = SemanticsLoader(Args(use_semantics=mode))
    loader.clamp_host_batch = True
    raw, sem = loader.get_semantics(batch)
    assert sem.is_cuda and raw.is_cuda
    assert digest(raw.cpu().numpy()) == golden_digests[f"sem/{mode}/raw"]
    assert digest(sem.cpu().numpy()) == golden_digests[f"sem/{mode}/out"]
    if "ade20k-places" in mode:  # the reference clamps the caller's batch tensor in place
        assert int(batch["semantics"].max()) <= 100 and int(batch["semantics"].min()) >= 0


@pytest.mark.parametrize("mode", INST_MODES)
def test_instance_loader_golden(mode, golden_digests):
    lab, areas = inst_labels(mode)
    raw, emb, ar = InstanceSegmentationLoader(Args(use_instance_segmentation=mode)).get_instance_segmentation(
        {"instance_labels": lab.clone(), "instance_areas": areas.clone()})
    assert digest(raw.cpu().numpy()) == golden_digests[f"inst/{mode}/raw"]
    assert digest(emb.cpu().numpy()) == golden_digests[f"inst/{mode}/emb"]
    assert digest(ar.cpu().numpy()) == golden_digests[f"inst/{mode}/areas"]


def test_gather_full_size_and_ragged():
    """Config 2 image size (416x544, 4 images to keep the CPU oracle quick) against the oracle, plus ragged sizes that take the scalar path."""
    table64 = load_table("ade20k_places_classes_glove_twitter_27b_25d_embeddings.npy")
    for (b, h, w) in [(4, 416, 544), (3, 37, 53), (1, 1, 1), (2, 5, 4)]:
        lab, _ = synthetic.label_maps(b, h, w, seed=100 + h, n_rect=(1, 6) if h < 40 else (20, 60))
        raw_ref, ref = oracle.semantics_loader("glove-25d-ade20k-places", lab.numpy(), table64)
        labd = lab.to(DEV)
        out = ops.gather_embed(labd, torch.from_numpy(table64).float().to(DEV), background=100, write_back=True)
        assert np.array_equal(out.cpu().numpy(), ref)
        assert np.array_equal(labd.cpu().numpy(), raw_ref)
    # float64 output (instance embedding / 150-class table) and the 300-d table that does not fit shared memory
    lab, _ = synthetic.label_maps(2, 64, 96, seed=7, lo=0, hi=149, inject=())
    for name in ("ade20k_150_classes_glove_twitter_27b_25d_embeddings.npy", "ade20k_150_classes_glove_840b_300d_embeddings.npy"):
        t = load_table(name)
        out = ops.gather_embed(lab.to(DEV), torch.from_numpy(t).to(DEV), background=None)
        assert out.dtype == torch.float64
        assert np.array_equal(out.cpu().numpy(), oracle.gather_rows(t, lab.numpy()))


def test_gather_into_encoder_input_matches_planar():
    """SemanticsLoader.bind_encoder_input: the embedding planes gathered straight into the channels_last encoder input (with
    the stem's SAME padding) are bit-identical to the planar gather, the returned tensors keep the reference's shapes, and the
    model's output does not change by a bit."""
    mode = "glove-25d-ade20k-places"
    m = make_model(insertion_point="input", semantics_mode=mode, instance_segmentation_mode=None).to(DEV).channels_last_()
    lab, _ = sem_labels(mode, 2, 352, 384, seed=150, n_rect=(20, 40))
    plain, bound = SemanticsLoader(Args(use_semantics=mode)), SemanticsLoader(Args(use_semantics=mode))
    assert bound.bind_encoder_input(m)
    raw1, sem1 = plain.get_semantics({"semantics": lab.clone()})
    raw2, sem2 = bound.get_semantics({"semantics": lab.clone()})
    assert sem2.shape == sem1.shape and torch.equal(sem2, sem1) and torch.equal(raw2, raw1)
    assert hasattr(sem2, "_mde_encoder_input")
    x = synthetic.image(2, 352, 384, seed=151).to(DEV)
    with torch.no_grad():
        e1, p1 = m(x, semantics=sem1)
        e2, p2 = m(x, semantics=sem2)
    assert torch.equal(p1, p2) and torch.equal(e1, e2)
    # with the device image in the batch dict the same kernel also writes the RGB planes (nothing left for the model to copy)
    raw3, sem3 = bound.get_semantics({"semantics": lab.clone(), "image": x})
    assert torch.equal(sem3, sem1) and sem3._mde_encoder_input[2] is x
    assert torch.equal(sem3._mde_encoder_input[0][:, :3, :352, :384], x)
    with torch.no_grad():
        e3, p3 = m(x, semantics=sem3)
    assert torch.equal(p1, p3) and torch.equal(e1, e3)
    # a loader bound to a model it does not fit declines
    other = make_model(insertion_point="input", semantics_mode=None, instance_segmentation_mode=None).to(DEV).channels_last_()
    assert not SemanticsLoader(Args(use_semantics=mode)).bind_encoder_input(other)


@pytest.mark.parametrize("dtype", [torch.int64, torch.int32, torch.uint8])
@pytest.mark.parametrize("shape,pads,c_after", [((3, 37, 50), (1, 2, 2, 1), 0), ((2, 64, 96), (0, 1, 0, 1), 0),
                                                ((2, 19, 33), (0, 0, 0, 0), 4)])
def test_gather_embed_nhwc_kernels_bit_exact(dtype, shape, pads, c_after):
    """ops.gather_embed_nhwc with the image planes fused in: dense pixel records (c_after = 0: the warp-staged kernel, ragged
    row ends, every label type) and records inside a wider pitch (c_after > 0: the per-piece kernel) -- bit-exact against
    torch indexing, borders zero, labels_out clamped."""
    b, h, w = shape
    rng = np.random.default_rng(170)
    table = torch.from_numpy(load_table("ade20k_places_classes_glove_twitter_27b_25d_embeddings.npy")).float().to(DEV)
    rows, d = table.shape
    hi = 255 if dtype == torch.uint8 else 400
    lab = torch.from_numpy(rng.integers(0 if dtype == torch.uint8 else -3, hi, size=(b, 1, h, w))).to(dtype).to(DEV)
    img = torch.from_numpy(rng.standard_normal((b, 3, h, w)).astype(np.float32)).to(DEV)
    lab_out = torch.empty((b, 1, h, w), dtype=torch.int64, device=DEV)
    buf, view = ops.gather_embed_nhwc(lab, table, rows - 1, 3, c_after, pads, labels_out=lab_out, image=img)
    l64 = lab.long()
    clamped = torch.where((l64 < 0) | (l64 > rows - 1), torch.full_like(l64, rows - 1), l64)
    ref = table[clamped[:, 0]].permute(0, 3, 1, 2)
    pt, pb, pl, pr = pads
    assert torch.equal(lab_out, clamped)
    assert torch.equal(view, ref) and torch.equal(buf[:, :3, pt:pt + h, pl:pl + w], img)
    inner = torch.zeros_like(buf, dtype=torch.bool)
    inner[:, :, pt:pt + h, pl:pl + w] = True
    if any(pads):
        assert float(buf[:, :3 + d][~inner[:, :3 + d]].abs().max()) == 0.0


def test_gather_out_of_range_raises():
    lab = torch.zeros(1, 1, 8, 8, dtype=torch.int64)
    lab[0, 0, 3, 3] = 150
    t = torch.from_numpy(load_table("ade20k_150_classes_glove_twitter_27b_25d_embeddings.npy")).to(DEV)
    with pytest.raises(IndexError):
        ops.gather_embed(lab.to(DEV), t, background=None)


def test_gather_empty():
    t = torch.from_numpy(load_table("ade20k_classes_abs_sizes.npy")).float().to(DEV)
    out = ops.gather_embed(torch.zeros(0, 1, 4, 4, dtype=torch.int64, device=DEV), t, background=100)
    assert out.shape == (0, 3, 4, 4)


def test_class_area_fraction():
    lab, _ = synthetic.label_maps(3, 120, 160, seed=5, lo=0, hi=149, inject=())
    out = ops.class_area_fraction(lab.to(DEV), 150)
    assert np.array_equal(out.cpu().numpy(), oracle.class_area_fraction(lab.numpy()))


# ------------------------------------------------------------------------------------------------------------
# A3 aux MLP and insertion
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cin,hw", [(1, (64, 96)), (3, (64, 96)), (3, (17, 23))])
def test_aux_mlp_forward_backward(cin, hw):
    seq = torch.nn.Sequential(torch.nn.Conv2d(cin, 10, 1), torch.nn.ReLU(), torch.nn.Conv2d(10, 10, 1), torch.nn.ReLU())
    synthetic.fill_state_dict(seq, seed=3)
    with torch.no_grad():
        for p in seq.parameters():
            p.mul_(8.0)  # make both ReLU branches common
    rng = np.random.default_rng(4)
    x = torch.from_numpy(rng.standard_normal((2, cin, *hw)).astype(np.float32) * 50)
    div = 7.0
    ref = seq(x / div)
    gout = torch.from_numpy(rng.standard_normal(ref.shape).astype(np.float32))
    ref.backward(gout)
    ref_grads = [p.grad.clone() for p in seq.parameters()]
    seq_d = torch.nn.Sequential(torch.nn.Conv2d(cin, 10, 1), torch.nn.ReLU(), torch.nn.Conv2d(10, 10, 1), torch.nn.ReLU())
    seq_d.load_state_dict(seq.state_dict())
    seq_d.to(DEV)
    out = ops.aux_mlp(x.to(DEV), seq_d[0].weight, seq_d[0].bias, seq_d[2].weight, seq_d[2].bias, in_div=div)
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().numpy(), rtol=1e-5, atol=1e-5)
    out.backward(gout.to(DEV))
    for p, g in zip(seq_d.parameters(), ref_grads):
        np.testing.assert_allclose(p.grad.cpu().numpy().reshape(g.shape), g.numpy(), rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "areas", "hsizes"])
def test_input_insertion_golden(name, golden):
    cases = {
        "cfg2": dict(semantics_mode="glove-25d-ade20k-places", instance_segmentation_mode=None),
        "cfg3": dict(semantics_mode="glove-25d", instance_segmentation_mode="ade20k_swin_human_sizes"),
        "areas": dict(semantics_mode="glove-25d-inst-areas", instance_segmentation_mode="coco"),
        "hsizes": dict(semantics_mode="glove-25d-ade20k-places-human-sizes", instance_segmentation_mode="ade20k_swin"),
    }
    kw = cases[name]
    m = make_model(insertion_point="input", image="rgb", **kw).to(DEV)
    h = w = 32
    img = synthetic.image(2, h, w, seed=41).to(DEV)
    slab, _ = sem_labels(kw["semantics_mode"], 2, h, w, seed=42, n_rect=(5, 10))
    _, sem = SemanticsLoader(Args(use_semantics=kw["semantics_mode"])).get_semantics({"semantics": slab})
    il = ia = None
    if kw["instance_segmentation_mode"]:
        ilab, iar = inst_labels(kw["instance_segmentation_mode"], 2, h, w, seed=43, n_rect=(5, 10))
        _, il, ia = InstanceSegmentationLoader(Args(use_instance_segmentation=kw["instance_segmentation_mode"])) \
            .get_instance_segmentation({"instance_labels": ilab, "instance_areas": iar})
    with torch.no_grad():
        x = m._concat_external(img, m._external_channels(sem, il, ia, h * w))
    np.testing.assert_allclose(x.cpu().numpy(), golden[f"insert/{name}"], rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------------------------
# head pieces
# ------------------------------------------------------------------------------------------------------------
def _head_state():
    m = make_model(insertion_point="input", semantics_mode=None, instance_segmentation_mode=None)
    return m, {k: v for k, v in m.state_dict().items()}


def test_patch_transformer_golden(golden):
    """patch-embed + 4 encoder layers (our fp32 kernels) vs the reference module's tokens [S, N, E]."""
    m, _ = _head_state()
    m.to(DEV)
    x = synthetic.decoder_features(2, 128, 176, 192, seed=21).to(DEV)
    with torch.no_grad():
        tgt = m.adaptive_bins_layer.patch_transformer(x.clone())
    np.testing.assert_allclose(tgt.cpu().numpy(), golden["head/tgt"], rtol=1e-3, atol=2e-4)


def test_encoder_layers_tc_vs_simt_and_fp64(golden):
    """3xTF32 tcgen05 encoder layers vs the exact-fp32 SIMT kernels and vs a float64 evaluation of the torch layers."""
    m, _ = _head_state()
    pt = m.adaptive_bins_layer.patch_transformer.to(DEV)
    rng = np.random.default_rng(99)
    for s_len, nb in [(221, 3), (132, 2), (300, 1)]:
        tok = torch.from_numpy(rng.standard_normal((s_len, nb, 128)).astype(np.float32)).to(DEV)
        layers = list(pt.transformer_encoder.layers)
        with torch.no_grad():
            out_tc = ops.encoder_layers_tc(tok, layers, pt._prepared_layers(layers))
            cur, ws = tok, None
            for layer in layers:
                cur, ws = ops.encoder_layer(cur, layer, ws)
            ref64 = pt.transformer_encoder.double().eval()(tok.double())
            pt.transformer_encoder.float()
        scale = float(ref64.abs().max())
        e_tc = float((out_tc.double() - ref64).abs().max()) / scale
        e_simt = float((cur.double() - ref64).abs().max()) / scale
        assert e_tc < 2e-5 and e_simt < 2e-5, (e_tc, e_simt)


@pytest.mark.parametrize("shape", [(2, 128, 176, 192), (1, 128, 208, 272), (3, 128, 240, 320)])
def test_patch_embed_tc(shape):
    """tcgen05 split-K patch-embedding GEMM (split-bf16 NHWC input, three bf16 products) vs the fp32 conv + positional rows
    of the reference."""
    m, _ = _head_state()
    pt = m.adaptive_bins_layer.patch_transformer.to(DEV)
    x = synthetic.decoder_features(*shape, seed=77)
    with torch.no_grad():
        ref = pt.embedding_convPxP.cpu()(x).flatten(2) + pt.positional_encodings.cpu()[: (shape[2] // 16) * (shape[3] // 16)].T
        ref = ref.permute(2, 0, 1)
        pt.to(DEV)
        xp = ops.split_bf16(x.to(DEV))
        assert ops.patch_embed_supported(xp, pt.embedding_convPxP)
        tok = ops.patch_embed(xp, pt._prepared_weight(), pt.embedding_convPxP.bias, pt.positional_encodings, 16)
    assert tok.shape == ref.shape
    err = float((tok.cpu() - ref).abs().max()) / float(ref.abs().max())
    assert err < 3e-5, err


@pytest.mark.parametrize("cfg", [
    dict(shape=(2, 128, 48, 64), cout=128),                       # head conv3x3: NT = 2, 16-px patches, exact tiling
    dict(shape=(1, 128, 26, 34), cout=128, pair_out=True),        # ragged: edge tiles zero-filled / clipped by TMA
    dict(shape=(2, 176, 40, 24), cout=80, affine=True, slope=0.01),   # decoder up4-like: N tile 80, C % 32 != 0
    dict(shape=(1, 344, 13, 17), cout=160, affine=True, slope=0.01, pair_out=True),  # up3-like: N tile 160 (one patch per CTA)
    dict(shape=(1, 680, 9, 11), cout=320, affine=True, slope=0.01),   # up2-like: two N tiles of 160
    dict(shape=(1, 96, 7, 5), cout=640, affine=True, slope=0.01, pair_out=True),     # five N tiles of 128
    dict(shape=(1, 80, 16, 16), cout=128, pair_out=True),
    dict(shape=(1, 1392, 15, 19), cout=640, affine=True, slope=0.01),  # B1 up1: 44 K chunks, ragged last chunk
    dict(shape=(1, 552, 9, 11), cout=256, affine=True, slope=0.01, pair_out=True),  # B5 up3: 256 = 2 N tiles of 128 (a 256-wide stage does not fit twice)
    dict(shape=(1, 64, 5, 7), cout=1024, affine=True, slope=0.01),     # B5 up1 width: 8 N tiles of 128
])
def test_conv3x3_tc(cfg):
    """tcgen05 implicit-GEMM 3x3 conv (NHWC split-bf16 pairs, three bf16 products per K step) vs an fp64 conv of the
    reference's Conv2d(+BN eval affine+LeakyReLU): fp32-grade, in both output formats."""
    rng = np.random.default_rng(91)
    b, c, h, w = cfg["shape"]
    cout = cfg["cout"]
    x = torch.from_numpy(rng.standard_normal((b, c, h, w)).astype(np.float32))
    wt = torch.from_numpy((rng.standard_normal((cout, c, 3, 3)) / np.sqrt(9 * c)).astype(np.float32))
    bias = torch.from_numpy(rng.standard_normal(cout).astype(np.float32))
    ref = torch.nn.functional.conv2d(x.double(), wt.double(), None, padding=1)
    if cfg.get("affine"):
        scale = torch.from_numpy((0.5 + rng.random(cout)).astype(np.float32))
        ref = ref * scale.double().view(1, -1, 1, 1) + bias.double().view(1, -1, 1, 1)
    else:
        scale = None
        ref = ref + bias.double().view(1, -1, 1, 1)
    slope = cfg.get("slope", 1.0)
    ref = torch.where(ref > 0, ref, ref * slope)
    pair_out = cfg.get("pair_out", False)
    xp = ops.split_bf16(x.to(DEV).contiguous(memory_format=torch.channels_last))
    assert ops.conv3x3_supported(xp, cout, pair_out)
    out = ops.conv3x3_nhwc(xp, ops.prepare_conv3x3_weight(wt.to(DEV)), None if scale is None else scale.to(DEV),
                           bias.to(DEV), slope=slope, pair_out=pair_out)
    if pair_out:
        assert isinstance(out, ops.SplitBF16)
        out = out.float()
    assert out.shape == ref.shape and out.is_contiguous(memory_format=torch.channels_last)
    err = float((out.cpu().double() - ref).abs().max()) / float(ref.abs().max())
    assert err < 5e-5, err


@pytest.mark.parametrize("cfg", [
    dict(shape=(2, 128, 24, 40), cout=128),                 # head conv
    dict(shape=(2, 176, 20, 28), cout=80),                  # decoder up4 conv_a (C_in = 176: dgrad with a 176-wide output)
    dict(shape=(1, 344, 13, 17), cout=160),                 # up3 conv_a: dgrad output width 344 (ragged N tiles)
    dict(shape=(3, 80, 9, 11), cout=128, bias=False),       # conv3-like, bias-free
])
def test_conv3x3_autograd(cfg):
    """Training form of the 3x3 conv (ops.conv3x3_autograd): forward on split-bf16 pairs, dgrad = the same kernel on the
    flipped / transposed filter, wgrad = tap-shifted TF32 NT GEMM over the padded pixel axis -- vs float64 autograd of
    F.conv2d (the reference's nn.Conv2d, models/miniViT.py:16, unet_adaptive_bins.py:43-48)."""
    rng = np.random.default_rng(140)
    b, c, h, w = cfg["shape"]
    cout = cfg["cout"]
    x = torch.from_numpy(rng.standard_normal((b, c, h, w)).astype(np.float32))
    wt = torch.from_numpy((rng.standard_normal((cout, c, 3, 3)) / np.sqrt(9 * c)).astype(np.float32))
    bias = torch.from_numpy(rng.standard_normal(cout).astype(np.float32)) if cfg.get("bias", True) else None
    g = torch.from_numpy(rng.standard_normal((b, cout, h, w)).astype(np.float32))
    xr, wr = x.double().requires_grad_(True), wt.double().requires_grad_(True)
    br = bias.double().requires_grad_(True) if bias is not None else None
    yr = torch.nn.functional.conv2d(xr, wr, br, padding=1)
    yr.backward(g.double())
    xd = x.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    wd = wt.to(DEV).requires_grad_(True)
    bd = bias.to(DEV).requires_grad_(True) if bias is not None else None
    conv = torch.nn.Conv2d(c, cout, 3, padding=1)
    assert ops.conv3x3_train_supported(xd, conv)
    y = ops.conv3x3_autograd(xd, wd, bd)
    assert float((y.detach().cpu().double() - yr.detach()).abs().max()) < 5e-5 * float(yr.abs().max())
    y.backward(g.to(DEV))
    for name, a, r in (("x", xd.grad, xr.grad), ("w", wd.grad, wr.grad)) + ((("b", bd.grad, br.grad),) if bias is not None else ()):
        err = float((a.cpu().double() - r).abs().max()) / float(r.abs().max())
        assert err < (1e-4 if name == "x" else 2e-3), (name, err)


@pytest.mark.parametrize("cfg", [
    dict(shape=(2, 16, 40, 56), cout=96, act=1),                    # EfficientNet expansion: K = 16 (one partial K chunk), SiLU
    dict(shape=(1, 96, 26, 34), cout=24, res=True),                 # projection with residual, N = 24 (ragged N tile)
    dict(shape=(2, 1280, 15, 19), cout=1280),                       # DecoderBN.conv2 (input already padded): 20 K chunks
    dict(shape=(1, 672, 13, 17), cout=112, res=True),
    dict(shape=(1, 1152, 7, 9), cout=320, act=1),
    dict(shape=(3, 40, 9, 11), cout=240, act=1),                    # K = 40: not a multiple of the 32-float TMA box
])
def test_pointwise_conv_x3(cfg):
    """1x1 conv as a tcgen05 GEMM whose fp32 activations are split into bf16 pairs in shared memory (ops.pointwise_conv):
    bias + SiLU + residual epilogue, vs a float64 evaluation -- fp32-grade."""
    rng = np.random.default_rng(180)
    b, c, h, w = cfg["shape"]
    cout = cfg["cout"]
    x = torch.from_numpy(rng.standard_normal((b, c, h, w)).astype(np.float32))
    wt = torch.from_numpy((rng.standard_normal((cout, c, 1, 1)) / np.sqrt(c)).astype(np.float32))
    bias = torch.from_numpy(rng.standard_normal(cout).astype(np.float32))
    res = torch.from_numpy(rng.standard_normal((b, cout, h, w)).astype(np.float32)) if cfg.get("res") else None
    ref = torch.nn.functional.conv2d(x.double(), wt.double(), bias.double())
    if cfg.get("act"):
        ref = ref * torch.sigmoid(ref)
    if res is not None:
        ref = ref + res.double()
    xd = x.to(DEV).contiguous(memory_format=torch.channels_last)
    assert ops.pointwise_supported(xd, c, cout)
    out = ops.pointwise_conv(xd, ops.prepare_pointwise_weight(wt.to(DEV)), bias.to(DEV), cfg.get("act", 0),
                             None if res is None else res.to(DEV).contiguous(memory_format=torch.channels_last))
    assert out.shape == ref.shape and out.is_contiguous(memory_format=torch.channels_last)
    err = float((out.cpu().double() - ref).abs().max()) / float(ref.abs().max())
    assert err < 5e-5, err


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [
    dict(shape=(2, 96, 23, 31), cout=24, gate=True),
    dict(shape=(3, 144, 13, 17), cout=40, gate=True, res=True),       # 3 images inside / across 128-row tiles
    dict(shape=(2, 16, 24, 40), cout=96, act=1, pads=(0, 1, 0, 1)),
    dict(shape=(2, 24, 13, 17), cout=144, act=1, pads=(1, 2, 2, 1), gate=True),
])
def test_pointwise_conv_gate_and_padded_output(cfg):
    """ops.pointwise_conv with the squeeze-excite gate multiplied into its input inside the GEMM and / or the result written
    inside a zero border (TensorFlow-SAME padding of the stride-2 depthwise conv that follows), vs float64."""
    rng = np.random.default_rng(181)
    b, c, h, w = cfg["shape"]
    cout = cfg["cout"]
    x = torch.from_numpy(rng.standard_normal((b, c, h, w)).astype(np.float32))
    wt = torch.from_numpy((rng.standard_normal((cout, c, 1, 1)) / np.sqrt(c)).astype(np.float32))
    bias = torch.from_numpy(rng.standard_normal(cout).astype(np.float32))
    gate = torch.from_numpy(rng.random((b, c)).astype(np.float32)) if cfg.get("gate") else None
    res = torch.from_numpy(rng.standard_normal((b, cout, h, w)).astype(np.float32)) if cfg.get("res") else None
    xin = x if gate is None else x * gate[:, :, None, None]  # fp32 product, as the reference forms it
    ref = torch.nn.functional.conv2d(xin.double(), wt.double(), bias.double())
    if cfg.get("act"):
        ref = ref * torch.sigmoid(ref)
    if res is not None:
        ref = ref + res.double()
    pads = cfg.get("pads")
    if pads:
        ref = torch.nn.functional.pad(ref, (pads[2], pads[3], pads[0], pads[1]))
    xd = x.to(DEV).contiguous(memory_format=torch.channels_last)
    out = ops.pointwise_conv(xd, ops.prepare_pointwise_weight(wt.to(DEV)), bias.to(DEV), cfg.get("act", 0),
                             None if res is None else res.to(DEV).contiguous(memory_format=torch.channels_last),
                             gate=None if gate is None else gate.to(DEV), out_pads=pads)
    assert out.shape == ref.shape and out.is_contiguous(memory_format=torch.channels_last)
    err = float((out.cpu().double() - ref).abs().max()) / float(ref.abs().max())
    assert err < 5e-5, err
    if pads:  # the border is exactly zero
        inner = torch.zeros_like(out, dtype=torch.bool)
        inner[:, :, pads[0]:pads[0] + h, pads[2]:pads[2] + w] = True
        assert float(out[~inner].abs().max()) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("shape,r", [((2, 96, 23, 31), 4), ((16, 1152, 13, 17), 48), ((3, 40, 5, 7), 10)])
def test_squeeze_excite_kernels(shape, r):
    """bias + SiLU + slab sums in one pass (ops.bias_act_pool_nhwc_) and the fused gate kernel (ops.se_gate) vs torch."""
    rng = np.random.default_rng(182)
    b, c, h, w = shape
    x = torch.from_numpy(rng.standard_normal(shape).astype(np.float32)).to(DEV).contiguous(memory_format=torch.channels_last)
    bias = torch.from_numpy(rng.standard_normal(c).astype(np.float32)).to(DEV)
    w1 = torch.from_numpy((rng.standard_normal((r, c)) / np.sqrt(c)).astype(np.float32)).to(DEV)
    b1 = torch.from_numpy(rng.standard_normal(r).astype(np.float32)).to(DEV)
    w2 = torch.from_numpy((rng.standard_normal((c, r)) / np.sqrt(r)).astype(np.float32)).to(DEV)
    b2 = torch.from_numpy(rng.standard_normal(c).astype(np.float32)).to(DEV)
    ref_y = torch.nn.functional.silu(x.double() + bias.double()[None, :, None, None])
    m = ref_y.mean((2, 3))
    hdn = torch.nn.functional.silu(m @ w1.double().T + b1.double())
    ref_gate = torch.sigmoid(hdn @ w2.double().T + b2.double())
    y, partial = ops.bias_act_pool_nhwc_(x.clone(memory_format=torch.channels_last), bias, 1)
    assert float((y.double() - ref_y).abs().max()) < 1e-5
    assert partial.shape[0] == b and partial.shape[2] == c
    assert float((partial.double().sum(1) / (h * w) - m).abs().max()) < 1e-5
    gate = ops.se_gate(partial, h * w, w1, b1, w2, b2)
    assert float((gate.double() - ref_gate).abs().max()) < 1e-5


@pytest.mark.parametrize("cfg", [
    dict(shape=(2, 96, 23, 31), k=3, stride=1), dict(shape=(2, 144, 24, 32), k=5, stride=2), dict(shape=(3, 32, 17, 19), k=3, stride=2),
    dict(shape=(2, 1152, 7, 9), k=5, stride=1), dict(shape=(1, 40, 9, 33), k=5, stride=2),
])
def test_depthwise_bias_act_pool(cfg):
    """ops.depthwise_bias_act_pool (depthwise conv + bias + SiLU + slab sums in one pass) vs float64 torch: stride 1 with
    symmetric padding and stride 2 with TensorFlow-SAME (asymmetric) padding, k = 3 / 5, C up to 1152 (two channel passes)."""
    rng = np.random.default_rng(183)
    b, c, h, w = cfg["shape"]
    k, s = cfg["k"], cfg["stride"]
    x = torch.from_numpy(rng.standard_normal((b, c, h, w)).astype(np.float32))
    wt = torch.from_numpy((rng.standard_normal((c, 1, k, k)) / k).astype(np.float32))
    bias = torch.from_numpy(rng.standard_normal(c).astype(np.float32))
    if s == 1:
        pt = pb = pl = pr = k // 2
    else:  # TensorFlow SAME: total = max((ceil(n / s) - 1) * s + k - n, 0), the extra pixel goes to the bottom / right
        th = max((-(-h // s) - 1) * s + k - h, 0)
        tw = max((-(-w // s) - 1) * s + k - w, 0)
        pt, pb, pl, pr = th // 2, th - th // 2, tw // 2, tw - tw // 2
    ref = torch.nn.functional.conv2d(torch.nn.functional.pad(x.double(), (pl, pr, pt, pb)), wt.double(), bias.double(), stride=s,
                                     groups=c)
    ref = ref * torch.sigmoid(ref)
    xd = x.to(DEV).contiguous(memory_format=torch.channels_last)
    assert ops.depthwise_supported(xd, c, (k, k), (s, s))
    y, partial = ops.depthwise_bias_act_pool(xd, wt[:, 0].permute(1, 2, 0).contiguous().to(DEV), bias.to(DEV), 1, s, pt, pl,
                                             ref.shape[-2:])
    assert y.shape == ref.shape and y.is_contiguous(memory_format=torch.channels_last)
    assert float((y.cpu().double() - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))
    assert float((partial.cpu().double().sum(1) - ref.sum((2, 3))).abs().max()) < 1e-4 * max(1.0, float(ref.sum((2, 3)).abs().max()))


@pytest.mark.parametrize("cfg", [dict(shape=(2, 28, 47, 63), cout=32), dict(shape=(1, 76, 32, 40), cout=32, prepad=True),
                                 dict(shape=(2, 4, 33, 34), cout=48, act=0)])
def test_stem_conv3x3s2(cfg):
    """ops.stem_conv3x3s2 (exact-fp32 3x3 / stride-2 stem with bias + SiLU, TensorFlow-SAME padding by bounds or pre-padded input)
    vs float64 torch; odd sizes exercise the asymmetric padding and the single-pixel tail of a row."""
    rng = np.random.default_rng(184)
    b, c, h, w = cfg["shape"]
    cout, act = cfg["cout"], cfg.get("act", 1)
    x = torch.from_numpy(rng.standard_normal((b, c, h, w)).astype(np.float32))
    wt = torch.from_numpy((rng.standard_normal((cout, c, 3, 3)) / np.sqrt(9 * c)).astype(np.float32))
    bias = torch.from_numpy(rng.standard_normal(cout).astype(np.float32))
    th, tw = max((-(-h // 2) - 1) * 2 + 3 - h, 0), max((-(-w // 2) - 1) * 2 + 3 - w, 0)
    pt, pb, pl, pr = th // 2, th - th // 2, tw // 2, tw - tw // 2
    xp = torch.nn.functional.pad(x, (pl, pr, pt, pb))
    ref = torch.nn.functional.conv2d(xp.double(), wt.double(), bias.double(), stride=2)
    if act:
        ref = ref * torch.sigmoid(ref)
    wk = ops.prepare_stem_weight(wt.to(DEV))
    if cfg.get("prepad"):
        xin = xp.to(DEV).contiguous(memory_format=torch.channels_last)
        out = ops.stem_conv3x3s2(xin, wk, bias.to(DEV), act, 0, 0, ref.shape[-2:])
    else:
        xin = x.to(DEV).contiguous(memory_format=torch.channels_last)
        assert ops.stem_conv_supported(xin, c, cout)
        out = ops.stem_conv3x3s2(xin, wk, bias.to(DEV), act, pt, pl, ref.shape[-2:])
    assert out.shape == ref.shape and out.is_contiguous(memory_format=torch.channels_last)
    assert float((out.cpu().double() - ref).abs().max()) < 1e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.gpu
def test_encoder_inference_walk_matches_generic_walk():
    """Encoder._forward_inference (folded stem, squeeze-excite gate inside the projection GEMM, padded expansion output,
    conv_head on the tcgen05 GEMM) yields the same decoder taps as walking the backbone module by module in exact fp32."""
    from mde_biological_vision_systems_b200.models import UnetAdaptiveBins
    torch.manual_seed(3)
    model = UnetAdaptiveBins.build(n_bins=256, min_val=1e-3, max_val=10.0, norm="linear", encoder_name="efficientnet-b1",
                                   semantics_mode=None, instance_segmentation_mode=None).to(DEV).eval()
    g = torch.Generator().manual_seed(5)
    for mod in model.encoder.modules():  # non-trivial BatchNorm statistics so that the folds are exercised
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.copy_(0.1 * torch.randn(mod.num_features, generator=g))
            mod.running_var.copy_(0.5 + torch.rand(mod.num_features, generator=g))
            mod.weight.data.copy_(0.5 + torch.rand(mod.num_features, generator=g))
            mod.bias.data.copy_(0.1 * torch.randn(mod.num_features, generator=g))
    x = torch.rand(2, 3, 96, 128, device=DEV).contiguous(memory_format=torch.channels_last)
    enc = model.encoder
    from mde_biological_vision_systems_b200.models import efficientnet as E
    with ops.exact_fp32_library():
        with torch.no_grad():
            fast = enc(x)
            E.DEPTHWISE_IMPL = "fused"  # the opt-in one-pass depthwise kernel gives the same taps
            try:
                fused = enc(x)
            finally:
                E.DEPTHWISE_IMPL = "cudnn"
        for i in (4, 5, 6, 8, 11):
            assert float((fused[i] - fast[i]).abs().max()) < 2e-5 * float(fast[i].abs().max()), i
        with torch.enable_grad():  # autograd on: every block takes its plain torch path (bn(conv(x)), SqueezeExcite.forward)
            ref = [x]
            for name, child in enc.original_model._modules.items():  # the reference's walk (unet_adaptive_bins.py:118-128)
                for stage in (child._modules.values() if name == "blocks" else (child,)):
                    ref.append(stage(ref[-1]).detach())
    assert len(fast) == len(ref) and fast[1] is None
    for i in (3, 4, 5, 6, 8, 11):
        err = float((fast[i].double() - ref[i].double()).abs().max()) / float(ref[i].abs().max())
        assert err < 2e-5, (i, err)


def test_conv3x3_tf32_form():
    """The single-pass TF32 form of the same kernel (operands pre-rounded by the caller, as its contract says)."""
    rng = np.random.default_rng(94)
    x = ops.round_tf32(torch.from_numpy(rng.standard_normal((2, 128, 24, 40)).astype(np.float32)).to(DEV))
    wt = torch.from_numpy((rng.standard_normal((128, 128, 3, 3)) / 34.0).astype(np.float32)).to(DEV)
    wp = ops.prepare_conv3x3_weight_tf32(wt)
    ref = torch.nn.functional.conv2d(x.cpu().double(), wp.permute(2, 3, 1, 0).cpu().double(), None, padding=1)
    out = ops.conv3x3_nhwc_tf32(x.contiguous(memory_format=torch.channels_last), wp, round_tf32=True)
    err = float((out.cpu().double() - ref).abs().max()) / float(ref.abs().max())
    assert err < 5e-4, err
    assert torch.equal(out, ops.round_tf32(out.contiguous()).contiguous(memory_format=torch.channels_last))


def test_conv3x3_small():
    """Direct fp32 kernel for the noAdaBins decoder's 1-channel conv3 (unet_adaptive_bins.py:78-80)."""
    rng = np.random.default_rng(97)
    for (b, c, h, w, cout) in [(2, 80, 24, 40, 1), (1, 16, 7, 9, 3)]:
        x = torch.from_numpy(rng.standard_normal((b, c, h, w)).astype(np.float32))
        wt = torch.from_numpy((rng.standard_normal((cout, c, 3, 3)) / np.sqrt(9 * c)).astype(np.float32))
        bias = torch.from_numpy(rng.standard_normal(cout).astype(np.float32))
        ref = torch.nn.functional.conv2d(x, wt, bias, padding=1)
        out = ops.conv3x3_small(x.to(DEV).contiguous(memory_format=torch.channels_last), wt.to(DEV), bias.to(DEV))
        assert float((out.cpu() - ref).abs().max()) < 1e-5 * float(ref.abs().max()) + 1e-6


def test_split_bf16_roundtrip():
    """fp32 -> (hi, mid) bf16 planes -> fp32: 16 significant bits, NCHW and channels_last sources give the same pair."""
    rng = np.random.default_rng(98)
    for shape in [(2, 128, 16, 24), (1, 72, 5, 7), (3, 8, 9, 2)]:
        x = torch.from_numpy((rng.standard_normal(shape) * 10 ** rng.uniform(-3, 3, shape)).astype(np.float32)).to(DEV)
        p1 = ops.split_bf16(x)
        p2 = ops.split_bf16(x.contiguous(memory_format=torch.channels_last))
        assert torch.equal(p1.planes, p2.planes)
        hi = x.permute(0, 2, 3, 1).to(torch.bfloat16)
        assert torch.equal(p1.planes[0], hi)
        assert torch.equal(p1.planes[1], (x.permute(0, 2, 3, 1) - hi.float()).to(torch.bfloat16))
        back = p1.float()
        assert back.is_contiguous(memory_format=torch.channels_last)
        assert float(((back - x).abs() / x.abs()).max()) < 2.0 ** -16


@pytest.mark.parametrize("cfg", [
    dict(batch=2, m=128, n=512, k=256, splits=1),     # d feat^T = W'^T gl shape family
    dict(batch=3, m=256, n=128, k=4096, splits=8),    # d W' = gl^T feat (split-K, atomic accumulation)
    dict(batch=1, m=200, n=72, k=100, splits=1),      # ragged M / N / K: TMA zero fill + guarded stores
    dict(batch=2, m=64, n=300, k=40, splits=3),
])
def test_gemm_nt_tc(cfg):
    rng = np.random.default_rng(93)
    a = torch.from_numpy(rng.standard_normal((cfg["batch"], cfg["m"], cfg["k"])).astype(np.float32))
    b = torch.from_numpy(rng.standard_normal((cfg["batch"], cfg["n"], cfg["k"])).astype(np.float32))
    ref = torch.matmul(a.double(), b.double().transpose(1, 2))
    out = ops.gemm_nt(a.to(DEV), b.to(DEV), splits=cfg["splits"], alpha=0.5).cpu().double()
    err = float((out - 0.5 * ref).abs().max()) / float(ref.abs().max())
    assert err < 1e-3, err


def test_regressor_bins(golden):
    m, sd = _head_state()
    tgt = torch.from_numpy(golden["head/tgt"])
    r = m.adaptive_bins_layer.regressor.to(DEV)
    for norm, split in [("linear", False), ("softmax", False), ("sigmoid", False), ("linear", True), ("softmax", True)]:
        wn, edges, centers, y_raw = ops.regressor_bins(tgt[0].to(DEV), r[0].weight, r[0].bias, r[2].weight, r[2].bias,
                                                       r[4].weight, r[4].bias, norm, 1e-3, 10.0, split=split)
        y = oracle.regressor(tgt[0], sd)
        wn_ref = oracle.normalise_widths(y, norm)
        e_ref, c_ref = oracle.bins_from_widths(wn_ref, 1e-3, 10.0)
        assert rel_err(wn.cpu(), wn_ref) < 1e-4
        assert rel_err(edges.cpu(), e_ref) < 1e-4
        assert rel_err(centers.cpu(), c_ref) < 1e-4
    wn, edges, _, _ = ops.regressor_bins(tgt[0].to(DEV), r[0].weight, r[0].bias, r[2].weight, r[2].bias, r[4].weight,
                                         r[4].bias, "linear", 1e-3, 10.0)
    assert rel_err(wn.cpu(), golden["head/widths"]) < 1e-4
    assert rel_err(edges.cpu(), golden["head/edges"]) < 1e-4


@pytest.mark.parametrize("shape", [(2, 256, 48, 64), (1, 256, 13, 17), (2, 100, 31, 20), (1, 7, 5, 3)])
def test_bins_pred_streaming(shape):
    rng = np.random.default_rng(9)
    b, n, h, w = shape
    logits = torch.from_numpy((4 * rng.standard_normal(shape)).astype(np.float32))
    centers = torch.from_numpy(np.sort(rng.random((b, n)).astype(np.float32) * 10, axis=1))
    ref = oracle.softmax_bins_pred(logits, centers)
    out = ops.bins_pred(logits.to(DEV), centers.to(DEV))
    assert rel_err(out.cpu(), ref) < 1e-5


@pytest.mark.parametrize("impl", ["simt", "tc"])
def test_range_attention(impl):
    rng = np.random.default_rng(10)
    b, h, w = 2, 48, 64
    x = torch.from_numpy(rng.standard_normal((b, 128, h, w)).astype(np.float32))
    q = torch.from_numpy(rng.standard_normal((b, 128, 128)).astype(np.float32))
    ref = oracle.pixelwise_dot(x.double(), q.double()).float()
    out = PixelWiseDotProduct(impl=impl)(x.to(DEV), q.to(DEV)).cpu()
    scale = float(ref.abs().max())
    err = float((out - ref).abs().max()) / scale
    assert err < (1e-5 if impl == "simt" else 3e-5), err


def test_range_attention_simt_ragged():
    rng = np.random.default_rng(11)
    x = torch.from_numpy(rng.standard_normal((2, 40, 7, 9)).astype(np.float32))
    q = torch.from_numpy(rng.standard_normal((2, 33, 40)).astype(np.float32))
    ref = oracle.pixelwise_dot(x, q)
    out = ops.range_attention(x.to(DEV), q.to(DEV), impl="simt").cpu()
    np.testing.assert_allclose(out.numpy(), ref.numpy(), rtol=1e-4, atol=1e-4)


def test_conv1x1():
    rng = np.random.default_rng(12)
    ram = torch.from_numpy(rng.standard_normal((2, 128, 24, 40)).astype(np.float32))
    wt = torch.from_numpy(rng.standard_normal((256, 128, 1, 1)).astype(np.float32) * 0.1)
    bias = torch.from_numpy(rng.standard_normal(256).astype(np.float32))
    ref = torch.nn.functional.conv2d(ram, wt, bias)
    out = ops.conv1x1(ram.to(DEV), wt.to(DEV), bias.to(DEV)).cpu()
    np.testing.assert_allclose(out.numpy(), ref.numpy(), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("fused", [True, False])
def test_head_golden(fused, golden):
    """mViT + conv_out + bins on the golden unet_out (logits of magnitude ~10): reference-module outputs within 1e-3
    relative on every pixel, for the fused tcgen05 path (patch embedding, conv3x3 and chain on split-bf16 pairs) and for the
    un-fused exact path."""
    m, _ = _head_state()
    m.to(DEV)
    m.fused_head = fused
    x = synthetic.decoder_features(2, 128, 176, 192, seed=21).to(DEV)
    with torch.no_grad():
        edges, pred = m._head(x)
    assert rel_err(edges.cpu(), golden["head/edges"]) < REL_DEPTH
    mx, p999 = rel_stats(pred.cpu(), golden["head/pred"])
    print("head golden (fused=%s): pred max %.3e p99.9 %.3e" % (fused, mx, p999))
    assert_depth_close(pred.cpu(), golden["head/pred"], tol=2e-4 if fused else REL_DEPTH)


@pytest.mark.parametrize("scale", [1.5, 4.0])
def test_head_large_logits(scale):
    """The fused head against the fp32 oracle when the logits are large (|logit| up to ~40 / ~100: sharply peaked
    softmax, the regime of a trained network) -- the case a single TF32 pass misses by 1e-2 (scripts/precision_study_head.py)."""
    m, sd = _head_state()
    x = synthetic.decoder_features(1, 128, 176, 192, seed=23, scale=scale)  # 132 tokens >= 1 + 128 queries
    with torch.no_grad():
        e_ref, p_ref = oracle.head(x.double(), {k: v.double() for k, v in sd.items() if v.is_floating_point()}, 1e-3, 10.0)
        m.to(DEV)
        edges, pred = m._head(x.to(DEV))
    assert rel_err(edges.cpu(), e_ref) < REL_DEPTH
    assert_depth_close(pred.cpu(), p_ref)


@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_head_chain_backward(layout):
    """Hand-written backward of the fused chain (backward-epilogue chain + two tcgen05 GEMMs) vs float64 autograd through
    the reference formulation: PixelWiseDotProduct -> conv_out -> Softmax(dim=1) -> sum(out * centres)."""
    rng = np.random.default_rng(95)
    b, k, h, w, nb = 2, 128, 48, 64, 256
    feat = torch.from_numpy(rng.standard_normal((b, k, h, w)).astype(np.float32))
    q = torch.from_numpy((rng.standard_normal((b, 128, k)) * 0.3).astype(np.float32))
    w_out = torch.from_numpy((rng.standard_normal((nb, 128, 1, 1)) * 0.2).astype(np.float32))
    b_out = torch.from_numpy(rng.standard_normal(nb).astype(np.float32))
    widths = torch.from_numpy(rng.random((b, nb)).astype(np.float32) + 0.1)
    centers = torch.cumsum(widths / widths.sum(1, keepdim=True) * 10, dim=1)
    g = torch.from_numpy(rng.standard_normal((b, 1, h, w)).astype(np.float32))
    ref_in = [t.double().requires_grad_(True) for t in (feat, q, w_out, b_out, centers)]
    ram = oracle.pixelwise_dot(ref_in[0], ref_in[1])
    sm = torch.softmax(torch.nn.functional.conv2d(ram, ref_in[2], ref_in[3]), dim=1)
    pred_ref = (sm * ref_in[4].view(b, nb, 1, 1)).sum(1, keepdim=True)
    (pred_ref * g.double()).sum().backward()
    dev_in = [t.to(DEV).requires_grad_(True) for t in (feat, q, w_out, b_out, centers)]
    x = dev_in[0]
    if layout == "nhwc":
        x = x.contiguous(memory_format=torch.channels_last)
    pred = ops.head_chain_autograd(x, dev_in[1], dev_in[2], dev_in[3], dev_in[4])
    # random operands give logits of std ~8 here (the trained-shape golden case is ~3)
    assert rel_err(pred.detach().cpu(), pred_ref.detach().float()) < REL_DEPTH
    (pred * g.to(DEV)).sum().backward()
    for name, a, r in zip(("feat", "queries", "w_out", "b_out", "centers"), dev_in, ref_in):
        ga, gr = a.grad.cpu().double(), r.grad
        err = float((ga - gr).abs().max()) / float(gr.abs().max())
        assert err < 5e-3, (name, err)


def test_mvit_forward_surface(golden):
    m, _ = _head_state()
    m.to(DEV)
    x = synthetic.decoder_features(2, 128, 176, 192, seed=21).to(DEV)
    with torch.no_grad():
        widths, ram = m.adaptive_bins_layer(x)
    assert rel_err(widths.cpu(), golden["head/widths"]) < REL_DEPTH
    sub = ram[:, :, ::16, ::16].cpu().numpy()
    scale = np.abs(golden["head/ram_sub"]).max()
    assert np.abs(sub - golden["head/ram_sub"]).max() / scale < 1e-4


def test_full_model_golden(golden):
    """Whole UnetAdaptiveBins (backbone + decoder passthrough, head on our kernels) vs the reference model."""
    m = make_model(insertion_point="input", semantics_mode=None, instance_segmentation_mode=None).to(DEV)
    x = synthetic.image(1, 352, 384, seed=31).to(DEV)
    with torch.no_grad():
        edges, pred = m(x)
    assert rel_err(edges.cpu(), golden["full/edges"]) < REL_DEPTH
    assert_depth_close(pred.cpu(), golden["full/pred"])
    m.fused_head = False
    with torch.no_grad():
        _, pred = m(x)
    assert_depth_close(pred.cpu(), golden["full/pred"])


@pytest.mark.parametrize("shape", [((2, 5, 15, 19), (2, 3, 26, 34)), ((1, 4, 13, 17), (1, 2, 27, 35)), ((2, 3, 8, 8), (2, 1, 8, 8))])
def test_upsample_concat_forward_backward(shape):
    """DecoderBN up-sampling step vs F.interpolate(bilinear, align_corners=True) + cat, values and gradients."""
    rng = np.random.default_rng(31)
    xs, ss = shape
    x = torch.from_numpy(rng.standard_normal(xs).astype(np.float32)).requires_grad_(True)
    skip = torch.from_numpy(rng.standard_normal(ss).astype(np.float32)).requires_grad_(True)
    ref = torch.cat((torch.nn.functional.interpolate(x, size=ss[-2:], mode="bilinear", align_corners=True), skip), 1)
    g = torch.from_numpy(rng.standard_normal(ref.shape).astype(np.float32))
    ref.backward(g)
    xd = x.detach().to(DEV).requires_grad_(True)
    sd = skip.detach().to(DEV).requires_grad_(True)
    out = ops.upsample_concat(xd, sd)
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().numpy(), rtol=1e-5, atol=1e-5)
    out.backward(g.to(DEV))
    np.testing.assert_allclose(xd.grad.cpu().numpy(), x.grad.numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(sd.grad.cpu().numpy(), skip.grad.numpy(), rtol=0, atol=0)


@pytest.mark.parametrize("skip_cl", [False, True])
def test_upsample_concat_nhwc(skip_cl):
    """channels_last resize + concat (feeds the tcgen05 decoder convs) vs F.interpolate(align_corners=True) + cat."""
    rng = np.random.default_rng(35)
    for xs, ss in [((2, 8, 15, 19), (2, 12, 26, 34)), ((1, 64, 13, 17), (1, 4, 27, 35)), ((2, 4, 8, 8), (2, 8, 8, 8))]:
        x = torch.from_numpy(rng.standard_normal(xs).astype(np.float32))
        skip = torch.from_numpy(rng.standard_normal(ss).astype(np.float32))
        ref = torch.cat((torch.nn.functional.interpolate(x, size=ss[-2:], mode="bilinear", align_corners=True), skip), 1)
        sd = skip.to(DEV)
        if skip_cl:
            sd = sd.contiguous(memory_format=torch.channels_last)
        out = ops.upsample_concat_nhwc(x.to(DEV).contiguous(memory_format=torch.channels_last), sd)
        assert out.is_contiguous(memory_format=torch.channels_last)
        np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), rtol=1e-5, atol=1e-5)


def test_upsample_concat_nhwc_backward():
    rng = np.random.default_rng(36)
    for xs, ss in [((2, 8, 15, 19), (2, 12, 26, 34)), ((1, 16, 13, 17), (1, 4, 26, 34)), ((2, 4, 8, 8), (2, 8, 8, 8)),
                   ((1, 4, 5, 7), (1, 4, 15, 21))]:
        x = torch.from_numpy(rng.standard_normal(xs).astype(np.float32)).requires_grad_(True)
        skip = torch.from_numpy(rng.standard_normal(ss).astype(np.float32)).requires_grad_(True)
        ref = torch.cat((torch.nn.functional.interpolate(x, size=ss[-2:], mode="bilinear", align_corners=True), skip), 1)
        g = torch.from_numpy(rng.standard_normal(ref.shape).astype(np.float32))
        ref.backward(g)
        xd = x.detach().to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        sd = skip.detach().to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        out = ops.upsample_concat_nhwc(xd, sd)
        np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().numpy(), rtol=1e-5, atol=1e-5)
        out.backward(g.to(DEV))
        np.testing.assert_allclose(xd.grad.cpu().numpy(), x.grad.numpy(), rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(sd.grad.cpu().numpy(), skip.grad.numpy(), rtol=0, atol=0)


def test_channels_last_model_matches_nchw():
    """UnetAdaptiveBins.channels_last_() (what build() returns) only changes strides: same outputs, same gradients."""
    kw = dict(insertion_point="input", semantics_mode=None, instance_segmentation_mode=None)
    m1, m2 = make_model(**kw).to(DEV), make_model(**kw).to(DEV).channels_last_()
    x = synthetic.image(2, 352, 384, seed=38).to(DEV)
    with torch.no_grad():
        e1, p1 = m1(x)
        e2, p2 = m2(x)
    assert rel_err(e2.cpu(), e1.cpu()) < 1e-4
    assert_depth_close(p2.cpu(), p1.cpu())
    # one training-mode forward/backward in each layout (stock BatchNorm with batch statistics, our NHWC / NCHW resize+concat)
    depth = synthetic.depth(2, 352, 384, seed=39).to(DEV)
    grads = []
    for m in (m1, m2):
        m.train()
        for mod in m.modules():  # no randomness: nn.Dropout modules and the attention-weight dropout of nn.MultiheadAttention
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
            if isinstance(mod, torch.nn.MultiheadAttention):
                mod.dropout = 0.0
        torch.manual_seed(0)
        m.zero_grad(set_to_none=True)
        with ops.exact_fp32_library():  # compare the two layouts, not cuDNN's TF32 kernel choices per layout
            e, p = m(x)
            loss = SILogLoss()(p, depth, mask=depth > 1e-3) + 0.1 * BinsChamferLoss()(e, depth)
            loss.backward()
        grads.append((float(loss.detach()), m.decoder.up4._net[0].weight.grad.detach().cpu().clone(),
                      m.encoder.original_model.conv_stem.weight.grad.detach().cpu().clone()))
    assert abs(grads[0][0] - grads[1][0]) <= 1e-3 * abs(grads[0][0])
    # train-mode BatchNorm at batch 2 makes these gradients ill-conditioned: against an fp64 evaluation BOTH layouts sit at
    # 0.2-2 % of the largest entry (scripts/debug_cl.py, measured on B200), so the two fp32 paths are compared at 5 %
    for a, b in zip(grads[0][1:], grads[1][1:]):
        assert float((a - b).abs().max()) <= 5e-2 * float(a.abs().max()) + 1e-7


def test_decoder_tc_vs_oracle():
    """(f)1 DecoderBN on our kernels (channels_last, resize + concat written as split-bf16 pairs, tcgen05 conv3x3 with
    BatchNorm(eval) + LeakyReLU folded into the epilogue) vs the CPU oracle's decoder_bn (models/unet_adaptive_bins.py:39-100)
    on the same encoder features; the stock torch modules (cuDNN) are checked alongside."""
    m = make_model(insertion_point="input", semantics_mode=None, instance_segmentation_mode=None)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = synthetic.image(2, 160, 192, seed=37)
    with torch.no_grad():
        feats_cpu = oracle.encoder_features(m.encoder.original_model, x)
        ref = oracle.decoder_bn(feats_cpu, sd)
        m.to(DEV).channels_last_()
        feats = [f.to(DEV).contiguous(memory_format=torch.channels_last) for f in feats_cpu]
        out = m.decoder(feats)
        assert isinstance(out, ops.SplitBF16), "the tcgen05 decoder path did not run"
        out = out.float()
        m.decoder.conv_impl = "cudnn"
        with ops.exact_fp32_library():
            stock = m.decoder(feats)
    assert out.shape == ref.shape
    err = float((out.cpu() - ref).abs().max()) / float(ref.abs().max())
    err_stock = float((stock.cpu() - ref).abs().max()) / float(ref.abs().max())
    print("decoder vs oracle: ours %.3e  stock cuDNN fp32 %.3e" % (err, err_stock))
    assert err < 5e-5, err


def test_upsample_concat_nhwc_pair():
    """The pair-writing form of the resize + concat step equals split_bf16 of the fp32 form, bit for bit."""
    rng = np.random.default_rng(40)
    for xs, ss in [((2, 8, 15, 19), (2, 16, 26, 34)), ((1, 64, 13, 17), (1, 8, 27, 35))]:
        x = torch.from_numpy(rng.standard_normal(xs).astype(np.float32)).to(DEV).contiguous(memory_format=torch.channels_last)
        skip = torch.from_numpy(rng.standard_normal(ss).astype(np.float32)).to(DEV).contiguous(memory_format=torch.channels_last)
        ref = ops.split_bf16(ops.upsample_concat_nhwc(x, skip))
        out = ops.upsample_concat_nhwc_pair(x, skip)
        assert torch.equal(out.planes, ref.planes)
        padded = ops.upsample_concat_nhwc_pair(x, skip, pad_to=32)  # zero channels up to the next multiple of 32
        c = ref.planes.shape[-1]
        assert padded.planes.shape[-1] == -(-c // 32) * 32
        assert torch.equal(padded.planes[..., :c], ref.planes) and float(padded.planes[..., c:].float().abs().max()) == 0.0


def test_nchw_to_nhwc():
    rng = np.random.default_rng(33)
    for shape in [(2, 128, 16, 24), (1, 70, 5, 7), (3, 3, 9, 2)]:
        x = torch.from_numpy(rng.standard_normal(shape).astype(np.float32)).to(DEV)
        y = ops.to_channels_last(x)
        assert y.is_contiguous(memory_format=torch.channels_last) and torch.equal(y, x)


def test_concat_channels_last():
    rng = np.random.default_rng(34)
    a = torch.from_numpy(rng.standard_normal((2, 3, 17, 23)).astype(np.float32)).to(DEV)
    b = torch.from_numpy(rng.standard_normal((2, 25, 17, 23)).astype(np.float32)).to(DEV)
    c = torch.from_numpy(rng.standard_normal((2, 70, 17, 23)).astype(np.float32)).to(DEV)
    out = ops.concat_channels_last([a, b, c])
    assert out.is_contiguous(memory_format=torch.channels_last) and torch.equal(out, torch.cat((a, b, c), 1))


def test_noadabins_epilogue():
    x = torch.from_numpy(np.random.default_rng(3).standard_normal((2, 1, 24, 32)).astype(np.float32))
    assert np.array_equal(ops.relu_eps(x.to(DEV)).cpu().numpy(), oracle.noadabins_epilogue(x).numpy())


# ------------------------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------------------------
def test_losses_golden(golden):
    silog, chamfer = SILogLoss(), BinsChamferLoss()
    for name, (b, h, w) in {"a": (2, 104, 136), "b": (3, 64, 96)}.items():
        depth = synthetic.depth(b, h, w, seed=52)
        pred = torch.from_numpy(golden[f"loss/{name}/pred"])
        edges = torch.from_numpy(golden[f"loss/{name}/edges"])
        d = depth.to(DEV)
        s = silog(pred.to(DEV), d, mask=(d > 1e-3), interpolate=True)
        assert rel_err(s.cpu(), golden[f"loss/{name}/silog"]) < REL_LOSS
        up = torch.nn.functional.interpolate(pred, depth.shape[-2:], mode="nearest")
        s2 = silog(up.to(DEV), d.clamp_min(0.2), mask=None, interpolate=False)
        assert rel_err(s2.cpu(), golden[f"loss/{name}/silog_nomask_noint"]) < REL_LOSS
        c = chamfer(edges.to(DEV), d)
        assert rel_err(c.cpu(), golden[f"loss/{name}/chamfer"]) < REL_LOSS


def test_losses_edge_cases(golden):
    depth = synthetic.depth(2, 64, 96, seed=53, all_valid=True)
    depth[1, :, 1:, :] = 0.0
    depth[1, :, 0, 5:] = 0.0
    pred = torch.from_numpy(golden["loss/edge/pred"]).to(DEV)
    edges = torch.linspace(1e-3, 10, 257).repeat(2, 1).contiguous().to(DEV)
    d = depth.to(DEV)
    assert rel_err(SILogLoss()(pred, d, mask=d > 1e-3).cpu(), golden["loss/edge/silog"]) < REL_LOSS
    assert rel_err(BinsChamferLoss()(edges, d).cpu(), golden["loss/edge/chamfer"]) < REL_LOSS
    d[0] = 0.0  # image without a single valid target -> NaN like the reference
    assert torch.isnan(BinsChamferLoss()(edges, d))


def test_silog_backward():
    rng = np.random.default_rng(21)
    depth = synthetic.depth(2, 48, 64, seed=22)
    pred = torch.from_numpy((0.5 + 5 * rng.random((2, 1, 24, 32))).astype(np.float32)).requires_grad_(True)
    oracle.silog(pred, depth, mask=depth > 1e-3, interpolate=True).backward()
    pd = pred.detach().to(DEV).requires_grad_(True)
    SILogLoss()(pd, depth.to(DEV), mask=(depth > 1e-3).to(DEV), interpolate=True).backward()
    g_ref = pred.grad.numpy()
    np.testing.assert_allclose(pd.grad.cpu().numpy(), g_ref, rtol=2e-3, atol=2e-3 * np.abs(g_ref).max())


def test_chamfer_backward():
    rng = np.random.default_rng(23)
    depth = synthetic.depth(2, 48, 64, seed=24)
    widths = torch.from_numpy(rng.random((2, 256), dtype=np.float32) + 0.1)
    widths = widths / widths.sum(1, keepdim=True) * 9.999
    edges = torch.cumsum(torch.nn.functional.pad(widths, (1, 0), value=1e-3), dim=1)
    e_ref = edges.clone().requires_grad_(True)
    # autograd through the brute-force oracle (min picks the same neighbours)
    oracle.bins_chamfer(e_ref, depth).backward()
    e_dev = edges.to(DEV).requires_grad_(True)
    BinsChamferLoss()(e_dev, depth.to(DEV)).backward()
    g_ref = e_ref.grad.numpy()
    np.testing.assert_allclose(e_dev.grad.cpu().numpy(), g_ref, rtol=2e-3, atol=2e-3 * np.abs(g_ref).max())


def test_depth_losses_fused_matches_oracle_and_separate():
    """The fused SILog + chamfer kernel (one pass over the depth map, masks derived in registers) against the oracle and the
    two drop-in modules, forward and backward, incl. ragged sizes (HW % 4 != 0) and an image with very few valid pixels."""
    rng = np.random.default_rng(120)
    for (b, h, w, H, W) in [(3, 104, 136, 208, 272), (2, 13, 17, 27, 35), (1, 8, 8, 8, 8)]:
        depth = synthetic.depth(b, H, W, seed=121 + H)
        if b > 1:
            depth[1, :, 2:, :] = 0.0  # image 1: only two rows of valid pixels
        pred = torch.from_numpy((0.3 + 9 * rng.random((b, 1, h, w), dtype=np.float32)))
        widths = torch.from_numpy(rng.random((b, 256)).astype(np.float32) + 0.05)
        edges = torch.cat((torch.full((b, 1), 1e-3), 1e-3 + torch.cumsum(widths / widths.sum(1, keepdim=True) * 9.999, 1)), 1)
        interp = (h, w) != (H, W)
        p_ref, e_ref = pred.clone().requires_grad_(True), edges.clone().requires_grad_(True)
        s_ref = oracle.silog(p_ref, depth, mask=depth > 1e-3, interpolate=interp)
        c_ref = oracle.bins_chamfer(e_ref, depth)
        (s_ref + 0.1 * c_ref).backward()
        pd, ed, dd = pred.to(DEV).requires_grad_(True), edges.to(DEV).requires_grad_(True), depth.to(DEV)
        s, c = DepthLosses(1e-3)(pd, ed, dd, interpolate=interp)
        assert abs(float(s) - float(s_ref)) <= REL_LOSS * abs(float(s_ref)), (float(s), float(s_ref))
        assert abs(float(c) - float(c_ref)) <= REL_LOSS * abs(float(c_ref)), (float(c), float(c_ref))
        (s + 0.1 * c).backward()
        assert float((pd.grad.cpu() - p_ref.grad).abs().max()) <= 1e-3 * float(p_ref.grad.abs().max()) + 1e-9
        assert float((ed.grad.cpu() - e_ref.grad).abs().max()) <= 1e-3 * float(e_ref.grad.abs().max()) + 1e-9
        # the separate drop-in modules give the same numbers (same kernel, other template flags); bit-reproducible run to run
        s2 = SILogLoss()(pred.to(DEV), dd, mask=dd > 1e-3, interpolate=interp)
        c2 = BinsChamferLoss()(edges.to(DEV), dd)
        assert abs(float(s2) - float(s)) <= 1e-6 * abs(float(s)) and abs(float(c2) - float(c)) <= 1e-6 * abs(float(c))
        s3, c3 = DepthLosses(1e-3)(pred.to(DEV), edges.to(DEV), dd, interpolate=interp)
        assert float(s3) == float(s) and float(c3) == float(c)


def test_chamfer_unsorted_centres_is_nan_not_wrong():
    """The kernel's nearest-neighbour search relies on ascending centres (always true for the model's edges); an unsorted
    vector is detected on the device and reported as NaN instead of a silently wrong loss."""
    depth = synthetic.depth(2, 32, 48, seed=130).to(DEV)
    edges = torch.linspace(1e-3, 10, 257).repeat(2, 1)
    assert torch.isfinite(BinsChamferLoss()(edges.to(DEV), depth))
    edges[1, 100], edges[1, 140] = edges[1, 140].clone(), edges[1, 100].clone()
    assert torch.isnan(BinsChamferLoss()(edges.to(DEV), depth))


# ------------------------------------------------------------------------------------------------------------
# full-size properties (config 2: B = 16, 416 x 544, n_bins = 256)
# ------------------------------------------------------------------------------------------------------------
def test_full_size_properties():
    b, H, W, h, w = 16, 416, 544, 208, 272
    depth = synthetic.depth(b, H, W, seed=61).to(DEV)
    rng = np.random.default_rng(62)
    pred = torch.from_numpy((0.3 + 9 * rng.random((b, 1, h, w), dtype=np.float32))).to(DEV)
    mask = depth > 1e-3
    silog, chamfer = SILogLoss(), BinsChamferLoss()
    s1 = float(silog(pred, depth, mask=mask))
    # (1) scale both by k: log-difference unchanged -> identical loss
    s2 = float(silog(pred * 3.0, depth * 3.0, mask=mask))
    assert abs(s1 - s2) <= 1e-5 * abs(s1)
    # (2) against torch on the same device (library reference for the full size)
    up = torch.nn.functional.interpolate(pred, (H, W), mode="bilinear", align_corners=True)
    g = torch.log(up[mask]) - torch.log(depth[mask])
    ref = float(10 * torch.sqrt(torch.var(g.double()) + 0.15 * g.double().mean() ** 2))
    assert abs(s1 - ref) <= REL_LOSS * abs(ref)
    # (3) chamfer: shuffling the pixels of each image does not change the loss; scaling by k scales it by k^2
    edges = torch.linspace(1e-3, 10, 257, device=DEV).repeat(b, 1).contiguous()
    c1 = float(chamfer(edges, depth))
    perm = torch.randperm(H * W, device=DEV)
    c2 = float(chamfer(edges, depth.flatten(1)[:, perm].reshape(b, 1, H, W).contiguous()))
    assert abs(c1 - c2) <= 1e-5 * abs(c1)
    # uniform bins of width d, targets uniform: per-image independent check on image 0 against the oracle
    c0 = float(chamfer(edges[:1].contiguous(), depth[:1].contiguous()))
    r0 = float(oracle.bins_chamfer(edges[:1].cpu(), depth[:1].cpu()))
    assert abs(c0 - r0) <= REL_LOSS * abs(r0)
    # (4) fused head == un-fused three-kernel path on a full-size feature map (both ours; different roundings)
    m = make_model(insertion_point="input", semantics_mode=None, instance_segmentation_mode=None).to(DEV)
    x = synthetic.decoder_features(4, 128, h, w, seed=63).to(DEV)
    with torch.no_grad():
        m.fused_head = True
        e1, p1 = m._head(x)
        m.fused_head = False
        e2, p2 = m._head(x)
    assert rel_err(e1.cpu(), e2.cpu()) < 1e-5  # same patch embedding / transformer / regressor kernels on both paths
    assert_depth_close(p1.cpu(), p2.cpu())
    assert float(p1.min()) >= 1e-3 and float(p1.max()) <= 10.0  # a convex combination of the bin centres


# ------------------------------------------------------------------------------------------------------------
# "next" rows (f)2-(f)4
# ------------------------------------------------------------------------------------------------------------
def test_eval_metrics_golden():
    """Fused evaluation epilogue + metrics vs the reference's utils.compute_errors (tests/golden/make_golden_eval.py)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden_eval as mg
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_eval.npz"))
    from mde_biological_vision_systems_b200 import evaluation
    for name, (b, h, w, H, W, lo, hi, garg, eigen, ds) in mg.CASES.items():
        pred, gt = mg.case_inputs(name)
        box = evaluation.crop_box(H, W, garg, eigen, ds)
        assert box == oracle.eval_crop_box(H, W, garg, eigen, ds)
        out = ops.eval_metrics(pred.to(DEV), gt.to(DEV), lo, hi, box).cpu().numpy().astype(np.float64)
        ref = gold[name]
        assert np.array_equal(out[:, 9], ref[:, 9]), name          # valid-pixel counts: exact
        np.testing.assert_allclose(out[:, :3], ref[:, :3], rtol=0, atol=1.5 / ref[:, 9].min())  # threshold counts
        np.testing.assert_allclose(out[:, 3:9], ref[:, 3:9], rtol=1e-4, err_msg=name)
    # non-finite predictions: reference = ATen's CUDA bilinear kernel (where the reference runs it) + the numpy epilogue
    pred, gt = mg.case_inputs("nyu_eigen")
    pred[0, 0, 100, 200] = float("nan")
    pred[1, 0, 50, 60] = float("inf")
    pred[1, 0, 70, 90] = float("-inf")
    up = torch.nn.functional.interpolate(pred.to(DEV), gt.shape[-2:], mode="bilinear", align_corners=True).cpu()
    out = ops.eval_metrics(pred.to(DEV), gt.to(DEV), 1e-3, 10.0, (45, 471, 41, 601)).cpu().numpy().astype(np.float64)
    for i in range(2):
        g, p = oracle.eval_mask_and_clip(up[i:i + 1], gt[i:i + 1], 1e-3, 10.0, False, True, "nyu")
        m = oracle.compute_errors(g, p)
        np.testing.assert_allclose(out[i, :9], [m[k] for k in mg.KEYS], rtol=2e-4, atol=2e-5)
        assert out[i, 9] == g.size
    # an image without any valid pixel -> NaN like numpy's mean of an empty array
    gt = torch.zeros(1, 1, 8, 8, device=DEV)
    assert torch.isnan(ops.eval_metrics(torch.ones(1, 1, 4, 4, device=DEV), gt, 1e-3, 10.0)[0, 0])


def test_flip_average():
    rng = np.random.default_rng(71)
    a = torch.from_numpy((12 * rng.random((2, 1, 9, 14)) - 1).astype(np.float32))
    b = torch.from_numpy((12 * rng.random((2, 1, 9, 14)) - 1).astype(np.float32))
    ref = 0.5 * (np.clip(a.numpy(), 1e-3, 10) + np.clip(b.numpy()[..., ::-1], 1e-3, 10))
    assert np.array_equal(ops.flip_average(a.to(DEV), b.to(DEV), 1e-3, 10).cpu().numpy(), ref)


@pytest.mark.parametrize("dtype", [torch.int32, torch.uint8])
def test_gather_compact_label_formats(dtype, golden_digests):
    """(f)3: int32 / uint8 label maps (the on-disk formats) gather to the same bits as the int64 batch tensors."""
    mode = "glove-25d-ade20k-places"
    lab, _ = sem_labels(mode)
    wire = lab.to(torch.uint8) if dtype == torch.uint8 else lab.to(torch.int32)  # -1 -> 255 like astype(np.ubyte)
    raw, sem = SemanticsLoader(Args(use_semantics=mode)).get_semantics({"semantics": wire})
    ref_raw, ref = oracle.semantics_loader(mode, wire.long().numpy(), load_table("ade20k_places_classes_glove_twitter_27b_25d_embeddings.npy"))
    assert raw.dtype == torch.int64 and np.array_equal(raw.cpu().numpy(), ref_raw)
    assert np.array_equal(sem.cpu().numpy(), ref)
    if dtype == torch.int32:  # int32 keeps the sign, so the result is the golden int64 one
        assert digest(sem.cpu().numpy()) == golden_digests[f"sem/{mode}/out"]
        ilab, iar = inst_labels("ade20k_swin")
        out = InstanceSegmentationLoader(Args(use_instance_segmentation="ade20k_swin")).get_instance_segmentation(
            {"instance_labels": ilab.to(torch.int32), "instance_areas": iar.to(torch.int32)})
        assert digest(out[0].cpu().numpy()) == golden_digests["inst/ade20k_swin/raw"]
        assert digest(out[1].cpu().numpy()) == golden_digests["inst/ade20k_swin/emb"]
        assert digest(out[2].cpu().numpy()) == golden_digests["inst/ade20k_swin/areas"]


def test_one_hot_semantics():
    """(f)4: one-hot-ade20k-places = the K3 gather with a 101 x 101 identity table (no reference implementation)."""
    mode = "one-hot-ade20k-places"
    lab, _ = sem_labels(mode)
    raw, sem = SemanticsLoader(Args(use_semantics=mode)).get_semantics({"semantics": lab.clone()})
    assert sem.shape == (lab.shape[0], 101, lab.shape[2], lab.shape[3]) and sem.dtype == torch.float32
    clamped = lab.clone()
    clamped[(clamped > 100) | (clamped < 0)] = 100
    ref = torch.nn.functional.one_hot(clamped[:, 0], 101).permute(0, 3, 1, 2).float()
    assert torch.equal(sem.cpu(), ref) and torch.equal(raw.cpu(), clamped)
    from mde_biological_vision_systems_b200.models import UnetAdaptiveBins as U
    assert U.get_num_channels_to_add("efficientnet-b1", mode, None, "rgb") == 101


# ------------------------------------------------------------------------------------------------------------
# BASELINE configs 4 and 5 as parity cases
# ------------------------------------------------------------------------------------------------------------
def test_config4_b5_head_vs_oracle():
    """EfficientNet-B5 AdaBins (original, before-attn without extras): decoder width 2048, head on our kernels."""
    m = make_model(encoder_name="efficientnet-b5", insertion_point="before-attn", semantics_mode=None,
                   instance_segmentation_mode=None).to(DEV)
    x = synthetic.image(1, 352, 384, seed=81).to(DEV)
    with torch.no_grad():
        with ops.exact_fp32_library():
            unet_out = _as_f32(m.decoder(m.encoder(x)))
        edges, pred = m(x)
    assert unet_out.shape == (1, 128, 176, 192) and pred.shape == (1, 1, 176, 192) and edges.shape == (1, 257)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    e_ref, p_ref = oracle.head(unet_out.cpu().contiguous(), sd, 1e-3, 10.0)
    assert rel_err(edges.cpu(), e_ref) < REL_DEPTH
    assert_depth_close(pred.cpu(), p_ref)


def test_config5_noadabins_480x640():
    """noAdaBins EfficientNet-B1 at 480x640: (None, relu(unet_out) + 1e-4), one channel at half resolution."""
    m = make_model(encoder_name="efficientnet-b1-noAdaBins", insertion_point="input", semantics_mode=None,
                   instance_segmentation_mode=None).to(DEV)
    x = synthetic.image(2, 480, 640, seed=82).to(DEV)
    with torch.no_grad():
        with ops.exact_fp32_library():
            feats = m.encoder(x)
            unet_out = m.decoder(feats)
        edges, pred = m(x)
    assert edges is None and pred.shape == (2, 1, 240, 320)
    assert np.array_equal(pred.cpu().numpy(), oracle.noadabins_epilogue(unet_out.cpu()).numpy())
    # the decoder itself (tcgen05 convs + the direct 1-channel conv3) against the CPU oracle on the same encoder features
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    ref = oracle.decoder_bn([None if f is None else f.cpu().contiguous() for f in feats], sd)  # None: entries no consumer reads
    assert float((unet_out.cpu() - ref).abs().max()) < 5e-5 * float(ref.abs().max())


# ------------------------------------------------------------------------------------------------------------
# section 8(e): SyncBatchNorm kernels (world size 1 == plain training-mode batch norm, the all-reduce is the identity)
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(4, 16, 33, 47), (2, 96, 20, 24), (3, 1280, 5, 7), (2, 40, 64, 80)])
def test_sync_batchnorm_kernels_vs_torch(shape):
    from mde_biological_vision_systems_b200.parallel import SyncBatchNorm2d
    rng = np.random.default_rng(97)
    b, c, h, w = shape
    x = torch.from_numpy((rng.standard_normal(shape) * 2 + 3).astype(np.float32))
    g = torch.from_numpy(rng.standard_normal(shape).astype(np.float32))
    ref = torch.nn.BatchNorm2d(c).double()
    with torch.no_grad():
        ref.weight.copy_(torch.from_numpy(1 + 0.3 * rng.standard_normal(c)))
        ref.bias.copy_(torch.from_numpy(0.2 * rng.standard_normal(c)))
    ours = SyncBatchNorm2d(c).to(DEV)
    ours.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    ours.force_kernels = True
    ref.train(), ours.train()
    xr = x.double().requires_grad_(True)
    yr = ref(xr)
    yr.backward(g.double())
    xd = x.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    yd = ours(xd)
    yd.backward(g.to(DEV))
    np.testing.assert_allclose(yd.detach().cpu().numpy(), yr.detach().numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(xd.grad.cpu().numpy(), xr.grad.numpy(), rtol=1e-3, atol=1e-4 * float(xr.grad.abs().max()) + 1e-6)
    np.testing.assert_allclose(ours.weight.grad.cpu().numpy(), ref.weight.grad.numpy(), rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(ours.bias.grad.cpu().numpy(), ref.bias.grad.numpy(), rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(ours.running_mean.cpu().numpy(), ref.running_mean.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ours.running_var.cpu().numpy(), ref.running_var.numpy(), rtol=1e-4, atol=1e-6)
    assert int(ours.num_batches_tracked) == 1


def test_autocast_bf16_training_step():
    """BASELINE config 4 runs DDP training under bf16 autocast: the stock torch bodies then produce bf16 activations, the
    hand-written kernels stay fp32 (inputs are widened at the operator boundary).  north_star tolerance for bf16: 2e-2."""
    kw = dict(insertion_point="input", semantics_mode=None, instance_segmentation_mode=None)
    m = make_model(**kw).to(DEV).channels_last_()
    x = synthetic.image(2, 352, 384, seed=83).to(DEV)
    depth = synthetic.depth(2, 352, 384, seed=84).to(DEV)
    with torch.no_grad():
        e32, p32 = m(x)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            e16, p16 = m(x)
    assert p16.dtype == torch.float32 and e16.dtype == torch.float32
    rel = ((p16 - p32).abs() / p32.abs()).flatten()
    assert float(rel.mean()) < 2e-2 and float(rel.kthvalue(int(rel.numel() * 0.99)).values) < 1e-1, (float(rel.mean()), float(rel.max()))
    assert rel_err(e16.cpu(), e32.cpu()) < 5e-2
    m.train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        e, p = m(x)
        loss = SILogLoss()(p, depth, mask=depth > 1e-3) + 0.1 * BinsChamferLoss()(e, depth)
    loss.backward()
    assert torch.isfinite(loss)
    g = m.decoder.up4._net[0].weight.grad
    assert g is not None and torch.isfinite(g).all() and float(g.abs().max()) > 0


def test_before_attn_insertion_vs_oracle():
    """A3': external info inserted before the attention head (nearest down-sampling of the embedding planes, concatenated
    onto the decoder output, unet_adaptive_bins.py:244-282): the head then sees 128 + 25 channels."""
    mode = "glove-25d-ade20k-places"
    m = make_model(insertion_point="before-attn", semantics_mode=mode, instance_segmentation_mode=None).to(DEV)
    assert m.num_decoded_channels == 153
    x = synthetic.image(1, 352, 384, seed=85).to(DEV)
    lab, _ = sem_labels(mode, 1, 352, 384, seed=86, n_rect=(20, 40))
    _, sem = SemanticsLoader(Args(use_semantics=mode)).get_semantics({"semantics": lab})
    with torch.no_grad():
        edges, pred = m(x, semantics=sem)
        with ops.exact_fp32_library():
            unet_out = _as_f32(m.decoder(m.encoder(x)))
        small = torch.nn.functional.interpolate(sem, size=unet_out.shape[-2:], mode="nearest").float()
        cat = torch.cat((unet_out, small), dim=1)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    e_ref, p_ref = oracle.head(cat.cpu().contiguous(), sd, 1e-3, 10.0)
    assert pred.shape == (1, 1, 176, 192)
    assert rel_err(edges.cpu(), e_ref) < REL_DEPTH
    assert_depth_close(pred.cpu(), p_ref)


def test_config3_full_path_vs_oracle_and_training_step():
    """BASELINE config 3 end to end: GloVe-25d semantics + ADE20K-Swin instance embeddings / areas / human sizes at the input
    (73 input channels, two trainable 1x1-conv MLPs), channels_last model.  Inference vs the oracle restatement of the
    whole reference path (loaders -> insertion -> encoder walk -> DecoderBN -> head); then one training step."""
    smode, imode = "glove-25d", "ade20k_swin_human_sizes"
    m = make_model(insertion_point="input", semantics_mode=smode, instance_segmentation_mode=imode).to(DEV).channels_last_()
    b, h, w = 1, 352, 384
    x = synthetic.image(b, h, w, seed=87)
    slab, _ = sem_labels(smode, b, h, w, seed=88, n_rect=(20, 40))
    ilab, iar = inst_labels(imode, b, h, w, seed=89, n_rect=(20, 40))
    _, sem = SemanticsLoader(Args(use_semantics=smode)).get_semantics({"semantics": slab})
    _, il, ia = InstanceSegmentationLoader(Args(use_instance_segmentation=imode)).get_instance_segmentation(
        {"instance_labels": ilab, "instance_areas": iar})
    with torch.no_grad():
        edges, pred = m(x.to(DEV), semantics=sem, instance_labels=il, instance_areas=ia)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    _, sem_ref = oracle.semantics_loader(smode, slab.numpy(), load_table("ade20k_150_classes_glove_twitter_27b_25d_embeddings.npy"))
    _, il_ref, ia_ref = oracle.instance_loader(imode, ilab.numpy(), iar.numpy(),
                                               load_table("ade20k_places_classes_glove_twitter_27b_25d_embeddings.npy"), 100,
                                               load_table("ade20k_classes_abs_sizes.npy"))
    cpu = make_model(insertion_point="input", semantics_mode=smode, instance_segmentation_mode=imode)  # same weights, CPU
    with torch.no_grad():
        xin = oracle.input_insertion(sd, x, smode, imode, "rgb", semantics=torch.from_numpy(sem_ref),
                                     instance_labels=torch.from_numpy(il_ref), instance_areas=torch.from_numpy(ia_ref))
        assert xin.shape[1] == 73
        unet = oracle.decoder_bn(oracle.encoder_features(cpu.encoder.original_model, xin), sd)
        e_ref, p_ref = oracle.head(unet, sd, 1e-3, 10.0)
    assert rel_err(edges.cpu(), e_ref) < REL_DEPTH
    assert_depth_close(pred.cpu(), p_ref)
    # one training step: gradients reach the aux MLPs through the channels_last input concatenation
    depth = synthetic.depth(b, h, w, seed=90).to(DEV)
    m.train()
    e, p = m(x.to(DEV), semantics=sem, instance_labels=il, instance_areas=ia)
    loss = SILogLoss()(p, depth, mask=depth > 1e-3) + 0.1 * BinsChamferLoss()(e, depth)
    loss.backward()
    for name in ("instance_areas_fc.0.weight", "instance_absolute_sizes_fc.2.bias", "decoder.conv3.weight",
                 "adaptive_bins_layer.conv3x3.weight", "conv_out.0.weight"):
        g = dict(m.named_parameters())[name].grad
        assert g is not None and torch.isfinite(g).all() and float(g.abs().max()) > 0, name


# ------------------------------------------------------------------------------------------------------------
# BASELINE config 2 at the benchmarked size, in the benchmarked mode
# ------------------------------------------------------------------------------------------------------------
def test_config2_full_size_bench_mode_vs_oracle():
    """B = 16, 416 x 544, GloVe-25d ADE20K-places semantics at the input; channels_last model, PyTorch's default backend flags
    (cudnn.allow_tf32 = True), the whole step (loader gather -> model -> SILog + chamfer) replayed as ONE CUDA graph
    (graphs.GraphedStep) exactly as bench.py runs it -- against the CPU oracle of the whole reference path: `pred` and
    `bin_edges` within 1e-3 relative on every pixel, both losses within 1e-4."""
    from mde_biological_vision_systems_b200.graphs import GraphedStep
    assert torch.backends.cudnn.allow_tf32, "this test must run under PyTorch's default flags (the mode bench.py measures)"
    mode = "glove-25d-ade20k-places"
    b, h, w = 16, 416, 544
    cpu = make_model(insertion_point="input", semantics_mode=mode, instance_segmentation_mode=None)
    sd = {k: v.detach().clone() for k, v in cpu.state_dict().items()}
    img = synthetic.image(b, h, w, seed=0)
    depth = synthetic.depth(b, h, w, seed=1)
    lab, _ = synthetic.label_maps(b, h, w, seed=2)
    table = load_table("ade20k_places_classes_glove_twitter_27b_25d_embeddings.npy")
    with torch.no_grad():
        _, sem_ref = oracle.semantics_loader(mode, lab.numpy(), table)
        xin = oracle.input_insertion(sd, img, mode, None, "rgb", semantics=torch.from_numpy(sem_ref))
        e_ref, p_ref, l1_ref, l2_ref = oracle.forward_and_losses(
            lambda t: oracle.decoder_bn(oracle.encoder_features(cpu.encoder.original_model, t), sd), sd, xin, depth, 1e-3, 10.0)
    m = make_model(insertion_point="input", semantics_mode=mode, instance_segmentation_mode=None).to(DEV).channels_last_()
    loader = SemanticsLoader(Args(use_semantics=mode))
    assert loader.bind_encoder_input(m)  # as bench.py: embeddings gathered straight into the NHWC encoder input
    both = DepthLosses(1e-3)

    def step(image, depth, semantics):
        _, sem = loader.get_semantics({"semantics": semantics})
        edges, pred = m(image, semantics=sem)
        l_dense, l_bins = both(pred, edges, depth, interpolate=True)  # train.py:414-419, one pass over the depth map
        return edges, pred, l_dense, l_bins

    resident = {"image": img.to(DEV), "depth": depth.to(DEV), "semantics": lab.to(DEV)}
    l0 = ops.launch_count()
    g = GraphedStep(step, resident)
    assert ops.launch_count() > l0
    edges, pred, l1, l2 = g(**resident)
    torch.cuda.synchronize()
    mx, p999 = rel_stats(pred.cpu(), p_ref)
    print("config 2 full size, graph replay: pred max %.3e p99.9 %.3e  edges %.3e  silog %.6f/%.6f  chamfer %.6f/%.6f" % (
        mx, p999, rel_err(edges.cpu(), e_ref), float(l1), float(l1_ref), float(l2), float(l2_ref)))
    assert rel_err(edges.cpu(), e_ref) < REL_DEPTH
    assert_depth_close(pred.cpu(), p_ref)
    assert abs(float(l1) - float(l1_ref)) <= REL_LOSS * abs(float(l1_ref))
    assert abs(float(l2) - float(l2_ref)) <= REL_LOSS * abs(float(l2_ref))


def test_bf16_mode_within_2e_2_of_fp32_reference():
    """north star: "depth maps and bin edges within 1e-3 relative in fp32/TF32 (2e-2 in bf16)".  model.precision = "bf16" runs
    the decoder convolutions, the head convolution and the fused chain with ONE bf16 product per K step (hi planes only) and
    the library bodies at their TF32 default; pred / edges vs the CPU fp32 oracle of the reference path within 2e-2, and the
    mode really is a different computation from the fp32-grade default."""
    mode = "glove-25d-ade20k-places"
    b, h, w = 4, 416, 544
    cpu = make_model(insertion_point="input", semantics_mode=mode, instance_segmentation_mode=None)
    sd = {k: v.detach().clone() for k, v in cpu.state_dict().items()}
    img = synthetic.image(b, h, w, seed=0)
    lab, _ = synthetic.label_maps(b, h, w, seed=2)
    table = load_table("ade20k_places_classes_glove_twitter_27b_25d_embeddings.npy")
    with torch.no_grad():
        _, sem_ref = oracle.semantics_loader(mode, lab.numpy(), table)
        xin = oracle.input_insertion(sd, img, mode, None, "rgb", semantics=torch.from_numpy(sem_ref))
        e_ref, p_ref = oracle.head(oracle.decoder_bn(oracle.encoder_features(cpu.encoder.original_model, xin), sd), sd, 1e-3, 10.0,
                                   "linear")
    m = make_model(insertion_point="input", semantics_mode=mode, instance_segmentation_mode=None).to(DEV).channels_last_()
    loader = SemanticsLoader(Args(use_semantics=mode))
    with torch.no_grad():
        _, sem = loader.get_semantics({"semantics": lab.to(DEV)})
        e32, p32 = m(img.to(DEV), semantics=sem)
        m.precision = "bf16"
        e16, p16 = m(img.to(DEV), semantics=sem)
    mx, p999 = rel_stats(p16.cpu(), p_ref)
    print("bf16 mode: pred max %.3e p99.9 %.3e  edges %.3e   (fp32 mode: pred max %.3e)" % (
        mx, p999, rel_err(e16.cpu(), e_ref), rel_stats(p32.cpu(), p_ref)[0]))
    assert rel_err(e16.cpu(), e_ref) < 2e-2
    assert_depth_close(p16.cpu(), p_ref, 2e-2)
    assert not torch.equal(p16, p32)
    m.precision = "fp64"
    with pytest.raises(ValueError):
        m(img.to(DEV), semantics=sem)


@pytest.mark.parametrize("cfg", [dict(b=2, c=128, h=24, w=40, cout=128), dict(b=1, c=352, h=13, w=17, cout=160, pair=True)])
def test_conv3x3_single_bf16_product(cfg):
    """products = 1 (ops.bf16_products): the kernel multiplies the hi planes only -- equal, up to fp32 accumulation, to a float64
    convolution of the bf16-rounded operands."""
    rng = np.random.default_rng(191)
    b, c, h, w, cout = cfg["b"], cfg["c"], cfg["h"], cfg["w"], cfg["cout"]
    x = torch.from_numpy(rng.standard_normal((b, c, h, w)).astype(np.float32)).to(DEV)
    wt = torch.from_numpy((rng.standard_normal((cout, c, 3, 3)) / np.sqrt(9 * c)).astype(np.float32)).to(DEV)
    xp, wp = ops.split_bf16(x), ops.prepare_conv3x3_weight(wt)
    ref = torch.nn.functional.conv2d(xp.planes[0].float().permute(0, 3, 1, 2).cpu().double(),
                                     wt.bfloat16().float().cpu().double(), None, padding=1)
    with ops.bf16_products():
        assert ops.products() == 1
        out = ops.conv3x3_nhwc(xp, wp, pair_out=bool(cfg.get("pair")))
    assert ops.products() == 3
    out = out.float() if cfg.get("pair") else out
    err = float((out.cpu().double() - ref).abs().max()) / float(ref.abs().max())
    assert err < (1e-4 if cfg.get("pair") else 2e-5), err  # a pair output is itself rounded to ~2^-17
    full = ops.conv3x3_nhwc(xp, wp)
    assert float((full.cpu().double() - ref).abs().max()) / float(ref.abs().max()) > 1e-4  # the default is a different product


def test_torch_library_ops_match_ctypes_layer():
    """torch.ops.mde.* (torch.library registration over the same C ABI) return what the ctypes layer returns, their
    registered autograd formulas give the same gradients, and torch.library.opcheck accepts schema / fake / autograd
    registration of the loss operator."""
    from mde_biological_vision_systems_b200 import torch_ops
    rng = np.random.default_rng(160)
    b, h, w = 2, 48, 64
    x = torch.from_numpy(rng.standard_normal((b, 128, h, w)).astype(np.float32)).to(DEV)
    planes = torch.ops.mde.split_bf16(x)
    assert torch.equal(planes, ops.split_bf16(x).planes)
    wt = torch.from_numpy((rng.standard_normal((128, 128, 3, 3)) / 34.0).astype(np.float32)).to(DEV)
    wp = ops.prepare_conv3x3_weight(wt)
    assert torch.equal(torch.ops.mde.conv3x3_x3(planes, wp, None, None, 1.0, False), ops.conv3x3_nhwc(ops.SplitBF16(planes), wp))
    depth = synthetic.depth(b, 2 * h, 2 * w, seed=161).to(DEV)
    pred = torch.from_numpy((0.3 + 9 * rng.random((b, 1, h, w), dtype=np.float32))).to(DEV)
    edges = torch.linspace(1e-3, 10, 257, device=DEV).repeat(b, 1).contiguous()
    p1, e1 = pred.clone().requires_grad_(True), edges.clone().requires_grad_(True)
    p2, e2 = pred.clone().requires_grad_(True), edges.clone().requires_grad_(True)
    s1, c1 = torch_ops.depth_losses(p1, e1, depth)
    s2, c2 = ops.depth_losses(p2, e2, depth)
    assert float(s1) == float(s2) and float(c1) == float(c2)
    (s1 + 0.1 * c1).backward()
    (s2 + 0.1 * c2).backward()
    assert torch.allclose(p1.grad, p2.grad, rtol=1e-5, atol=1e-9) and torch.equal(e1.grad, e2.grad)
    l1 = torch_ops.silog(pred, depth, depth > 1e-3, True)
    assert float(l1) == float(SILogLoss()(pred, depth, mask=depth > 1e-3))
    torch.library.opcheck(torch.ops.mde.silog_fwd.default, (pred, depth, depth > 1e-3, True),
                          test_utils=("test_schema", "test_faketensor"))


def test_training_step_with_own_conv_kernels_matches_cudnn():
    """train_conv_impl = "tc" routes every 3x3 convolution of the decoder and the head (forward, dgrad, wgrad) through
    ops.conv3x3_autograd inside the real model: same loss and gradients as the stock cuDNN modules."""
    kw = dict(insertion_point="input", semantics_mode=None, instance_segmentation_mode=None)
    x = synthetic.image(2, 352, 384, seed=170).to(DEV)   # 22 x 24 = 528 tokens >= 1 + 128 queries
    depth = synthetic.depth(2, 352, 384, seed=171).to(DEV)
    out = {}
    for impl in ("cudnn", "tc"):
        m = make_model(**kw).to(DEV).channels_last_()
        m.train()
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
            if isinstance(mod, torch.nn.MultiheadAttention):
                mod.dropout = 0.0
            if hasattr(mod, "train_conv_impl"):
                mod.train_conv_impl = impl
        l0 = ops.launch_count()
        with ops.exact_fp32_library():
            e, p = m(x)
            s, c = DepthLosses(1e-3)(p, e, depth)
            (s + 0.1 * c).backward()
        out[impl] = (float(s.detach()), m.decoder.up3._net[0].weight.grad.clone(), m.decoder.conv3.weight.grad.clone(),
                     m.adaptive_bins_layer.conv3x3.weight.grad.clone(), ops.launch_count() - l0)
    assert out["tc"][4] > out["cudnn"][4] + 20  # the own-kernel path really ran
    assert abs(out["tc"][0] - out["cudnn"][0]) <= 1e-3 * abs(out["cudnn"][0])
    for a, b in zip(out["tc"][1:4], out["cudnn"][1:4]):
        assert float((a - b).abs().max()) <= 3e-2 * float(b.abs().max()) + 1e-8  # train-mode BatchNorm at batch 2 (see above)


def test_graphed_train_step_matches_eager():
    """training.GraphedTrainStep (zero_grad + loader gather + forward + SILog / chamfer + backward replayed as ONE CUDA graph,
    then the eager all-reduce / clip / AdamW / OneCycle) follows the eager TrainStep: same loss trajectory over several
    iterations from the same initial weights (dropout off: the two launch modes draw different masks), weights really
    change between replays, and the captured step contains this package's launches."""
    from argparse import Namespace
    from mde_biological_vision_systems_b200.training import GraphedTrainStep, TrainStep
    mode = "glove-25d-ade20k-places"
    b, h, w = 2, 352, 384
    lab, _ = sem_labels(mode, b, h, w, seed=201, n_rect=(20, 40))
    batch = {"image": synthetic.image(b, h, w, seed=200).pin_memory(), "depth": synthetic.depth(b, h, w, seed=202).pin_memory(),
             "semantics": lab.pin_memory()}
    losses = {}
    for kind in ("eager", "graph"):
        torch.manual_seed(0)
        m = make_model(insertion_point="input", semantics_mode=mode, instance_segmentation_mode=None).to(DEV).channels_last_()
        m.train()
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
            if isinstance(mod, torch.nn.MultiheadAttention):
                mod.dropout = 0.0
        loader = SemanticsLoader(Namespace(use_semantics=mode), device=DEV)
        loader.bind_encoder_input(m)
        kw = dict(semantics_loader=loader, total_steps=100, cudnn_benchmark=False)
        if kind == "eager":
            st = TrainStep(m, **kw)
            losses[kind] = [float(st(batch, DEV)) for _ in range(5)]
        else:
            st = GraphedTrainStep(m, batch, DEV, warmup=2, **kw)   # two eager iterations inside, then the capture
            assert st.captured_launches > 20
            w0 = m.conv_out[0].weight.detach().clone()
            losses[kind] = [None, None] + [float(st(batch)) for _ in range(3)]
            assert not torch.equal(w0, m.conv_out[0].weight.detach())
            st.prefetch(batch)   # the overlapped host -> device copy of the next batch, consumed by a call without arguments
            staged = float(st())
            assert abs(staged - losses["eager"][4]) <= 2e-2 * abs(losses["eager"][4]) and staged != losses[kind][4]
    for i in range(2, 5):  # iterations 3..5 of both runs (the graphed run's first two were its eager warm-up)
        assert abs(losses["graph"][i] - losses["eager"][i]) <= 2e-3 * abs(losses["eager"][i]), losses
    assert losses["eager"][4] != losses["eager"][2]


def test_depthwise_engine_is_exact():
    """models/efficientnet.py leaves cuDNN's TF32 switch on for DEPTHWISE convolutions inside the exact-fp32 region (it only
    selects the NHWC engine; depthwise math is fp32 FMAs either way): the two settings must give bit-identical outputs."""
    rng = np.random.default_rng(190)
    for (c, k, s, hw) in [(96, 3, 2, (104, 136)), (240, 5, 1, (52, 68)), (672, 5, 2, (26, 34)), (32, 3, 1, (208, 272))]:
        x = torch.from_numpy(rng.standard_normal((2, c, *hw)).astype(np.float32)).to(DEV).contiguous(memory_format=torch.channels_last)
        w = torch.from_numpy(rng.standard_normal((c, 1, k, k)).astype(np.float32)).to(DEV)
        outs = []
        for flag in (False, True):
            torch.backends.cudnn.allow_tf32 = flag
            outs.append(torch.nn.functional.conv2d(x, w, None, s, k // 2, 1, c))
        torch.backends.cudnn.allow_tf32 = True
        assert torch.equal(outs[0], outs[1]), (c, k, s)
