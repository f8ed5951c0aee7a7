"""Drop-in for the reference's models/unet_adaptive_bins.py.

Keeps ``UnetAdaptiveBins.build()/forward() -> (bin_edges, pred)``, ``get_1x_lr_params/get_10x_lr_params``,
``get_num_channels_to_add`` and every state_dict key (/root/reference/models/unet_adaptive_bins.py:119-395), while
the head runs on the sm_100a kernels of this package:

* external-info insertion at the input (:194-235): the 1x1-conv MLPs run as one streaming kernel each
  (ops.aux_mlp) writing straight into the concatenated encoder input;
* mViT + conv_out + softmax + bins (:285-302): ops.fold_queries + ops.head_chain (TMA + tcgen05, nothing but
  ``pred`` written) or, when the fused kernel's shape constraints do not hold, range attention -> conv1x1 ->
  ops.bins_pred (streaming);
* DecoderBN (:39-100) in inference: resize + concat and every 3x3 conv (BatchNorm(eval) + LeakyReLU in
  the epilogue) on our kernels, activations handed from kernel to kernel as split-bf16 pairs (ops.SplitBF16);
* noAdaBins epilogue (:240-242): ops.relu_eps.

Precision: every tensor-core product of the inference path is formed from split-bf16 pairs (three bf16 products, fp32
accumulation: fp32-grade), and the library passthrough bodies (EfficientNet encoder, decoder conv2) run in true fp32
(``backbone_tf32 = False``), so `pred` / `bin_edges` stay within 1e-3 of the fp32 reference on every pixel.

The EfficientNet encoder is outside the hot path and stays a plain PyTorch/cuDNN module (as do the
decoder / head bodies in training mode); it only exists so the surface and the checkpoints match.
"""
import contextlib
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .efficientnet import SamePadConv2d, build_backbone
from .miniViT import mViT

Conv2dSame = SamePadConv2d  # the reference exposes this name (unet_adaptive_bins.py:24-36)


def _block(cin, cout):
    return [nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=1), nn.BatchNorm2d(cout), nn.LeakyReLU()]


class UpSampleBN(nn.Module):
    """bilinear(align_corners) up-sample to the skip's size, concat, 2 x (conv3x3-BN-LeakyReLU); keys ``_net.{0,1,3,4}``."""

    def __init__(self, skip_input, output_features):
        super().__init__()
        self._net = nn.Sequential(*_block(skip_input, output_features), *_block(output_features, output_features))
        # training / autograd: "cudnn" = stock modules (default); "tc" = our conv kernels (ops.conv3x3_autograd: forward and
        # dgrad on the bf16x3 tcgen05 kernel, wgrad on the tap-shifted TF32 NT GEMM) -- parity-tested, but measured 3x slower
        # than cuDNN's TF32 kernels on B200 (29 vs 9.5 ms of the config-2 training step: the nine taps re-read both operands),
        # see DESIGN.md section 6
        self.train_conv_impl = "cudnn"

    def _folded(self):
        """Per conv block: ([dx][dy][Cout][C] split-bf16 filter, scale, shift, slope) with the eval-mode BatchNorm folded into
        a per-channel affine (scale = gamma / sqrt(var + eps), shift = beta + (conv bias - mean) * scale); cached per
        parameter / running-statistics version."""
        tensors = []
        for i in (0, 3):
            conv, bn = self._net[i], self._net[i + 1]
            tensors += [conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var]
        key = tuple(t._version for t in tensors) + (tensors[0].device,)
        cached = getattr(self, "_mde_folded", None)
        if cached is None or cached[0] != key:
            blocks = []
            with torch.no_grad():
                for i in (0, 3):
                    conv, bn = self._net[i], self._net[i + 1]
                    scale = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
                    shift = bn.bias + (conv.bias - bn.running_mean) * scale
                    # the concatenated input of the first conv is written with its channel pitch rounded up to 32 (64-byte
                    # operand rows stay sector-aligned: 680- and 344-channel rows otherwise start at 16-byte offsets)
                    w = ops.prepare_conv3x3_weight(conv.weight, cin_pad_to=32 if i == 0 else 1)
                    blocks.append((w, scale.contiguous(), shift.contiguous(), self._net[i + 2].negative_slope))
            cached = (key, blocks)
            self._mde_folded = cached
        return cached[1]

    def tc_supported(self, c_up, skip):
        c_in = self._net[0].in_channels
        return (c_up % 4 == 0 and skip.shape[1] % 4 == 0 and c_up + skip.shape[1] == c_in and c_in % 8 == 0
                and ops.conv3x3_cout_ok(self._net[0].out_channels, pair_out=True)
                and ops.conv3x3_cout_ok(self._net[3].out_channels, pair_out=True))

    def forward_tc(self, x_cl, concat_with, pair_out=False, name="up"):
        """Inference path on our kernels, channels_last throughout: resize + concat (one streaming pass, written as a
        split-bf16 pair) -> 2 x tcgen05 conv3x3 (three bf16 products per K step) with BatchNorm(eval) + LeakyReLU in the
        epilogue.  Returns fp32 channels_last (input of the next resize) or an ops.SplitBF16 (``pair_out``)."""
        (w1, s1, b1, a1), (w2, s2, b2, a2) = self._folded()
        y = ops.upsample_concat_nhwc_pair(x_cl, concat_with, pad_to=32)
        y = ops.conv3x3_nhwc(y, w1, s1, b1, slope=a1, pair_out=True, name=name + ".conv_a")
        return ops.conv3x3_nhwc(y, w2, s2, b2, slope=a2, pair_out=pair_out, name=name + ".conv_b")

    def _net_train(self, y):
        """The block with its two 3x3 convolutions on our kernels (forward, dgrad, wgrad: ops.conv3x3_autograd); BatchNorm
        (batch statistics; SyncBatchNorm2d kernels when converted) and LeakyReLU stay modules."""
        for i in (0, 3):
            conv = self._net[i]
            if ops.conv3x3_train_supported(y, conv):
                y = ops.conv3x3_autograd(y, conv.weight, conv.bias)
            else:
                y = conv(y)
            y = self._net[i + 2](self._net[i + 1](y))
        return y

    def forward(self, x, concat_with):
        if x.is_cuda:  # fused resize + concat kernel (ATen's align_corners bilinear kernel dominates the step otherwise)
            if x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous() \
                    and x.shape[1] % 4 == 0 and concat_with.shape[1] % 4 == 0:
                y = ops.upsample_concat_nhwc(x, concat_with)  # channels_last model: stay in NHWC
                if self.train_conv_impl == "tc" and y.dtype == torch.float32 and not torch.is_autocast_enabled():
                    return self._net_train(y)
                return self._net(y)
            return self._net(ops.upsample_concat(x, concat_with))
        raise ops._lib.MdeError("UpSampleBN runs on the B200 kernels only (no CPU path)")


class DecoderBN(nn.Module):
    _SKIP_EXTRA = {2048: (64, 24, 16, 8), 1280: (0, 0, 0, 0)}  # B5 / B1 skip widths differ from the B1 defaults

    def __init__(self, num_features=2048, num_classes=1, bottleneck_features=2048, mode="AdaBins"):
        super().__init__()
        f = int(num_features)
        extra = self._SKIP_EXTRA[f]
        self.conv2 = nn.Conv2d(bottleneck_features, f, kernel_size=1, stride=1, padding=1)
        self.up1 = UpSampleBN(f + 112 + extra[0], f // 2)
        self.up2 = UpSampleBN(f // 2 + 40 + extra[1], f // 4)
        self.up3 = UpSampleBN(f // 4 + 24 + extra[2], f // 8)
        self.up4 = UpSampleBN(f // 8 + 16 + extra[3], f // 16)
        self.mode = mode
        self.conv3 = nn.Conv2d(f // 16, num_classes if mode == "AdaBins" else 1, kernel_size=3, stride=1, padding=1)

        self.conv_impl = "auto"  # "auto" / "tc": the tcgen05 conv3x3 path in inference; "cudnn": stock modules (fp32)

    def _use_tc(self, feats):
        if self.conv_impl == "cudnn" or self.training or torch.is_grad_enabled():
            return False
        s0, s1, s2, s3, bottleneck = feats
        if not (bottleneck.is_cuda and bottleneck.dtype == torch.float32):
            return False
        c = self.conv2.out_channels
        for up, skip in ((self.up1, s3), (self.up2, s2), (self.up3, s1), (self.up4, s0)):
            if not up.tc_supported(c, skip):
                return False
            c = up._net[3].out_channels
        cout = self.conv3.out_channels
        return cout <= 4 or ops.conv3x3_cout_ok(cout, pair_out=True)

    def _prepared_conv2(self):
        w = self.conv2.weight
        cached = getattr(self, "_mde_w2_prep", None)
        if cached is None or cached[0] != w._version or cached[1].device != w.device:
            cached = (w._version, ops.prepare_pointwise_weight(w))
            self._mde_w2_prep = cached
        return cached[1]

    def _prepared_conv3(self):
        w = self.conv3.weight
        cached = getattr(self, "_mde_w3_prep", None)
        if cached is None or cached[0] != w._version or cached[1].device != w.device:
            cached = (w._version, ops.prepare_conv3x3_weight(w))
            self._mde_w3_prep = cached
        return cached[1]

    def forward(self, features):
        """-> unet_out: an fp32 [B,C,h,w] tensor, or (inference on our kernels, AdaBins mode) the same feature map as an
        ops.SplitBF16, the operand format of the head's tensor-core kernels."""
        s0, s1, s2, s3, bottleneck = features[4], features[5], features[6], features[8], features[11]
        if self._use_tc((s0, s1, s2, s3, bottleneck)):
            # (f)1: the whole decoder on our kernels, channels_last.  conv2 is a 1x1 conv with padding 1 (:61): pad the 13x17
            # input by one zero pixel and run the point-wise GEMM -- border pixels come out as the bias, as in the padded conv
            xb = F.pad(bottleneck, (1, 1, 1, 1)).contiguous(memory_format=torch.channels_last)
            if ops.pointwise_supported(xb, self.conv2.in_channels, self.conv2.out_channels):
                y = ops.pointwise_conv(xb, self._prepared_conv2(), self.conv2.bias, name="decoder.conv2")
            else:
                with ops.exact_fp32_library():
                    y = self.conv2(bottleneck.contiguous(memory_format=torch.channels_last))
                y = y.contiguous(memory_format=torch.channels_last)
            small = self.conv3.out_channels <= 4  # noAdaBins: 1 output channel, direct fp32 kernel
            ups = ((self.up1, s3), (self.up2, s2), (self.up3, s1), (self.up4, s0))
            for i, (up, skip) in enumerate(ups):
                y = up.forward_tc(y, skip, pair_out=(i == 3 and not small), name="up%d" % (i + 1))
            if small:
                return ops.conv3x3_small(y, self.conv3.weight, self.conv3.bias)
            return ops.conv3x3_nhwc(y, self._prepared_conv3(), None, self.conv3.bias, pair_out=True, name="decoder.conv3")
        y = self.conv2(bottleneck)
        for up, skip in ((self.up1, s3), (self.up2, s2), (self.up3, s1), (self.up4, s0)):
            y = up(y, skip)
        if self.up4.train_conv_impl == "tc" and not torch.is_autocast_enabled() and y.is_cuda \
                and y.is_contiguous(memory_format=torch.channels_last) and ops.conv3x3_train_supported(y, self.conv3):
            return ops.conv3x3_autograd(y, self.conv3.weight, self.conv3.bias)
        return self.conv3(y)


class Encoder(nn.Module):
    """Walks the backbone's children in order and records every intermediate (the decoder indexes the list)."""

    def __init__(self, backend):
        super().__init__()
        self.original_model = backend

    def forward(self, x, stem_prepadded=False):
        """``stem_prepadded``: x already carries the zero border the TensorFlow-SAME stem convolution would add.
        Inference on CUDA (no autograd, eval-mode BatchNorm) takes a shorter walk that leaves the entries no consumer reads as
        ``None``: conv_stem + bn1 + act1 run as one folded convolution (entries 1 and 2 are None, entry 3 is the activation),
        conv_head as the fp32-grade tcgen05 1x1 GEMM when the library's TF32 is off, and whatever follows conv_head (bn2, act2,
        pooling, classifier -- the decoder's deepest tap is conv_head's own output, entry 11) is not computed."""
        om = self.original_model
        if (x.is_cuda and not torch.is_grad_enabled() and not om.training and isinstance(getattr(om, 'bn1', None), nn.BatchNorm2d)
                and list(om._modules)[:5] == ['conv_stem', 'bn1', 'act1', 'blocks', 'conv_head']
                and isinstance(om.act1, nn.SiLU)):
            return self._forward_inference(x, stem_prepadded)
        feats = [x]
        for name, child in self.original_model._modules.items():
            stages = child._modules.values() if name == 'blocks' else (child,)
            for stage in stages:
                if name == 'conv_stem' and stem_prepadded and isinstance(stage, SamePadConv2d):
                    feats.append(stage.forward_with(feats[-1], stage.weight, stage.bias, prepadded=True))
                else:
                    feats.append(stage(feats[-1]))
        return feats

    def _forward_inference(self, x, stem_prepadded):
        from .efficientnet import conv_bn
        om = self.original_model
        y = conv_bn(om.conv_stem, om.bn1, x, om.act1, prepadded=stem_prepadded and isinstance(om.conv_stem, SamePadConv2d))
        feats = [x, None, None, y]
        for stage in om.blocks._modules.values():
            feats.append(stage(feats[-1]))
        head, y = om.conv_head, feats[-1]
        if (not torch.backends.cudnn.allow_tf32 and head.kernel_size == (1, 1) and head.bias is None
                and ops.pointwise_supported(y, head.in_channels, head.out_channels)):
            key = head.weight._version
            cached = getattr(head, "_mde_pair", None)
            if cached is None or cached[0] != key or cached[1].device != y.device:
                cached = (key, ops.prepare_pointwise_weight(head.weight.detach().flatten(1)))
                head._mde_pair = cached
            feats.append(ops.pointwise_conv(y, cached[1], None, 0, None, name="pointwise"))
        else:
            feats.append(head(y))
        feats.extend([None] * (len(om._modules) - 5))
        return feats


def _mlp(cin):
    return nn.Sequential(nn.Conv2d(cin, 10, kernel_size=1), nn.ReLU(), nn.Conv2d(10, 10, kernel_size=1), nn.ReLU())


class UnetAdaptiveBins(nn.Module):
    def __init__(self, backend, n_bins=100, min_val=0.1, max_val=10, norm='linear', encoder_name="efficientnet-b5",
                 semantics_mode=None, instance_segmentation_mode=None, insertion_point="before-attn", image="rgb"):
        super().__init__()
        self.num_classes = n_bins
        self.min_val = min_val
        self.max_val = max_val
        self.encoder = Encoder(backend)
        self.semantics_mode = semantics_mode
        self.instance_segmentation_mode = instance_segmentation_mode
        self.insertion_point = insertion_point
        self.image = image
        self.encoder_name = encoder_name
        self.image_pre_encode = None
        self.fused_head = True  # False: range attention -> conv1x1 -> streaming bins (three kernels)
        # False (default): the cuDNN passthrough bodies run in true fp32 during inference (the tolerance contract is against
        # the fp32 reference); True: leave torch.backends.cudnn.allow_tf32 as the caller set it (PyTorch's default: TF32)
        self.backbone_tf32 = False
        # "fp32" (default): the tensor-core kernels form fp32-grade products (three bf16 MMAs per K step) -- depth and bin edges
        # within 1e-3 of the fp32 reference.  "bf16": ONE bf16 product per K step in the decoder convolutions, the head
        # convolution and the fused range-attention chain, and the passthrough bodies at the library's default (TF32) -- the
        # north star's bf16 mode, depth and bin edges within 2e-2 (inference only; tests/test_gpu_parity.py).
        self.precision = "fp32"

        self.num_decoded_channels = 128
        extra = UnetAdaptiveBins.get_num_channels_to_add(encoder_name, semantics_mode, instance_segmentation_mode, image)
        if insertion_point == "before-attn":
            self.num_decoded_channels += extra

        if semantics_mode is not None:
            if semantics_mode == "glove-25d-inst-areas":
                self.semantics_areas_fc = _mlp(1)
            if "human-sizes" in semantics_mode:
                self.semantics_absolute_sizes_fc = _mlp(3)
        if instance_segmentation_mode is not None:
            self.instance_areas_fc = _mlp(1)
            if "human_sizes" in instance_segmentation_mode:
                self.instance_absolute_sizes_fc = _mlp(3)

        adabins = "noAdaBins" not in encoder_name
        if adabins:
            self.adaptive_bins_layer = mViT(self.num_decoded_channels, n_query_channels=128, patch_size=16,
                                            dim_out=n_bins, embedding_dim=128, norm=norm)
        if "efficientnet-b5" in encoder_name:
            self.decoder = DecoderBN(num_classes=128, num_features=2048, bottleneck_features=2048)
        elif "efficientnet-b1" in encoder_name:
            self.decoder = DecoderBN(num_classes=128, num_features=1280, bottleneck_features=1280,
                                     mode="AdaBins" if adabins else "noAdaBins")
        if adabins:
            self.conv_out = nn.Sequential(nn.Conv2d(128, n_bins, kernel_size=1, stride=1, padding=0), nn.Softmax(dim=1))

    def channels_last_(self):
        """Keep the encoder / decoder weights and activations in channels_last (NHWC): cuDNN then runs its NHWC conv and
        batch-norm kernels without per-layer layout conversions (training step: 102 -> 78 ms of kernels at B = 16), and the
        decoder hands the head an NHWC feature map, which is what the tcgen05 kernels consume.  Tensor shapes, values and
        state_dict keys are unchanged -- only strides differ."""
        self.encoder.to(memory_format=torch.channels_last)
        self.decoder.to(memory_format=torch.channels_last)
        self._channels_last = True
        return self

    def stem_pads(self, h, w):
        """(top, bottom, left, right) zeros the TensorFlow-SAME stem convolution adds to an h x w input ((0, 0, 0, 0) for a
        plain stem)."""
        stem = getattr(self.encoder.original_model, "conv_stem", None)
        if isinstance(stem, SamePadConv2d) and self.image != "none":
            return tuple(int(v) for v in stem.same_pads(h, w))
        return (0, 0, 0, 0)

    # ---- external-info insertion -------------------------------------------------------------------------------
    @staticmethod
    def _run_mlp(seq, x, in_div, out):
        return ops.aux_mlp(x, seq[0].weight, seq[0].bias, seq[2].weight, seq[2].bias, in_div=in_div, out=out)

    def _external_channels(self, semantics, instance_labels, instance_areas, hw):
        """The channel groups the reference concatenates (unet_adaptive_bins.py:194-228 / :244-282), in order, as
        ("copy", tensor) or ("mlp", module, input tensor, divisor) items."""
        items = []
        if semantics is not None:
            if self.semantics_mode == "glove-25d-inst-areas":
                items.append(("copy", semantics[:, 0:25]))
                items.append(("mlp", self.semantics_areas_fc, semantics[:, 25:26], 1.0))
            elif "human-sizes" in self.semantics_mode:
                items.append(("copy", semantics[:, 0:-3]))
                items.append(("mlp", self.semantics_absolute_sizes_fc, semantics[:, -3:], 1.0))
            else:
                items.append(("copy", semantics))
        if instance_labels is not None:
            items.append(("copy", instance_labels))
        if instance_areas is not None:
            if "human_sizes" in self.instance_segmentation_mode:
                items.append(("mlp", self.instance_areas_fc, instance_areas[:, 0:1], float(hw)))
                items.append(("mlp", self.instance_absolute_sizes_fc, instance_areas[:, 1:4], 1.0))
            else:
                items.append(("mlp", self.instance_areas_fc, instance_areas, float(hw)))
        return items

    def _concat_external(self, x, items, pads=None):
        """One allocation for the widened tensor; pass-through groups are copied (with the .float() cast fused into
        the copy), MLP groups are written in place by the streaming kernel.  ``pads`` (top, bottom, left, right) is honoured
        only by the channels_last fast path (callers detect it by the grown spatial size)."""
        if not items:
            return x
        bound = getattr(items[0][1], "_mde_encoder_input", None) if len(items) == 1 and items[0][0] == "copy" else None
        if bound is not None and getattr(self, "_channels_last", False) and x.is_cuda \
                and not (torch.is_grad_enabled() and x.requires_grad):
            # the loader gathered the embedding planes straight into the encoder's NHWC input (SemanticsLoader.
            # bind_encoder_input): only the image planes are missing
            buf, bpads, filled = bound
            want = (0, 0, 0, 0) if pads is None else tuple(int(v) for v in pads)
            if tuple(bpads) == want and buf.shape[1] == x.shape[1] + items[0][1].shape[1] \
                    and buf.shape[0] == x.shape[0] and buf.shape[2] == x.shape[2] + want[0] + want[1]:
                if filled is not None and filled.data_ptr() == x.data_ptr() and filled._version == x._version:
                    return buf  # the loader's kernel already wrote these very image planes
                return ops.fill_channels_nhwc(buf, x, 0, want)
        if getattr(self, "_channels_last", False) and x.is_cuda and all(it[0] == "copy" for it in items) \
                and not (torch.is_grad_enabled() and x.requires_grad):
            # channels_last model, pass-through groups only (config 2): the planar sources are transposed straight into
            # their channel slices of the NHWC encoder input (no planar concatenation, no second layout pass)
            return ops.concat_channels_last([x] + [it[1] for it in items], pads)
        widths = [it[1].shape[1] if it[0] == "copy" else 10 for it in items]
        b, c, h, w = x.shape
        out = torch.empty((b, c + sum(widths), h, w), dtype=torch.float32, device=x.device)
        out[:, :c].copy_(x)
        ch = c
        for it, wd in zip(items, widths):
            dst = out[:, ch:ch + wd]
            if it[0] == "copy":
                dst.copy_(it[1])
            else:
                self._run_mlp(it[1], it[2], it[3], dst)
            ch += wd
        return out

    # ---- head -------------------------------------------------------------------------------------------------
    def _head(self, unet_out):
        """unet_out: fp32 tensor or ops.SplitBF16 (inference, from the tcgen05 decoder)."""
        head = self.adaptive_bins_layer
        conv = self.conv_out[0]
        fusable = self.fused_head and unet_out.is_cuda and self.num_classes == 256 \
            and (unet_out.shape[2] * unet_out.shape[3]) % 128 == 0 and head.conv3x3.out_channels == 128 \
            and head.n_query_channels == 128
        needs_grad = torch.is_grad_enabled() and (getattr(unet_out, "requires_grad", False) or conv.weight.requires_grad)
        if needs_grad and not fusable:
            raise RuntimeError("training needs the fused head (n_bins = 256, 128 query channels, h*w % 128 == 0)")
        if fusable and not needs_grad:
            # inference: the 3x3 conv runs bias-free on the tcgen05 kernel and hands its output to the fused chain as a
            # split-bf16 pair (K-major operand, consumed in place); the conv bias is folded into the chain's per-image bias.
            # (Shapes the tcgen05 conv does not cover -- e.g. the 153-channel before-attn input -- take the library conv in
            # true fp32 and the chain splits its output.)
            tgt, feat = head.tokens_and_features(unet_out, bias_free=True, pair_out=True)
            _, bin_edges, centers, _ = head.bin_widths(tgt, self.min_val, self.max_val)
            queries = tgt[1:head.n_query_channels + 1].permute(1, 0, 2)
            wf, biasf = ops.fold_queries(conv.weight, conv.bias, queries, feat_bias=head.conv3x3.bias)
            return bin_edges, ops.head_chain(feat, wf, biasf, centers)
        tgt, feat = head.tokens_and_features(unet_out)
        _, bin_edges, centers, _ = head.bin_widths(tgt, self.min_val, self.max_val)
        queries = tgt[1:head.n_query_channels + 1].permute(1, 0, 2)  # [N, 128, E] view
        if needs_grad:
            pred = ops.head_chain_autograd(feat, queries, conv.weight, conv.bias, centers)
        else:
            ram = ops.range_attention(feat, queries, impl="simt")
            pred = ops.bins_pred(ops.conv1x1(ram, conv.weight, conv.bias), centers)
        return bin_edges, pred

    def forward(self, x, semantics=None, instance_labels=None, instance_areas=None, **kwargs):
        stem_prepadded = False
        if self.insertion_point == "input":
            items = self._external_channels(semantics, instance_labels, instance_areas, x.shape[2] * x.shape[3])
            stem = getattr(self.encoder.original_model, "conv_stem", None)
            pads = None
            if isinstance(stem, SamePadConv2d) and self.image != "none":
                pads = stem.same_pads(x.shape[2], x.shape[3])  # written by the concatenation itself instead of an F.pad copy
            hw = x.shape[-2:]
            x = self._concat_external(x, items, pads)
            stem_prepadded = x.shape[-2:] != hw
        if self.image == "none":
            if x.shape[1] <= 3:
                sys.exit("Error: Add more auxiliary information at input if using no image")
            x = x[:, 3:, :, :]

        if getattr(self, "_channels_last", False) and x.is_cuda and x.dtype == torch.float32:
            if torch.is_grad_enabled() and x.requires_grad:  # trainable aux MLP channels: keep the autograd edge
                x = x.contiguous(memory_format=torch.channels_last)
            else:
                x = ops.to_channels_last(x)
        if self.precision not in ("fp32", "bf16"):
            raise ValueError("UnetAdaptiveBins.precision must be 'fp32' or 'bf16'")
        bf16 = self.precision == "bf16" and not (self.training or torch.is_grad_enabled())
        with ops.bf16_products(bf16):
            return self._forward_body(x, stem_prepadded, bf16, semantics, instance_labels, instance_areas, kwargs)

    def _forward_body(self, x, stem_prepadded, bf16, semantics, instance_labels, instance_areas, kwargs):
        exact = not (bf16 or self.backbone_tf32 or self.training or torch.is_grad_enabled())
        with (ops.exact_fp32_library() if exact else contextlib.nullcontext()):
            feats = self.encoder(x, stem_prepadded=stem_prepadded)
        unet_out = self.decoder(feats, **kwargs)

        if "noAdaBins" in self.encoder_name:
            return None, ops.relu_eps(unet_out, 0.0001)

        if self.insertion_point == "before-attn":
            if isinstance(unet_out, ops.SplitBF16):
                unet_out = unet_out.float()
            size = unet_out.shape[-2:]
            near = lambda t: None if t is None else F.interpolate(t, size=size, mode='nearest').float()
            # NB: the reference's human-sizes branch here concatenates onto x instead of unet_out (a bug that makes
            # that mode unusable, unet_adaptive_bins.py:253-259); this implementation concatenates onto unet_out.
            items = self._external_channels(near(semantics), near(instance_labels), near(instance_areas),
                                            x.shape[2] * x.shape[3])
            unet_out = self._concat_external(unet_out, items)

        return self._head(unet_out)

    def get_1x_lr_params(self):  # lr/10 learning rate
        return self.encoder.parameters()

    def get_10x_lr_params(self):  # lr learning rate
        modules = [self.decoder] if "noAdaBins" in self.encoder_name else \
            [self.decoder, self.adaptive_bins_layer, self.conv_out]
        for m in modules:
            yield from m.parameters()

    @classmethod
    def build(cls, n_bins, encoder_name="efficientnet-b5", insertion_point="before-attn", **kwargs):
        """Same call as the reference's build() (:315-360).  The backbone is the offline geffnet-shaped generator of
        models/efficientnet.py (random init; a geffnet ``tf_efficientnet_b{1,5}_ap`` state_dict loads into it)."""
        if "efficientnet-b5" in encoder_name:
            basemodel_name = 'tf_efficientnet_b5_ap'
        elif "efficientnet-b1" in encoder_name:
            basemodel_name = 'tf_efficientnet_b1_ap'
        else:
            sys.exit("Error [models/unet_adaptive_bins.py]: encoder not recognised")
        basemodel = build_backbone(basemodel_name)
        basemodel.global_pool = nn.Identity()
        basemodel.classifier = nn.Identity()

        if insertion_point == "input":
            extra = UnetAdaptiveBins.get_num_channels_to_add(
                encoder_name=encoder_name, semantics_mode=kwargs.get('semantics_mode'),
                instance_segmentation_mode=kwargs.get('instance_segmentation_mode'), image=kwargs.get('image', 'rgb'))
            stem = basemodel.conv_stem
            rgb_weights = stem.weight.detach().clone()
            keep_rgb = kwargs.get('image', 'rgb') != "none"
            if not keep_rgb and extra < 1:
                sys.exit("Too few input channels - add more inputs")
            # the reference hard-codes 32 stem filters here (:345); B5's 48 would not fit the next layer
            new_stem = Conv2dSame((3 if keep_rgb else 0) + extra, 32, kernel_size=(3, 3), stride=(2, 2), bias=False)
            if keep_rgb and rgb_weights.shape[0] == 32:
                with torch.no_grad():
                    new_stem.weight[:, 0:3] = rgb_weights
            basemodel.conv_stem = new_stem

        return cls(basemodel, n_bins=n_bins, encoder_name=encoder_name, insertion_point=insertion_point,
                   **kwargs).channels_last_()

    @staticmethod
    def get_num_channels_to_add(encoder_name, semantics_mode, instance_segmentation_mode, image):
        """Channel bookkeeping of the reference (:363-395): returns how many channels the external info adds."""
        n = 0
        if semantics_mode is not None:
            if "raw" in semantics_mode:
                n += 1
            elif semantics_mode == "glove":
                n += 300
            elif "glove-25d" in semantics_mode:
                n += 25
            elif "one-hot" in semantics_mode:
                n += 101  # (f)4 extension: one plane per ADE20K-places class (no reference implementation exists)
            else:
                sys.exit("Error [models/unet_adaptive_bins.py]: semantics mode not recognised")
            n += 10 * (("inst-areas" in semantics_mode) + ("human-sizes" in semantics_mode))
        if instance_segmentation_mode is not None:
            if instance_segmentation_mode == "raw":
                n += 1
            elif instance_segmentation_mode == "coco" or "ade20k_swin" in instance_segmentation_mode:
                n += 35  # 25 embedding channels + 10 from the area MLP
            if "human_sizes" in instance_segmentation_mode:
                n += 10
        return n
