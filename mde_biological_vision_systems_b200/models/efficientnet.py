"""Offline, random-init EfficientNet backbones shaped like geffnet's ``tf_efficientnet_b{1,5}_ap``.

The reference obtains its encoder with ``torch.hub.load('rwightman/gen-efficientnet-pytorch', ...)``
(/root/reference/models/unet_adaptive_bins.py:318-324), which needs the network and an un-pinned
third-party repository.  This module rebuilds the same *module tree* (child order ``conv_stem, bn1,
act1, blocks[0..6], conv_head, bn2, act2, global_pool, classifier`` and geffnet parameter names) so that

* ``Encoder`` (which walks ``_modules`` in order and keeps every intermediate, reference
  ``unet_adaptive_bins.py:103-116``) sees the feature list the decoder indexes ([4],[5],[6],[8],[11]);
* a real geffnet checkpoint can be dropped in with ``load_state_dict``.

The backbone is *outside* the hot path: it is plain PyTorch/cuDNN passthrough and both
sides of every parity test receive the same backbone object.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

# (block type, repeats, kernel, stride, expansion, out channels) of EfficientNet-B0; every stage has se 0.25
_B0_STAGES = (
    ("ds", 1, 3, 1, 1, 16),
    ("ir", 2, 3, 2, 6, 24),
    ("ir", 2, 5, 2, 6, 40),
    ("ir", 3, 3, 2, 6, 80),
    ("ir", 3, 5, 1, 6, 112),
    ("ir", 4, 5, 2, 6, 192),
    ("ir", 1, 3, 1, 6, 320),
)
_SCALING = {  # name -> (width multiplier, depth multiplier)
    "tf_efficientnet_b0_ap": (1.0, 1.0),
    "tf_efficientnet_b1_ap": (1.0, 1.1),
    "tf_efficientnet_b5_ap": (1.6, 2.2),
}
_BN_EPS = 1e-3  # TF-ported nets use eps 1e-3


def _round_channels(ch, mult, divisor=8):
    ch = ch * mult
    new = max(divisor, int(ch + divisor / 2) // divisor * divisor)
    if new < 0.9 * ch:
        new += divisor
    return new


def _same_pad_amount(size, k, s):
    return max((math.ceil(size / s) - 1) * s + k - size, 0)


class SamePadConv2d(nn.Conv2d):
    """TensorFlow 'SAME' convolution (asymmetric padding computed from the input size)."""

    def __init__(self, cin, cout, kernel_size, stride=1, groups=1, bias=False):
        super().__init__(cin, cout, kernel_size, stride, 0, 1, groups, bias)

    def forward(self, x):
        return self.forward_with(x, self.weight, self.bias)

    def same_pads(self, h, w):
        """(top, bottom, left, right) zeros TensorFlow's SAME rule adds for an h x w input."""
        ph = _same_pad_amount(h, self.kernel_size[0], self.stride[0])
        pw = _same_pad_amount(w, self.kernel_size[1], self.stride[1])
        return ph // 2, ph - ph // 2, pw // 2, pw - pw // 2

    def forward_with(self, x, weight, bias, prepadded=False):
        if not prepadded:
            pt, pb, pl, pr = self.same_pads(x.shape[-2], x.shape[-1])
            if pt or pb or pl or pr:
                x = F.pad(x, (pl, pr, pt, pb))
        return F.conv2d(x, weight, bias, self.stride, 0, 1, self.groups)


def _folded_conv_bn(conv, bn):
    """Eval-mode BatchNorm folded into the preceding convolution: w' = w * gamma / sqrt(var + eps) per output channel,
    b' = beta + (conv bias - mean) * gamma / sqrt(var + eps); cached on the conv module per parameter / buffer version."""
    tensors = (conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var)
    key = tuple(t._version for t in tensors) + (conv.weight.device, conv.weight.stride())
    cached = getattr(conv, "_mde_fold", None)
    if cached is None or cached[0] != key:
        with torch.no_grad():
            scale = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
            w = conv.weight * scale.view(-1, 1, 1, 1)
            b = bn.bias - bn.running_mean * scale
            if conv.bias is not None:
                b = b + conv.bias * scale
        cached = (key, w, b)
        conv._mde_fold = cached
    return cached[1], cached[2]


def _pointwise_exact(conv, x, act, residual, out_pads):
    from .. import ops
    return (not torch.backends.cudnn.allow_tf32 and conv.kernel_size == (1, 1) and conv.stride == (1, 1) and conv.groups == 1
            and conv.padding in ((0, 0), 0) and (act is None or isinstance(act, nn.SiLU))
            and (out_pads is None or not any(out_pads) or residual is None)
            and ops.pointwise_supported(x, conv.in_channels, conv.out_channels)
            and (residual is None or (residual.dtype == torch.float32 and residual.is_contiguous(memory_format=torch.channels_last))))


# "cudnn" (default): the depthwise convolutions stay on the library's NHWC fp32 kernel, followed by ONE bias + SiLU + pooling pass
# of ours (ops.bias_act_pool_nhwc_).  "fused": ops.depthwise_bias_act_pool does convolution + bias + SiLU + pooling in one pass.
# Measured on B200 at config 2 (23 layers): cudnn + pool 1.68 ms, fused 2.15 ms -- the one-output-per-thread mapping of the fused
# kernel issues ~240 instructions per output float4 and re-reads every input row k times through L2; until it computes a strip of
# outputs per thread (sharing taps) the library kernel is the faster body, so the fused form is opt-in (tested either way).
DEPTHWISE_IMPL = "cudnn"


def _depthwise_module_ok(conv, act):
    return (DEPTHWISE_IMPL == "fused" and conv.groups > 1 and conv.groups == conv.in_channels == conv.out_channels and (act is None or isinstance(act, nn.SiLU))
            and conv.kernel_size in ((3, 3), (5, 5)) and conv.stride in ((1, 1), (2, 2)) and conv.dilation == (1, 1)
            and conv.in_channels % 4 == 0
            and (isinstance(conv, SamePadConv2d) or (isinstance(conv.padding, tuple) and conv.padding_mode == "zeros")))


def _depthwise_ours(conv, x, act):
    from .. import ops
    return _depthwise_module_ok(conv, act) and ops.depthwise_supported(x, conv.in_channels, conv.kernel_size, conv.stride,
                                                                       conv.dilation)


def _folded_depthwise_kkc(conv, w_folded):
    """the BatchNorm-folded depthwise filter [C,1,k,k] as [k,k,C] (channel innermost), cached next to the fold"""
    key = conv._mde_fold[0]
    cached = getattr(conv, "_mde_fold_kkc", None)
    if cached is None or cached[0] != key:
        cached = (key, w_folded.detach()[:, 0].permute(1, 2, 0).contiguous().float())
        conv._mde_fold_kkc = cached
    return cached[1]


def _folded_pointwise_pair(conv, bn, w_folded):
    """split-bf16 pair of the BatchNorm-folded 1x1 filter, cached next to the fold (same version key)."""
    from .. import ops
    key = conv._mde_fold[0]
    cached = getattr(conv, "_mde_fold_pair", None)
    if cached is None or cached[0] != key:
        cached = (key, ops.prepare_pointwise_weight(w_folded))
        conv._mde_fold_pair = cached
    return cached[1]


class _depthwise_engine:
    """A depthwise convolution (groups == channels) has no channel reduction and cuDNN computes it with fp32 FMAs whatever
    the TF32 switch says -- but with the switch OFF (exact mode) the library falls back to an NCHW engine behind two layout
    conversions per call (0.55 ms + slower kernels per config-2 step).  For depthwise convs only, the switch is therefore left
    on, which selects the direct NHWC fp32 kernel; tests/test_gpu_parity.py::test_depthwise_engine_is_exact checks that the
    two settings give bit-identical outputs on this cuDNN."""

    def __init__(self, conv, x):
        self.on = bool(x.is_cuda and conv.groups > 1 and conv.groups == conv.in_channels and not torch.backends.cudnn.allow_tf32)

    def __enter__(self):
        if self.on:
            torch.backends.cudnn.allow_tf32 = True

    def __exit__(self, *exc):
        if self.on:
            torch.backends.cudnn.allow_tf32 = False
        return False


def conv_bn(conv, bn, x, act=None, residual=None, out_pads=None, prepadded=False, gate=None, pool=False):
    """act(bn(conv(x))) (+ residual); in inference (eval-mode statistics, no autograd) as ONE convolution with the folded
    filter -- the 69 per-block BatchNorm passes of EfficientNet-B1 are 3 ms of a 14 ms step otherwise -- followed, on
    channels_last CUDA tensors, by one fused bias + SiLU (+ residual) pass (ops.bias_act_nhwc_) instead of separate bias,
    activation and add kernels.  Same values up to fp32 rounding.
    ``out_pads`` (top, bottom, left, right): on the fused path the result is written zero-padded (the SAME padding of the
    stride-2 convolution that consumes it; callers detect it by the grown spatial size and pass ``prepadded`` on).
    ``gate`` [B, C_in]: the convolution's input is x * gate[:, :, None, None] (squeeze-excite), multiplied inside the GEMM
    where that kernel runs.  ``pool``: return (y, slab sums of y or None) -- the spatial mean the following squeeze-excite
    needs, taken from the bias/activation pass."""
    def ret(y, partial=None):
        return (y, partial) if pool else y

    if bn.training or torch.is_grad_enabled() or not isinstance(bn, nn.BatchNorm2d) or not bn.track_running_stats \
            or not bn.affine:
        if gate is not None:
            x = x * gate[:, :, None, None]
        y = bn(conv(x))
        y = act(y) if act is not None else y
        return ret(y + residual if residual is not None else y)
    from .. import ops
    w, b = _folded_conv_bn(conv, bn)
    if _pointwise_exact(conv, x, act, residual, out_pads):
        # exact mode (the caller switched the library's TF32 off: UnetAdaptiveBins inference): the 1x1 convolution, its folded
        # BatchNorm bias, SiLU, the squeeze-excite gate on its input and the residual run as ONE tcgen05 GEMM with fp32-grade
        # products (ops.pointwise_conv) instead of the library's legacy fp32 kernels + a bias/activation pass
        return ret(ops.pointwise_conv(x, _folded_pointwise_pair(conv, bn, w), b, 1 if act is not None else 0, residual,
                                      gate=gate, out_pads=out_pads))
    if gate is not None:
        x = x * gate[:, :, None, None]
    if (not torch.backends.cudnn.allow_tf32 and residual is None and out_pads is None and not pool and conv.groups == 1
            and conv.kernel_size == (3, 3) and conv.stride == (2, 2) and conv.dilation == (1, 1)
            and (act is None or isinstance(act, nn.SiLU)) and isinstance(conv, SamePadConv2d)
            and ops.stem_conv_supported(x, conv.in_channels, conv.out_channels)):
        # the stem in exact mode: the library's exact-fp32 NHWC engine takes 0.83 ms for this one layer at config 2; a direct fp32
        # kernel of ours (ops.stem_conv3x3s2: bias + SiLU in the epilogue, SAME padding by bounds) does it in a fraction
        h_in, w_in = x.shape[-2:]
        pt, pb, pl, pr = (0, 0, 0, 0) if prepadded else conv.same_pads(h_in, w_in)
        out_hw = ((h_in + pt + pb - 3) // 2 + 1, (w_in + pl + pr - 3) // 2 + 1)
        key = conv._mde_fold[0]
        cached = getattr(conv, "_mde_fold_tcc", None)
        if cached is None or cached[0] != key:
            cached = (key, ops.prepare_stem_weight(w))
            conv._mde_fold_tcc = cached
        return ops.stem_conv3x3s2(x, cached[1], b, 1 if act is not None else 0, pt, pl, out_hw)
    if pool and residual is None and gate is None and out_pads is None and _depthwise_ours(conv, x, act):
        # the depthwise convolution of an MBConv block: convolution + folded bias + SiLU + the squeeze-excite pooling in one
        # pass of our own fp32 kernel (ops.depthwise_bias_act_pool) instead of the library kernel, a bias/activation pass and a
        # mean pass over the same tensor
        k, s = conv.kernel_size[0], conv.stride[0]
        h_in, w_in = x.shape[-2:]
        if isinstance(conv, SamePadConv2d):
            pt, pb, pl, pr = (0, 0, 0, 0) if prepadded else conv.same_pads(h_in, w_in)
        else:
            pt = pb = conv.padding[0]
            pl = pr = conv.padding[1]
        out_hw = ((h_in + pt + pb - k) // s + 1, (w_in + pl + pr - k) // s + 1)
        return ops.depthwise_bias_act_pool(x, _folded_depthwise_kkc(conv, w), b, 1 if act is not None else 0, s, pt, pl, out_hw)
    fused = (act is None or isinstance(act, nn.SiLU)) and x.is_cuda
    with _depthwise_engine(conv, x):
        y = conv.forward_with(x, w, None if fused else b, prepadded) if isinstance(conv, SamePadConv2d) \
            else conv._conv_forward(x, w, None if fused else b)
    if fused and ops.bias_act_supported(y, residual):
        if out_pads is not None and residual is None and any(out_pads):
            return ret(ops.bias_act_pad_nhwc(y, b, 1 if act is not None else 0, out_pads))
        if pool and residual is None:
            return ops.bias_act_pool_nhwc_(y, b, 1 if act is not None else 0)
        return ret(ops.bias_act_nhwc_(y, b, 1 if act is not None else 0, residual))
    if fused:
        y = y + b.view(1, -1, 1, 1)
    y = act(y) if act is not None else y
    return ret(y + residual if residual is not None else y)


def _conv(cin, cout, k, stride=1, groups=1, bias=False):
    if stride == 1:  # symmetric padding is exact 'SAME' for odd kernels at stride 1
        return nn.Conv2d(cin, cout, k, 1, k // 2, groups=groups, bias=bias)
    return SamePadConv2d(cin, cout, k, stride, groups, bias)


class SqueezeExcite(nn.Module):
    def __init__(self, channels, reduced):
        super().__init__()
        self.conv_reduce = nn.Conv2d(channels, reduced, 1, bias=True)
        self.act1 = nn.SiLU(inplace=True)
        self.conv_expand = nn.Conv2d(reduced, channels, 1, bias=True)

    def forward(self, x):
        g = x.mean((2, 3), keepdim=True)
        g = self.conv_expand(self.act1(self.conv_reduce(g)))
        return x * torch.sigmoid(g)

    def gate_from(self, partial, hw):
        """[B, C] gate sigmoid(conv_expand(silu(conv_reduce(mean)))) from the slab sums the preceding bias/activation pass left
        (inference; the two 1x1 convolutions act on a [B, C, 1, 1] tensor, i.e. they are two tiny dense layers -- as library
        convolutions in exact fp32 they cost ~48 us each, 2.2 ms per config-2 step)."""
        from .. import ops
        r, e = self.conv_reduce, self.conv_expand
        return ops.se_gate(partial, hw, r.weight.flatten(1), r.bias, e.weight.flatten(1), e.bias)


def _se_project(se, conv, bn, y, partial, residual):
    """squeeze-excite + the 1x1 projection that follows it; with the slab sums at hand the gate goes into the GEMM"""
    if partial is not None:
        return conv_bn(conv, bn, y, None, residual, gate=se.gate_from(partial, y.shape[2] * y.shape[3]))
    return conv_bn(conv, bn, se(y), None, residual)


class DepthwiseSeparableConv(nn.Module):
    def __init__(self, cin, cout, k, stride, se_ratio=0.25):
        super().__init__()
        self.has_residual = stride == 1 and cin == cout
        self.conv_dw = _conv(cin, cin, k, stride, groups=cin)
        self.bn1 = nn.BatchNorm2d(cin, eps=_BN_EPS)
        self.act1 = nn.SiLU(inplace=True)
        self.se = SqueezeExcite(cin, max(1, int(cin * se_ratio)))
        self.conv_pw = nn.Conv2d(cin, cout, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout, eps=_BN_EPS)
        self.act2 = nn.Identity()

    def forward(self, x):
        y, partial = conv_bn(self.conv_dw, self.bn1, x, self.act1, pool=True)
        return _se_project(self.se, self.conv_pw, self.bn2, y, partial, x if self.has_residual else None)


class InvertedResidual(nn.Module):
    def __init__(self, cin, cout, k, stride, expand, se_ratio=0.25):
        super().__init__()
        mid = cin * expand
        self.has_residual = stride == 1 and cin == cout
        self.conv_pw = nn.Conv2d(cin, mid, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(mid, eps=_BN_EPS)
        self.act1 = nn.SiLU(inplace=True)
        self.conv_dw = _conv(mid, mid, k, stride, groups=mid)
        self.bn2 = nn.BatchNorm2d(mid, eps=_BN_EPS)
        self.act2 = nn.SiLU(inplace=True)
        self.se = SqueezeExcite(mid, max(1, int(cin * se_ratio)))
        self.conv_pwl = nn.Conv2d(mid, cout, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(cout, eps=_BN_EPS)

    def forward(self, x):
        # a stride-2 SAME depthwise conv follows the 1x1 expansion: let the expansion's epilogue write its output padded
        pads = self.conv_dw.same_pads(x.shape[-2], x.shape[-1]) if isinstance(self.conv_dw, SamePadConv2d) else None
        if (pads is not None and not torch.is_grad_enabled() and not self.bn2.training and _depthwise_module_ok(self.conv_dw, self.act2)
                and x.is_cuda and x.dtype == torch.float32 and x.is_contiguous(memory_format=torch.channels_last)):
            pads = None  # our depthwise kernel applies the SAME padding through its bounds test: nothing to pre-pad
        y = conv_bn(self.conv_pw, self.bn1, x, self.act1, out_pads=pads)
        y, partial = conv_bn(self.conv_dw, self.bn2, y, self.act2, prepadded=(y.shape[-2:] != x.shape[-2:]), pool=True)
        return _se_project(self.se, self.conv_pwl, self.bn3, y, partial, x if self.has_residual else None)


class GenEfficientNet(nn.Module):
    """Module tree in geffnet order; ``forward`` is only used stand-alone (the Encoder walks children)."""

    def __init__(self, width=1.0, depth=1.0, in_chans=3, num_classes=1000):
        super().__init__()
        stem = _round_channels(32, width)
        self.conv_stem = SamePadConv2d(in_chans, stem, 3, 2)
        self.bn1 = nn.BatchNorm2d(stem, eps=_BN_EPS)
        self.act1 = nn.SiLU(inplace=True)
        stages, cin = [], stem
        for kind, reps, k, stride, expand, ch in _B0_STAGES:
            cout = _round_channels(ch, width)
            blocks = []
            for r in range(int(math.ceil(reps * depth))):
                s = stride if r == 0 else 1
                if kind == "ds":
                    blocks.append(DepthwiseSeparableConv(cin, cout, k, s))
                else:
                    blocks.append(InvertedResidual(cin, cout, k, s, expand))
                cin = cout
            stages.append(nn.Sequential(*blocks))
        self.blocks = nn.Sequential(*stages)
        head = _round_channels(1280, width)
        self.conv_head = nn.Conv2d(cin, head, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(head, eps=_BN_EPS)
        self.act2 = nn.SiLU(inplace=True)
        self.global_pool = nn.AdaptiveAvgPool2d(1)
        self.classifier = nn.Linear(head, num_classes)
        self.num_features = head

    def forward(self, x):
        x = self.act1(self.bn1(self.conv_stem(x)))
        x = self.act2(self.bn2(self.conv_head(self.blocks(x))))
        return self.classifier(self.global_pool(x).flatten(1))


def build_backbone(name, seed=None):
    """Random-init stand-in for ``torch.hub.load('rwightman/gen-efficientnet-pytorch', name)``."""
    width, depth = _SCALING[name]
    if seed is None:
        return GenEfficientNet(width, depth)
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(seed)
        return GenEfficientNet(width, depth)
