"""Drop-in for the reference's models/miniViT.py (mViT head), /root/reference/models/miniViT.py:7-45."""
import torch
import torch.nn as nn

from .. import ops
from .layers import PatchTransformerEncoder, PixelWiseDotProduct


def _channels_last_weight(conv):
    """NHWC copy of a conv weight, cached per parameter version (cuDNN otherwise re-lays-out the weights every call:
    2 x 126 us for the 16.8 MB patch-embedding filter)."""
    w = conv.weight
    cached = getattr(conv, "_mde_w_cl", None)
    if cached is None or cached[0] != w._version or cached[1].device != w.device:
        cached = (w._version, w.detach().contiguous(memory_format=torch.channels_last))
        conv._mde_w_cl = cached
    return cached[1]


class mViT(nn.Module):
    def __init__(self, in_channels, n_query_channels=128, patch_size=16, dim_out=256, embedding_dim=128, num_heads=4,
                 norm='linear'):
        super().__init__()
        self.norm = norm
        self.n_query_channels = n_query_channels
        self.patch_transformer = PatchTransformerEncoder(in_channels, patch_size, embedding_dim, num_heads)
        self.dot_product_layer = PixelWiseDotProduct()
        self.conv3x3 = nn.Conv2d(in_channels, embedding_dim, kernel_size=3, stride=1, padding=1)
        self.regressor = nn.Sequential(nn.Linear(embedding_dim, 256), nn.LeakyReLU(), nn.Linear(256, 256),
                                       nn.LeakyReLU(), nn.Linear(256, dim_out))
        # 3x3 conv of the inference path: "tc" = tcgen05 implicit GEMM on split-bf16 pairs (three bf16 products per K step,
        # fp32-grade), "cudnn" = library conv in true fp32, "auto" = "tc" whenever the shape allows
        self.conv3x3_impl = "auto"
        self.train_conv_impl = "cudnn"  # training: "cudnn" (default) or "tc" (ops.conv3x3_autograd), see UpSampleBN

    def _conv3x3_uses_tc(self, x, pair_out=False):
        return self.conv3x3_impl != "cudnn" and ops.conv3x3_supported(x, self.conv3x3.out_channels, pair_out)

    def _prepared_conv3x3(self):
        """[dx][dy][Cout][C] split-bf16 filter for the tcgen05 conv, cached per parameter version."""
        w = self.conv3x3.weight
        cached = getattr(self, "_mde_w3_prep", None)
        if cached is None or cached[0] != w._version or cached[1].device != w.device:
            cached = (w._version, ops.prepare_conv3x3_weight(w))
            self._mde_w3_prep = cached
        return cached[1]

    # -- pieces shared by the reference-shaped forward() and the fused path of UnetAdaptiveBins ------------------
    def tokens_and_features(self, x, bias_free=False, pair_out=False):
        """-> (tgt [S,N,E], conv3x3(x) [N,E,h,w]).  x: fp32 tensor or ops.SplitBF16.  With ``bias_free`` the 3x3 conv runs
        without its bias (the caller folds it into the fused chain, ops.fold_queries(feat_bias=...)), which saves a full pass
        over the feature map; with ``pair_out`` (inference, tcgen05 conv only) the features come back as an ops.SplitBF16,
        the operand format of the fused chain."""
        # the reference clones x first (miniViT.py:25); nothing below writes to x, so the 29 MB/img copy is skipped
        c = self.conv3x3
        needs_grad = torch.is_grad_enabled() and (getattr(x, "requires_grad", False) or c.weight.requires_grad)
        if not needs_grad and x.is_cuda and self._conv3x3_uses_tc(x, pair_out):
            x = ops.split_bf16(x)  # once, for both consumers
            tgt = self.patch_transformer(x)
            feat = ops.conv3x3_nhwc(x, self._prepared_conv3x3(), None, None if bias_free else c.bias, pair_out=pair_out,
                                    name="conv3x3_head")
            return tgt, feat
        tgt = self.patch_transformer(x)
        if isinstance(x, ops.SplitBF16):
            x = x.float()
        if needs_grad:
            if self.train_conv_impl == "tc" and not torch.is_autocast_enabled() and ops.conv3x3_train_supported(x, c):
                return tgt, ops.conv3x3_autograd(x, c.weight, None if bias_free else c.bias)  # fwd, dgrad, wgrad on our kernels
            return tgt, (torch.nn.functional.conv2d(x, c.weight, None, c.stride, c.padding) if bias_free else c(x))
        with ops.exact_fp32_library():
            if bias_free:
                return tgt, torch.nn.functional.conv2d(x, _channels_last_weight(c), None, c.stride, c.padding)
            return tgt, c(x)

    def bin_widths(self, tgt, min_val=None, max_val=None):
        """regressor + normalisation on token 0 (miniViT.py:35-45); with min/max also edges and centres.
        -> (widths_normed, edges, centers, y_raw)"""
        r = self.regressor
        lo = 0.0 if min_val is None else min_val
        hi = 1.0 if max_val is None else max_val
        if torch.is_grad_enabled() and (tgt.requires_grad or r[0].weight.requires_grad):
            # training: this [B,256]-sized piece goes through autograd (same arithmetic as the kernel)
            y = self.regressor(tgt[0])
            if self.norm == 'linear':
                wn = torch.relu(y) + 0.1
                wn = wn / wn.sum(dim=1, keepdim=True)
            elif self.norm == 'softmax':
                wn = torch.softmax(y, dim=1)
            else:
                wn = torch.sigmoid(y)
                wn = wn / wn.sum(dim=1, keepdim=True)
            widths = torch.nn.functional.pad((hi - lo) * wn, (1, 0), mode='constant', value=lo)
            edges = torch.cumsum(widths, dim=1)
            return wn, edges, 0.5 * (edges[:, :-1] + edges[:, 1:]), y
        return ops.regressor_bins(tgt[0], r[0].weight, r[0].bias, r[2].weight, r[2].bias, r[4].weight, r[4].bias,
                                  self.norm, lo, hi)

    def forward(self, x):
        """-> (bin_widths_normed [N, dim_out], range_attention_maps [N, n_query, h, w]) as the reference."""
        tgt, feat = self.tokens_and_features(x)
        queries = tgt[1:self.n_query_channels + 1, ...].permute(1, 0, 2)
        range_attention_maps = self.dot_product_layer(feat, queries)
        widths_normed = self.bin_widths(tgt)[0]
        return widths_normed, range_attention_maps
