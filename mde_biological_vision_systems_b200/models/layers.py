"""Drop-in for the reference's models/layers.py (PatchTransformerEncoder, PixelWiseDotProduct).

Same constructor signatures, forward contracts and state_dict keys (``embedding_convPxP.*``,
``positional_encodings``, ``transformer_encoder.layers.{i}.*``) so reference checkpoints load unchanged
(/root/reference/models/layers.py:5-36, model_io.py:36-72).
"""
import torch
import torch.nn as nn

from .. import ops


class PatchTransformerEncoder(nn.Module):
    """conv k=patch s=patch -> + positional rows -> 4 post-LN encoder layers; returns [S, N, E]
    (reference layers.py:5-24).  The parameter containers are the stock torch modules so that key names, shapes
    and default initialisation are identical to the reference's."""

    def __init__(self, in_channels, patch_size=10, embedding_dim=128, num_heads=4):
        super().__init__()
        layer = nn.TransformerEncoderLayer(embedding_dim, num_heads, dim_feedforward=1024)
        self.transformer_encoder = nn.TransformerEncoder(layer, num_layers=4, enable_nested_tensor=False)
        self.embedding_convPxP = nn.Conv2d(in_channels, embedding_dim, kernel_size=patch_size, stride=patch_size,
                                           padding=0)
        self.positional_encodings = nn.Parameter(torch.rand(500, embedding_dim), requires_grad=True)
        self.use_tc_patch_embed = True  # False: cuDNN fp32 conv for the patch embedding (the transformer stays on our kernels)
        self.use_tc_layers = True       # False: exact-fp32 SIMT linear kernels instead of the 3xTF32 tcgen05 GEMMs

    def _needs_autograd(self, x):
        return torch.is_grad_enabled() and (getattr(x, "requires_grad", False) or any(p.requires_grad for p in self.parameters()))

    def _prepared_weight(self):
        """NHWC filter as a split-bf16 pair for the tcgen05 patch GEMM, cached per parameter version."""
        w = self.embedding_convPxP.weight
        cached = getattr(self, "_mde_w_prep", None)
        if cached is None or cached[0] != w._version or cached[1].device != w.device:
            cached = (w._version, ops.prepare_patch_weight(w))
            self._mde_w_prep = cached
        return cached[1]

    def forward(self, x):
        """x: fp32 [N,C,h,w] (either memory format) or the same feature map as an ops.SplitBF16 (inference only)."""
        conv = self.embedding_convPxP
        if not x.is_cuda:
            raise ops._lib.MdeError("PatchTransformerEncoder runs on the B200 kernels only (no CPU path)")
        if self._needs_autograd(x) or (self.training and self.transformer_encoder.layers[0].dropout.p > 0):
            # training: dropout and the backward pass run through the stock torch layers
            if isinstance(x, ops.SplitBF16):
                x = x.float()
            emb = conv(x).flatten(2)  # [N, E, S]
            emb = emb + self.positional_encodings[: emb.shape[2], :].T.unsqueeze(0)
            return self.transformer_encoder(emb.permute(2, 0, 1))
        if self.use_tc_patch_embed and ops.patch_embed_supported(x, conv):
            # TMA + tcgen05 split-K GEMM on split-bf16 pairs writes the [S, N, E] tokens (bias and positional rows added)
            tokens = ops.patch_embed(x, self._prepared_weight(), conv.bias, self.positional_encodings, conv.kernel_size[0])
        else:
            if isinstance(x, ops.SplitBF16):
                x = x.float()
            with ops.exact_fp32_library():
                emb = conv(x).flatten(2)
            emb = emb + self.positional_encodings[: emb.shape[2], :].T.unsqueeze(0)
            tokens = emb.permute(2, 0, 1).contiguous()
        layers = list(self.transformer_encoder.layers)
        if self.use_tc_layers and self._tc_layers_supported(layers):
            return ops.encoder_layers_tc(tokens, layers, self._prepared_layers(layers))
        ws = None
        for layer in layers:
            tokens, ws = ops.encoder_layer(tokens, layer, ws)
        return tokens

    @staticmethod
    def _tc_layers_supported(layers):
        l0 = layers[0]
        return (l0.self_attn.embed_dim == 128 and l0.self_attn.num_heads == 4 and l0.linear1.out_features % 4 == 0
                and not l0.norm_first and l0.self_attn.in_proj_weight is not None)

    def _prepared_layers(self, layers):
        """[w_hi | w_hi | w_lo] operands of the four linear maps of every layer, cached per parameter version."""
        params = [p for layer in layers for p in (layer.self_attn.in_proj_weight, layer.self_attn.out_proj.weight,
                                                  layer.linear1.weight, layer.linear2.weight)]
        key = tuple(p._version for p in params) + (params[0].device,)
        cached = getattr(self, "_mde_l3_prep", None)
        if cached is None or cached[0] != key:
            prep = [tuple(ops.prepare_linear3(p) for p in params[4 * i: 4 * i + 4]) for i in range(len(layers))]
            cached = (key, prep)
            self._mde_l3_prep = cached
        return cached[1]


class PixelWiseDotProduct(nn.Module):
    """y[n, cout, h, w] = sum_c x[n, c, h, w] * K[n, cout, c]  (reference layers.py:27-36) on the tcgen05 (three bf16
    products per K step) / SIMT contraction kernels (ops.range_attention)."""

    def __init__(self, impl="auto"):
        super().__init__()
        self.impl = impl

    def forward(self, x, K):
        n, c, h, w = x.size()
        _, cout, ck = K.size()
        assert c == ck, "Number of channels in x and Embedding dimension (at dim 2) of K matrix must match"
        return ops.range_attention(x, K, impl=self.impl)
