"""CPU study behind DESIGN.md section 5: which tensor-core operand format keeps `pred` within north_star's 1e-3 of the
fp32 reference on EVERY pixel when the decoder convs, the patch embedding, the head conv and the fused chain all run on
the tensor cores?  Emulates the operand splits in torch CPU (products accumulated in fp32, as the tensor core does).

    python scripts/precision_study.py [--batch 2] [--hw 416 544]
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import adabins_oracle as oracle  # noqa: E402
from mde_biological_vision_systems_b200 import synthetic  # noqa: E402


def tf32_rna(x):
    i = x.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)


def tf32_trunc(x):
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def split(x, fmt, n):
    parts, r = [], x
    for _ in range(n):
        if fmt == "bf16":
            p = r.to(torch.bfloat16).float()
        elif fmt == "fp16":
            p = r.half().float()
        elif fmt == "tf32":
            p = tf32_rna(r)
        elif fmt == "tf32t":
            p = tf32_trunc(r)
        else:
            raise ValueError(fmt)
        parts.append(p)
        r = r - p
    return parts


class Scheme:
    """fmt + number of split planes + which (i, j) plane products are kept."""

    def __init__(self, name, fmt=None, planes=1, order=0):
        self.name, self.fmt, self.planes, self.order = name, fmt, planes, order

    def bilinear(self, f, a, b):
        if self.fmt is None:
            return f(a, b)
        pa, pb = split(a, self.fmt, self.planes), split(b, self.fmt, self.planes)
        out = None
        for i in range(self.planes):
            for j in range(self.planes):
                if i + j <= self.order:
                    t = f(pa[i], pb[j])
                    out = t if out is None else out + t
        return out


def run(scheme, model, sd, x, what):
    """what: set of {"decoder", "head"} pieces that use the scheme; the rest is exact fp32."""
    dec = scheme if "decoder" in what else Scheme("exact")
    hd = scheme if "head" in what else Scheme("exact")
    feats = oracle.encoder_features(model.encoder.original_model, x)
    s0, s1, s2, s3, bott = feats[4], feats[5], feats[6], feats[8], feats[11]
    p = "decoder."
    y = F.conv2d(bott, sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1)

    def cbl(y, pre, ci, bi):
        y = dec.bilinear(lambda a, b: F.conv2d(a, b, None, padding=1), y, sd[f"{pre}_net.{ci}.weight"]) \
            + sd[f"{pre}_net.{ci}.bias"][None, :, None, None]
        y = F.batch_norm(y, sd[f"{pre}_net.{bi}.running_mean"], sd[f"{pre}_net.{bi}.running_var"], sd[f"{pre}_net.{bi}.weight"],
                         sd[f"{pre}_net.{bi}.bias"], False, 0.1, 1e-5)
        return F.leaky_relu(y, 0.01)

    for name, skip in (("up1.", s3), ("up2.", s2), ("up3.", s1), ("up4.", s0)):
        y = F.interpolate(y, size=skip.shape[-2:], mode="bilinear", align_corners=True)
        y = torch.cat([y, skip], dim=1)
        y = cbl(y, p + name, 0, 1)
        y = cbl(y, p + name, 3, 4)
    unet = dec.bilinear(lambda a, b: F.conv2d(a, b, None, padding=1), y, sd[p + "conv3.weight"]) \
        + sd[p + "conv3.bias"][None, :, None, None]
    # head
    hp = "adaptive_bins_layer."
    pt = hp + "patch_transformer."
    emb = hd.bilinear(lambda a, b: F.conv2d(a, b, None, stride=16), unet, sd[pt + "embedding_convPxP.weight"]) \
        + sd[pt + "embedding_convPxP.bias"][None, :, None, None]
    emb = emb.flatten(2)
    emb = emb + sd[pt + "positional_encodings"][: emb.shape[2], :].T.unsqueeze(0)
    t = emb.permute(2, 0, 1)
    for i in range(4):
        t = oracle.encoder_layer(t, sd, f"{pt}transformer_encoder.layers.{i}.")
    feat = hd.bilinear(lambda a, b: F.conv2d(a, b, None, padding=1), unet, sd[hp + "conv3x3.weight"])  # bias folded below
    queries = t[1:129].permute(1, 0, 2)                                   # [B,128,E]
    wo = sd["conv_out.0.weight"].reshape(256, 128)
    wf = torch.matmul(wo.unsqueeze(0), queries)                           # [B,256,128] exact fp32 fold
    biasf = sd["conv_out.0.bias"][None] + torch.einsum("bjk,k->bj", wf, sd[hp + "conv3x3.bias"])
    b_, c_, h_, w_ = feat.shape
    fm = feat.permute(0, 2, 3, 1).reshape(b_, h_ * w_, c_)
    logits = hd.bilinear(lambda a, b: torch.matmul(a, b.transpose(1, 2)), fm, wf) + biasf[:, None, :]
    sm = torch.softmax(logits, dim=2)
    edges, centers = oracle.bins_from_widths(oracle.normalise_widths(oracle.regressor(t[0], sd)), 1e-3, 10.0)
    pred = (sm * centers[:, None, :]).sum(2).reshape(b_, 1, h_, w_)
    return edges, pred, logits


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--hw", type=int, nargs=2, default=[416, 544])
    args = ap.parse_args()
    from helpers import make_model
    torch.set_num_threads(os.cpu_count())
    model = make_model()
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    x = synthetic.image(args.batch, args.hw[0], args.hw[1], seed=0)
    with torch.no_grad():
        e_ref, p_ref = oracle.head(oracle.decoder_bn(oracle.encoder_features(model.encoder.original_model, x), sd), sd, 1e-3, 10.0)
        e0, p0, lg = run(Scheme("exact"), model, sd, x, set())
        print("restructured-exact vs oracle: pred max rel %.2e  edges %.2e   |logit| max %.1f  std %.2f" % (
            float(((p0 - p_ref).abs() / p_ref.abs()).max()), float(((e0 - e_ref).abs() / e_ref.abs()).max()),
            float(lg.abs().max()), float(lg.std())))
        schemes = [Scheme("tf32 x1 (RNA)", "tf32", 1, 0), Scheme("tf32 x1 (trunc)", "tf32t", 1, 0),
                   Scheme("bf16 x3", "bf16", 2, 1), Scheme("bf16 x6", "bf16", 3, 2),
                   Scheme("fp16 x3", "fp16", 2, 1), Scheme("tf32 x3", "tf32", 2, 1)]
        for sc in schemes:
            for what in ({"head"}, {"decoder", "head"}):
                e, p, _ = run(sc, model, sd, x, what)
                r = ((p - p_ref).abs() / p_ref.abs()).flatten()
                print("%-16s %-16s pred rel: max %.2e  p99.9 %.2e  mean %.2e | edges max %.2e" % (
                    sc.name, "+".join(sorted(what)), float(r.max()), float(r.kthvalue(int(r.numel() * 0.999)).values),
                    float(r.mean()), float(((e - e_ref).abs() / e_ref.abs()).max())))


if __name__ == "__main__":
    main()
