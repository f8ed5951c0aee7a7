"""Head-only companion of precision_study.py: the golden head case (synthetic decoder features, logits of magnitude ~10)
and scaled-up variants, per operand scheme.  python scripts/precision_study_head.py"""
import os, sys
import torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from oracle import adabins_oracle as oracle
from mde_biological_vision_systems_b200 import synthetic
from mde_biological_vision_systems_b200.models.miniViT import mViT
from precision_study import Scheme

torch.set_num_threads(os.cpu_count())
head = synthetic.fill_state_dict(mViT(128, n_query_channels=128, patch_size=16, dim_out=256, embedding_dim=128), 7).eval()
conv_out = synthetic.fill_state_dict(torch.nn.Conv2d(128, 256, 1), 8)
sd = {"adaptive_bins_layer." + k: v.detach() for k, v in head.state_dict().items()}
sd.update({"conv_out.0.weight": conv_out.weight.detach(), "conv_out.0.bias": conv_out.bias.detach()})
hp = "adaptive_bins_layer."; pt = hp + "patch_transformer."
for scale in (0.5, 1.5, 4.0):
    x = synthetic.decoder_features(2, 128, 208, 272, seed=3, scale=scale)
    with torch.no_grad():
        e_ref, p_ref = oracle.head(x.double(), {k: v.double() for k, v in sd.items()}, 1e-3, 10.0)
        e32, p32 = oracle.head(x, sd, 1e-3, 10.0)
        r = ((p32.double() - p_ref).abs() / p_ref.abs()).flatten()
        print("scale %.1f: fp32 oracle vs fp64: pred max %.2e" % (scale, float(r.max())))
        for sc in [Scheme("tf32 x1 (RNA)", "tf32", 1, 0), Scheme("tf32 x1 (trunc)", "tf32t", 1, 0), Scheme("bf16 x3", "bf16", 2, 1),
                   Scheme("bf16 x6", "bf16", 3, 2), Scheme("fp16 x3", "fp16", 2, 1), Scheme("tf32 x3", "tf32", 2, 1)]:
            emb = sc.bilinear(lambda a, b: F.conv2d(a, b, None, stride=16), x, sd[pt + "embedding_convPxP.weight"]) + sd[pt + "embedding_convPxP.bias"][None, :, None, None]
            emb = emb.flatten(2); emb = emb + sd[pt + "positional_encodings"][: emb.shape[2], :].T.unsqueeze(0)
            t = emb.permute(2, 0, 1)
            for i in range(4):
                t = oracle.encoder_layer(t, sd, f"{pt}transformer_encoder.layers.{i}.")
            feat = sc.bilinear(lambda a, b: F.conv2d(a, b, None, padding=1), x, sd[hp + "conv3x3.weight"])
            q = t[1:129].permute(1, 0, 2)
            wf = torch.matmul(sd["conv_out.0.weight"].reshape(256, 128).unsqueeze(0), q)
            biasf = sd["conv_out.0.bias"][None] + torch.einsum("bjk,k->bj", wf, sd[hp + "conv3x3.bias"])
            b_, c_, h_, w_ = feat.shape
            fm = feat.permute(0, 2, 3, 1).reshape(b_, h_ * w_, c_)
            logits = sc.bilinear(lambda a, b: torch.matmul(a, b.transpose(1, 2)), fm, wf) + biasf[:, None, :]
            edges, centers = oracle.bins_from_widths(oracle.normalise_widths(oracle.regressor(t[0], sd)), 1e-3, 10.0)
            pred = (torch.softmax(logits, 2) * centers[:, None, :]).sum(2).reshape(b_, 1, h_, w_)
            r = ((pred.double() - p_ref).abs() / p_ref.abs()).flatten()
            print("  %-16s |logit| max %.1f std %.2f  pred rel: max %.2e p99.9 %.2e mean %.2e | edges %.2e" % (
                sc.name, float(logits.abs().max()), float(logits.std()), float(r.max()),
                float(r.kthvalue(int(r.numel() * 0.999)).values), float(r.mean()), float(((edges.double() - e_ref).abs() / e_ref.abs()).max())))
