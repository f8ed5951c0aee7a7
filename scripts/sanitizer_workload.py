"""Small-shape pass over every hand-written tensor-core / streaming kernel, for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python scripts/sanitizer_workload.py
Each kernel runs once on inputs small enough for the sanitizer's ~50x slowdown; results are checked for finiteness only
(parity is tests/test_gpu_parity.py's job)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from mde_biological_vision_systems_b200 import ops, synthetic
from mde_biological_vision_systems_b200.loss import DepthLosses
from helpers import make_model

dev = "cuda:0"
rng = np.random.default_rng(0)
t = lambda *s: torch.from_numpy(rng.standard_normal(s).astype(np.float32)).to(dev)
done = []

def ok(name, *tensors):
    torch.cuda.synchronize()
    for x in tensors:
        x = x.float() if isinstance(x, ops.SplitBF16) else x
        assert torch.isfinite(x).all(), name
    done.append(name)

x = t(1, 128, 16, 24)
xp = ops.split_bf16(x); ok("split_bf16", xp)
ok("split_bf16_nchw == nhwc", ops.split_bf16(x.contiguous(memory_format=torch.channels_last)))
w = ops.prepare_conv3x3_weight(t(128, 128, 3, 3) / 34)
ok("conv3x3_x3 f32", ops.conv3x3_nhwc(xp, w, None, t(128), slope=0.01))
ok("conv3x3_x3 pair", ops.conv3x3_nhwc(xp, w, None, None, pair_out=True))
ok("conv3x3 ragged N", ops.conv3x3_nhwc(ops.split_bf16(t(1, 40, 9, 11)), ops.prepare_conv3x3_weight(t(344, 40, 3, 3) / 19)))
ok("conv3x3_small", ops.conv3x3_small(t(1, 80, 9, 11).contiguous(memory_format=torch.channels_last), t(1, 80, 3, 3), t(1)))
ok("pointwise", ops.pointwise_conv(t(1, 40, 9, 11).contiguous(memory_format=torch.channels_last),
                                   ops.prepare_pointwise_weight(t(240, 40, 1, 1)), t(240), 1))
up = ops.upsample_concat_nhwc_pair(t(1, 8, 7, 9).contiguous(memory_format=torch.channels_last),
                                   t(1, 16, 14, 18).contiguous(memory_format=torch.channels_last)); ok("upsample pair", up)
m = make_model(insertion_point="input", semantics_mode=None, instance_segmentation_mode=None).to(dev)
feat = synthetic.decoder_features(1, 128, 176, 192, seed=1).to(dev)
with torch.no_grad():
    edges, pred = m._head(feat); ok("head (patch_embed, encoder layers, conv3x3, fold, chain)", edges, pred)
    m.fused_head = False
    e2, p2 = m._head(feat); ok("head un-fused (range attention simt, conv1x1, bins_pred)", p2)
    q = t(1, 128, 128)
    ok("range_attention tc", ops.range_attention(feat, q, impl="tc"))
depth = synthetic.depth(1, 352, 384, seed=2).to(dev)
pr, ed = pred.clone().requires_grad_(True), edges.clone().requires_grad_(True)
s, c = DepthLosses(1e-3)(pr, ed, depth); (s + 0.1 * c).backward(); ok("depth_losses fwd+bwd", s, c, pr.grad, ed.grad)
xg = t(1, 16, 9, 11).contiguous(memory_format=torch.channels_last).requires_grad_(True)
wg = (t(24, 16, 3, 3) / 12).requires_grad_(True)
y = ops.conv3x3_autograd(xg, wg, None); y.sum().backward(); ok("conv3x3_autograd (dgrad, wgrad)", y, xg.grad, wg.grad)
fq = t(1, 128, 48, 64).requires_grad_(True)
qq = (t(1, 128, 128) * 0.3).requires_grad_(True)
wo, bo = (t(256, 128, 1, 1) * 0.2).requires_grad_(True), t(256).requires_grad_(True)
cen = torch.cumsum(torch.rand(1, 256, device=dev) + 0.1, 1).requires_grad_(True)
pp = ops.head_chain_autograd(fq, qq, wo, bo, cen); pp.sum().backward(); ok("head_chain fwd_train + bwd + gemm_nt", pp, fq.grad, qq.grad)
lab, _ = synthetic.label_maps(1, 32, 48, seed=3)
tab = torch.from_numpy(np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data",
                       "ade20k_places_classes_glove_twitter_27b_25d_embeddings.npy"))).float().to(dev)
ok("gather planar", ops.gather_embed(lab.to(dev), tab, background=100))
buf, view = ops.gather_embed_nhwc(lab.to(dev), tab, 100, c_before=3, pads=(0, 1, 0, 1), image=t(1, 3, 32, 48)); ok("gather nhwc", buf)
print("sanitizer workload ok:", len(done), "groups:", "; ".join(done))
