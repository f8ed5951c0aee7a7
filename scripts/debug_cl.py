import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import make_model
from mde_biological_vision_systems_b200 import synthetic, ops
from mde_biological_vision_systems_b200.loss import BinsChamferLoss, SILogLoss
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda:0"
kw = dict(insertion_point="input", semantics_mode=None, instance_segmentation_mode=None)
x = synthetic.image(2, 352, 384, seed=38).to(DEV)
depth = synthetic.depth(2, 352, 384, seed=39).to(DEV)

def run(cl, mode="full"):
    m = make_model(**kw).to(DEV)
    if cl:
        m.channels_last_()
    m.train()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    m.zero_grad(set_to_none=True)
    if mode == "decoder":
        feats = m.encoder(ops.to_channels_last(x) if cl else x)
        out = m.decoder(feats)
        loss = (out * out).mean()
    else:
        e, p = m(x)
        loss = SILogLoss()(p, depth, mask=depth > 1e-3) + 0.1 * BinsChamferLoss()(e, depth)
    loss.backward()
    return float(loss.detach()), {k: v.grad.detach().cpu().clone() for k, v in m.named_parameters() if v.grad is not None}

for mode in ("decoder", "full"):
    l1, g1 = run(False, mode)
    l1b, g1b = run(False, mode)
    l2, g2 = run(True, mode)
    print(mode, "loss", l1, l1b, l2)
    rows = []
    for k in g1:
        a, b, c = g1[k], g1b[k], g2[k]
        s = float(a.abs().max()) + 1e-30
        rows.append((float((a - c).abs().max()) / s, float((a - b).abs().max()) / s, k))
    rows = [r for r in rows if r[1] < 1e-2]
    rows.sort(reverse=True)
    for r in rows[:14]:
        print("  cl-vs-nchw %.3e   rerun %.3e   %s" % r)

# float64 truth for the decoder-only objective (stock resize + cat, everything in double)
import torch.nn.functional as F
from mde_biological_vision_systems_b200.models import unet_adaptive_bins as uab
orig_forward = uab.UpSampleBN.forward

def stock_forward(self, x, concat_with):
    x = F.interpolate(x, size=concat_with.shape[-2:], mode='bilinear', align_corners=True)
    return self._net(torch.cat((x, concat_with), dim=1))

uab.UpSampleBN.forward = stock_forward
m = make_model(**kw).to(DEV).double().train()
m.zero_grad(set_to_none=True)
out = m.decoder(m.encoder(x.double()))
loss = (out * out).mean()
loss.backward()
g64 = {k: v.grad.detach().float().cpu() for k, v in m.named_parameters() if v.grad is not None}
uab.UpSampleBN.forward = orig_forward
l1, g1 = run(False, "decoder")
l2, g2 = run(True, "decoder")
uab.UpSampleBN.forward = stock_forward
l3, g3 = run(False, "decoder")   # NCHW, stock resize+cat in fp32
print("fp64 loss", float(loss), "nchw", l1, "cl", l2, "nchw-stock", l3)
for k in ["decoder.up4._net.0.weight", "decoder.up2._net.0.weight", "decoder.up1._net.0.weight", "decoder.up3._net.0.weight",
          "decoder.conv2.weight", "encoder.original_model.blocks.0.0.conv_dw.weight", "encoder.original_model.conv_stem.weight"]:
    s = float(g64[k].abs().max())
    print("  %-55s nchw %.3e  cl %.3e  nchw-stock %.3e" % (k, float((g1[k] - g64[k]).abs().max()) / s,
          float((g2[k] - g64[k]).abs().max()) / s, float((g3[k] - g64[k]).abs().max()) / s))
