"""CPU ORACLE -- test infrastructure, NOT product code.

A plain restatement (numpy / torch-CPU tensor algebra, no nn.Module, no CUDA) of the reference's
AdaBins head + loss + external-info path, the hot path named by BASELINE.json:north_star.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this file; the product package ``mde_biological_vision_systems_b200`` never does.

Every function cites the reference lines it follows (paths relative to /root/reference).

Pinning status
--------------
* Everything except chamfer is pinned against outputs of the *reference modules themselves*, imported
  from /root/reference and run in the build container by ``tests/golden/make_golden.py`` (the vectors
  live in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them).
* ``chamfer_distance`` lives in pytorch3d 0.6.1 (environment.yml:104), which is not vendored and not
  installable here.  It is restated from the published algorithm (pytorch3d/loss/chamfer.py @ v0.6.1:
  K=1 nearest neighbour in squared L2, both directions, padded points masked by ``y_lengths``,
  point_reduction="mean", batch_reduction="mean").  The reference's own ``BinsChamferLoss`` wrapper
  (loss.py:33-46: centres, ``>= 1e-3`` target mask, pad_sequence) IS exercised through the reference
  code with this restatement injected as ``pytorch3d.loss.chamfer_distance`` -- so for chamfer:
  **parity unpinned below the pytorch3d call boundary**.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# A1 / A2: external-info loaders (ExternalInfoLoaders/SemanticsLoader.py, InstanceSegmentationLoader.py)
# ----------------------------------------------------------------------------------------------


def clamp_labels(labels, rows, background):
    """SemanticsLoader.py:115-118 (>100 -> 100, <0 -> 100) and InstanceSegmentationLoader.py:103-104
    (<0 -> bg, >rows-1 -> bg).  Both reduce to: anything outside [0, rows-1] becomes ``background``."""
    lab = np.asarray(labels).copy()
    lab[(lab < 0) | (lab > rows - 1)] = background
    return lab


def gather_rows(table, labels):
    """``table.index_select(0, raw.view(-1)).view(B,1,H,W,D).squeeze(1).permute(0,3,1,2).contiguous()``
    (SemanticsLoader.py:125-127, InstanceSegmentationLoader.py:107-109).  labels [B,1,H,W] int64."""
    table = np.asarray(table)
    lab = np.asarray(labels)
    if lab.min() < 0 or lab.max() >= table.shape[0]:
        raise IndexError("label out of range for table (reference index_select raises)")
    out = table[lab[:, 0]]  # [B,H,W,D]
    return np.ascontiguousarray(np.transpose(out, (0, 3, 1, 2)))


def class_area_fraction(labels):
    """SemanticsLoader.get_semantics_inst_areas (SemanticsLoader.py:88-99): per image, every pixel gets
    (#pixels of its class in that image) / (H*W) as float64."""
    lab = np.asarray(labels)
    b, _, h, w = lab.shape
    out = np.empty(lab.shape, dtype=np.float64)
    total = h * w
    for im in range(b):
        vals, inv, counts = np.unique(lab[im], return_inverse=True, return_counts=True)
        out[im] = (counts[inv.reshape(lab[im].shape)] / total)
    return out


def semantics_loader(mode, labels, table, sizes_table=None):
    """SemanticsLoader.get_semantics (SemanticsLoader.py:102-145).  Returns (raw_clamped, semantics).
    ``mode`` is the reference's --use_semantics string; tables are the float64 .npy contents."""
    raw = np.asarray(labels).copy()
    places = "ade20k-places" in mode
    if places:
        raw = clamp_labels(raw, 101, 100)  # :115-118
    if "raw" in mode:
        sem = raw.astype(np.float32)  # :121-122
    else:
        sem = gather_rows(table, raw)  # float64, :125-127
        if places:
            sem = sem.astype(np.float32)  # :128-129
    if "inst-areas" in mode:
        sem = np.concatenate((sem, class_area_fraction(raw)), axis=1)  # :134-136 (promotes to float64)
    if sizes_table is not None:
        sz = gather_rows(sizes_table, raw).astype(np.float32)  # :139-142
        sem = np.concatenate((sem, sz), axis=1)  # :143
    return raw, sem


def instance_loader(mode, labels, areas, table, background, sizes_table=None):
    """InstanceSegmentationLoader.get_instance_segmentation (InstanceSegmentationLoader.py:89-121).
    Returns (raw_clamped, embedding float64 [B,25,H,W], areas float32 [B,1 or 4,H,W])."""
    raw = clamp_labels(labels, np.asarray(table).shape[0], background)  # :103-104
    emb = gather_rows(table, raw)  # :107-109 (stays float64)
    ar = np.asarray(areas).astype(np.float32)  # :112
    if sizes_table is not None:
        sz = gather_rows(sizes_table, raw).astype(np.float32)  # :115-118
        ar = np.concatenate((ar, sz), axis=1)  # :119
    return raw, emb, ar


# ----------------------------------------------------------------------------------------------
# A3: input insertion (models/unet_adaptive_bins.py:194-235) incl. the 1x1-conv MLPs (:144-174)
# ----------------------------------------------------------------------------------------------


def aux_mlp(x, sd, prefix):
    """nn.Sequential(Conv2d(c,10,1), ReLU, Conv2d(10,10,1), ReLU) -- unet_adaptive_bins.py:146-174."""
    x = F.relu(F.conv2d(x, sd[prefix + ".0.weight"], sd[prefix + ".0.bias"]))
    return F.relu(F.conv2d(x, sd[prefix + ".2.weight"], sd[prefix + ".2.bias"]))


def input_insertion(sd, x, semantics_mode=None, instance_mode=None, image="rgb",
                    semantics=None, instance_labels=None, instance_areas=None):
    """UnetAdaptiveBins.forward, insertion_point == "input" branch (unet_adaptive_bins.py:194-235)."""
    if semantics is not None:
        if semantics_mode == "glove-25d-inst-areas":  # :196-202
            x = torch.cat((x, semantics[:, 0:25].float()), 1)
            x = torch.cat((x, aux_mlp(semantics[:, 25:26].float(), sd, "semantics_areas_fc")), 1)
        elif "human-sizes" in semantics_mode:  # :203-209
            x = torch.cat((x, semantics[:, 0:-3].float()), 1)
            x = torch.cat((x, aux_mlp(semantics[:, -3:].float(), sd, "semantics_absolute_sizes_fc")), 1)
        else:  # :210-211
            x = torch.cat((x, semantics.float()), 1)
    if instance_labels is not None:  # :212-213
        x = torch.cat((x, instance_labels.float()), 1)
    if instance_areas is not None:
        hw = x.shape[2] * x.shape[3]
        if "human_sizes" in instance_mode:  # :215-224
            x = torch.cat((x, aux_mlp(instance_areas[:, 0:1] / hw, sd, "instance_areas_fc")), 1)
            x = torch.cat((x, aux_mlp(instance_areas[:, 1:4], sd, "instance_absolute_sizes_fc")), 1)
        else:  # :225-228
            x = torch.cat((x, aux_mlp(instance_areas / hw, sd, "instance_areas_fc")), 1)
    if image == "none":  # :231-235
        x = x[:, 3:]
    return x


# ----------------------------------------------------------------------------------------------
# A4: PatchTransformerEncoder (models/layers.py:5-24) -- nn.TransformerEncoderLayer defaults restated
# ----------------------------------------------------------------------------------------------


def encoder_layer(x, sd, p, nhead=4, eps=1e-5):
    """One post-LN nn.TransformerEncoderLayer(d_model=128, nhead=4, dim_feedforward=1024), eval mode
    (layers.py:8; torch defaults: ReLU, norm_first=False, layer_norm_eps=1e-5).  x: [S,B,E]."""
    s, b, e = x.shape
    hd = e // nhead
    qkv = F.linear(x, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"])
    q, k, v = qkv.split(e, dim=-1)

    def heads(t):  # [S,B,E] -> [B,H,S,hd]
        return t.reshape(s, b, nhead, hd).permute(1, 2, 0, 3)

    q, k, v = heads(q) / math.sqrt(hd), heads(k), heads(v)
    att = torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v  # [B,H,S,hd]
    att = att.permute(2, 0, 1, 3).reshape(s, b, e)
    att = F.linear(att, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
    x = F.layer_norm(x + att, (e,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps)
    ff = F.linear(F.relu(F.linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"])),
                  sd[p + "linear2.weight"], sd[p + "linear2.bias"])
    return F.layer_norm(x + ff, (e,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps)


def patch_transformer(x, sd, p="adaptive_bins_layer.patch_transformer.", patch=16, layers=4):
    """PatchTransformerEncoder.forward (layers.py:16-24): conv k16 s16, + positional rows, 4 layers."""
    emb = F.conv2d(x, sd[p + "embedding_convPxP.weight"], sd[p + "embedding_convPxP.bias"], stride=patch)
    emb = emb.flatten(2)  # [B,E,S]
    emb = emb + sd[p + "positional_encodings"][: emb.shape[2], :].T.unsqueeze(0)
    t = emb.permute(2, 0, 1)  # [S,B,E]
    for i in range(layers):
        t = encoder_layer(t, sd, f"{p}transformer_encoder.layers.{i}.")
    return t


# ----------------------------------------------------------------------------------------------
# A5-A8: mViT head, range attention, conv_out + softmax, bin pipeline
# ----------------------------------------------------------------------------------------------


def pixelwise_dot(x, queries):
    """PixelWiseDotProduct.forward (layers.py:31-36): y[b,n,h,w] = sum_k x[b,k,h,w] * K[b,n,k]."""
    return torch.einsum("bkhw,bnk->bnhw", x, queries)


def regressor(t0, sd, p="adaptive_bins_layer.regressor."):
    """Linear-LeakyReLU-Linear-LeakyReLU-Linear (miniViT.py:17-21), slope 0.01."""
    y = F.leaky_relu(F.linear(t0, sd[p + "0.weight"], sd[p + "0.bias"]), 0.01)
    y = F.leaky_relu(F.linear(y, sd[p + "2.weight"], sd[p + "2.bias"]), 0.01)
    return F.linear(y, sd[p + "4.weight"], sd[p + "4.bias"])


def normalise_widths(y, norm="linear"):
    """miniViT.py:36-44."""
    if norm == "linear":
        y = torch.relu(y) + 0.1
    elif norm == "softmax":
        return torch.softmax(y, dim=1)
    else:
        y = torch.sigmoid(y)
    return y / y.sum(dim=1, keepdim=True)


def mvit(x, sd, norm="linear", n_query=128):
    """mViT.forward (miniViT.py:23-45) -> (bin_widths_normed [B,n_bins], range_attention_maps)."""
    p = "adaptive_bins_layer."
    tgt = patch_transformer(x, sd)
    xc = F.conv2d(x, sd[p + "conv3x3.weight"], sd[p + "conv3x3.bias"], padding=1)
    queries = tgt[1:n_query + 1].permute(1, 0, 2)
    ram = pixelwise_dot(xc, queries)
    return normalise_widths(regressor(tgt[0], sd), norm), ram


def bins_from_widths(widths_normed, min_val, max_val):
    """unet_adaptive_bins.py:292-296: scale, left-pad with min_val, cumsum -> edges; midpoints."""
    widths = (max_val - min_val) * widths_normed
    widths = F.pad(widths, (1, 0), mode="constant", value=min_val)
    edges = torch.cumsum(widths, dim=1)
    centers = 0.5 * (edges[:, :-1] + edges[:, 1:])
    return edges, centers


def head(unet_out, sd, min_val, max_val, norm="linear"):
    """unet_adaptive_bins.py:285-302: mViT -> conv_out(1x1)+Softmax(dim=1) -> bins -> centre-weighted sum.
    Returns (bin_edges [B,n_bins+1], pred [B,1,h,w])."""
    widths_normed, ram = mvit(unet_out, sd, norm)
    out = torch.softmax(F.conv2d(ram, sd["conv_out.0.weight"], sd["conv_out.0.bias"]), dim=1)
    edges, centers = bins_from_widths(widths_normed, min_val, max_val)
    pred = torch.sum(out * centers[:, :, None, None], dim=1, keepdim=True)
    return edges, pred


def softmax_bins_pred(logits, centers):
    """The streaming slice of the above: softmax over dim 1 then centre-weighted sum (:286,:300)."""
    return torch.sum(torch.softmax(logits, dim=1) * centers[:, :, None, None], dim=1, keepdim=True)


def noadabins_epilogue(unet_out):
    """unet_adaptive_bins.py:240-242."""
    return F.relu(unet_out) + 0.0001


# ----------------------------------------------------------------------------------------------
# A9 / A10: losses (loss.py)
# ----------------------------------------------------------------------------------------------


def silog(pred, target, mask=None, interpolate=True):
    """SILogLoss.forward (loss.py:12-25)."""
    if interpolate:
        pred = F.interpolate(pred, target.shape[-2:], mode="bilinear", align_corners=True)
    if mask is not None:
        pred, target = pred[mask], target[mask]
    g = torch.log(pred) - torch.log(target)
    dg = torch.var(g) + 0.15 * torch.pow(torch.mean(g), 2)
    return 10 * torch.sqrt(dg)


def chamfer_distance(x, y, y_lengths):
    """pytorch3d.loss.chamfer_distance @ v0.6.1 with the reference's arguments (loss.py:45): x [N,P1,D]
    (all P1 valid), y [N,P2,D] zero-padded, y_lengths [N]; squared-L2 K=1 NN both ways; padded y never
    matches and contributes 0; per-cloud mean then batch mean; returns (cham_x + cham_y, None)."""
    n, p1, _ = x.shape
    cham_x = x.new_zeros(n)
    cham_y = x.new_zeros(n)
    for i in range(n):
        li = int(y_lengths[i])
        yi = y[i, :li]
        if li > 0:
            # brute force in chunks to bound memory; (a-b)^2 summed over D, as knn_points does
            dx = torch.full((p1,), float("inf"), dtype=x.dtype)
            dy = torch.empty(li, dtype=x.dtype)
            for s in range(0, li, 32768):
                d = ((x[i][:, None, :] - yi[None, s:s + 32768, :]) ** 2).sum(-1)  # [P1, chunk]
                dx = torch.minimum(dx, d.min(dim=1).values)
                dy[s:s + 32768] = d.min(dim=0).values
            cham_x[i] = dx.sum() / p1
            cham_y[i] = dy.sum() / li
        else:
            cham_x[i] = 0.0
            cham_y[i] = float("nan")  # 0 / 0 in the reference (sum of no points / length 0)
    return cham_x.sum() / n + cham_y.sum() / n, None


def bins_chamfer(bins, target_depth_maps):
    """BinsChamferLoss.forward (loss.py:33-46)."""
    centers = 0.5 * (bins[:, 1:] + bins[:, :-1])
    n, p = centers.shape
    pts = target_depth_maps.flatten(1)
    keep = pts.ge(1e-3)
    lists = [t[m] for t, m in zip(pts, keep)]
    lengths = torch.tensor([len(t) for t in lists], dtype=torch.long)
    padded = torch.nn.utils.rnn.pad_sequence(lists, batch_first=True).unsqueeze(2)
    loss, _ = chamfer_distance(centers.view(n, p, 1), padded, lengths)
    return loss


# ----------------------------------------------------------------------------------------------
# "next" row (f)2: evaluation metrics (utils.py:119-139)
# ----------------------------------------------------------------------------------------------


def compute_errors(gt, pred):
    gt = np.asarray(gt, dtype=np.float64)
    pred = np.asarray(pred, dtype=np.float64)
    thresh = np.maximum(gt / pred, pred / gt)
    err = np.log(pred) - np.log(gt)
    return dict(
        a1=(thresh < 1.25).mean(), a2=(thresh < 1.25 ** 2).mean(), a3=(thresh < 1.25 ** 3).mean(),
        abs_rel=np.mean(np.abs(gt - pred) / gt), rmse=np.sqrt(((gt - pred) ** 2).mean()),
        log_10=np.abs(np.log10(gt) - np.log10(pred)).mean(),
        rmse_log=np.sqrt(((np.log(gt) - np.log(pred)) ** 2).mean()),
        silog=np.sqrt(np.mean(err ** 2) - np.mean(err) ** 2) * 100,
        sq_rel=np.mean(((gt - pred) ** 2) / gt))


def eval_crop_box(h, w, garg_crop=False, eigen_crop=False, dataset="nyu"):
    """(y0, y1, x0, x1) of evaluate.py:136-148 / train.py:552-564; the full frame without a crop flag."""
    if garg_crop:
        return int(0.40810811 * h), int(0.99189189 * h), int(0.03594771 * w), int(0.96405229 * w)
    if eigen_crop:
        if dataset == "kitti":
            return int(0.3324324 * h), int(0.91351351 * h), int(0.0359477 * w), int(0.96405229 * w)
        return 45, 471, 41, 601
    return 0, h, 0, w


def eval_epilogue(pred, gt, min_depth_eval, max_depth_eval, garg_crop=False, eigen_crop=False, dataset="nyu"):
    """evaluate.py:59-71 + :128-150 for ONE image: pred [1,1,h,w] tensor, gt [1,1,H,W] tensor ->
    (gt[valid], pred[valid]) float32 vectors ready for compute_errors.  (When a crop flag is set the reference ANDs the
    validity mask with the crop box; without one it would AND with an undefined eval_mask -- the full frame is used.)"""
    p = torch.nn.functional.interpolate(pred, gt.shape[-2:], mode="bilinear", align_corners=True)
    return eval_mask_and_clip(p, gt, min_depth_eval, max_depth_eval, garg_crop, eigen_crop, dataset)


def eval_mask_and_clip(p, gt, min_depth_eval, max_depth_eval, garg_crop=False, eigen_crop=False, dataset="nyu"):
    """The part of eval_epilogue after the up-sampling (p already has gt's size)."""
    p = p.squeeze().cpu().numpy().copy()
    p[p < min_depth_eval] = min_depth_eval
    p[p > max_depth_eval] = max_depth_eval
    p[np.isinf(p)] = max_depth_eval
    p[np.isnan(p)] = min_depth_eval
    g = gt.squeeze().cpu().numpy()
    valid = np.logical_and(g > min_depth_eval, g < max_depth_eval)
    y0, y1, x0, x1 = eval_crop_box(g.shape[0], g.shape[1], garg_crop, eigen_crop, dataset)
    box = np.zeros(valid.shape, dtype=bool)
    box[y0:y1, x0:x1] = True
    valid = np.logical_and(valid, box)
    return g[valid], p[valid]


def flip_tta(model_fn, image, min_depth, max_depth):
    """infer.py:108-118: average of the clipped prediction and the clipped, un-mirrored prediction of the mirrored image
    (both at the model's output resolution)."""
    pred = np.clip(model_fn(image).cpu().numpy(), min_depth, max_depth)
    pred_lr = np.clip(model_fn(torch.flip(image, dims=[-1])).cpu().numpy()[..., ::-1], min_depth, max_depth)
    return 0.5 * (pred + pred_lr)


# ----------------------------------------------------------------------------------------------
# Encoder walk and DecoderBN (models/unet_adaptive_bins.py:39-116), functional, eval-mode BatchNorm.  The backbone itself
# is third-party (geffnet): any torch module with the geffnet child order is walked as the reference's Encoder does.
# ----------------------------------------------------------------------------------------------


def encoder_features(backbone, x):
    """Encoder.forward (:108-116): every child's output (the seven 'blocks' stages individually) appended to a list."""
    feats = [x]
    for name, child in backbone._modules.items():
        stages = child._modules.values() if name == "blocks" else (child,)
        for stage in stages:
            feats.append(stage(feats[-1]))
    return feats


def _conv_bn_lrelu(x, sd, p, conv_i, bn_i):
    x = F.conv2d(x, sd[f"{p}_net.{conv_i}.weight"], sd[f"{p}_net.{conv_i}.bias"], padding=1)
    x = F.batch_norm(x, sd[f"{p}_net.{bn_i}.running_mean"], sd[f"{p}_net.{bn_i}.running_var"], sd[f"{p}_net.{bn_i}.weight"],
                     sd[f"{p}_net.{bn_i}.bias"], False, 0.1, 1e-5)
    return F.leaky_relu(x, 0.01)


def decoder_bn(features, sd, p="decoder."):
    """DecoderBN.forward (:89-100) with UpSampleBN (:51-54): conv2 (1x1, padding 1) -> 4 x [bilinear(align_corners) to
    the skip's size, cat, 2 x conv3x3-BN-LeakyReLU] -> conv3."""
    s0, s1, s2, s3, bottleneck = features[4], features[5], features[6], features[8], features[11]
    y = F.conv2d(bottleneck, sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1)
    for name, skip in (("up1.", s3), ("up2.", s2), ("up3.", s1), ("up4.", s0)):
        y = F.interpolate(y, size=skip.shape[-2:], mode="bilinear", align_corners=True)
        y = torch.cat([y, skip], dim=1)
        y = _conv_bn_lrelu(y, sd, p + name, 0, 1)
        y = _conv_bn_lrelu(y, sd, p + name, 3, 4)
    return F.conv2d(y, sd[p + "conv3.weight"], sd[p + "conv3.bias"], padding=1)


# ----------------------------------------------------------------------------------------------
# Whole-path helper used by the CPU baseline: forward + SILog + chamfer given a backbone callable
# ----------------------------------------------------------------------------------------------


def forward_and_losses(backbone_decoder, sd, x, depth, min_val, max_val, min_depth=1e-3, norm="linear"):
    """train.py:405-425 restated for the AdaBins variants: unet_out = decoder(encoder(x)); head; losses."""
    unet_out = backbone_decoder(x)
    edges, pred = head(unet_out, sd, min_val, max_val, norm)
    mask = depth > min_depth
    return edges, pred, silog(pred, depth, mask=mask, interpolate=True), bins_chamfer(edges, depth)
