"""Seeded synthetic inputs and weights (numpy Generator streams, so they reproduce on any box).

Follows SURVEY.md section 8(d): ImageNet-normalised RGB, metric depth with ~15 % invalid (zero) holes,
piecewise-constant label maps with injected out-of-range values, instance areas = true pixel counts.
Used by tests/, bench.py and tests/golden/make_golden.py so every side sees identical tensors.
"""
import numpy as np
import torch

_MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32).reshape(1, 3, 1, 1)
_STD = np.array([0.229, 0.224, 0.225], dtype=np.float32).reshape(1, 3, 1, 1)


def image(b, h, w, seed=0):
    """dataloader.py:530,535 contract: fp32 [B,3,H,W], ImageNet-normalised."""
    rng = np.random.default_rng(seed)
    x = rng.random((b, 3, h, w), dtype=np.float32)
    return torch.from_numpy((x - _MEAN) / _STD)


def depth(b, h, w, seed=1, hole_frac=0.15, max_depth=10.0, all_valid=False):
    """fp32 [B,1,H,W] metres, 0 = invalid (dataloader.py:209); rectangular holes cover ~hole_frac."""
    rng = np.random.default_rng(seed)
    d = (0.5 + (max_depth - 0.5) * rng.random((b, 1, h, w), dtype=np.float32)).astype(np.float32)
    if not all_valid:
        for i in range(b):
            covered = 0
            while covered < hole_frac * h * w:
                hh = int(rng.integers(max(2, h // 16), max(3, h // 4)))
                ww = int(rng.integers(max(2, w // 16), max(3, w // 4)))
                y0 = int(rng.integers(0, h - hh + 1))
                x0 = int(rng.integers(0, w - ww + 1))
                d[i, 0, y0:y0 + hh, x0:x0 + ww] = 0.0
                covered += hh * ww
    return torch.from_numpy(d)


def label_maps(b, h, w, seed=2, lo=-1, hi=100, inject=(-7, 101, 255, 1000), n_rect=(20, 60)):
    """int64 labels [B,1,H,W] made of random rectangles with values in [lo, hi] plus injected
    out-of-range values, and int64 instance areas = pixel count of the (last-painted) rectangle a pixel
    belongs to, 0 for background (Swin tools/nyud2_inference.py:111-124 semantics)."""
    rng = np.random.default_rng(seed)
    lab = np.full((b, 1, h, w), lo if lo < 0 else hi, dtype=np.int64)
    owner = np.zeros((b, h, w), dtype=np.int64)  # rectangle id per pixel, 0 = background
    for i in range(b):
        n = int(rng.integers(n_rect[0], n_rect[1] + 1))
        for r in range(1, n + 1):
            hh = min(h, int(rng.integers(max(2, h // 20), max(3, h // 3))))
            ww = min(w, int(rng.integers(max(2, w // 20), max(3, w // 3))))
            y0 = int(rng.integers(0, h - hh + 1))
            x0 = int(rng.integers(0, w - ww + 1))
            if inject and rng.random() < 0.1:
                v = int(inject[int(rng.integers(0, len(inject)))])
            else:
                v = int(rng.integers(lo, hi + 1))
            lab[i, 0, y0:y0 + hh, x0:x0 + ww] = v
            owner[i, y0:y0 + hh, x0:x0 + ww] = r
    areas = np.zeros((b, 1, h, w), dtype=np.int64)
    for i in range(b):
        counts = np.bincount(owner[i].ravel())
        counts[0] = 0
        areas[i, 0] = counts[owner[i]]
    return torch.from_numpy(lab), torch.from_numpy(areas)


def decoder_features(b, c, h, w, seed=3, scale=0.5):
    """Stand-in for ``unet_out`` [B,128,h,w] when the head is exercised without the backbone."""
    rng = np.random.default_rng(seed)
    return torch.from_numpy((scale * rng.standard_normal((b, c, h, w), dtype=np.float32)).astype(np.float32))


def fill_state_dict(module, seed=0):
    """Overwrite every parameter/buffer of ``module`` in place from a numpy stream (sorted key order), with
    torch-default-like scales: weights U(-1/sqrt(fan_in), +), biases U(-1/sqrt(fan_in_of_weight)) approximated
    by U(-0.05, 0.05), norm weights 1 + 0.1 N(0,1), running_var in [0.5, 1.5], positional rows U(0,1)."""
    rng = np.random.default_rng(seed)
    sd = module.state_dict()
    with torch.no_grad():
        for name in sorted(sd.keys()):
            t = sd[name]
            if not torch.is_floating_point(t):
                continue
            shape = tuple(t.shape)
            if name.endswith("positional_encodings"):
                v = rng.random(shape, dtype=np.float32)
            elif name.endswith("running_var"):
                v = 0.5 + rng.random(shape, dtype=np.float32)
            elif name.endswith("running_mean"):
                v = 0.1 * rng.standard_normal(shape, dtype=np.float32)
            elif t.dim() >= 2:
                fan_in = int(np.prod(shape[1:]))
                a = 1.0 / np.sqrt(fan_in)
                v = (2 * rng.random(shape, dtype=np.float32) - 1) * a
            elif "norm" in name and name.endswith("weight") or (".bn" in name and name.endswith("weight")) \
                    or ("_net.1.weight" in name) or ("_net.4.weight" in name):
                v = 1.0 + 0.1 * rng.standard_normal(shape, dtype=np.float32)
            else:
                v = (2 * rng.random(shape, dtype=np.float32) - 1) * 0.05
            t.copy_(torch.from_numpy(np.asarray(v, dtype=np.float32)).to(t.dtype))
    return module
