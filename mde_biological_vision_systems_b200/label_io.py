"""On-disk label formats -> batch tensors ("next" row (f)3).

Host-side readers for the files the reference's data loader opens (dataloader.py:98-161 train / :248-297 test), with the
same sentinel handling, returning COMPACT integer arrays (uint8 semantics, int32 instance labels / areas) instead of the
int64 tensors the reference builds: the loaders of this package gather on the GPU straight from these types
(ops.gather_embed, label_dtype MDE_U8 / MDE_I32), so a 416 x 544 label map crosses PCIe as 0.23 / 0.9 MB instead of
1.8 MB of int64 -- or the 22.6 MB of 25-channel floats the reference ships.

File formats (producers: semantic-segmentation-pytorch/test.py:30-32, Swin-Transformer-Object-Detection/tools/
nyud2_inference.py:104-128, misc_scripts/maskrcnn_inference_nyud2.py:189-199):
  semantic_seg_*.npy                       int array [H,W], ADE-150 argmax
  instance_labels_ade20k_swin_*.npz        'arr_0': int32 [H,W] class per pixel (-1 none) or a 0-d object array (None)
                                           when the detector produced nothing
  instance_areas_ade20k_swin(_bbox)_*.npz  'arr_0': int32 [H,W] pixel count of the instance covering the pixel (0 none)
  instance_{labels,areas}_coco_*.npy       same, plain .npy
"""
import numpy as np
import torch


def _arr0(path):
    z = np.load(path, allow_pickle=True)
    return z["arr_0"] if hasattr(z, "files") else z


def load_semantic_labels(path, mode, image_hw):
    """-> uint8 [H,W].  Mirrors dataloader.py:121-133: ADE-150 maps are cast with astype(np.ubyte); the ade20k-places
    modes read the Swin instance-label .npz, replace a missing prediction by -1 everywhere and cast to ubyte, so -1
    arrives as 255 (and is clamped to the background class by the loader, SemanticsLoader.py:115-118)."""
    if "ade20k-places" not in mode:
        return np.load(path).astype(np.ubyte)
    raw = _arr0(path)
    if raw.ndim != 2:
        raw = np.full(image_hw, -1, dtype=np.int32)
    return raw.astype(np.ubyte)


def load_instance_maps(labels_path, areas_path, mode, image_hw):
    """-> (labels int32 [H,W], areas int32 [H,W]).  Mirrors dataloader.py:136-150: a 0-d / non-2-D array means the
    detector found nothing: labels become -1, areas 0."""
    labels, areas = _arr0(labels_path), _arr0(areas_path)
    if "ade20k_swin" in mode:
        if labels.ndim != 2:
            labels = np.full(image_hw, -1, dtype=np.int32)
        if areas.ndim != 2:
            areas = np.zeros(image_hw, dtype=np.int32)
    return labels.astype(np.int32, copy=False), areas.astype(np.int32, copy=False)


def collate(samples, pin=True):
    """list of {'semantics': uint8 [H,W], 'instance_labels': int32 [H,W], 'instance_areas': int32 [H,W]} (any subset)
    -> batch dict of [B,1,H,W] tensors in pinned host memory, ready for SemanticsLoader / InstanceSegmentationLoader."""
    out = {}
    for key in ("semantics", "instance_labels", "instance_areas"):
        if samples and key in samples[0]:
            t = torch.from_numpy(np.stack([np.ascontiguousarray(s[key]) for s in samples])[:, None])
            out[key] = t.pin_memory() if pin and torch.cuda.is_available() else t
    return out
