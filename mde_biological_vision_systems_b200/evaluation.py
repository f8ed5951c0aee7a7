"""Evaluation epilogue and metrics on the device ("next" row (f)2 of SURVEY.md section 8).

Mirrors the reference's evaluate.py:50-71,128-152 / train.py:543-568 / utils.py:75-139 / infer.py:105-130: the
prediction is up-sampled to the ground truth's size, clipped, masked (valid depth range + Garg / Eigen crop) and reduced
to a1 a2 a3 abs_rel rmse log_10 rmse_log silog sq_rel per image -- in ONE kernel per batch (ops.eval_metrics) instead of
``.cpu().numpy()`` per image, plus the mirror test-time augmentation (ops.flip_average).
"""
import torch

from . import ops


def crop_box(height, width, garg_crop=False, eigen_crop=False, dataset="nyu"):
    """(y0, y1, x0, x1) of the evaluation crop (evaluate.py:136-148); the full frame when no crop flag is set."""
    if garg_crop:
        return (int(0.40810811 * height), int(0.99189189 * height), int(0.03594771 * width), int(0.96405229 * width))
    if eigen_crop:
        if dataset == 'kitti':
            return (int(0.3324324 * height), int(0.91351351 * height), int(0.0359477 * width), int(0.96405229 * width))
        return (45, 471, 41, 601)
    return (0, height, 0, width)


def compute_errors(gt, pred, args):
    """Batched, on-device counterpart of the reference loop body: gt [B,1,H,W], pred [B,1,h,w] (model output) ->
    list of B dicts with the reference's keys (images without a valid pixel give NaNs, as numpy's empty mean does)."""
    box = crop_box(gt.shape[-2], gt.shape[-1], getattr(args, "garg_crop", False), getattr(args, "eigen_crop", False),
                   getattr(args, "dataset", "nyu"))
    rows = ops.eval_metrics(pred, gt, args.min_depth_eval, args.max_depth_eval, box).cpu()
    return [dict(zip(ops.METRIC_KEYS, (float(v) for v in row[:9]))) for row in rows]


class RunningAverage:  # utils.py:60-72
    def __init__(self):
        self.avg = 0
        self.count = 0

    def append(self, value):
        self.avg = (value + self.count * self.avg) / (self.count + 1)
        self.count += 1

    def get_value(self):
        return self.avg


class RunningAverageDict:  # utils.py:75-89
    def __init__(self):
        self._dict = None

    def update(self, new_dict):
        if self._dict is None:
            self._dict = {key: RunningAverage() for key in new_dict}
        for key, value in new_dict.items():
            self._dict[key].append(value)

    def get_value(self):
        return {key: value.get_value() for key, value in self._dict.items()}


@torch.no_grad()
def predict_flip_tta(model, image, min_depth, max_depth, **model_kwargs):
    """infer.py:105-118: average of the prediction and the un-mirrored prediction of the mirrored image, both clipped
    to [min_depth, max_depth]; returns (bin_edges, averaged prediction at the model's output resolution).  External-info
    tensors in ``model_kwargs`` (semantics=, instance_labels=, instance_areas=) are mirrored with the image."""
    edges, pred = model(image, **model_kwargs)
    flipped = {k: (None if v is None else torch.flip(v, dims=[-1])) for k, v in model_kwargs.items()}
    _, pred_lr = model(torch.flip(image, dims=[-1]), **flipped)
    return edges, ops.flip_average(pred, pred_lr, min_depth, max_depth)
