"""Golden vectors for the evaluation metrics ("next" row (f)2): runs the REFERENCE's own utils.compute_errors.

utils.py imports matplotlib (absent here), so the function's source is cut out of /root/reference/utils.py with ast and
executed with numpy only -- the arithmetic is the reference's, unmodified.  The masking / clipping epilogue around it is
inline code in evaluate.py:59-71,128-150 (not a callable), restated by oracle.eval_epilogue.

    python tests/golden/make_golden_eval.py        (build container only; writes tests/golden/golden_eval.npz)
"""
import ast
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import adabins_oracle as oracle  # noqa: E402
from mde_biological_vision_systems_b200 import synthetic  # noqa: E402

KEYS = ["a1", "a2", "a3", "abs_rel", "rmse", "log_10", "rmse_log", "silog", "sq_rel"]
CASES = {  # name -> (B, h, w, H, W, min_eval, max_eval, garg, eigen, dataset)
    "nyu_eigen": (2, 240, 320, 480, 640, 1e-3, 10.0, False, True, "nyu"),
    "kitti_garg": (2, 176, 608, 352, 1216, 1e-3, 80.0, True, False, "kitti"),
    "kitti_eigen": (1, 176, 608, 352, 1216, 1e-3, 80.0, False, True, "kitti"),
    "nocrop_fullres": (2, 96, 128, 96, 128, 1e-3, 10.0, False, False, "nyu"),
}


def case_inputs(name):
    b, h, w, H, W, lo, hi, garg, eigen, ds = CASES[name]
    seed = 700 + sorted(CASES).index(name)
    rng = np.random.default_rng(seed)
    gt = synthetic.depth(b, H, W, seed=seed + 50) * (hi / 10.0)
    pred = torch.from_numpy((0.2 + 1.1 * hi * rng.random((b, 1, h, w), dtype=np.float32)).astype(np.float32))
    # out-of-range values exercise the clipping; non-finite ones are tested separately on the GPU (ATen's CPU and CUDA
    # bilinear kernels propagate NaN / inf differently, and the reference interpolates on the device)
    pred[0, 0, 3, 5] = 1e6
    pred[-1, 0, 11, 2] = -1.0
    return pred, gt


def main():
    src = open("/root/reference/utils.py").read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "compute_errors")
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "utils.py", "exec"), ns)
    ref_compute_errors = ns["compute_errors"]
    out = {}
    for name, (b, h, w, H, W, lo, hi, garg, eigen, ds) in CASES.items():
        pred, gt = case_inputs(name)
        rows = []
        for i in range(b):
            g, p = oracle.eval_epilogue(pred[i:i + 1], gt[i:i + 1], lo, hi, garg, eigen, ds)
            m = ref_compute_errors(g, p)
            rows.append([float(m[k]) for k in KEYS] + [float(g.size)])
        out[name] = np.asarray(rows, dtype=np.float64)
        print(name, out[name])
    np.savez_compressed(os.path.join(HERE, "golden_eval.npz"), **out)


if __name__ == "__main__":
    main()
