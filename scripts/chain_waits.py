"""Where does each role of head_chain_kernel wait?  Prints mean cycles per tile per role (config 2 shapes)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mde_biological_vision_systems_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
B, h, w = 16, 208, 272
torch.manual_seed(0)
for nhwc in (True, False):
    x = torch.randn(B, 128, h, w, device="cuda")
    if nhwc:
        x = ops.to_channels_last(x)
    q = torch.randn(B, 128, 128, device="cuda") * 0.1
    wout = torch.randn(256, 128, device="cuda") * 0.09
    bias = torch.randn(256, device="cuda") * 0.05
    centers = torch.sort(torch.rand(B, 256, device="cuda") * 10, dim=1).values.contiguous()
    wf, biasf = ops.fold_queries(wout, bias, q)
    for _ in range(3):
        ops.head_chain(x, wf, biasf, centers)
    prof = torch.zeros(148, 8, dtype=torch.int64, device="cuda")
    lib.mde_tc_debug_profile(ctypes.c_void_p(prof.data_ptr()))
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    ops.head_chain(x, wf, biasf, centers)
    e.record()
    torch.cuda.synchronize()
    lib.mde_tc_debug_profile(None)
    tiles = B * h * w / 128 / 148
    p = prof.double().mean(0) / tiles
    names = ["prod wait-empty", "prod total", "mma wait-full", "mma wait-acc-empty", "mma wait-weights", "mma total",
             "epi wait-acc-full", "epi total"]
    print("layout", "NHWC" if nhwc else "NCHW", "kernel us", s.elapsed_time(e) * 1e3, "tiles/CTA", tiles)
    for n, v in zip(names, p.tolist()):
        print(f"   {n:20s} {v:9.0f} cycles/tile")
