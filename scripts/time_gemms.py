"""Times the tcgen05 NT GEMM at the transformer's shapes (CUDA events, 20 reps after warm-up)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mde_biological_vision_systems_b200 import _lib, ops
lib = _lib.load()
dev = "cuda:0"
M = 3536
def run(name, K3, N, splits, act, split_out, ldc):
    a = torch.randn(M, K3, device=dev); b = torch.randn(N, K3, device=dev); bias = torch.randn(N, device=dev)
    c = torch.zeros(M, ldc, device=dev)
    def go():
        return lib.mde_gemm_nt_tf32_ex(ops._p(a), K3, 0, ops._p(b), K3, 0, ops._p(c), ldc, 0, 1, M, N, K3, splits, 1.0, ops._p(bias), act, split_out, ops._s())
    for _ in range(5): assert go() == 0
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): go()
    e.record(); torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5): go()
        torch.cuda.synchronize()
    dev_us = [ev.device_time_total / ev.count for ev in prof.key_averages() if "gemm_nt" in ev.key]
    print(f"{name:28s} K'={K3:5d} N={N:5d} splits={splits} : {s.elapsed_time(e) / 20 * 1e3:7.1f} us back-to-back, kernel {dev_us[0] if dev_us else -1:6.1f} us")
run("qkv", 384, 384, 1, 0, 0, 384)
run("out_proj", 384, 128, 1, 0, 0, 128)
run("ffn1 (relu, split out)", 384, 1024, 1, 1, 1, 3072)
run("ffn1 (plain out)", 384, 1024, 1, 1, 0, 1024)
for sp in (1, 2, 4, 8):
    run("ffn2", 3072, 128, sp, 0, 0, 128)
for tn in ("256", "128", "64"):
    os.environ["MDE_GEMM_TN"] = tn
    run("ffn1 split out, tn=" + tn, 384, 1024, 1, 1, 1, 3072)
    run("ffn2 tn=" + tn, 3072, 128, 1, 0, 0, 128)
    run("qkv tn=" + tn, 384, 384, 1, 0, 0, 384)
os.environ.pop("MDE_GEMM_TN")
print("-- fixed-cost probes")
M = 128
run("1 CTA, 1 chunk", 32, 32, 1, 0, 0, 32)
run("1 CTA, 12 chunks", 384, 32, 1, 0, 0, 32)
M = 3536
run("28 CTAs, 1 chunk", 32, 32, 1, 0, 0, 32)
run("28 CTAs, 12 chunks", 384, 32, 1, 0, 0, 32)
run("112 CTAs, 12 chunks", 384, 128, 1, 0, 0, 128)
import time
a = torch.randn(M, 384, device=dev); b = torch.randn(128, 384, device=dev); c = torch.zeros(M, 128, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200):
    lib.mde_gemm_nt_tf32_ex(ops._p(a), 384, 0, ops._p(b), 384, 0, ops._p(c), 128, 0, 1, M, 128, 384, 1, 1.0, None, 0, 0, ops._s())
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host time per call {(t1 - t0) / 200 * 1e6:.1f} us, incl. drain {(t2 - t0) / 200 * 1e6:.1f} us")
x = torch.randn(1 << 20, device=dev)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); s.record()
for _ in range(200): x.add_(1.0)
e.record(); torch.cuda.synchronize()
print(f"reference: trivial torch kernel back-to-back {s.elapsed_time(e) / 200 * 1e3:.1f} us each")
