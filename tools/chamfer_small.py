"""small-cloud Chamfer forward: pre-filtered search (mode 2) vs exact packed search (mode 1), B = 32; CUDA events after warm-up"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vn_pointcloudcompletion_b200 as V
from vn_pointcloudcompletion_b200 import _lib


def timeit(fn, iters=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


g = torch.Generator(device="cuda").manual_seed(0)
print("| N | M | mode 2 (pre-filtered) ms | mode 1 (exact packed) ms |\n|---|---|---:|---:|")
for N, M in ((1024, 1024), (2048, 2048), (4096, 4096), (8192, 8192), (1024, 16384), (256, 256)):
    a = torch.rand(32, N, 3, device="cuda", generator=g) - 0.5
    b = torch.rand(32, M, 3, device="cuda", generator=g) - 0.5
    t = {}
    for mode in (2, 1):
        _lib.raw("vnpcc_chamfer_set_packed_math", mode)
        t[mode] = timeit(lambda: V.chamfer_3DFunction.apply(a, b))
    _lib.raw("vnpcc_chamfer_set_packed_math", 2)
    print(f"| {N} | {M} | {t[2]:.3f} | {t[1]:.3f} |")
