"""development aid: forward Chamfer time vs the work-item planner's per-item overhead constant (tuning knob 4)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import torch
import vn_pointcloudcompletion_b200 as V
from vn_pointcloudcompletion_b200 import _lib

g = torch.Generator(device="cuda").manual_seed(0)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for (N, M) in [(16384, 16384), (1024, 16384), (2048, 2048), (4096, 4096), (8192, 8192)]:
    a = torch.rand(32, N, 3, device="cuda", generator=g) - 0.5
    b = torch.rand(32, M, 3, device="cuda", generator=g) - 0.5
    res = []
    for ov in (0, 1):
        _lib.raw("vnpcc_set_tuning", 6, ov)
        plan = (ctypes.c_int * 4)()
        _lib.load().vnpcc_debug_chamfer_plan(32, N, M, plan)
        t = timeit(lambda: V.chamfer_3DFunction.apply(a, b))
        res.append(f"variant={ov}: {t:.3f} ms (splits {plan[1]} x {plan[2]}, {2 * 32 * N * M / t / 1e9:.0f} Gpairs/s)")
    print(f"N={N} M={M}: " + "; ".join(res), flush=True)
_lib.raw("vnpcc_set_tuning", 6, 0)
