"""development aid: rows GEMM with the statistics epilogue vs the plain kernel (+ the separate statistics pass) at the two training shapes"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vn_pointcloudcompletion_b200 as V
from vn_pointcloudcompletion_b200 import _lib, ops

V.set_gemm_mode("tf32")


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


for (R, K, Cout, Cs, nb) in [(1572864, 256, 512, 256, 0), (196608, 512, 2048, 1024, 32), (196608, 512, 2048, 1024, 0)]:
    x = torch.randn(R, K, device="cuda")
    w = torch.randn(Cout, K, device="cuda") / K ** 0.5
    bias = torch.randn(nb * 3, Cout, device="cuda") if nb else None
    rps = R // nb if nb else 0
    y = torch.empty(R, Cout, device="cuda")
    sums = torch.empty(2 * Cs, device="cuda", dtype=torch.float64)
    t_plain = timeit(lambda: ops.gemm_rows(x, w, False, bias, rps, out=y))
    t_stats = {}
    for var in (1, 2, 3):
        for nomath in (0, 1):
            _lib.raw("vnpcc_set_tuning", 2, var)
            _lib.raw("vnpcc_set_tuning", 3, nomath)
            t_stats[(var, nomath)] = round(timeit(lambda: ops.gemm_rows(x, w, False, bias, rps, out=y, stats=(sums, Cs))), 3)
    _lib.raw("vnpcc_set_tuning", 2, 0)
    _lib.raw("vnpcc_set_tuning", 3, 0)
    t_pass = timeit(lambda: _lib.call("vnpcc_vn_norm_stats", y, Cout, R // 3, Cs, sums, _lib.stream()))
    gb = 4.0 * (R * K + R * Cout) / 1e9
    print(f"R={R} K={K} Cout={Cout} bias={nb > 0}: plain {t_plain:.3f} ms ({gb / t_plain:.0f} GB/s, {2e-9 * R * K * Cout / t_plain:.0f} TF/s)  "
          f"stats-epilogue (variant 1 STG/4st, 2 STG/3st, 3 TMA/3st; nomath) {t_stats} ms  separate pass {t_pass:.3f} ms", flush=True)
