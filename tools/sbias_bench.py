"""development aid: BatchNorm backward pass 2 with / without the fused per-sample bias sums, next to the separate rows_sample_sum pass"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vn_pointcloudcompletion_b200 import _lib, ops


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


B, N, C = 32, 2048, 1024
P = B * N
R = 3 * P
pd = torch.randn(R, 2 * C, device="cuda")
gpd = torch.randn(R, 2 * C, device="cuda")
stat = torch.cat([torch.rand(C, device="cuda") + 1, torch.rand(C, device="cuda") + 0.5])
gamma, beta = torch.rand(C, device="cuda") + 0.5, torch.rand(C, device="cuda")
sums = torch.randn(2 * C, device="cuda", dtype=torch.float64) * 100
gg, gbb = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
gb = torch.empty(B * 3, 2 * C, device="cuda")
st = _lib.stream()
t0 = timeit(lambda: _lib.call("vnpcc_vn_bn_bwd2", gpd, 2 * C, pd, 2 * C, P, C, stat, gamma, beta, sums, float(P), 1, gg, gbb, st))
t1 = timeit(lambda: _lib.call("vnpcc_vn_bn_bwd2_sbias", gpd, 2 * C, pd, 2 * C, P, C, stat, gamma, beta, sums, float(P), 1, gg, gbb, gpd[:, C:], 2 * C,
                              gb, 2 * C, N, st))
t2 = timeit(lambda: ops.rows_sample_sum(gpd, B, N))
print(f"bwd2 {t0:.3f} ms, bwd2+sbias {t1:.3f} ms, rows_sample_sum {t2:.3f} ms  ->  {t0 + t2:.3f} vs {t1:.3f}")
# correctness of the fused sums (fresh gradient: the timing loops above transformed gpd in place many times)
gpd = torch.randn(R, 2 * C, device="cuda")
g0 = gpd.clone()
_lib.call("vnpcc_vn_bn_bwd2", g0, 2 * C, pd, 2 * C, P, C, stat, gamma, beta, sums, float(P), 1, gg, gbb, st)
ref = ops.rows_sample_sum(g0, B, N)
g1 = gpd.clone()
_lib.call("vnpcc_vn_bn_bwd2_sbias", g1, 2 * C, pd, 2 * C, P, C, stat, gamma, beta, sums, float(P), 1, gg, gbb, g1[:, C:], 2 * C, gb, 2 * C, N, st)
print("gp equal", torch.equal(g0, g1), "gb rel err", float((gb - ref).norm() / ref.norm()))
