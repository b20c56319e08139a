"""development aid: decoder-tail backward, fused (sums pre-pass + tail_dgrad_tf32_kernel) vs unfused (bwd1 -> bwd2 -> dgrad GEMM), per kernel"""
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


P, C, Cin = 32 * 16384, 256, 256
R = 3 * P
pd = torch.randn(R, 2 * C, device="cuda")
gy = torch.randn(R, device="cuda")
stat = torch.cat([torch.rand(C, device="cuda") + 1, torch.rand(C, device="cuda") + 0.5])
gamma, beta = torch.rand(C, device="cuda") + 0.5, torch.rand(C, device="cuda")
w2 = torch.randn(C, device="cuda") / 16
wcat = torch.randn(2 * C, Cin, device="cuda") / 16
wt = wcat.t().contiguous()
sums = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
gw2 = torch.zeros(C, device="cuda", dtype=torch.float64)
gpd = torch.empty(R, 2 * C, device="cuda")
gh = torch.empty(R, Cin, device="cuda")
gg, gb = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
st = _lib.stream()
h = torch.randn(R, Cin, device="cuda")
gW = torch.empty(2 * C, Cin, device="cuda")
t_fused = timeit(lambda: _lib.call("vnpcc_tail_bwd_tf32", gy, pd, 2 * C, P, C, stat, gamma, beta, 0.2, w2, wt, 2 * C, Cin, 1, sums, gw2, gpd, 2 * C, gh, Cin, None, 0, None, 0, st))
gh1, gpd1 = gh.clone(), gpd.clone()
_lib.raw("vnpcc_set_tuning", 7, 2)
t_pair = timeit(lambda: _lib.call("vnpcc_tail_bwd_tf32", gy, pd, 2 * C, P, C, stat, gamma, beta, 0.2, w2, wt, 2 * C, Cin, 1, sums, gw2, gpd, 2 * C, gh, Cin, None, 0, None, 0, st))
_lib.raw("vnpcc_set_tuning", 7, 0)
print(f"CTA pairs: pre-pass + tail_dgrad {t_pair:.3f} ms (one SM per tile {t_fused:.3f}); gh equal {torch.equal(gh, gh1)}, gpd equal {torch.equal(gpd, gpd1)}, "
      f"max diff gh {float((gh - gh1).abs().max()):.2e}", flush=True)
t_fused_w = timeit(lambda: _lib.call("vnpcc_tail_bwd_tf32", gy, pd, 2 * C, P, C, stat, gamma, beta, 0.2, w2, wt, 2 * C, Cin, 1, sums, gw2, None, 0, gh, Cin, h, Cin, gW, Cin, st))
t_wg = timeit(lambda: ops.gemm_wgrad(gpd, h, out=gW))
t_b1 = timeit(lambda: _lib.call("vnpcc_bn_leaky_dot_bwd1", gy, pd, 2 * C, pd[:, C:], 2 * C, gpd, 2 * C, gpd[:, C:], 2 * C, P, C, stat, gamma, beta, 0.2, sums, w2, gw2, st))
t_b2 = timeit(lambda: _lib.call("vnpcc_vn_bn_bwd2", gpd, 2 * C, pd, 2 * C, P, C, stat, gamma, beta, sums, float(P), 1, gg, gb, st))
gpd.normal_()
t_dg = timeit(lambda: ops.gemm_rows(gpd, wcat, True, out=gh))
print(f"fused (pre-pass + tail_dgrad writing gpd) {t_fused:.3f} ms + wgrad GEMM {t_wg:.3f} = {t_fused + t_wg:.3f} ms")
print(f"fused (pre-pass + tail_dgrad + tail_wgrad, gpd never stored) {t_fused_w:.3f} ms")
print(f"unfused: bwd1 {t_b1:.3f} + bwd2 {t_b2:.3f} + dgrad {t_dg:.3f} + wgrad {t_wg:.3f} = {t_b1 + t_b2 + t_dg + t_wg:.3f} ms")
