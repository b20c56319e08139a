"""development aid: the few-row GEMMs of the per-sample heads (R = 96): split-K rows GEMM vs one-CTA-per-tile, tcgen05 wgrad vs the SIMT kernel"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vn_pointcloudcompletion_b200 as V
from vn_pointcloudcompletion_b200 import _lib, ops
from stream_bench import timed

V.set_gemm_mode("tf32")
st = _lib.stream()
for R, K, Cout in ((96, 1024, 1024), (96, 1024, 3072), (96, 1024, 512), (96, 512, 1024), (96, 2048, 1024)):
    x = torch.randn(R, K, device="cuda")
    w = torch.randn(Cout, K, device="cuda")
    gy = torch.randn(R, Cout, device="cuda")
    y = torch.empty(R, Cout, device="cuda")
    g = torch.empty(Cout, K, device="cuda")
    res = []
    for legacy in (1, 0):
        _lib.raw("vnpcc_set_tuning", 0, legacy)
        res.append(timed(lambda: _lib.call("vnpcc_gemm_rows_tf32", x, K, w, K, y, Cout, R, K, Cout, None, 0, 0, st), 50) * 1e3)
    _lib.raw("vnpcc_set_tuning", 0, 0)
    t_tc = timed(lambda: _lib.call("vnpcc_gemm_wgrad_tf32", gy, Cout, x, K, g, K, R, Cout, K, None, 0, st), 50) * 1e3
    t_simt = timed(lambda: _lib.call("vnpcc_gemm_wgrad_fp32", gy, Cout, x, K, g, K, R, Cout, K, 0, st), 50) * 1e3
    print(f"R={R} K={K} Cout={Cout}: rows GEMM one CTA per tile {res[0]:.1f} us, split-K {res[1]:.1f} us | wgrad tcgen05 {t_tc:.1f} us, SIMT {t_simt:.1f} us")
