"""development aid: the fused VNLinear -> VNMaxPool GEMM (gemm_vn_fused_kernel, MODE_POOL) at the encoder shape, one SM per tile vs CTA pairs"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vn_pointcloudcompletion_b200 as V
from vn_pointcloudcompletion_b200 import _lib, ops
from stream_bench import timed

V.set_gemm_mode("tf32")
G, N, K, C = 32, 2048, 1024, 2048
x = torch.randn(G * N * 3, K, device="cuda")
w = torch.randn(C, K, device="cuda") / K ** 0.5
wdir = torch.randn(C, C, device="cuda") / C ** 0.5
res = {}
for knob in (1, 4):      # knob 2: 1 = one SM per tile everywhere, 4 = CTA pairs also for the fused kernels
    _lib.raw("vnpcc_set_tuning", 2, knob)
    with torch.no_grad():
        fn = lambda: ops.linear_maxpool_rows(x, w, wdir, G, N)
        out, idx = fn()
        t = timed(fn, 10)
    res[knob] = (t, out.clone(), idx.clone())
_lib.raw("vnpcc_set_tuning", 2, 0)
fl = 2.0 * G * N * 3 * K * 2 * C
print(f"one SM per tile {res[1][0]:.3f} ms ({fl / res[1][0] / 1e9:.0f} TF/s incl. the small kernels around it), CTA pairs {res[4][0]:.3f} ms "
      f"({fl / res[4][0] / 1e9:.0f} TF/s); same selections: {torch.equal(res[4][2], res[1][2])}, same pooled rows: {torch.equal(res[4][1], res[1][1])}")
