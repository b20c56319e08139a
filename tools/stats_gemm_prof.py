"""development aid: a few launches of the decoder-shape rows GEMM (plain, then with the statistics epilogue) for ncu"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vn_pointcloudcompletion_b200 as V
from vn_pointcloudcompletion_b200 import ops
V.set_gemm_mode("tf32")
R, K, Cout, Cs = 1572864, 256, 512, 256
x = torch.randn(R, K, device="cuda")
w = torch.randn(Cout, K, device="cuda") / 16
y = torch.empty(R, Cout, device="cuda")
sums = torch.empty(2 * Cs, device="cuda", dtype=torch.float64)
for _ in range(2):
    ops.gemm_rows(x, w, out=y)
    ops.gemm_rows(x, w, out=y, stats=(sums, Cs))
torch.cuda.synchronize()
print("ok")
