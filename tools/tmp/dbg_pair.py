import os, sys
sys.path.insert(0, "/root/repo")
import torch
import vn_pointcloudcompletion_b200 as V
from vn_pointcloudcompletion_b200 import _lib, ops
V.set_gemm_mode("tf32")
variant = int(sys.argv[1])
for (R, K, Cout) in [(15000, 64, 256), (20000, 256, 512)]:
    x = torch.randn(R, K, device="cuda"); w = torch.randn(Cout, K, device="cuda")
    y0 = ops.gemm_rows(x, w); torch.cuda.synchronize()
    _lib.raw("vnpcc_set_tuning", 2, 4); _lib.raw("vnpcc_set_tuning", 3, variant)
    y1 = ops.gemm_rows(x, w); torch.cuda.synchronize(); print("variant", variant, "pair ok", torch.equal(y0, y1), flush=True)
    _lib.raw("vnpcc_set_tuning", 2, 0); _lib.raw("vnpcc_set_tuning", 3, 0)
