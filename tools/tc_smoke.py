"""first contact with the tcgen05 kernels on the GPU box: tiny shapes, errors printed, run under `timeout`"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vn_pointcloudcompletion_b200 import _lib
def tf32(a): return (a.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
def dev(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()
st = torch.cuda.current_stream().cuda_stream
for (R, K, Cout) in [(256, 32, 128), (512, 128, 256), (768, 512, 2048)]:
    rng = np.random.RandomState(0)
    x = tf32(rng.standard_normal((R, K)).astype(np.float32)); w = tf32(rng.standard_normal((Cout, K)).astype(np.float32))
    xd, wd = dev(x), dev(w); y = torch.zeros((R, Cout), device="cuda")
    rc = _lib.raw("vnpcc_gemm_rows_tf32", xd.data_ptr(), K, wd.data_ptr(), K, y.data_ptr(), Cout, R, K, Cout, None, 0, 0, st)
    torch.cuda.synchronize()
    ref = x.astype(np.float64) @ w.astype(np.float64).T
    err = np.abs(y.cpu().numpy() - ref).max()
    print("rows", (R, K, Cout), "rc", rc, "max err", err, "ref max", np.abs(ref).max(), flush=True)
for (R, K, Cout) in [(256, 32, 128), (1024, 256, 128), (4096, 256, 256)]:
    rng = np.random.RandomState(1)
    x = tf32(rng.standard_normal((R, K)).astype(np.float32)); gy = tf32(rng.standard_normal((R, Cout)).astype(np.float32))
    xd, gd = dev(x), dev(gy); g = torch.zeros((Cout, K), device="cuda")
    rc = _lib.raw("vnpcc_gemm_wgrad_tf32", gd.data_ptr(), Cout, xd.data_ptr(), K, g.data_ptr(), K, R, Cout, K, None, 0, st)
    torch.cuda.synchronize()
    ref = gy.astype(np.float64).T @ x.astype(np.float64)
    err = np.abs(g.cpu().numpy() - ref).max()
    print("wgrad", (R, K, Cout), "rc", rc, "max err", err, "ref max", np.abs(ref).max(), flush=True)
