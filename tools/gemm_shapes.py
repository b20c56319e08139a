"""per-shape timing of the TF32 tcgen05 GEMMs of the train step (B=32): rows GEMM (fwd / dgrad) and weight gradient.
Prints TFLOP/s and the HBM GB/s implied by reading X once and writing Y once."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vn_pointcloudcompletion_b200 as V
from vn_pointcloudcompletion_b200 import ops

def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

V.set_gemm_mode("tf32")
Re, Rd = 196608, 1572864
shapes = [("enc first_conv[1] fwd", Re, 128, 512), ("enc maxpool1 dir", Re, 512, 512), ("enc second_conv[0] fwd (stacked, K_eff)", Re, 512, 2048),
          ("enc second_conv[0] dgrad", Re, 2048, 512), ("enc first_conv[1] dgrad", Re, 512, 128),
          ("dec final_conv[1] fwd (stacked)", Rd, 256, 512), ("dec final_conv[1] dgrad", Rd, 512, 256)]
print("| GEMM | R | K | Cout | ms | TFLOP/s | GB/s (X read + Y write) |")
print("|---|---:|---:|---:|---:|---:|---:|")
for name, R, K, Co in shapes:
    x = torch.randn(R, K, device="cuda"); w = torch.randn(Co, K, device="cuda") * 0.05
    y = torch.empty(R, Co, device="cuda")
    t = timeit(lambda: ops.gemm_rows(x, w, out=y))
    print(f"| {name} | {R} | {K} | {Co} | {t:.3f} | {2.0 * R * K * Co / t / 1e9:.0f} | {(R * K + R * Co) * 4 / t / 1e6:.0f} |")
    del x, w, y
print()
print("| wgrad | R | Cout | K | ms | TFLOP/s | GB/s (dY + X read) |")
print("|---|---:|---:|---:|---:|---:|---:|")
for name, R, Co, K in [("enc second_conv[0]", Re, 2048, 512), ("enc first_conv[1]", Re, 512, 128), ("dec final_conv[1]", Rd, 512, 256)]:
    dy = torch.randn(R, Co, device="cuda"); x = torch.randn(R, K, device="cuda")
    t = timeit(lambda: ops.gemm_wgrad(dy, x))
    print(f"| {name} | {R} | {Co} | {K} | {t:.3f} | {2.0 * R * K * Co / t / 1e9:.0f} | {(R * K + R * Co) * 4 / t / 1e6:.0f} |")
    del dy, x
