"""FP32-pipe micro-benchmark (csrc/microbench.cu): scalar FFMA, packed FFMA2, and the Chamfer instruction mix."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vn_pointcloudcompletion_b200 import _lib
lib = _lib.load()
scratch = torch.zeros(16, device="cuda")
for mode, name in ((0, "FFMA scalar"), (1, "FFMA2 packed"), (2, "chamfer mix (3 FADD2+FMUL2+2 FFMA2+FMNMX3 per 2 pairs)")):
    ms = ctypes.c_float(); ops = ctypes.c_double()
    rc = lib.vnpcc_measure_fp32_peak(mode, 20000, scratch.data_ptr(), ctypes.addressof(ms), ctypes.addressof(ops), torch.cuda.current_stream().cuda_stream)
    rate = ops.value / (ms.value * 1e-3)
    print(f"{name}: {ms.value:.3f} ms, {rate/1e12:.2f} T lane-ops/s = {100*rate/(148*128*1.965e9):.1f}% of 148x128x1.965GHz")
