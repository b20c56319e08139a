"""development aid: what HBM delivers on this box for read-only, write-only and mixed streams (torch kernels over 3.2 GB tensors),
the yardstick for the write-heavy decoder GEMM (1.6 GB in, 3.2 GB out) and the read-only weight-gradient GEMM"""
import torch
from stream_bench import timed

n = 800 * 1024 * 1024          # 3.2 GB of fp32
a = torch.empty(n, device="cuda").normal_()
b = torch.empty(n, device="cuda")
h = a[: n // 2]
gb = 4 * n / 1e9
t = timed(lambda: b.zero_(), 10)
print(f"write-only  fill 3.2 GB            : {t:.3f} ms  {gb / t:.0f} GB/s")
t = timed(lambda: a.sum(), 10)
print(f"read-only   sum  3.2 GB            : {t:.3f} ms  {gb / t:.0f} GB/s")
t = timed(lambda: b.copy_(a), 10)
print(f"copy        3.2 GB -> 3.2 GB       : {t:.3f} ms  {2 * gb / t:.0f} GB/s")
t = timed(lambda: torch.cat([h, h], out=b), 10)
print(f"1 read : 2 written (cat of a half) : {t:.3f} ms  {1.5 * gb / t:.0f} GB/s (DRAM traffic; the half may be re-read from L2)")
t = timed(lambda: torch.add(a[: n // 2], a[n // 2:], out=b[: n // 2]), 10)
print(f"2 read : 1 written (add)           : {t:.3f} ms  {1.5 * gb / t:.0f} GB/s")
