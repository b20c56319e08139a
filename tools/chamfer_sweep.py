"""BASELINE config 4: Chamfer L1 fwd+bwd sweep, B=32, N=M in {1k..64k} plus the training shapes, this implementation vs
the reference kernel recompiled for sm_100a (oracle/_ref).  Prints a markdown table (Gpairs/s = directed pairs of the
forward / forward time; % = 6 FP32 instr per pair / (148 SMs x 128 lanes x SM clock))."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vn_pointcloudcompletion_b200 as V
from oracle import ref_chamfer as RC

def timeit(fn, iters):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def main():
    B = 32
    clk = float(os.environ.get("SM_MHZ", "1965")) * 1e6
    peak = 148 * 128 * clk / 6
    shapes = [(n, n) for n in (1024, 2048, 4096, 8192, 16384, 32768, 65536)] + [(1024, 16384), (16384, 16384)]
    print("| N | M | ours fwd ms | ours Gpairs/s | % FP32 peak | ours fwd+bwd (L1) ms | reference kernel fwd ms | ref Gpairs/s | speed-up fwd | idx/dist equal |")
    print("|---|---|---:|---:|---:|---:|---:|---:|---:|---|")
    g = torch.Generator(device="cuda").manual_seed(0)
    for N, M in shapes:
        a = (torch.rand(B, N, 3, device="cuda", generator=g) - 0.5).requires_grad_(True)
        b = torch.rand(B, M, 3, device="cuda", generator=g) - 0.5
        pairs = 2.0 * B * N * M
        iters = 20 if N * M <= 16384 * 16384 else 3
        t_f = timeit(lambda: V.chamfer_3DFunction.apply(a.detach(), b), iters)
        def fb():
            a.grad = None
            V.cd_loss_L1(a, b).backward()
        t_fb = timeit(fb, iters)
        if RC.available():
            t_r = timeit(lambda: RC.forward(a.detach(), b), max(1, iters // 4))
            r = RC.forward(a.detach(), b); o = V.chamfer_3DFunction.apply(a.detach(), b)
            eq = all(torch.equal(x, y) for x, y in zip(o, r))
        else:
            t_r, eq = float("nan"), "n/a"
        print(f"| {N} | {M} | {t_f:.3f} | {pairs / t_f / 1e6:.0f} | {100 * pairs / (t_f * 1e-3) / peak:.1f}% | {t_fb:.3f} | {t_r:.3f} | {pairs / t_r / 1e6:.0f} | {t_r / t_f:.1f}x | {eq} |")

if __name__ == "__main__":
    main()
