"""A/B timing of the fused small-K layer (csrc/vn_fused.cu) and of the BatchNorm-on-norm / leaky streaming kernels (csrc/vn_stream.cu) at the
decoder shapes of the BASELINE train step (B = 32, 16384 dense points, 256 channels), per development knob of vnpcc_set_tuning:
  knob 0: 1 = legacy fixed grids, 0 = occupancy-sized single-wave grids
  knob 1: register budget / channels per thread of the fused small-K backward (1, 2, 3)
Gradients of every variant are compared with variant (0=0, 1=1).  CUDA events after warm-up; prints a markdown table.

    python tools/stream_bench.py [--out gpurun_out/stream_bench.md]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn

import vn_pointcloudcompletion_b200 as V
from vn_pointcloudcompletion_b200 import _lib, ops


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--batch", type=int, default=32)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    V.set_gemm_mode("tf32")
    B, N, C = args.batch, 16384, 256
    g = torch.Generator(device="cpu").manual_seed(0)
    x = (torch.rand(B * N * 3, 2, generator=g) - 0.5).to(dev).requires_grad_(True)
    w = (torch.randn(2 * C, 2, generator=g) * 0.5).to(dev).requires_grad_(True)
    bias = (torch.randn(B * 3, 2 * C, generator=g) * 0.3).to(dev).requires_grad_(True)
    bn = nn.BatchNorm1d(C).to(dev).train()
    gout = torch.randn(B * N * 3, C, generator=g).to(dev)
    lines = ["| kernel group | knob0 (legacy grid) | knob1 (fold bwd variant) | ms | max rel diff of (gx, gw, gbias) vs baseline |", "|---|---|---|---:|---|"]

    def fold_fwd():
        with torch.no_grad():
            return ops.smallk_bn_leaky(x, w, bias, bn, True, 0.2, B, N, 1)

    def fold_fwd_bwd():
        for t in (x, w, bias, bn.weight, bn.bias):
            t.grad = None
        h = ops.smallk_bn_leaky(x, w, bias, bn, True, 0.2, B, N, 1)
        h.backward(gout)
        return x.grad, w.grad, bias.grad

    ref = None
    for legacy, variant in ((1, 1), (0, 1), (0, 2), (0, 3)):
        _lib.raw("vnpcc_set_tuning", 0, legacy)
        _lib.raw("vnpcc_set_tuning", 1, variant)
        t_f = timed(fold_fwd)
        t_fb = timed(fold_fwd_bwd)
        grads = [t.clone() for t in fold_fwd_bwd()]
        if ref is None:
            ref = grads
            diff = "-"
        else:
            diff = ", ".join(f"{float((a - b).abs().max() / b.abs().max()):.1e}" for a, b in zip(grads, ref))
        lines.append(f"| fused small-K layer: stats + forward | {legacy} | {variant} | {t_f:.3f} | |")
        lines.append(f"| fused small-K layer: forward + backward (sums + main) | {legacy} | {variant} | {t_fb:.3f} | {diff} |")
    _lib.raw("vnpcc_set_tuning", 1, 0)

    # streaming BatchNorm-on-norm + leaky layer on stacked (p | d) rows: forward (stats + apply) and backward (bwd1 + bwd2)
    R = B * N * 3
    pd = torch.randn(R, 2 * C, generator=g).to(dev).requires_grad_(True)
    w2 = (torch.randn(1, C, generator=g) * 0.1).to(dev).requires_grad_(True)
    res = torch.rand(R, generator=g).to(dev)
    gy = torch.randn(R, generator=g).to(dev)
    for legacy in (1, 0):
        _lib.raw("vnpcc_set_tuning", 0, legacy)

        def bn_fwd():
            with torch.no_grad():
                return ops.bn_leaky(pd, None, bn, True, 0.2, stacked=True)

        def bn_fwd_bwd():
            pd.grad = None
            o = ops.bn_leaky(pd, None, bn, True, 0.2, stacked=True)
            o.backward(gout)

        lines.append(f"| streaming BN + leaky [R, 512] -> [R, 256]: stats + forward | {legacy} | - | {timed(bn_fwd):.3f} | |")
        lines.append(f"| streaming BN + leaky: forward + backward | {legacy} | - | {timed(bn_fwd_bwd):.3f} | |")

        def dot_fwd_bwd():
            pd.grad = None
            w2.grad = None
            y = ops.bn_leaky_dot(pd, bn, True, 0.2, w2, res)
            y.backward(gy)

        lines.append(f"| fused tail BN + leaky + VNLinear(256, 1) + residual: forward + backward | {legacy} | - | {timed(dot_fwd_bwd):.3f} | |")
    _lib.raw("vnpcc_set_tuning", 0, 0)
    txt = "\n".join(lines)
    print(txt)
    if args.out:
        with open(args.out, "w") as f:
            f.write(txt + "\n")


if __name__ == "__main__":
    main()
