"""development aid: fused small-K decoder layer (csrc/vn_fused.cu), backward variants of vnpcc_set_tuning knob 1 at the BASELINE shape"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
import vn_pointcloudcompletion_b200 as V
from vn_pointcloudcompletion_b200 import _lib, ops
from stream_bench import timed

V.set_gemm_mode("tf32")
dev = torch.device("cuda", 0)
B, N, C = 32, 16384, 256
g = torch.Generator(device="cpu").manual_seed(0)
x = (torch.rand(B * N * 3, 2, generator=g) - 0.5).to(dev).requires_grad_(True)
w = (torch.randn(2 * C, 2, generator=g) * 0.5).to(dev).requires_grad_(True)
bias = (torch.randn(B * 3, 2 * C, generator=g) * 0.3).to(dev).requires_grad_(True)
bn = nn.BatchNorm1d(C).to(dev).train()
gout = torch.randn(B * N * 3, C, generator=g).to(dev)


def fwd():
    with torch.no_grad():
        return ops.smallk_bn_leaky(x, w, bias, bn, True, 0.2, B, N, 1)


def fwd_bwd():
    for t in (x, w, bias, bn.weight, bn.bias):
        t.grad = None
    h = ops.smallk_bn_leaky(x, w, bias, bn, True, 0.2, B, N, 1)
    h.backward(gout)
    return x.grad, w.grad, bias.grad, bn.weight.grad, bn.bias.grad


for fv in (1, 0, 3, 4):      # knob 8: 1 = fp64 statistics pass + forward at the compiler's register count, 0 = default, 3 / 4 = forward at 3 / 4 CTAs per SM
    _lib.raw("vnpcc_set_tuning", 8, fv)
    print(f"forward variant {fv}: statistics + forward {timed(fwd):.3f} ms")
_lib.raw("vnpcc_set_tuning", 8, 0)
t_f = timed(fwd)
ref = None
for variant in (3, 4):
    _lib.raw("vnpcc_set_tuning", 1, variant)
    t = timed(fwd_bwd)
    grads = [a.clone() for a in fwd_bwd()]
    if ref is None:
        ref = grads
    diff = ", ".join(f"{float((a - b).abs().max() / b.abs().max()):.1e}" for a, b in zip(grads, ref))
    print(f"variant {variant}: forward {t_f:.3f} ms, backward {t - t_f:.3f} ms   (max rel diff vs variant 3: {diff})")
_lib.raw("vnpcc_set_tuning", 1, 0)
