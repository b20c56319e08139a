"""SURVEY 8f row f2 measurement: VN attention core, VNLayerNorm and the Attention_VN_FoldingNet train step on one B200.

    python tools/attn_bench.py [--batch 32]
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import torch
import vn_pointcloudcompletion_b200 as V
from vn_pointcloudcompletion_b200 import ops
from vn_pointcloudcompletion_b200.synthetic import make_batch
from vn_pointcloudcompletion_b200.trainer import DataParallelTrainer


def timeit(fn, iters=10, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--profile-step", action="store_true", help="only 2 train steps of the attention network (for an ncu launch list)")
    a = ap.parse_args()
    if a.profile_step:
        V.set_gemm_mode("tf32")
        cfg = SimpleNamespace(num_coarse=1024, latent_dim=2048, only_coarse=False, device="cuda", enc_pretrained="none")
        torch.manual_seed(0)
        net = V.PCNNet(cfg, enc_type="vn_pointnet", dec_type="attention_vn_foldingnet").train()
        p, c, R = make_batch(a.batch, 2048, 16384, seed=1234)
        pt, ct, Rt = (torch.from_numpy(z).cuda() for z in (p, c, R))
        tr = DataParallelTrainer(net, lr=1e-4)
        for _ in range(2):
            tr.train_step(pt, ct, Rt)
        torch.cuda.synchronize()
        return
    B, N, H, D = a.batch, 1024, 8, 48
    C = H * D
    print(f"B = {B}, tokens N = {N}, heads = {H}, head features = 3 x {D}\n")
    print("| op | B200 ms | rate |")
    print("|---|---:|---|")
    qkv = (torch.randn(B * N * 3, 3 * C, device="cuda") * 0.2).requires_grad_(True)
    fl = 4.0 * B * H * N * N * 3 * D
    t = timeit(lambda: ops.vn_attention(qkv.detach(), B, N, H, 1.0))
    print(f"| vn_attention forward (fp32 SIMT, flash-style) | {t:.3f} | {fl / t / 1e9:.1f} TFLOP/s |")
    V.set_gemm_mode("tf32")
    t = timeit(lambda: ops.vn_attention(qkv.detach(), B, N, H, 1.0))
    V.set_gemm_mode("fp32")
    print(f"| vn_attention forward (tcgen05 / TMEM, TF32 operands, two-pass softmax) | {t:.3f} | {fl / t / 1e9:.1f} TFLOP/s algorithmic ({1.5 * fl / t / 1e9:.1f} executed: Q K^T runs twice) |")
    go = torch.randn(B * N * 3, C, device="cuda")
    def fb():
        qkv.grad = None
        ops.vn_attention(qkv, B, N, H, 1.0).backward(go)
    t2 = timeit(fb)
    print(f"| vn_attention forward + backward (fp32 SIMT) | {t2:.3f} | {3.5 * fl / t2 / 1e9:.1f} TFLOP/s (7 tile products) |")
    V.set_gemm_mode("tf32")
    t3 = timeit(fb)
    V.set_gemm_mode("fp32")
    print(f"| vn_attention forward + backward (tcgen05: fwd, dV, dQ, dK kernels) | {t3:.3f} | {3.5 * fl / t3 / 1e9:.1f} TFLOP/s algorithmic |")
    x = torch.randn(B * N * 3, C, device="cuda")
    ln = torch.nn.LayerNorm(C).cuda()
    with torch.no_grad():
        t = timeit(lambda: ops.vn_layernorm(x, ln))
    print(f"| VNLayerNorm forward [{B * N * 3}, {C}] | {t:.3f} | {2 * x.numel() * 4 / t / 1e6:.0f} GB/s |")
    V.set_gemm_mode("tf32")
    cfg = SimpleNamespace(num_coarse=1024, latent_dim=2048, only_coarse=False, device="cuda", enc_pretrained="none")
    torch.manual_seed(0)
    net = V.PCNNet(cfg, enc_type="vn_pointnet", dec_type="attention_vn_foldingnet").train()
    p, c, R = make_batch(B, 2048, 16384, seed=1234)
    pt, ct, Rt = (torch.from_numpy(z).cuda() for z in (p, c, R))
    tr = DataParallelTrainer(net, lr=1e-4)
    t = timeit(lambda: tr.train_step(pt, ct, Rt), 5)
    print(f"| PCNNet(vn_pointnet + attention_vn_foldingnet) train step, TF32 mode (tcgen05 GEMMs and attention) | {t:.3f} | {B / t * 1e3:.0f} samples/s |")


if __name__ == "__main__":
    main()
