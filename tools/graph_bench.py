"""SURVEY 8f row f1 measurement: the point-set graph kernels and the VN_DGCNN_fps encoder on one B200, with the CPU oracle
(oracle/graph_oracle.c, OpenMP on all host cores) timed beside them.  Prints a markdown table.

    python tools/graph_bench.py [--batch 32] [--no-cpu]
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import numpy as np
import torch
import vn_pointcloudcompletion_b200 as V
from vn_pointcloudcompletion_b200 import graph_ops as G
from vn_pointcloudcompletion_b200.synthetic import make_batch
from vn_pointcloudcompletion_b200.trainer import DataParallelTrainer


def timeit(fn, iters=20, warm=3):
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
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--profile-enc", action="store_true", help="only 3 forward+backward passes of the encoder (for an ncu launch list)")
    a = ap.parse_args()
    B = a.batch
    p, c, R = make_batch(B, 2048, 16384, seed=1234)
    pt, ct, Rt = (torch.from_numpy(x).cuda() for x in (p, c, R))
    rows = []
    if a.profile_enc:
        V.set_gemm_mode("tf32")
        cfg = SimpleNamespace(num_coarse=1024, latent_dim=512, only_coarse=False, device="cuda", enc_pretrained="none")
        torch.manual_seed(0)
        enc = V.VN_DGCNN_fps(cfg).cuda().train()
        for _ in range(3):
            for q in enc.parameters(): q.grad = None
            co, gf = enc(pt)
            (co.sum() + gf.sum()).backward()
        torch.cuda.synchronize()
        return
    t = timeit(lambda: G.knn3d(pt, pt, 16))
    pairs = B * 2048.0 * 2048
    rows.append(("knn3d k=16, 2048 x 2048", t, f"{pairs / t / 1e6:.0f} Gpairs/s"))
    t512 = timeit(lambda: G.fps(pt, 512))
    rows.append(("fps 2048 -> 512", t512, f"{511 / t512:.0f} selections/ms per sample (sequential), {B} samples in parallel"))
    sub = pt[:, :512].contiguous()
    t = timeit(lambda: G.knn3d(sub, sub, 16))
    rows.append(("knn3d k=16, 512 x 512", t, ""))
    t = timeit(lambda: G.fps(sub, 128))
    rows.append(("fps 512 -> 128", t, ""))
    V.set_gemm_mode("tf32")
    cfg = SimpleNamespace(num_coarse=1024, latent_dim=512, only_coarse=False, device="cuda", enc_pretrained="none")
    torch.manual_seed(0)
    net = V.PCNNet(cfg, enc_type="vn_dgcnn_fps", dec_type="vn_foldingnet").train()
    enc = net.encoder
    def enc_fwd():
        with torch.no_grad():
            enc(pt)
    rows.append(("VN_DGCNN_fps forward (no grad), TF32", timeit(enc_fwd), ""))
    def enc_fb():
        for q in enc.parameters(): q.grad = None
        co, gf = enc(pt)
        (co.sum() + gf.sum()).backward()
    rows.append(("VN_DGCNN_fps forward + backward, TF32", timeit(enc_fb, 10), ""))
    tr = DataParallelTrainer(net, lr=1e-4)
    t = timeit(lambda: tr.train_step(pt, ct, Rt), 10)
    rows.append(("PCNNet(vn_dgcnn_fps + vn_foldingnet, latent 512) train step, TF32", t, f"{B / t * 1e3:.0f} samples/s"))
    cpu = {}
    if not a.no_cpu:
        from oracle import graph_oracle as GO
        t0 = time.perf_counter(); oi, _ = GO.knn3d(p, p, 16); cpu["knn3d k=16, 2048 x 2048"] = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter(); of = GO.fps(p, 512); cpu["fps 2048 -> 512"] = (time.perf_counter() - t0) * 1e3
        assert np.array_equal(G.knn3d(pt, pt, 16).cpu().numpy(), oi) and np.array_equal(G.fps(pt, 512).cpu().numpy(), of)
    print(f"B = {B}; CPU = oracle/graph_oracle.c with OpenMP on {os.cpu_count()} host cores (bit-identical indices asserted)\n")
    print("| op | B200 ms | note | CPU oracle ms |")
    print("|---|---:|---|---:|")
    for name, t, note in rows:
        print(f"| {name} | {t:.3f} | {note} | {cpu.get(name, float('nan')):.1f} |")


if __name__ == "__main__":
    main()
