"""tests/eager_port.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Plain-PyTorch (ATen, autograd) restatement of the reference's VN-PCN forward: the same operator chain the reference executes
(nn.Linear on the transposed tensor, torch.norm + BatchNorm on the norms, mask arithmetic, the materialised concatenations), written
functionally over a state_dict with the reference's keys.  It runs on the GPU box, where /root/reference does not exist, and serves two
purposes:
  * tests/test_gpu_fullsize.py: an independent fp32 reference at the FULL BASELINE shapes (2048 -> 1024 / 16384 points), where the numpy
    oracle takes minutes: outputs and autograd gradients of the CUDA path are compared with it on the same weights and inputs;
  * `python tests/eager_port.py --time`: the secondary yardstick of SURVEY 8d, "the reference's eager-PyTorch model on one B200"
    (kind: port), in fp32 and with torch's TF32 matmul flag, next to which bench.py's number can be read.
It is pinned to the reference itself on CPU by tests/test_oracle_golden.py::test_eager_port_matches_reference_golden.

Reference semantics followed (file:line under /root/reference):
  VNLinear models/vn_layers.py:17-22, VNLinearLeakyReLU :60-74, VNLinearAndLeakyReLU :96-104 (+ VNLeakyReLU :33-43), VNBatchNorm :116-127,
  VNMaxPool :158-167, VN_PointNet.forward models/pcn.py:163-184, VN_FoldingNet.forward :364-389, PCNNet.forward models/model.py:52-64,
  cd_loss_L1 metrics/loss.py:20-31.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import torch
import torch.nn.functional as F

EPS = 1e-6


def vn_linear(x, W):
    return F.linear(x.transpose(1, -1), W).transpose(1, -1)


def _leaky(p, d, ns):
    dot = (p * d).sum(2, keepdim=True)
    mask = (dot >= 0).to(p.dtype)
    dsq = (d * d).sum(2, keepdim=True)
    return ns * p + (1 - ns) * (mask * p + (1 - mask) * (p - (dot / (dsq + EPS)) * d))


def vn_batchnorm(x, P, prefix, training):
    norm = torch.norm(x, dim=2) + EPS
    nb = F.batch_norm(norm, P[prefix + "running_mean"], P[prefix + "running_var"], P[prefix + "weight"], P[prefix + "bias"], training, 0.1, 1e-5)
    return x / norm.unsqueeze(2) * nb.unsqueeze(2)


def vn_linear_leaky_relu(x, P, prefix, training, ns=0.2):
    p = vn_batchnorm(vn_linear(x, P[prefix + "map_to_feat.weight"]), P, prefix + "batchnorm.bn.", training)
    d = vn_linear(x, P[prefix + "map_to_dir.weight"])
    return _leaky(p, d, ns)


def vn_linear_and_leaky_relu(x, P, prefix, ns=0.2):      # use_batchnorm='none' (the only form on this path)
    x = vn_linear(x, P[prefix + "linear.map_to_feat.weight"])
    return _leaky(x, vn_linear(x, P[prefix + "leaky_relu.map_to_dir.weight"]), ns)


def vn_max_pool(x, Wd, forced_idx=None):
    """[B, C, 3, N] -> ([B, C, 3], idx [B, C]); first maximum wins (torch.max)"""
    if forced_idx is None:
        d = vn_linear(x, Wd)
        idx = (x * d).sum(2).max(dim=-1)[1]
    else:
        idx = forced_idx
    return x.gather(3, idx[:, :, None, None].expand(-1, -1, 3, 1)).squeeze(-1), idx


def encoder(P, xyz, training=True, forced_idx=(None, None), prefix="encoder."):
    B, N, _ = xyz.shape
    f = vn_linear_leaky_relu(xyz.transpose(2, 1).unsqueeze(1), P, prefix + "first_conv.0.", training)
    f = vn_linear(f, P[prefix + "first_conv.1.map_to_feat.weight"])
    g, idx1 = vn_max_pool(f, P[prefix + "maxpool1.map_to_dir.weight"], forced_idx[0])
    f = torch.cat([g.unsqueeze(-1).expand(-1, -1, -1, N), f], dim=1)
    f = vn_linear_leaky_relu(f, P, prefix + "second_conv.0.", training)
    f = vn_linear(f, P[prefix + "second_conv.1.map_to_feat.weight"])
    fg, idx2 = vn_max_pool(f, P[prefix + "maxpool2.map_to_dir.weight"], forced_idx[1])
    fg = fg.unsqueeze(-1)
    m = vn_linear_and_leaky_relu(fg, P, prefix + "mlp.0.")
    m = vn_linear_and_leaky_relu(m, P, prefix + "mlp.1.")
    m = vn_linear(m, P[prefix + "mlp.2.map_to_feat.weight"])
    nc = m.shape[1]
    return m.reshape(-1, nc, 3).contiguous(), fg, (idx1, idx2)


def decoder(P, coarse, fg, R=None, training=True, grid_size=4, prefix="decoder."):
    B, nc, _ = coarse.shape
    S = grid_size * grid_size
    nd = nc * S
    lin = torch.linspace(-0.05, 0.05, steps=grid_size, dtype=coarse.dtype, device=coarse.device)
    a = lin.view(1, grid_size).expand(grid_size, grid_size).reshape(1, -1)
    b = lin.view(grid_size, 1).expand(grid_size, grid_size).reshape(1, -1)
    seed = torch.cat([a, b, torch.zeros_like(a)], dim=0).reshape(1, 1, 3, -1)
    if R is not None:      # rot.transform_points(p) = p @ R on the [1, S, 3] seed points (models/pcn.py:367-370)
        seed = (seed.squeeze(1).transpose(1, 2) @ R).transpose(1, 2).unsqueeze(1)
    point_feat = coarse.unsqueeze(2).expand(-1, -1, S, -1).reshape(-1, nd, 3).transpose(2, 1).unsqueeze(1)
    seed = seed.unsqueeze(3).expand(B, -1, -1, nc, -1).reshape(B, -1, 3, nd)
    feat = torch.cat([fg.expand(-1, -1, -1, nd), seed, point_feat], dim=1)
    h = vn_linear_leaky_relu(feat, P, prefix + "final_conv.0.", training)
    h = vn_linear_leaky_relu(h, P, prefix + "final_conv.1.", training)
    fine = vn_linear(h, P[prefix + "final_conv.2.map_to_feat.weight"]) + point_feat
    return fine.squeeze(1).transpose(1, 2).contiguous()


def pcn_forward(P, xyz, R=None, training=True, forced_idx=(None, None)):
    coarse, fg, idx = encoder(P, xyz, training, forced_idx)
    return coarse, decoder(P, coarse, fg, R, training), idx


def cd_loss_l1(chamfer, a, b):
    """metrics/loss.py:20-31 on top of a chamfer(xyz1, xyz2) -> (dist1, dist2, ...) callable"""
    d = chamfer(a, b)
    return (torch.sqrt(d[0]).mean() + torch.sqrt(d[1]).mean()) / 2.0


def params_from_module(net, requires_grad=False):
    """state_dict of a PCNNet (reference keys) as independent leaf tensors"""
    P = {}
    for k, v in net.state_dict().items():
        t = v.detach().clone()
        if requires_grad and t.is_floating_point() and "running_" not in k:
            t.requires_grad_(True)
        P[k] = t
    return P


def _time_main():
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from types import SimpleNamespace

    import vn_pointcloudcompletion_b200 as V
    from vn_pointcloudcompletion_b200.synthetic import make_batch
    ap = argparse.ArgumentParser()
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    cfg = SimpleNamespace(num_coarse=1024, latent_dim=2048, only_coarse=False, device=dev, enc_pretrained="none")
    torch.manual_seed(0)
    net = V.PCNNet(cfg).train()
    P = params_from_module(net, requires_grad=True)
    del net
    leaves = [t for t in P.values() if t.requires_grad]
    opt = torch.optim.Adam(leaves, lr=1e-4)
    chamfer = V.chamfer_3DFunction.apply      # bit-identical to the reference kernel (tests/test_gpu_chamfer.py) and 2.3x faster: favours the yardstick
    print("| eager PyTorch port of the reference step (fwd + L1-CD + bwd + Adam) | batch | ms / step | samples/s | peak memory GB |\n|---|---:|---:|---:|---:|")
    for tf32 in (False, True):
        batch = args.batch
        while batch >= 1:
            p, c, R = (torch.from_numpy(a).to(dev) for a in make_batch(batch, 2048, 16384, seed=1234))
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            torch.cuda.reset_peak_memory_stats()

            def step():
                opt.zero_grad(set_to_none=True)
                coarse, fine, _ = pcn_forward(P, p, R, True)
                loss = cd_loss_l1(chamfer, coarse, c) + cd_loss_l1(chamfer, fine, c)
                loss.backward()
                opt.step()
                return loss
            try:
                step()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    step()
                torch.cuda.synchronize()
            except torch.cuda.OutOfMemoryError:      # the materialised [B, 2050, 3, 16384] concatenation and its transposed copies
                opt.zero_grad(set_to_none=True)
                torch.cuda.empty_cache()
                batch //= 2
                continue
            dt = (time.perf_counter() - t0) / args.steps
            print(f"| {'TF32 matmul flag' if tf32 else 'fp32'} | {batch} | {dt * 1e3:.1f} | {batch / dt:.1f} | {torch.cuda.max_memory_allocated() / 2**30:.1f} |")
            break


if __name__ == "__main__":
    _time_main()
