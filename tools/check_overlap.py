"""torchrun --nproc-per-node 2 tools/check_overlap.py : the overlapped two-bucket gradient exchange (trainer.OverlappedExchange, NCCL, launched
from the autograd thread during backward) against the plain single all-reduce, same weights and per-rank batches; prints the relative L2
difference of the exchanged flat gradient per step (the gradient is evaluated twice, hooks on and off; fp32 atomics make two evaluations of
the same gradient differ by ~1e-6)."""
import os
import sys
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import vn_pointcloudcompletion_b200 as V
from vn_pointcloudcompletion_b200.loss import cd_loss_L1
from vn_pointcloudcompletion_b200.model import Rotate
from vn_pointcloudcompletion_b200.synthetic import make_batch
from vn_pointcloudcompletion_b200.trainer import DataParallelTrainer, exchange_gradients

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
V.set_gemm_mode("tf32")
cfg = SimpleNamespace(num_coarse=1024, latent_dim=2048, only_coarse=False, device=dev, enc_pretrained="none")
torch.manual_seed(0)
net = V.PCNNet(cfg).train()
tr = DataParallelTrainer(net, lr=0.0, world_size=world)          # lr 0: the weights stay put, every step sees the same function
assert tr.exchange is not None
p, c, R = (torch.from_numpy(a).to(dev) for a in make_batch(4, 2048, 16384, seed=1234 + rank))
worst = 0.0


def backward_once():
    tr.opt.zero_grad()
    coarse, dense = net(p, Rotate(R))
    (cd_loss_L1(coarse, c) + cd_loss_L1(dense, c)).backward()


for step in range(4):
    # (a) the trainer's path: hooks on, the tail is reduced while backward is still running
    tr.exchange.enabled = True
    backward_once()
    early = tr.exchange.work is not None
    tr.exchange.finish()
    got = tr.opt.flat_g.clone()
    # (b) the same gradient again with the hooks off, then ONE plain all-reduce
    tr.exchange.enabled = False
    backward_once()
    want = tr.opt.flat_g.clone()
    exchange_gradients(want, world)
    err = float((got - want).norm() / want.norm())
    worst = max(worst, err)
    if rank == 0:
        print(f"step {step}: tail launched during backward = {early}, rel-L2(overlapped - plain) = {err:.2e}, |g| = {float(want.norm()):.4e}")
ok = worst < 1e-4
if rank == 0:
    print("OK" if ok else "MISMATCH")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
