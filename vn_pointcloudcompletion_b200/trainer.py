"""Batch-sharded data-parallel training step for PCNNet (the data-parallel extension of train.py:127-173; the
reference itself is single-GPU).

One process per GPU.  Parameters and gradients live in two flat fp32 buffers (every nn.Parameter is a view), so the
gradient exchange is ONE NCCL all-reduce over NVLink per step and the optimiser is ONE fused Adam kernel
(torch.optim.Adam semantics: train.py:70 Adam(lr, betas=(0.9, 0.999)); parameters whose gradient is None in the
reference -- the two VNMaxPool.map_to_dir weights, SURVEY.md B.3 -- see a zero gradient and therefore never move).
BatchNorm statistics stay rank-local, exactly the single-process semantics at the per-rank batch size.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops
from .loss import cd_loss_L1
from .model import Rotate


def exchange_gradients(flat_grad, world_size, process_group=None):
    """the ONE exchange step of the data-parallel path: sum the flat gradient buffer over ranks (NCCL over NVLink on
    the GPUs; any backend works -- the CPU tests run it over gloo).  Returns the scale (1/world) the optimiser applies,
    so that the mean is never materialised in a separate pass."""
    if world_size > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=process_group)
    return 1.0 / world_size


def rank_seed(base_seed, rank, step=0):
    """per-rank, per-step data seed: ranks draw disjoint synthetic shards (bench.py)"""
    return base_seed + rank + 1000 * step


class FlatAdam:
    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat_p = torch.empty(n, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(n, device=dev, dtype=torch.float32)
        self.m = torch.zeros(n, device=dev, dtype=torch.float32)
        self.v = torch.zeros(n, device=dev, dtype=torch.float32)
        o = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat_p[o:o + k].copy_(p.reshape(-1))
                p.data = self.flat_p[o:o + k].view_as(p)
                p.grad = self.flat_g[o:o + k].view_as(p)
                o += k
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.step_count = 0

    def zero_grad(self):
        self.flat_g.zero_()

    def step(self, grad_scale=1.0):
        self.step_count += 1
        ops.adam_step(self.flat_p, self.flat_g, self.m, self.v, self.lr, self.betas[0], self.betas[1], self.eps, self.wd,
                      self.step_count, grad_scale)

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.m, "exp_avg_sq": self.v, "lr": self.lr}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.m.copy_(sd["exp_avg"])
        self.v.copy_(sd["exp_avg_sq"])
        self.lr = sd.get("lr", self.lr)


class DataParallelTrainer:
    """train.py:127-173 for one rank: forward, L1-CD(coarse) + L1-CD(dense), backward, grad all-reduce, Adam."""

    def __init__(self, model, lr=1e-4, world_size=1, process_group=None):
        self.model = model
        self.opt = FlatAdam(model.parameters(), lr=lr)
        self.world = world_size
        self.pg = process_group

    def train_step(self, p, c, R=None):
        """p [B,2048,3] partial, c [B,16384,3] complete, R [B,3,3] rotation already applied to both (train.py:133-138).
        Returns the detached loss tensor (no host sync)."""
        self.opt.zero_grad()
        coarse, dense = self.model(p, Rotate(R) if R is not None else None)
        loss = cd_loss_L1(coarse, c)
        if dense is not None:
            loss = loss + cd_loss_L1(dense, c)
        loss.backward()
        scale = exchange_gradients(self.opt.flat_g, self.world, self.pg)
        self.opt.step(grad_scale=scale)
        return loss.detach()
