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


class OverlappedExchange:
    """The same sum as exchange_gradients, started early and without the bytes that are never written: the flat gradient buffer is cut at
    `split` (the offset of the first parameter of the LAST-executed layers: everything from there to the end of the buffer is complete
    early in the backward pass).  As soon as every parameter of that tail has received its gradient, its all-reduce is launched
    asynchronously and runs over NVLink under the rest of the backward pass; the head is reduced after backward.  For VN-PCN the tail is
    the encoder's mlp + the decoder: 80 % of the 90 MB, ready after ~45 % of the backward time.

    `params`: [(parameter, offset in flat_grad, numel)].  Some parameters never receive a gradient (VNMaxPool direction weights: 17.8 MB
    of the head; modules the reference's forward does not use), so WHICH parameters receive one is measured on the first step, which runs
    the plain exchange: afterwards only the ranges that are written are exchanged.  The graph is static; a step that deviates falls back
    to the plain exchange or fails loudly."""

    def __init__(self, flat_grad, params, split, world_size, process_group=None):
        self.flat = flat_grad
        self.split = int(split)
        self.world = world_size
        self.pg = process_group
        self.info = {id(p): (int(o), int(n)) for p, o, n in params}
        self.fired = set()            # ids of the parameters that received a gradient in this step
        self.expected = None          # after calibration: ids expected per step
        self.tail_ids = None
        self.head_ranges = None
        self.tail_range = None
        self.tail_seen = 0
        self.work = None
        self.enabled = True           # False: the hooks do nothing (tools/check_overlap.py evaluates the plain exchange beside this one)
        self.handles = [p.register_post_accumulate_grad_hook(self._on_grad) for p, _, _ in params] if world_size > 1 else []

    @staticmethod
    def _merge(ranges):
        out = []
        for a, b in sorted(ranges):
            if out and a <= out[-1][1]:
                out[-1][1] = max(out[-1][1], b)
            else:
                out.append([a, b])
        return [(a, b) for a, b in out]

    def _on_grad(self, param):
        if not self.enabled:
            return
        pid = id(param)
        self.fired.add(pid)
        if self.tail_ids is not None and pid in self.tail_ids:
            self.tail_seen += 1
            if self.tail_seen == len(self.tail_ids):
                # NCCL: the collective is enqueued on the communication stream behind everything the compute stream has done so far
                a, b = self.tail_range
                self.work = dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.SUM, group=self.pg, async_op=True)

    def finish(self):
        """call after backward(): reduces the head, joins the tail; returns the 1/world scale for the optimiser"""
        fired, work = self.fired, self.work
        self.fired, self.work, self.tail_seen = set(), None, 0
        if self.world > 1:
            if self.expected is None:          # calibration step: plain exchange, remember who received a gradient
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.pg)
                if fired:
                    self.expected = frozenset(fired)
                    self.tail_ids = frozenset(i for i in fired if self.info[i][0] >= self.split)
                    spans = [(self.info[i][0], self.info[i][0] + self.info[i][1]) for i in fired]
                    tail = [s for s in spans if s[0] >= self.split]
                    self.tail_range = (min(s[0] for s in tail), max(s[1] for s in tail)) if tail else None
                    self.head_ranges = self._merge([s for s in spans if s[0] < self.split])
                    if not tail:
                        self.tail_ids = None
            elif fired != self.expected:
                if work is not None:
                    raise RuntimeError("the set of parameters receiving a gradient changed after the tail bucket was reduced: the "
                                       "autograd graph changed; rebuild the trainer")
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.pg)
            else:
                for a, b in self.head_ranges:
                    dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.SUM, group=self.pg)
                if work is not None:
                    work.wait()
                elif self.tail_range is not None:      # no tail hook fired (cannot happen when fired == expected), be safe
                    dist.all_reduce(self.flat[self.tail_range[0]:self.tail_range[1]], op=dist.ReduceOp.SUM, group=self.pg)
        return 1.0 / self.world

    def remove(self):
        for h in self.handles:
            h.remove()
        self.handles = []


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

    def offset_of(self, param):
        """offset of a parameter's view inside the flat buffers"""
        return (param.data_ptr() - self.flat_p.data_ptr()) // 4

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

    def __init__(self, model, lr=1e-4, world_size=1, process_group=None, overlap=True):
        self.model = model
        self.opt = FlatAdam(model.parameters(), lr=lr)
        self.world = world_size
        self.pg = process_group
        # overlap the gradient exchange with the backward pass: the tail bucket starts at the first parameter of the layers that run last
        # in the forward pass (VN_PointNet.mlp, then the decoder); parameters are laid out in module order, so the tail is contiguous
        self.exchange = None
        tail_start = getattr(getattr(model, "encoder", None), "mlp", None)
        if tail_start is None:                 # encoders without a late mlp (VN_DGCNN_fps): the tail is the decoder
            tail_start = getattr(model, "decoder", None)
        if overlap and world_size > 1 and tail_start is not None:
            first = next(iter(tail_start.parameters()), None)
            if first is not None and first.requires_grad:
                split = self.opt.offset_of(first)
                plist = [(p, self.opt.offset_of(p), p.numel()) for p in self.opt.params]
                self.exchange = OverlappedExchange(self.opt.flat_g, plist, split, world_size, process_group)

    def train_step(self, p, c, R=None):
        """p [B,2048,3] partial, c [B,16384,3] complete, R [B,3,3] rotation already applied to both (train.py:133-138).
        Returns the detached loss tensor (no host sync)."""
        self.opt.zero_grad()
        coarse, dense = self.model(p, Rotate(R) if R is not None else None)
        loss = cd_loss_L1(coarse, c)
        if dense is not None:
            loss = loss + cd_loss_L1(dense, c)
        loss.backward()
        if self.exchange is not None:
            scale = self.exchange.finish()
        else:
            scale = exchange_gradients(self.opt.flat_g, self.world, self.pg)
        self.opt.step(grad_scale=scale)
        return loss.detach()
