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

    def __init__(self, flat_grad, params, split, world_size, process_group=None, extra_splits=()):
        """extra_splits: offsets below `split` that cut the head into further buckets [s_i, s_{i+1}); each is reduced asynchronously as
        soon as all of its parameters have received their gradient (backward reaches them in reverse buffer order), so that only the
        bucket of the FIRST-executed layers -- complete only when backward ends -- is reduced after backward."""
        self.flat = flat_grad
        self.split = int(split)
        self.extra = sorted(int(x) for x in extra_splits if 0 < int(x) < int(split))
        self.mid = []                 # after calibration: [{"range": (a, b), "ids": frozenset, "seen": 0, "work": None}], one per extra bucket
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
            return
        for m in self.mid:
            if pid in m["ids"]:
                m["seen"] += 1
                if m["seen"] == len(m["ids"]):
                    a, b = m["range"]
                    m["work"] = dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
                return

    def finish(self):
        """call after backward(): reduces the head, joins the tail; returns the 1/world scale for the optimiser"""
        fired, work = self.fired, self.work
        self.fired, self.work, self.tail_seen = set(), None, 0
        mid_works = [m["work"] for m in self.mid]
        for m in self.mid:
            m["seen"], m["work"] = 0, None
        if self.world > 1:
            if self.expected is None:          # calibration step: plain exchange, remember who received a gradient
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.pg)
                if fired:
                    self.expected = frozenset(fired)
                    self.tail_ids = frozenset(i for i in fired if self.info[i][0] >= self.split)
                    spans = [(self.info[i][0], self.info[i][0] + self.info[i][1]) for i in fired]
                    tail = [s for s in spans if s[0] >= self.split]
                    self.tail_range = (min(s[0] for s in tail), max(s[1] for s in tail)) if tail else None
                    # head buckets: [0, e_0) is reduced after backward, [e_i, e_{i+1}) ... [e_last, split) asynchronously when complete
                    edges = self.extra + [self.split]
                    first_edge = edges[0]
                    self.head_ranges = self._merge([s for s in spans if s[0] < first_edge])
                    self.mid = []
                    for lo, hi in zip(edges[:-1], edges[1:]):
                        ids = frozenset(i for i in fired if lo <= self.info[i][0] < hi)
                        if ids:
                            sp = [(self.info[i][0], self.info[i][0] + self.info[i][1]) for i in ids]
                            self.mid.append({"range": (min(x[0] for x in sp), max(x[1] for x in sp)), "ids": ids, "seen": 0, "work": None})
                    if not tail:
                        self.tail_ids = None
            elif fired != self.expected:
                if work is not None or any(w is not None for w in mid_works):
                    raise RuntimeError("the set of parameters receiving a gradient changed after the tail bucket was reduced: the "
                                       "autograd graph changed; rebuild the trainer")
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.pg)
            else:
                for a, b in self.head_ranges:
                    dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.SUM, group=self.pg)
                for m, w in zip(self.mid, mid_works):
                    if w is not None:
                        w.wait()
                    else:
                        dist.all_reduce(self.flat[m["range"][0]:m["range"][1]], op=dist.ReduceOp.SUM, group=self.pg)
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


class FlatAdam(torch.optim.Optimizer):
    """torch.optim.Adam (train.py:70: Adam(model.parameters(), lr, betas=(0.9, 0.999))) as ONE fused kernel over flat buffers.

    It IS a torch.optim.Optimizer: `param_groups[0]["lr"]` is what the kernel reads, so torch.optim.lr_scheduler.StepLR (train.py:93)
    drives it unchanged, and state_dict() / load_state_dict() speak torch.optim.Adam's format ({"state": {index: {"step", "exp_avg",
    "exp_avg_sq"}}, "param_groups": [...]}), so the reference's optimizer/optim_last.pth (train.py:72-80, 262-277) can be resumed from and
    is what a checkpoint written here looks like.  Like torch's Adam, parameters that never received a gradient (the two VNMaxPool
    direction weights) have no state entry.

    Every parameter is rebound as a view of `flat_p` and its .grad as a view of `flat_g`.  Moving the model afterwards (model.to(),
    .cuda()) or zero_grad(set_to_none=True) through another optimizer would silently detach them: step() checks the bindings and raises."""

    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        params = [p for p in params if p.requires_grad]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False, foreach=None,
                                      capturable=False, differentiable=False, fused=None))
        if len(self.param_groups) != 1:
            raise ValueError("FlatAdam takes one parameter group (the reference uses one)")
        self.params = self.param_groups[0]["params"]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat_p = torch.empty(n, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(n, device=dev, dtype=torch.float32)
        self.m = torch.zeros(n, device=dev, dtype=torch.float32)
        self.v = torch.zeros(n, device=dev, dtype=torch.float32)
        self._spans = []
        o = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat_p[o:o + k].copy_(p.reshape(-1))
                p.data = self.flat_p[o:o + k].view_as(p)
                p.grad = self.flat_g[o:o + k].view_as(p)
                self._spans.append((o, k))
                o += k
        self._bound = [(p, p.data_ptr(), p.grad.data_ptr()) for p in self.params]
        self.step_count = 0
        self._dev_state = None      # {lr, bias correction 1, sqrt(bias correction 2), step} on the device: set by make_capturable()
        self._dev_lr = None

    def make_capturable(self):
        """keep lr and the step counter in device memory (vnpcc_adam_step_dev) so that step() can be captured in a CUDA graph and replayed;
        the host-side step_count keeps mirroring it (state_dict, schedulers)"""
        if self._dev_state is None:
            self._dev_state = torch.tensor([self.lr, 0.0, 0.0, float(self.step_count)], device=self.flat_p.device, dtype=torch.float32)
            self._dev_lr = self.lr
        return self

    def sync_device_state(self):
        """push a changed learning rate (LR scheduler) or a restored step count to the device copy; called outside the captured region"""
        if self._dev_state is not None:
            if self._dev_lr != self.lr:
                self._dev_state[0:1].fill_(self.lr)
                self._dev_lr = self.lr

    # the hyper-parameters live in the param group (what LR schedulers edit)
    lr = property(lambda self: self.param_groups[0]["lr"])
    betas = property(lambda self: self.param_groups[0]["betas"])
    eps = property(lambda self: self.param_groups[0]["eps"])
    wd = property(lambda self: self.param_groups[0]["weight_decay"])

    def offset_of(self, param):
        """offset of a parameter's view inside the flat buffers"""
        return (param.data_ptr() - self.flat_p.data_ptr()) // 4

    def zero_grad(self, set_to_none=False):
        # never set_to_none: the .grad views ARE the flat gradient buffer the exchange and the fused kernel read
        self.flat_g.zero_()

    def _check_bindings(self):
        for p, dp, gp in self._bound:
            if p.data_ptr() != dp or p.grad is None or p.grad.data_ptr() != gp:
                raise RuntimeError("a parameter (or its .grad) no longer aliases FlatAdam's flat buffers: the model was moved / re-created "
                                   "or its gradients were set to None after the optimizer was built; rebuild the optimizer")

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        loss = closure() if closure is not None else None
        self._check_bindings()
        self.step_count += 1
        if self._dev_state is not None:
            if not torch.cuda.is_current_stream_capturing():
                self.sync_device_state()
            ops.adam_step_dev(self.flat_p, self.flat_g, self.m, self.v, self._dev_state, self.betas[0], self.betas[1], self.eps, self.wd,
                              grad_scale)
        else:
            ops.adam_step(self.flat_p, self.flat_g, self.m, self.v, self.lr, self.betas[0], self.betas[1], self.eps, self.wd,
                          self.step_count, grad_scale)
        return loss

    def state_dict(self):
        """torch.optim.Adam's format; exp_avg / exp_avg_sq are copies shaped like their parameters"""
        self.state.clear()
        if self.step_count > 0:
            touched = [bool(self.v[o:o + k].any()) or bool(self.m[o:o + k].any()) for o, k in self._spans]
            for p, (o, k), t in zip(self.params, self._spans, touched):
                if t:
                    self.state[p] = {"step": torch.tensor(float(self.step_count)), "exp_avg": self.m[o:o + k].view_as(p).clone(),
                                     "exp_avg_sq": self.v[o:o + k].view_as(p).clone()}
        sd = super().state_dict()
        self.state.clear()
        return sd

    def load_state_dict(self, state_dict):
        """accepts torch.optim.Adam's format (any torch version: `step` an int or a tensor) and this class's round-1 flat format"""
        if "exp_avg" in state_dict and "state" not in state_dict:      # round-1 flat format
            self.step_count = int(state_dict["step"])
            self.m.copy_(state_dict["exp_avg"])
            self.v.copy_(state_dict["exp_avg_sq"])
            if "lr" in state_dict:
                self.param_groups[0]["lr"] = state_dict["lr"]
            return
        super().load_state_dict(state_dict)
        self.m.zero_()
        self.v.zero_()
        steps = []
        for p, (o, k) in zip(self.params, self._spans):
            st = self.state.get(p)
            if st:
                self.m[o:o + k].copy_(st["exp_avg"].reshape(-1))
                self.v[o:o + k].copy_(st["exp_avg_sq"].reshape(-1))
                steps.append(int(float(st["step"])))
        if steps and min(steps) != max(steps):
            raise ValueError("parameters with different Adam step counts: FlatAdam keeps one step counter (the reference steps all "
                             "parameters together)")
        self.step_count = steps[0] if steps else 0
        self.state.clear()
        self.params = self.param_groups[0]["params"]
        if self._dev_state is not None:
            self._dev_state[3:4].fill_(float(self.step_count))


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
        self.exchange_off = False
        tail_start = getattr(getattr(model, "encoder", None), "mlp", None)
        if tail_start is None:                 # encoders without a late mlp (VN_DGCNN_fps): the tail is the decoder
            tail_start = getattr(model, "decoder", None)
        if overlap and world_size > 1 and tail_start is not None:
            first = next(iter(tail_start.parameters()), None)
            if first is not None and first.requires_grad:
                split = self.opt.offset_of(first)
                plist = [(p, self.opt.offset_of(p), p.numel()) for p in self.opt.params]
                # a second early bucket: VN_PointNet.second_conv[1] (8.4 MB) is complete ~1.5 ms before backward ends; what is left for
                # after backward is second_conv[0] + first_conv
                extra = []
                sc = getattr(getattr(model, "encoder", None), "second_conv", None)
                if sc is not None and len(sc) > 1:
                    q = next(iter(sc[1].parameters()), None)
                    if q is not None and q.requires_grad:
                        extra.append(self.opt.offset_of(q))
                self.exchange = OverlappedExchange(self.opt.flat_g, plist, split, world_size, process_group, extra_splits=extra)

    def make_scheduler(self, step_size=50, gamma=0.8):
        """train.py:93: StepLR(optimizer, step_size=50, gamma=0.8); call .step() once per epoch like the reference (train.py:186)"""
        self.scheduler = torch.optim.lr_scheduler.StepLR(self.opt, step_size=step_size, gamma=gamma)
        return self.scheduler

    def optimizer_checkpoint(self, epoch, best_metrics, best_epoch):
        """the dict train.py:262-277 saves as optimizer/optim_last.pth (and optim_best.pth)"""
        return {"epoch": epoch, "optim_state_dict": self.opt.state_dict(), "best_metrics": best_metrics, "best_epoch": best_epoch}

    def load_optimizer_checkpoint(self, ckpt):
        """resume like train.py:72-82; returns (start_epoch, best_metrics, best_epoch)"""
        self.opt.load_state_dict(ckpt["optim_state_dict"])
        return ckpt["epoch"] + 1, ckpt["best_metrics"], ckpt["best_epoch"]

    @torch.no_grad()
    def evaluate(self, batches):
        """The validation loop of train.py:199-242 sharded by batch: every rank runs eval-mode forwards over ITS batches
        (p, c, R-or-None), sums l1_cd(coarse, c) and l1_cd(dense, c) (sum-over-batch metrics, metrics/metric.py:19-23) and the sample
        count on the device, and ONE all-reduce of three scalars yields the dataset means on every rank: (coarse, dense, total)."""
        from .loss import l1_cd
        was_training = self.model.training
        self.model.eval()
        dev = self.opt.flat_p.device
        acc = torch.zeros(3, device=dev, dtype=torch.float64)
        for p, c, R in batches:
            coarse, dense = self.model(p, Rotate(R) if R is not None else None)
            acc[0] += l1_cd(coarse, c)
            if dense is not None:
                acc[1] += l1_cd(dense, c)
            acc[2] += p.shape[0]
        if self.world > 1:
            dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=self.pg)
        self.model.train(was_training)
        coarse_m, dense_m = (acc[0] / acc[2]).item(), (acc[1] / acc[2]).item()
        return coarse_m, dense_m, coarse_m + dense_m

    def capture(self, p, c, R=None, warmup=3):
        """Capture the whole train step (zero_grad, forward, both losses, backward, gradient exchange, Adam) for inputs of these shapes in
        ONE CUDA graph; train_step() then copies its inputs into the graph's static buffers and replays it: ~190 kernel launches, their
        Python / autograd dispatch and their launch gaps become one graph launch.  The warm-up steps are real optimisation steps.
        Anything that changes the launch sequence afterwards (gemm mode, tuning knobs, model.eval(), other input shapes) needs a new
        capture (release_graph() first).  Like every whole-step capture in PyTorch it needs the parameters' gradient accumulators to have
        been created on a side stream: call it before the model has run a backward pass on the default stream whose autograd graph is
        still alive (build the model, build the trainer, capture).  Returns the loss of the last warm-up step."""
        from . import _lib
        self.release_graph()
        self.opt.make_capturable()
        self._static_in = (p.clone(), c.clone(), R.clone() if R is not None else None)
        side = torch.cuda.Stream(device=p.device)
        side.wait_stream(torch.cuda.current_stream(p.device))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):      # lazy initialisation (function attributes, workspaces) must not happen under capture
                loss = self._step_eager(*self._static_in)
        torch.cuda.current_stream(p.device).wait_stream(side)
        torch.cuda.synchronize(p.device)
        steps_before = self.opt.step_count
        l0 = _lib.launch_count()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._static_loss = self._step_eager(*self._static_in)
        self.graph_launches = _lib.launch_count() - l0      # this library's kernel launches inside one replay
        # capturing enqueued nothing: undo the host-side bookkeeping of the captured step
        self.opt.step_count = steps_before
        self._graph = g
        return loss

    def release_graph(self):
        self._graph = None
        self._static_in = None
        self._static_loss = None

    def train_step(self, p, c, R=None):
        """p [B,2048,3] partial, c [B,16384,3] complete, R [B,3,3] rotation already applied to both (train.py:133-138).
        Returns the detached loss tensor (no host sync).  After capture() the step is a CUDA-graph replay."""
        g = getattr(self, "_graph", None)
        if g is not None and not self.exchange_off:
            sp, sc, sR = self._static_in
            if p.shape != sp.shape or c.shape != sc.shape or (R is None) != (sR is None):
                raise ValueError("train_step inputs differ in shape from the captured step: call capture() again (or release_graph())")
            if p.data_ptr() != sp.data_ptr():
                sp.copy_(p, non_blocking=True)
            if c.data_ptr() != sc.data_ptr():
                sc.copy_(c, non_blocking=True)
            if R is not None and R.data_ptr() != sR.data_ptr():
                sR.copy_(R, non_blocking=True)
            self.opt.sync_device_state()
            g.replay()
            self.opt.step_count += 1
            return self._static_loss
        return self._step_eager(p, c, R)

    def _step_eager(self, p, c, R=None):
        self.opt.sync_device_state()
        self.opt.zero_grad()
        coarse, dense = self.model(p, Rotate(R) if R is not None else None)
        loss = cd_loss_L1(coarse, c)
        if dense is not None:
            loss = loss + cd_loss_L1(dense, c)
        loss.backward()
        if self.exchange_off:          # measurement only (bench.py comm_exposed_ms): the step without its one exchange
            scale = 1.0
        elif self.exchange is not None:
            scale = self.exchange.finish()
        else:
            scale = exchange_gradients(self.opt.flat_g, self.world, self.pg)
        self.opt.step(grad_scale=scale)
        return loss.detach()
