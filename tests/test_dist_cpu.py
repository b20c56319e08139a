"""world_size-2 gloo tests (CPU) of the host-side data-parallel logic: the flat-gradient exchange of
vn_pointcloudcompletion_b200/trainer.py and the per-rank data sharding.  The kernels themselves need a GPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vn_pointcloudcompletion_b200.synthetic import make_batch
    from vn_pointcloudcompletion_b200.trainer import exchange_gradients, rank_seed
    g = torch.arange(1000, dtype=torch.float32) * (rank + 1)          # rank-dependent "gradient"
    scale = exchange_gradients(g, world)
    mean = g * scale
    p, c, R = make_batch(2, 64, 128, seed=rank_seed(1234, rank))
    q.put((rank, mean.numpy(), float(p.sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_exchange_and_sharding_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out.sort(key=lambda t: t[0])
    want = np.arange(1000, dtype=np.float32) * 1.5                      # mean of (1x, 2x)
    for _, mean, _ in out:
        np.testing.assert_allclose(mean, want, rtol=1e-6)
    assert out[0][2] != out[1][2]                                       # ranks see different shards


def test_exchange_is_identity_for_world1():
    from vn_pointcloudcompletion_b200.trainer import exchange_gradients
    g = torch.ones(8)
    assert exchange_gradients(g, 1) == 1.0 and torch.equal(g, torch.ones(8))


def _overlap_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch.nn as nn

    from vn_pointcloudcompletion_b200.trainer import OverlappedExchange
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(6, 5), nn.Tanh(), nn.Linear(5, 4), nn.Tanh(), nn.Linear(4, 3))
    unused = nn.Parameter(torch.ones(7))                      # sits in the tail bucket but never receives a gradient
    params = list(net.parameters()) + [unused]
    n = sum(p.numel() for p in params)
    flat = torch.zeros(n)
    offs, o = {}, 0
    for p in params:
        p.grad = flat[o:o + p.numel()].view_as(p)
        offs[id(p)] = o
        o += p.numel()
    split = offs[id(net[2].weight)]                          # tail = the last two Linear layers + the unused parameter
    ex = OverlappedExchange(flat, [(p, offs[id(p)], p.numel()) for p in params], split, world)
    results = []
    for step in range(3):                                    # step 0 calibrates (plain exchange), steps 1-2 overlap
        x = torch.randn(8, 6, generator=torch.Generator().manual_seed(100 * step + rank))
        # the local gradient, taken with the hooks off (once the tail's all-reduce is in flight the buffer is being overwritten in place)
        ex.enabled = False
        flat.zero_()
        net(x).square().sum().backward()
        local = flat.clone()
        ex.enabled = True
        flat.zero_()
        net(x).square().sum().backward()
        launched_early = ex.work is not None
        scale = ex.finish()
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        results.append((launched_early, float((flat * scale - sum(gathered) / world).abs().max()), len(ex.tail_ids), ex.tail_range, ex.head_ranges))
    q.put((rank, results))
    dist.barrier()
    dist.destroy_process_group()


def test_overlapped_exchange_world2():
    """the early (during-backward) all-reduce of the tail bucket + the late one of the head equal the plain mean over ranks; the number of
    gradient arrivals is calibrated on the first step (a tail parameter without gradient does not block the launch)"""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_overlap_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, results in out:
        assert [r[0] for r in results] == [False, True, True]
        assert all(r[1] < 1e-6 for r in results)
        assert results[-1][2] == 4                             # weight + bias of the two tail layers
        assert results[-1][3] == (35, 35 + 24 + 15) and results[-1][4] == [(0, 35)]      # the unused tail parameter is not exchanged


def _overlap_mid_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch.nn as nn

    from vn_pointcloudcompletion_b200.trainer import OverlappedExchange
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(6, 5), nn.Tanh(), nn.Linear(5, 4), nn.Tanh(), nn.Linear(4, 3))
    params = list(net.parameters())
    flat = torch.zeros(sum(p.numel() for p in params))
    offs, o = {}, 0
    for p in params:
        p.grad = flat[o:o + p.numel()].view_as(p)
        offs[id(p)] = o
        o += p.numel()
    # three buckets: layer 0 (after backward), layer 2 (middle, early), layer 4 (tail, earliest)
    ex = OverlappedExchange(flat, [(p, offs[id(p)], p.numel()) for p in params], offs[id(net[4].weight)], world,
                            extra_splits=[offs[id(net[2].weight)]])
    results = []
    for step in range(3):
        x = torch.randn(8, 6, generator=torch.Generator().manual_seed(100 * step + rank))
        ex.enabled = False
        flat.zero_()
        net(x).square().sum().backward()
        local = flat.clone()
        ex.enabled = True
        flat.zero_()
        net(x).square().sum().backward()
        early = (ex.work is not None, [m["work"] is not None for m in ex.mid])
        scale = ex.finish()
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        results.append((early, float((flat * scale - sum(gathered) / world).abs().max()), [m["range"] for m in ex.mid], ex.head_ranges))
    q.put((rank, results))
    dist.barrier()
    dist.destroy_process_group()


def test_overlapped_exchange_three_buckets_world2():
    """tail + one middle bucket are reduced during backward, only the bucket of the first-executed layer after it; result = plain mean"""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_overlap_mid_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, results in out:
        assert results[0][0] == (False, [])                            # calibration step: nothing early
        assert results[1][0] == (True, [True]) and results[2][0] == (True, [True])
        assert all(r[1] < 1e-6 for r in results)
        assert results[-1][2] == [(35, 35 + 24)] and results[-1][3] == [(0, 35)]


def _overlap_fallback_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch.nn as nn

    from vn_pointcloudcompletion_b200.trainer import OverlappedExchange
    torch.manual_seed(0)
    a, b, c = nn.Linear(4, 4), nn.Linear(4, 4), nn.Linear(4, 4)
    params = list(a.parameters()) + list(b.parameters()) + list(c.parameters())
    flat = torch.zeros(sum(p.numel() for p in params))
    offs, o = {}, 0
    for p in params:
        p.grad = flat[o:o + p.numel()].view_as(p)
        offs[id(p)] = o
        o += p.numel()
    ex = OverlappedExchange(flat, [(p, offs[id(p)], p.numel()) for p in params], offs[id(b.weight)], world)
    x = torch.randn(3, 4, generator=torch.Generator().manual_seed(rank))

    def run(full):
        flat.zero_()
        h = b(a(x))
        (c(h) if full else h).square().sum().backward()

    out = {}
    run(False)                                   # calibration on the SHORT graph: the tail is layer b only
    ex.finish()
    out["tail_ids"] = len(ex.tail_ids)
    # a step with fewer arrivals than calibrated (only layer a): nothing is launched early, finish() reduces everything
    flat.zero_()
    a(x).square().sum().backward()
    local = flat.clone()
    early = ex.work is not None
    scale = ex.finish()
    g = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(g, local)
    out["fewer"] = (early, float((flat * scale - sum(g) / world).abs().max()))
    # a step with MORE arrivals (layer c joins after the tail was launched): must fail loudly, not silently mix reduced and local values
    run(True)
    try:
        ex.finish()
        out["more"] = "no error"
    except RuntimeError as e:
        out["more"] = "raised" if "autograd graph changed" in str(e) else str(e)
    if ex.work is not None:
        ex.work.wait()
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_overlapped_exchange_deviating_steps_world2():
    """after calibration, a step with fewer gradient arrivals falls back to the plain exchange (correct mean), a step with more arrivals
    than calibrated raises instead of mixing already-reduced and local gradients"""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_overlap_fallback_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, o in out:
        assert o["tail_ids"] == 2
        assert o["fewer"][0] is False and o["fewer"][1] < 1e-6
        assert o["more"] == "raised"
