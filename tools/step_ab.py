"""development aid: A/B timing of the BASELINE train step inside ONE process (box-to-box clock differences of several per cent make
separate bench.py runs useless for 1 % effects).  Each --set is 'label:python statements' executed with `ops`, `_lib`, `V` in scope;
the settings are visited round-robin `--rounds` times, `--steps` timed steps each.

    python tools/step_ab.py --set "a:_lib.raw('vnpcc_set_tuning', 0, 1)" --set "b:_lib.raw('vnpcc_set_tuning', 0, 0)"
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import torch
import vn_pointcloudcompletion_b200 as V
from vn_pointcloudcompletion_b200 import _lib, ops
from vn_pointcloudcompletion_b200.synthetic import make_batch
from vn_pointcloudcompletion_b200.trainer import DataParallelTrainer

ap = argparse.ArgumentParser()
ap.add_argument("--set", action="append", default=[])
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--rounds", type=int, default=4)
args = ap.parse_args()
dev = torch.device("cuda", 0)
V.set_gemm_mode("tf32")
cfg = SimpleNamespace(num_coarse=1024, latent_dim=2048, only_coarse=False, device=dev, enc_pretrained="none")
torch.manual_seed(0)
net = V.PCNNet(cfg, enc_type="vn_pointnet", dec_type="vn_foldingnet").train()
trainer = DataParallelTrainer(net, lr=1e-4, world_size=1)
data = [tuple(torch.from_numpy(a).to(dev) for a in make_batch(32, 2048, 16384, seed=1234 + 1000 * i)) for i in range(2)]
settings = [s.split(":", 1) for s in args.set] or [["default", "pass"]]
scope = {"ops": ops, "_lib": _lib, "V": V, "trainer": trainer}
for i in range(4):
    trainer.train_step(*data[i % 2])
res = {k: [] for k, _ in settings}
for r in range(args.rounds):
    for label, code in settings:
        exec(code, scope)
        for i in range(2):
            trainer.train_step(*data[i % 2])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            trainer.train_step(*data[i % 2])
        e1.record()
        torch.cuda.synchronize()
        res[label].append(e0.elapsed_time(e1) / args.steps)
for k, v in res.items():
    print(f"{k:24s} " + " ".join(f"{t:7.3f}" for t in v) + f"   median {sorted(v)[len(v) // 2]:.3f} ms/step")
