"""development aid: the trainer test's loss trajectory under each Chamfer search mode and GEMM mode (same seeds)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import torch
import vn_pointcloudcompletion_b200 as V
from vn_pointcloudcompletion_b200 import _lib
from vn_pointcloudcompletion_b200.synthetic import make_batch
from vn_pointcloudcompletion_b200.trainer import DataParallelTrainer

for gm in ("tf32", "fp32"):
    for cm in (2, 1, 2):
        V.set_gemm_mode(gm)
        _lib.load().vnpcc_chamfer_set_packed_math(cm)
        cfg = SimpleNamespace(num_coarse=1024, latent_dim=2048, only_coarse=False, device="cuda", enc_pretrained="none")
        torch.manual_seed(0)
        net = V.PCNNet(cfg).train()
        tr = DataParallelTrainer(net, lr=float(sys.argv[1]) if len(sys.argv) > 1 else 1e-3, world_size=1)
        p, c, R = (torch.from_numpy(a).cuda() for a in make_batch(4, 256, 2048, seed=21))
        losses = [tr.train_step(p, c, R).item() for _ in range(12)]
        print(gm, "chamfer mode", cm, " ".join(f"{x:.4f}" for x in losses), flush=True)
