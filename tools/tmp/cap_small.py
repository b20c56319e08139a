import sys
sys.path.insert(0, "/root/repo")
from types import SimpleNamespace
import numpy as np, torch
import vn_pointcloudcompletion_b200 as V
from vn_pointcloudcompletion_b200.synthetic import make_batch
from vn_pointcloudcompletion_b200.trainer import DataParallelTrainer
V.set_gemm_mode("tf32")
cfg = SimpleNamespace(num_coarse=1024, latent_dim=2048, only_coarse=False, device="cuda:0", enc_pretrained="none")
torch.manual_seed(0)
net = V.PCNNet(cfg).train()
p, c, R = (torch.from_numpy(a).cuda() for a in make_batch(6, n_partial=64, n_gt=512, seed=3))
tr = DataParallelTrainer(net, lr=1e-4, world_size=1)
torch.autograd.set_detect_anomaly(True)
tr.capture(p, c, R, warmup=1)
print("captured", tr.graph_launches, float(tr.train_step(p, c, R)))
