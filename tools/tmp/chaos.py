import sys
sys.path.insert(0, "/root/repo")
from types import SimpleNamespace
import torch
import vn_pointcloudcompletion_b200 as V
from vn_pointcloudcompletion_b200.synthetic import make_batch
from vn_pointcloudcompletion_b200.trainer import DataParallelTrainer
V.set_gemm_mode("tf32")
cfg = SimpleNamespace(num_coarse=1024, latent_dim=2048, only_coarse=False, device="cuda", enc_pretrained="none")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
data = [tuple(torch.from_numpy(x).cuda() for x in make_batch(B, 512, 4096, seed=50 + i)) for i in range(2)]
for mode in ("eager", "eager", "graph", "graph"):
    torch.manual_seed(0)
    net = V.PCNNet(cfg).train()
    tr = DataParallelTrainer(net, lr=1e-4, world_size=1)
    losses = []
    start = 0
    if mode == "graph":
        losses.append(float(tr.capture(*data[0], warmup=1)))
        start = 1
    for i in range(start, 7):
        losses.append(float(tr.train_step(*data[i % 2])))
    print(mode, " ".join(f"{l:.5f}" for l in losses), flush=True)
