import sys, time, os
sys.path.insert(0, "/root/repo")
from types import SimpleNamespace
import torch
import vn_pointcloudcompletion_b200 as V
from vn_pointcloudcompletion_b200.synthetic import make_batch
from vn_pointcloudcompletion_b200.trainer import DataParallelTrainer
dev = torch.device("cuda", 0)
V.set_gemm_mode("tf32")
cfg = SimpleNamespace(num_coarse=1024, latent_dim=2048, only_coarse=False, device=dev, enc_pretrained="none")
torch.manual_seed(0)
net = V.PCNNet(cfg).train()
tr = DataParallelTrainer(net, lr=1e-4, world_size=1)
data = [tuple(torch.from_numpy(a).to(dev) for a in make_batch(32, 2048, 16384, seed=1234 + 1000 * i)) for i in range(2)]
for i in range(5): tr.train_step(*data[i % 2])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(20): tr.train_step(*data[i % 2])
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0) / 20:.2f} ms/step, total {1e3 * (t2 - t0) / 20:.2f} ms/step, cpus {os.cpu_count()}")
