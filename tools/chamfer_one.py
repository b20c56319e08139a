import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vn_pointcloudcompletion_b200 as V
g = torch.Generator(device="cuda").manual_seed(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
M = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
a = torch.rand(32, N, 3, device="cuda", generator=g) - 0.5
b = torch.rand(32, M, 3, device="cuda", generator=g) - 0.5
for _ in range(3):
    V.chamfer_3DFunction.apply(a, b)
torch.cuda.synchronize()
