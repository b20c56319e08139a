"""run by tests/test_gpu_compat_loop.py in a fresh interpreter with compat/ first on sys.path: the reference-shaped training loop
(train.py:60-93,127-186,252-277) written against the REFERENCE'S module paths -- models.model, metrics.loss, metrics.metric,
pytorch3d.transforms -- with plain torch.optim.Adam + StepLR, next to the package's DataParallelTrainer on a copy of the same model.
Prints one JSON line."""
import copy
import json
import sys
from types import SimpleNamespace

import torch
import torch.optim as Optim

from metrics.loss import cd_loss_L1
from metrics.metric import l1_cd, l2_cd, f_score
from models.model import PCNNet
from pytorch3d.transforms import Rotate, RotateAxisAngle, random_rotations
from utils.loss import calc_dcd
from utils.voxel_util import evaluate_iou

import vn_pointcloudcompletion_b200 as V
from vn_pointcloudcompletion_b200.synthetic import make_batch
from vn_pointcloudcompletion_b200.trainer import DataParallelTrainer

out = {}
V.set_gemm_mode(sys.argv[1] if len(sys.argv) > 1 else "fp32")
config = SimpleNamespace(num_coarse=1024, latent_dim=2048, only_coarse=False, device="cuda", enc_pretrained="none", lr=1e-4,
                         enc_type="vn_pointnet", dec_type="vn_foldingnet")
torch.manual_seed(0)
model = PCNNet(config, enc_type=config.enc_type, dec_type=config.dec_type)          # train.py:60
twin = copy.deepcopy(model)
before = [q.detach().clone() for q in model.parameters()]
optimizer = Optim.Adam(model.parameters(), lr=config.lr, betas=(0.9, 0.999))       # train.py:70
scheduler = Optim.lr_scheduler.StepLR(optimizer, step_size=2, gamma=0.8)           # train.py:93 (step_size shortened for the test)
trainer = DataParallelTrainer(twin, lr=config.lr, world_size=1)
tsched = trainer.make_scheduler(step_size=2, gamma=0.8)

p0, c0, _ = (torch.from_numpy(a) for a in make_batch(4, 256, 2048, seed=5))
gen = torch.Generator().manual_seed(1)
losses, tlosses = [], []
model.train()
for epoch in range(3):
    p, c = p0.to(config.device), c0.to(config.device)                              # train.py:128
    R = random_rotations(p.shape[0]) if epoch else RotateAxisAngle(angle=torch.rand(p.shape[0], generator=gen) * 360, axis="Z").R
    trot = Rotate(R=R).to(config.device)                                           # train.py:131-134
    p, c = trot.transform_points(p), trot.transform_points(c)                      # train.py:136-138
    optimizer.zero_grad()
    coarse_pred, dense_pred = model(p, trot)                                       # train.py:142
    loss = cd_loss_L1(coarse_pred, c) + cd_loss_L1(dense_pred, c)                  # train.py:146-164
    loss.backward()
    optimizer.step()
    scheduler.step()                                                               # train.py:186 (once per epoch)
    losses.append(loss.item())
    tlosses.append(trainer.train_step(p, c, trot.R).item())
    tsched.step()
    if epoch == 0:
        # after ONE step both models have seen bit-identical forwards; later steps diverge chaotically at random init (fp32 atomic-add
        # order in the backward kernels -> 1e-9 parameter differences -> flipped VNMaxPool near-ties)
        d1 = max(float((a - b).abs().max()) for a, b in zip(model.parameters(), twin.parameters()))
        # relative L2 distance of the two parameter UPDATES (Adam's first step is lr * g / (|g| + 1e-8): every element moves by ~lr, so
        # elements whose gradient is at the atomics' noise level differ by O(lr) -- they are compared through the optimizer state instead)
        num = sum(float(((a - q) - (b - q)).double().pow(2).sum()) for a, b, q in zip(model.parameters(), twin.parameters(), before))
        den = sum(float((a - q).double().pow(2).sum()) for a, q in zip(model.parameters(), before))
        out["step1_update_rel_l2"] = (num / den) ** 0.5
        d1b = max(float((a - b).abs().max()) for a, b in zip(model.buffers(), twin.buffers()) if a.dtype.is_floating_point)
        out["step1_param_diff"], out["step1_buffer_diff"] = d1, d1b
        s_ref, s_flat = optimizer.state_dict(), trainer.opt.state_dict()
        out["state_keys_equal"] = sorted(s_ref["state"].keys()) == sorted(s_flat["state"].keys())
        out["n_state"], out["n_params"] = len(s_flat["state"]), len(s_flat["param_groups"][0]["params"])
        worst = 0.0
        for k in s_ref["state"]:
            for key in ("exp_avg", "exp_avg_sq"):
                a, b = s_ref["state"][k][key], s_flat["state"][k][key]
                worst = max(worst, float((a - b).norm() / (a.norm() + 1e-30)))
            assert float(s_ref["state"][k]["step"]) == float(s_flat["state"][k]["step"]) == 1.0
        out["step1_state_rel_diff"] = worst
out["losses"], out["trainer_losses"] = losses, tlosses
out["lr"], out["trainer_lr"] = scheduler.get_last_lr()[0], tsched.get_last_lr()[0]
sd_ref, sd_flat = optimizer.state_dict(), trainer.opt.state_dict()
optimizer.load_state_dict(copy.deepcopy(sd_flat))          # FlatAdam checkpoint -> torch.optim.Adam
trainer.opt.load_state_dict(copy.deepcopy(sd_ref))         # torch.optim.Adam checkpoint (the reference's optim_last.pth) -> FlatAdam
out["resumed_step"] = trainer.opt.step_count
ck = trainer.optimizer_checkpoint(2, 0.5, 1)
out["ckpt_keys"] = sorted(ck.keys())

# validation / test loop pieces (train.py:199-226, test.py:54-78) through the reference's module paths
model.eval()
with torch.no_grad():
    coarse_pred, dense_pred = model(p, trot)
    out["l1_cd"], out["l2_cd"] = l1_cd(dense_pred, c).item(), l2_cd(dense_pred, c).item()
    out["f_score"] = f_score(dense_pred[0].detach().cpu().numpy(), c[0].detach().cpu().numpy())
    out["iou"] = evaluate_iou(dense_pred[0].detach().cpu().numpy(), c[0].detach().cpu().numpy())
    out["dcd"] = calc_dcd(coarse_pred, c, alpha=40, n_lambda=0.5)[0].mean().item()
    ev = trainer.evaluate([(p, c, trot.R)])
    out["trainer_eval"] = list(ev)
print(json.dumps(out))
