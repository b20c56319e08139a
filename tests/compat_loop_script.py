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
out["losses"], out["trainer_losses"] = losses, tlosses
out["lr"], out["trainer_lr"] = scheduler.get_last_lr()[0], tsched.get_last_lr()[0]
diff, scale = 0.0, 0.0
for (n, a), (_, b) in zip(model.named_parameters(), twin.named_parameters()):
    diff = max(diff, float((a - b).abs().max()))
    scale = max(scale, float(a.abs().max()))
out["max_param_diff"], out["max_param"] = diff, scale
for (n, a), (_, b) in zip(model.named_buffers(), twin.named_buffers()):
    if a.dtype.is_floating_point:
        diff = max(diff, float((a - b).abs().max()))
out["max_param_or_buffer_diff"] = diff

# optimizer checkpoints cross-load (train.py:72-80, 262-277)
sd_ref, sd_flat = optimizer.state_dict(), trainer.opt.state_dict()
out["state_keys_equal"] = sorted(sd_ref["state"].keys()) == sorted(sd_flat["state"].keys())
out["n_state"], out["n_params"] = len(sd_flat["state"]), len(sd_flat["param_groups"][0]["params"])
worst = 0.0
for k in sd_ref["state"]:
    for key in ("exp_avg", "exp_avg_sq"):
        a, b = sd_ref["state"][k][key], sd_flat["state"][k][key]
        worst = max(worst, float((a - b).abs().max() / (a.abs().max() + 1e-30)))
    assert float(sd_ref["state"][k]["step"]) == float(sd_flat["state"][k]["step"])
out["state_rel_diff"] = worst
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
