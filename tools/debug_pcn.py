"""debug: compare encoder/decoder intermediates with the numpy oracle (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import numpy as np, torch
import vn_pointcloudcompletion_b200 as V
from vn_pointcloudcompletion_b200 import ops
from oracle import vn_oracle as O
from vn_pointcloudcompletion_b200.synthetic import make_batch

V.set_gemm_mode("fp32")
cfg = SimpleNamespace(num_coarse=1024, latent_dim=2048, only_coarse=False, device="cuda", enc_pretrained="none")
torch.manual_seed(0)
net = V.PCNNet(cfg).train()
P = {k: v.detach().cpu().numpy().copy() for k, v in net.state_dict().items()}
p, c, R = make_batch(2, n_partial=128, n_gt=1024, seed=7)
orc = O.PCNNetOracle(P)
oc, of = orc.forward(p, R, training=True)
ch = orc.enc.cache
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
def rows_np(a):  # logical [B,C,3,N] -> rows
    return np.ascontiguousarray(np.transpose(a, (0, 3, 2, 1))).reshape(-1, a.shape[1])
def cmp(name, mine, ref):
    mine = mine.detach().cpu().numpy()
    err = np.abs(mine - ref).max() / (np.abs(ref).max() + 1e-30)
    print(f"{name:10s} max rel-to-max err {err:.3e}  shape {mine.shape}")
enc = net.encoder
B, N = 2, 128
x0 = dev(p).view(B * N * 3, 1)
f0 = enc.first_conv[0].forward_rows(x0); cmp("f0", f0, rows_np(ch["f0"]))
f1 = ops.linear_rows(f0, enc.first_conv[1].map_to_feat.weight); cmp("f1", f1, rows_np(ch["f1"]))
enc.maxpool1.forced_idx = dev(orc.enc.idx[0]).reshape(2, -1)
g1 = enc.maxpool1.forward_rows(f1, B, N)
og1 = np.take_along_axis(ch["f1"], np.broadcast_to(ch["idx1"][..., None], ch["f1"].shape[:3] + (1,)), axis=-1)[..., 0]  # [B,C,3]
cmp("g1", g1, np.transpose(og1, (0, 2, 1)).reshape(B * 3, -1))
l0 = enc.second_conv[0]
wcat = torch.cat([l0.map_to_feat.weight, l0.map_to_dir.weight], dim=0)
bias = ops.linear_rows(g1, wcat[:, :512])
pd = ops.linear_rows(f1, wcat[:, 512:], bias, 3 * N)
x, pp, dd, bnc = ch["c2"]
praw = O.vn_linear(ch["cat"], P["encoder.second_conv.0.map_to_feat.weight"])
cmp("p_raw", pd[:, :1024], rows_np(praw)); cmp("d", pd[:, 1024:], rows_np(dd))
f2 = ops.bn_leaky(pd, None, l0.batchnorm.bn, True, 0.2, stacked=True); cmp("f2", f2, rows_np(ch["f2"]))
f3 = ops.linear_rows(f2, enc.second_conv[1].map_to_feat.weight); cmp("f3", f3, rows_np(ch["f3"]))
enc.maxpool2.forced_idx = dev(orc.enc.idx[1]).reshape(2, -1)
fg = enc.maxpool2.forward_rows(f3, B, N); cmp("fg", fg, np.transpose(ch["fg"][..., 0], (0, 2, 1)).reshape(B * 3, -1))
m0 = enc.mlp[0].forward_rows(fg); cmp("m0", m0, np.transpose(ch["m0"][..., 0], (0, 2, 1)).reshape(B * 3, -1))
m1 = enc.mlp[1].forward_rows(m0); cmp("m1", m1, np.transpose(ch["m1"][..., 0], (0, 2, 1)).reshape(B * 3, -1))
# ---- decoder
dec = net.decoder
coarse_o, fg_o = oc, ch["fg"]
od = orc.dec
dc = od.cache
coarse_t = dev(oc)
fg_t = dev(fg_o)
S, nc, nd = 16, 1024, 16384
seed_pts = dec.folding_seed.to("cuda").squeeze(1).transpose(1, 2)
seed_pts = V.Rotate(dev(R)).transform_points(seed_pts).expand(B, S, 3)
local = torch.stack([seed_pts[:, None, :, :].expand(B, nc, S, 3), coarse_t[:, :, None, :].expand(B, nc, S, 3)], dim=-1).reshape(B * nd * 3, 2)
fg_rows = fg_t.squeeze(-1).transpose(1, 2).reshape(B * 3, -1)
l0, l1, l2 = dec.final_conv
wcat = torch.cat([l0.map_to_feat.weight, l0.map_to_dir.weight], dim=0)
bias = ops.linear_rows(fg_rows, wcat[:, :2048])
Wf, Wd = P["decoder.final_conv.0.map_to_feat.weight"], P["decoder.final_conv.0.map_to_dir.weight"]
x_, p_, d_, bnc_ = dc["c0"]
obias = np.concatenate([O.vn_linear(fg_o, Wf[:, :2048]), O.vn_linear(fg_o, Wd[:, :2048])], axis=1)  # [B,512,3,1]
cmp("bias", bias, np.transpose(obias[..., 0], (0, 2, 1)).reshape(B * 3, -1))
cmp("local", local, rows_np(x_[:, 2048:]))
pd = ops.linear_rows(local, wcat[:, 2048:], bias, 3 * nd)
cmp("d0", pd[:, 256:], rows_np(d_))
cmp("p0raw", pd[:, :256], rows_np(O.vn_linear(x_, Wf)))
h = ops.bn_leaky(pd, None, l0.batchnorm.bn, True, 0.2, stacked=True)
x1_, p1_, d1_, bnc1_ = dc["c1"]
cmp("h0", h, rows_np(x1_))
h1 = l1.forward_rows(h); cmp("h1", h1, rows_np(dc["h1"]))
fine = ops.rows_dot(h1, l2.map_to_feat.weight, local[:, 1]).view(B, nd, 3); cmp("fine", fine, of)
fine2 = ops.rows_dot(h1, l2.map_to_feat.weight, None).view(B, nd, 3); cmp("h2", fine2, of - np.repeat(oc, 16, axis=1))
