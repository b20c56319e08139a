"""debug: stage-by-stage fp32 vs tf32 comparison of Attention_VN_FoldingNet on the golden inputs"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import numpy as np, torch
import vn_pointcloudcompletion_b200 as V
from vn_pointcloudcompletion_b200 import ops
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/attn_small.npz"))
cfg = SimpleNamespace(num_coarse=1024, latent_dim=2048, only_coarse=False, device="cuda", enc_pretrained="none")
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
res = {}
for mode in ("fp32", "tf32"):
    V.set_gemm_mode(mode)
    torch.manual_seed(0)
    dec = V.Attention_VN_FoldingNet(cfg).cuda().train()
    st = {}
    coarse, fg = dev(g["dec.coarse"]), dev(g["dec.fg"])
    B, N, _ = coarse.shape
    S = 16
    centers = coarse.reshape(B, 3, N).transpose(1, 2).contiguous().view(B * N * 3, 1)
    fg_rows = fg.squeeze(-1).transpose(1, 2).reshape(B * 3, -1)
    dg = ops.linear_rows(fg_rows, dec.downsize_global.map_to_feat.weight); st["dg"] = dg
    ones = torch.ones((dg.shape[1], 1), device="cuda")
    tok = ops.linear_rows(centers, ones, dg, 3 * N); st["tok"] = tok
    for i, blk in enumerate(dec.transformer):
        n1 = blk.norm1.forward_rows(tok); st[f"b{i}.n1"] = n1
        a = blk.attn.forward_rows(n1, B, N); st[f"b{i}.attn"] = a
        x1 = ops.rows_add(tok, a)
        n2 = blk.norm2.forward_rows(x1); st[f"b{i}.n2"] = n2
        h3 = blk.conv3.forward_rows(n2); st[f"b{i}.h3"] = h3
        h4 = blk.conv4.forward_rows(h3); st[f"b{i}.h4"] = h4
        tok = ops.rows_add(x1, h4)
    T = B * N
    seed = dec.folding_seed.cuda().t().contiguous()
    local1 = seed.unsqueeze(0).expand(T, S, 3).reshape(T * S * 3, 1)
    fd1 = dec._fold(dec.vn_folding1, local1, tok, T, S, True); st["fd1"] = fd1
    with torch.no_grad():
        fd1n = dec._fold(dec.vn_folding1, local1, tok, T, S, True); st["fd1_nograd"] = fd1n
    res[mode] = {k: v.detach().float().cpu().numpy() for k, v in st.items()}
for k in res["fp32"]:
    a, b = res["fp32"][k], res["tf32"][k]
    print(f"{k:12s} max|fp32| {np.abs(a).max():.4f}  relL2 {np.linalg.norm(a - b) / np.linalg.norm(a):.3e}  max {np.abs(a - b).max():.3e}")
