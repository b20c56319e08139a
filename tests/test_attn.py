"""Transformer-refined VN folding decoder: VNLayerNorm, Attention, VN_Block, Attention_VN_FoldingNet.
CPU: the numpy oracle against the reference golden (tests/golden/attn_small.npz).  GPU: the CUDA path against the golden and
the oracle.  Tolerances: activations 1e-4 relative (fp32 mode); gradients relative L2 <= 5e-3 (conftest.assert_grad_close)."""
from types import SimpleNamespace

import numpy as np
import pytest

from conftest import assert_grad_close
from oracle import attn_oracle as AO

TOL = dict(rtol=1e-4, atol=2e-5)
# The full decoder stacks 2 LayerNorms + 4 BatchNorm-on-norm layers + 4 leaky projections per block on tokens that share
# most of their feature (the broadcast global feature): at random init the reference's own fp32 forward differs from a float64
# evaluation of the same network by up to 6e-4 (mean 5e-5) in `pts` (mask flips at <p,d> ~ 0 and norm-variance
# amplification; measured with the oracle in both dtypes).  Net-level outputs are therefore compared in relative L2
# (<= 1e-3) and max error (<= 1e-2 of the largest entry); every layer on its own is compared at 1e-4.
NET_L2, NET_MX = 1e-3, 1e-2
# Gradients of the full decoder are more sensitive still (a flipped leaky mask changes a whole token's contribution): the
# reference's fp32 autograd gradients differ from the float64 oracle's by 0.5 % - 3.7 % relative L2 per tensor on this fixture.
# Net-level gradient checks are therefore wiring checks at 12 % / 20 %; VN_Block, Attention and VNLayerNorm are checked on
# their own at the usual 5e-3.
NETG_L2, NETG_MX = 0.12, 0.2


def _digest(a):
    a = np.asarray(a, np.float64).ravel()
    return np.array([a.sum(), np.abs(a).sum(), (a * a).sum(), a[:: max(1, a.size // 97)][:64].sum()], np.float64)


def _params(g, key):
    pre = key + ".pre.sd."
    return {k[len(pre):]: g[k].copy() for k in g.files if k.startswith(pre)}


def _seeded_decoder(device="cpu"):
    import torch

    import vn_pointcloudcompletion_b200 as V
    cfg = SimpleNamespace(num_coarse=1024, latent_dim=2048, only_coarse=False, device=device, enc_pretrained="none")
    torch.manual_seed(0)
    return V.Attention_VN_FoldingNet(cfg)


# ------------------------------------------------------------------------------------------------ CPU: oracle vs golden
def test_layernorm_oracle(golden):
    g = golden("attn_small")
    P = _params(g, "VNLayerNorm")
    y, c = AO.vn_layernorm(g["VNLayerNorm.x"], P["layer_norm.weight"], P["layer_norm.bias"])
    np.testing.assert_allclose(y, g["VNLayerNorm.y"], **TOL)
    gx, gw, gb = AO.vn_layernorm_bwd(c, g["VNLayerNorm.gy"])
    assert_grad_close(gx, g["VNLayerNorm.gx"], "gx")
    assert_grad_close(gw, g["VNLayerNorm.grad.layer_norm.weight"], "gw")
    assert_grad_close(gb, g["VNLayerNorm.grad.layer_norm.bias"], "gb")


@pytest.mark.parametrize("key,H,scale", [("Attention", 4, 1.0), ("Attention_defscale", 2, 48 ** -0.5)])
def test_attention_oracle(golden, key, H, scale):
    g = golden("attn_small")
    P = _params(g, key)
    y, c = AO.attention(g[key + ".x"], P, "", H, scale)
    np.testing.assert_allclose(y, g[key + ".y"], **TOL)
    G = {}
    gx = AO.attention_bwd(c, P, "", g[key + ".gy"], G)
    assert_grad_close(gx, g[key + ".gx"], "gx")
    for n in ("proj_vnq", "proj_vnk", "proj_vnv", "proj_vn"):
        assert_grad_close(G[n + ".map_to_feat.weight"], g[f"{key}.grad.{n}.map_to_feat.weight"], n)
    assert g[key + ".grad.qkv.weight"].size == 0 and g[key + ".grad.proj.weight"].size == 0      # unused by the reference's forward


def test_block_oracle(golden):
    g = golden("attn_small")
    P = _params(g, "VN_Block")
    x = AO.block_tokens_to_vn(g["VN_Block.x"])
    y, c = AO.vn_block(x, P, "", 4, 1.0)
    np.testing.assert_allclose(AO.block_vn_to_tokens(y), g["VN_Block.y"], **TOL)
    G = {}
    gx = AO.vn_block_bwd(c, P, "", AO.block_tokens_to_vn(g["VN_Block.gy"]), G)
    assert_grad_close(AO.block_vn_to_tokens(gx), g["VN_Block.gx"], "gx")
    for k in g.files:
        if k.startswith("VN_Block.grad."):
            n = k[len("VN_Block.grad."):]
            if g[k].size == 0:
                assert n not in G, n
            else:
                assert_grad_close(G[n], g[k], n)


def test_decoder_state_dict_matches_reference(golden):
    g = golden("attn_small")
    sd = _seeded_decoder().state_dict()
    ref_keys = sorted(k[len("sd_digest.decoder."):] for k in g.files if k.startswith("sd_digest.decoder."))
    assert sorted(sd.keys()) == ref_keys
    for k in ref_keys:
        np.testing.assert_allclose(_digest(sd[k].float().numpy()), g["sd_digest.decoder." + k], rtol=1e-12, err_msg=k)


def test_decoder_oracle(golden):
    g = golden("attn_small")
    P = {"decoder." + k: v.detach().numpy().copy() for k, v in _seeded_decoder().state_dict().items()}
    orc = AO.AttnFoldingOracle(P)
    pts = orc.forward(g["dec.coarse"], g["dec.fg"])
    assert_grad_close(pts, g["dec.pts"], "pts", NET_L2, NET_MX)
    G, gc, gfg = orc.backward(g["dec.w"])
    assert_grad_close(gc, g["dec.gcoarse"], "gcoarse", NETG_L2, NETG_MX)
    assert_grad_close(gfg, g["dec.gfg"], "gfg", NETG_L2, NETG_MX)
    for k in g.files:
        if k.startswith("grad.decoder."):
            assert_grad_close(G[k[5:]], g[k], k, NETG_L2, NETG_MX)
        elif k.startswith("grad_head.decoder."):
            assert_grad_close(G[k[10:]].ravel()[:256], g[k], k, NETG_L2, NETG_MX)
        elif k.startswith("grad_none.decoder."):
            assert k[10:] not in G, k
        elif k.startswith("buf_post.decoder.") and not k.endswith("num_batches_tracked"):
            np.testing.assert_allclose(P[k[9:]], g[k], rtol=1e-4, atol=1e-6, err_msg=k)


# ------------------------------------------------------------------------------------------------ GPU
def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _load(mod, g, key):
    import torch
    pre = key + ".pre.sd."
    mod.load_state_dict({k[len(pre):]: torch.from_numpy(g[k].copy()) for k in g.files if k.startswith(pre)})
    return mod.cuda().train()


def _check_layer(g, key, mod, l2=5e-3, mx=2e-2, **tol):
    import torch
    x = _dev(g[key + ".x"]).requires_grad_(True)
    y = mod(x)
    (y * _dev(g[key + ".gy"])).sum().backward()
    torch.cuda.synchronize()
    if tol.get("robust"):      # TF32 operands: a few elements near a leaky-mask boundary move by more than any element-wise bound
        assert_grad_close(y.detach().cpu().numpy(), g[key + ".y"], key + ".y", 2e-2, 1e-1)
    else:
        np.testing.assert_allclose(y.detach().cpu().numpy(), g[key + ".y"], **(tol or TOL))
    assert_grad_close(x.grad.cpu().numpy(), g[key + ".gx"], key + ".gx", l2, mx)
    for n, p in mod.named_parameters():
        ref = g[f"{key}.grad.{n}"]
        if ref.size == 0:
            assert p.grad is None, n
        else:
            assert_grad_close(p.grad.cpu().numpy(), ref, n, l2, mx)


@pytest.mark.gpu
def test_layernorm_gpu(golden):
    import vn_pointcloudcompletion_b200 as V
    g = golden("attn_small")
    _check_layer(g, "VNLayerNorm", _load(V.VNLayerNorm(48), g, "VNLayerNorm"))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fp32", "tf32"])
def test_attention_gpu(golden, mode):
    import vn_pointcloudcompletion_b200 as V
    g = golden("attn_small")
    V.set_gemm_mode(mode)
    try:
        kw = {} if mode == "fp32" else dict(l2=5e-2, mx=2e-1, robust=True)
        _check_layer(g, "Attention", _load(V.Attention(64, num_heads=4, qk_scale=1), g, "Attention"), **kw)
        _check_layer(g, "Attention_defscale", _load(V.Attention(96, num_heads=2), g, "Attention_defscale"), **kw)
        _check_layer(g, "VN_Block", _load(V.VN_Block(dim=64, num_heads=4, mlp_ratio=1, qk_scale=1), g, "VN_Block"), **kw)
    finally:
        V.set_gemm_mode("fp32")


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,H,D", [(2, 64, 2, 16), (1, 200, 3, 32), (2, 130, 8, 48), (3, 1, 1, 16)])
def test_attention_core_vs_oracle(B, N, H, D):
    """the flash-style attention kernels against the numpy oracle on ragged token counts (tile = 64)"""
    import torch

    from vn_pointcloudcompletion_b200 import ops
    rng = np.random.RandomState(B * 100 + N)
    C = H * D
    q, k, v = (rng.standard_normal((B, C, 3, N)).astype(np.float32) * 0.3 for _ in range(3))
    gy = rng.standard_normal((B, C, 3, N)).astype(np.float32)
    o, c = AO.attention_core(q, k, v, H, 0.7)
    gq, gk, gv = AO.attention_core_bwd(c, gy)
    rows = lambda a: np.ascontiguousarray(a.transpose(0, 3, 2, 1)).reshape(B * N * 3, -1)      # noqa: E731
    qkv = _dev(np.concatenate([rows(q), rows(k), rows(v)], 1)).requires_grad_(True)
    out = ops.vn_attention(qkv, B, N, H, 0.7)
    (out * _dev(rows(gy))).sum().backward()
    torch.cuda.synchronize()
    np.testing.assert_allclose(out.detach().cpu().numpy(), rows(o), rtol=1e-4, atol=1e-5)
    got = qkv.grad.cpu().numpy()
    assert_grad_close(got[:, :C], rows(gq), "gq", 1e-4, 1e-3)
    assert_grad_close(got[:, C:2 * C], rows(gk), "gk", 1e-4, 1e-3)
    assert_grad_close(got[:, 2 * C:], rows(gv), "gv", 1e-4, 1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fp32", "tf32"])
def test_decoder_gpu(golden, mode):
    import torch

    import vn_pointcloudcompletion_b200 as V
    g = golden("attn_small")
    V.set_gemm_mode(mode)
    try:
        dec = _seeded_decoder("cuda").cuda().train()
        ci, fi = _dev(g["dec.coarse"]).requires_grad_(True), _dev(g["dec.fg"]).requires_grad_(True)
        pts = dec(ci, fi)
        (pts * _dev(g["dec.w"])).sum().backward()
        torch.cuda.synchronize()
        if mode == "tf32":
            # At random init this decoder amplifies a relative perturbation ~3x per layer (measured stage by stage with
            # tools/debug_attn.py: TF32 operand rounding 2e-3 after the first attention -> 0.35 after the first folding MLP; the
            # same mechanism turns fp32 rounding into the 6e-4 noted above).  A value comparison of the whole decoder in TF32
            # mode says nothing; its layers are compared one by one in test_attention_gpu[tf32].  Here: runs, finite, same shape.
            assert pts.shape == tuple(g["dec.pts"].shape) and torch.isfinite(pts).all()
            assert all(torch.isfinite(p.grad).all() for p in dec.parameters() if p.grad is not None)
            return
        assert_grad_close(pts.detach().cpu().numpy(), g["dec.pts"], "pts", NET_L2, NET_MX)
        l2, mx = NETG_L2, NETG_MX
        assert_grad_close(ci.grad.cpu().numpy(), g["dec.gcoarse"], "gcoarse", l2, mx)
        assert_grad_close(fi.grad.cpu().numpy(), g["dec.gfg"], "gfg", l2, mx)
        sd = dict(dec.named_parameters())
        for k in g.files:
            if k.startswith("grad.decoder."):
                assert_grad_close(sd[k[13:]].grad.cpu().numpy(), g[k], k, l2, mx)
            elif k.startswith("grad_head.decoder."):
                assert_grad_close(sd[k[18:]].grad.cpu().numpy().ravel()[:256], g[k], k, l2, mx)
            elif k.startswith("grad_none.decoder."):
                assert sd[k[18:]].grad is None, k
        if mode == "fp32":
            bufs = dict(dec.named_buffers())
            for k in g.files:
                if k.startswith("buf_post.decoder.") and not k.endswith("num_batches_tracked"):
                    np.testing.assert_allclose(bufs[k[17:]].cpu().numpy(), g[k], rtol=1e-4, atol=1e-6, err_msg=k)
            dec.eval()
            with torch.no_grad():
                pe = dec(ci.detach(), fi.detach())
            assert_grad_close(pe.cpu().numpy(), g["dec.eval_pts"], "eval_pts", NET_L2, NET_MX)
    finally:
        V.set_gemm_mode("fp32")


@pytest.mark.gpu
def test_pcnnet_pointnet_attention_decoder_trains():
    """PCNNet(enc_type='vn_pointnet', dec_type='attention_vn_foldingnet'): the pair that runs in the reference (SURVEY 8f f1 note)"""
    import torch

    import vn_pointcloudcompletion_b200 as V
    from vn_pointcloudcompletion_b200.synthetic import make_batch
    cfg = SimpleNamespace(num_coarse=1024, latent_dim=2048, only_coarse=False, device="cuda", enc_pretrained="none")
    torch.manual_seed(0)
    V.set_gemm_mode("tf32")
    try:
        net = V.PCNNet(cfg, enc_type="vn_pointnet", dec_type="attention_vn_foldingnet").train()
        p, c, R = make_batch(2, n_partial=512, n_gt=4096, seed=11)
        coarse, fine = net(_dev(p), V.Rotate(_dev(R)))
        assert coarse.shape == (2, 1024, 3) and fine.shape == (2, 16384, 3)
        loss = V.cd_loss_L1(coarse, _dev(c)) + V.cd_loss_L1(fine, _dev(c))
        loss.backward()
        assert np.isfinite(loss.item())
        used = [n for n, prm in net.named_parameters() if prm.grad is not None]
        assert any("transformer.1.attn.proj_vnq" in n for n in used) and any("vn_folding2.0.map_to_dir" in n for n in used)
        for n, prm in net.named_parameters():
            if prm.grad is not None:
                assert torch.isfinite(prm.grad).all(), n
    finally:
        V.set_gemm_mode("fp32")


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,H,ds_ws,store_p", [(1, 64, 1, True, True), (2, 130, 2, True, True), (1, 1024, 8, True, True), (2, 200, 3, True, True),
                                                 (2, 256, 2, False, False), (2, 256, 2, True, False), (1, 1024, 8, True, False), (2, 256, 2, True, True), (3, 128, 4, True, True)])
def test_attention_core_tf32_tensor_core(B, N, H, ds_ws, store_p):
    """csrc/attention_tc.cu (tcgen05 / TMEM, TF32 operands) against the numpy oracle.  Stated TF32 tolerance: operands carry a 10-bit
    mantissa, scores are sums of 144 products of O(1) features -> |dS| <~ 3e-3, i.e. a few 1e-3 relative error on softmax weights;
    outputs are compared at 1e-2 of the largest entry (rel-L2 5e-3), lse at 5e-3 absolute."""
    import torch

    import vn_pointcloudcompletion_b200 as V
    from vn_pointcloudcompletion_b200 import ops
    D = 48
    rng = np.random.RandomState(B * 100 + N)
    C = H * D
    q, k, v = (rng.standard_normal((B, C, 3, N)).astype(np.float32) * 0.3 for _ in range(3))
    o, c = AO.attention_core(q, k, v, H, 0.7)
    rows = lambda a: np.ascontiguousarray(a.transpose(0, 3, 2, 1)).reshape(B * N * 3, -1)      # noqa: E731
    qkv = _dev(np.concatenate([rows(q), rows(k), rows(v)], 1))
    gy = rng.standard_normal((B, C, 3, N)).astype(np.float32)
    gq, gk, gv = AO.attention_core_bwd(c, gy)
    qkv.requires_grad_(True)
    V.set_gemm_mode("tf32")
    # backward routes: store_p (N % 32 == 0): P kept by the forward, dV / dQ / dK as streaming GEMMs; else ds_ws: dK from the stored dS;
    # else the recomputing kernels
    ops._ATTN_DS_WORKSPACE, ops._ATTN_STORE_P = ds_ws, store_p
    try:
        out = ops.vn_attention(qkv, B, N, H, 0.7)
        torch.cuda.synchronize()
        assert ops._LAST_KERNEL[0] == "attention_fwd_tf32"
        out.backward(_dev(rows(gy)))
        torch.cuda.synchronize()
        assert ops._LAST_KERNEL[0] == "attention_bwd_tf32"
    finally:
        ops._ATTN_DS_WORKSPACE, ops._ATTN_STORE_P = True, True
        V.set_gemm_mode("fp32")
    assert_grad_close(out.detach().cpu().numpy(), rows(o), "out", 5e-3, 1e-2)
    got = qkv.grad.cpu().numpy()
    assert_grad_close(got[:, 2 * C:], rows(gv), "gv", 1e-2, 2e-2)
    assert_grad_close(got[:, :C], rows(gq), "gq", 1e-2, 2e-2)
    assert_grad_close(got[:, C:2 * C], rows(gk), "gk", 1e-2, 2e-2)
