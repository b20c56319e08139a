"""oracle/attn_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of the reference's transformer-refined VN folding decoder, forward and backward.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

Reference citations (file:line under /root/reference):
  VNLayerNorm                 models/vn_layers.py:129-150   (nn.LayerNorm over channels of the vector norms: eps 1e-5, affine)
  Attention                   models/transformer.py:73-105
  VN_Block (knn_index=None)   models/transformer.py:25-71
  Attention_VN_FoldingNet     models/pcn.py:392-520

Pinned by tests/golden/attn_small.npz (the reference's own classes run unmodified on CPU, outputs and autograd gradients).
"""
from __future__ import annotations

import numpy as np

from . import vn_oracle as O

LN_EPS = 1e-5


# ---------------------------------------------------------------------------------------------- VNLayerNorm
def vn_layernorm(x, w, b):
    """x [B,C,3,N] -> (y, cache)"""
    r = np.sqrt((x * x).sum(2))                                  # [B,C,N]
    n = r + O.EPS
    mean = n.mean(1, keepdims=True, dtype=np.float64).astype(x.dtype)
    var = n.var(1, keepdims=True, dtype=np.float64).astype(x.dtype)
    rstd = 1.0 / np.sqrt(var + LN_EPS)
    nh = (n - mean) * rstd
    l = nh * w[None, :, None] + b[None, :, None]
    y = x / n[:, :, None] * l[:, :, None]
    return y, (x, r, n, nh, l, rstd, w)


def vn_layernorm_bwd(cache, g):
    """returns (gx, gw, gb)"""
    x, r, n, nh, l, rstd, w = cache
    a = (g * x).sum(2) / n
    gw = (a * nh).sum((0, 2), dtype=np.float64).astype(x.dtype)
    gb = a.sum((0, 2), dtype=np.float64).astype(x.dtype)
    dnh = a * w[None, :, None]
    m1 = dnh.mean(1, keepdims=True, dtype=np.float64).astype(x.dtype)
    m2 = (dnh * nh).mean(1, keepdims=True, dtype=np.float64).astype(x.dtype)
    dn = rstd * (dnh - m1 - nh * m2) - a * l / n
    with np.errstate(divide="ignore", invalid="ignore"):
        unit = np.where(r[:, :, None] > 0, x / r[:, :, None], 0.0).astype(x.dtype)
    gx = g * (l / n)[:, :, None] + dn[:, :, None] * unit
    return gx, gw, gb


# ---------------------------------------------------------------------------------------------- attention
def _heads(t, H):
    """[B,C,3,N] -> [B,H,N,(C/H)*3]   (transformer.py:89)"""
    B, C, _, N = t.shape
    return t.reshape(B, H, C // H, 3, N).transpose(0, 1, 4, 2, 3).reshape(B, H, N, -1)


def _unheads(t, C):
    """[B,H,N,(C/H)*3] -> [B,C,3,N]   (transformer.py:98-99)"""
    B, H, N, _ = t.shape
    return t.transpose(0, 2, 1, 3).reshape(B, N, C, 3).transpose(0, 2, 3, 1)


def attention_core(q, k, v, H, scale):
    """q,k,v [B,C,3,N] -> (o [B,C,3,N], cache)"""
    C = q.shape[1]
    qh, kh, vh = _heads(q, H), _heads(k, H), _heads(v, H)
    s = np.matmul(qh, kh.transpose(0, 1, 3, 2)) * scale
    s = s - s.max(-1, keepdims=True)
    p = np.exp(s)
    p = p / p.sum(-1, keepdims=True)
    o = np.matmul(p, vh)
    return np.ascontiguousarray(_unheads(o, C)), (qh, kh, vh, p, H, scale, C)


def attention_core_bwd(cache, g):
    """returns (gq, gk, gv) in [B,C,3,N]"""
    qh, kh, vh, p, H, scale, C = cache
    go = _heads(g, H)
    gv = np.matmul(p.transpose(0, 1, 3, 2), go)
    gp = np.matmul(go, vh.transpose(0, 1, 3, 2))
    gs = p * (gp - (gp * p).sum(-1, keepdims=True)) * scale
    gq = np.matmul(gs, kh)
    gk = np.matmul(gs.transpose(0, 1, 3, 2), qh)
    return tuple(np.ascontiguousarray(_unheads(t, C)) for t in (gq, gk, gv))


def attention(x, P, pf, H, scale):
    """Attention.forward (transformer.py:87-103); P[pf + 'proj_vn{q,k,v,}.map_to_feat.weight']"""
    Wq, Wk, Wv, Wo = (P[pf + n + ".map_to_feat.weight"] for n in ("proj_vnq", "proj_vnk", "proj_vnv", "proj_vn"))
    q, k, v = O.vn_linear(x, Wq), O.vn_linear(x, Wk), O.vn_linear(x, Wv)
    o, c = attention_core(q, k, v, H, scale)
    return O.vn_linear(o, Wo), (x, o, c)


def attention_bwd(cache, P, pf, g, G):
    x, o, c = cache
    Wq, Wk, Wv, Wo = (P[pf + n + ".map_to_feat.weight"] for n in ("proj_vnq", "proj_vnk", "proj_vnv", "proj_vn"))
    go, G[pf + "proj_vn.map_to_feat.weight"] = O.vn_linear_bwd(o, Wo, g)
    gq, gk, gv = attention_core_bwd(c, go)
    gx = 0
    for n, W, gg in (("proj_vnq", Wq, gq), ("proj_vnk", Wk, gk), ("proj_vnv", Wv, gv)):
        gxi, G[pf + n + ".map_to_feat.weight"] = O.vn_linear_bwd(x, W, gg)
        gx = gx + gxi
    return gx


# ---------------------------------------------------------------------------------------------- VN_Block
def _vnll(P, name, x, training, update_running):
    bn = O.bn_from_params(P, name + ".batchnorm.bn")
    y, c = O.vn_linear_leaky_relu(x, P[name + ".map_to_feat.weight"], P[name + ".map_to_dir.weight"], bn, training, update_running=update_running)
    O.bn_to_params(P, name + ".batchnorm.bn", bn)
    return y, c


def _vnll_bwd(P, name, cache, g, G):
    r = O.vn_linear_leaky_relu_bwd(cache, P[name + ".map_to_feat.weight"], P[name + ".map_to_dir.weight"], g)
    G[name + ".map_to_feat.weight"], G[name + ".map_to_dir.weight"] = r["gWf"], r["gWd"]
    G[name + ".batchnorm.bn.weight"], G[name + ".batchnorm.bn.bias"] = r["gweight"], r["gbias"]
    return r["gx"]


def vn_block(x, P, pf, H, scale, training=True, update_running=True):
    """VN_Block.forward with knn_index=None on the VN view: x [B,C,3,N] -> (y, cache)"""
    n1, c1 = vn_layernorm(x, P[pf + "norm1.layer_norm.weight"], P[pf + "norm1.layer_norm.bias"])
    a, ca = attention(n1, P, pf + "attn.", H, scale)
    x1 = x + a
    n2, c2 = vn_layernorm(x1, P[pf + "norm2.layer_norm.weight"], P[pf + "norm2.layer_norm.bias"])
    h3, c3 = _vnll(P, pf + "conv3", n2, training, update_running)
    h4, c4 = _vnll(P, pf + "conv4", h3, training, update_running)
    return x1 + h4, (c1, ca, c2, c3, c4)


def vn_block_bwd(cache, P, pf, g, G):
    c1, ca, c2, c3, c4 = cache
    gh3 = _vnll_bwd(P, pf + "conv4", c4, g, G)
    gn2 = _vnll_bwd(P, pf + "conv3", c3, gh3, G)
    gx1, G[pf + "norm2.layer_norm.weight"], G[pf + "norm2.layer_norm.bias"] = vn_layernorm_bwd(c2, gn2)
    gx1 = gx1 + g
    gn1 = attention_bwd(ca, P, pf + "attn.", gx1, G)
    gx, G[pf + "norm1.layer_norm.weight"], G[pf + "norm1.layer_norm.bias"] = vn_layernorm_bwd(c1, gn1)
    return gx + gx1


def block_tokens_to_vn(x):
    """[B,N,C*3] -> [B,C,3,N]   (transformer.py:46-47)"""
    B, N, C3 = x.shape
    return np.swapaxes(x, 1, 2).reshape(B, C3 // 3, 3, N)


def block_vn_to_tokens(v):
    B, C, _, N = v.shape
    return np.ascontiguousarray(np.swapaxes(v.reshape(B, C * 3, N), 1, 2))


# ---------------------------------------------------------------------------------------------- decoder
def folding_seed_attn(grid_size=4, dtype=np.float32):
    """models/pcn.py:450-454: [3,S] grid in the xy plane, z = 0, extent +-1"""
    lin = np.linspace(-1.0, 1.0, grid_size, dtype=dtype)
    a = np.broadcast_to(lin[None, :], (grid_size, grid_size)).reshape(1, -1)
    b = np.broadcast_to(lin[:, None], (grid_size, grid_size)).reshape(1, -1)
    return np.concatenate([a, b, np.zeros_like(a)], 0).astype(dtype)


class AttnFoldingOracle:
    """Attention_VN_FoldingNet (models/pcn.py:392-520); prefix 'decoder.' in the PCNNet state_dict."""

    H, SCALE = 8, 1.0

    def __init__(self, P, prefix="decoder.", grid_size=4):
        self.P, self.pf, self.gs = P, prefix, grid_size

    def _fold_fwd(self, name, x, training, update_running):
        P, pf = self.P, self.pf
        h0, c0 = _vnll(P, pf + name + ".0", x, training, update_running)
        h1, c1 = _vnll(P, pf + name + ".1", h0, training, update_running)
        return O.vn_linear(h1, P[pf + name + ".2.map_to_feat.weight"]), (c0, c1, h1)

    def _fold_bwd(self, name, cache, g, G):
        P, pf = self.P, self.pf
        c0, c1, h1 = cache
        gh1, G[pf + name + ".2.map_to_feat.weight"] = O.vn_linear_bwd(h1, P[pf + name + ".2.map_to_feat.weight"], g)
        gh0 = _vnll_bwd(P, pf + name + ".1", c1, gh1, G)
        return _vnll_bwd(P, pf + name + ".0", c0, gh0, G)

    def forward(self, coarse, fg, training=True, update_running=True):
        P, pf = self.P, self.pf
        B, N, _ = coarse.shape
        S = self.gs ** 2
        dg = O.vn_linear(fg, P[pf + "downsize_global.map_to_feat.weight"])                 # [B,384,3,1]
        # pcn.py:466-470: expand(-1,384,-1,-1).reshape(bs,-1,N) on [B,384,N,3] re-interprets each sample's [N,3] block as [3,N]
        x = dg + np.ascontiguousarray(coarse).reshape(B, 1, 3, N)                           # [B,384,3,N]
        caches = []
        for i in range(2):
            x, c = vn_block(x, P, f"{pf}transformer.{i}.", self.H, self.SCALE, training, update_running)
            caches.append(c)
        C = x.shape[1]
        feat = np.ascontiguousarray(x.transpose(0, 3, 1, 2)).reshape(B * N, C, 3, 1)        # pcn.py:477-483
        feats = np.broadcast_to(feat, (B * N, C, 3, S))
        seed = np.broadcast_to(folding_seed_attn(self.gs, coarse.dtype).reshape(1, 1, 3, S), (B * N, 1, 3, S))
        fd1, cf1 = self._fold_fwd("vn_folding1", np.concatenate([seed, feats], 1), training, update_running)
        fd2, cf2 = self._fold_fwd("vn_folding2", np.concatenate([fd1, feats], 1), training, update_running)
        rel = fd2.reshape(B, N, 3, S)
        pts = np.swapaxes(rel + coarse[..., None], 2, 3).reshape(B, -1, 3)                  # pcn.py:492-493
        self.cache = (fg, caches, cf1, cf2, B, N, S, C)
        return np.ascontiguousarray(pts)

    def backward(self, g_pts):
        """returns (grads dict, g_coarse, g_fg)"""
        P, pf = self.P, self.pf
        fg, caches, cf1, cf2, B, N, S, C = self.cache
        G = {}
        grel = np.swapaxes(g_pts.reshape(B, N, S, 3), 2, 3)                                 # [B,N,3,S]
        g_coarse = grel.sum(-1, dtype=np.float64).astype(g_pts.dtype)
        gx2 = self._fold_bwd("vn_folding2", cf2, grel.reshape(B * N, 1, 3, S), G)
        gfeat = gx2[:, 1:].sum(-1, dtype=np.float64)
        gx1 = self._fold_bwd("vn_folding1", cf1, gx2[:, :1], G)
        gfeat = (gfeat + gx1[:, 1:].sum(-1, dtype=np.float64)).astype(g_pts.dtype)          # [B*N,C,3]
        gx = np.ascontiguousarray(gfeat.reshape(B, N, C, 3).transpose(0, 2, 3, 1))          # [B,C,3,N]
        for i in (1, 0):
            gx = vn_block_bwd(caches[i], P, f"{pf}transformer.{i}.", gx, G)
        g_coarse = g_coarse + gx.sum(1, dtype=np.float64).astype(gx.dtype).reshape(B, N, 3)
        gdg = gx.sum(-1, keepdims=True, dtype=np.float64).astype(gx.dtype)
        g_fg, G[pf + "downsize_global.map_to_feat.weight"] = O.vn_linear_bwd(fg, P[pf + "downsize_global.map_to_feat.weight"], gdg)
        return G, g_coarse, g_fg
