"""Drop-in VN_Block / Attention (reference: models/transformer.py:25-105): same constructor arguments, sub-module names and
state_dict keys (including the reference's unused `qkv` / `proj` nn.Linear parameters and the `conv1` / `conv2` layers of the
knn branch), executed on the row layout by the sm_100a kernels.

Token layout: VN_Block.forward takes and returns x [B, N, C*3] whose last axis is (channel, component) with the component
fastest, exactly as the reference (transformer.py:46-47,70); forward_rows works on rows (b, n, v) x C and is what
Attention_VN_FoldingNet chains, so the per-block [B, N, C*3] <-> [B, C, 3, N] transposes of the reference disappear.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .vn_layers import VNLayerNorm, VNLinear, VNLinearLeakyReLU, mean_pool


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)            # unused by the reference's forward; kept for state_dict parity
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)                               # unused as well
        self.proj_drop = nn.Dropout(proj_drop)
        self.proj_vnq = VNLinear(dim, dim)
        self.proj_vnk = VNLinear(dim, dim)
        self.proj_vnv = VNLinear(dim, dim)
        self.proj_vn = VNLinear(dim, dim)
        if attn_drop or proj_drop:
            raise NotImplementedError("dropout inside the VN attention (the reference instantiates it with 0)")

    def forward_rows(self, rows, B, N):
        """rows (b, n, v) x C -> rows (b, n, v) x C"""
        w = torch.cat([self.proj_vnq.map_to_feat.weight, self.proj_vnk.map_to_feat.weight, self.proj_vnv.map_to_feat.weight], dim=0)
        qkv = ops.linear_rows(rows, w)                                # one GEMM for q, k, v: the tokens are read once
        o = ops.vn_attention(qkv, B, N, self.num_heads, self.scale)
        return ops.linear_rows(o, self.proj_vn.map_to_feat.weight)

    def forward(self, vn_x):
        from .vn_layers import from_rows, to_rows
        rows, B, sp = to_rows(vn_x)
        return from_rows(self.forward_rows(rows, B, sp[0]), B, sp)


class VN_Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0.,
                 drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self.norm1 = VNLayerNorm(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = nn.Identity()
        self.norm2 = VNLayerNorm(dim)
        self.conv1 = VNLinearLeakyReLU(dim * 2, dim)
        self.conv2 = VNLinear(dim * 2, dim)
        self.conv3 = VNLinearLeakyReLU(dim, dim * 2, dim=4)
        self.conv4 = VNLinearLeakyReLU(dim * 2, dim, dim=4)
        self.pool1 = mean_pool

    def forward_rows(self, rows, B, N):
        x1 = self.attn.forward_rows(self.norm1.forward_rows(rows), B, N)
        rows = ops.rows_add(rows, x1)
        x2 = self.conv4.forward_rows(self.conv3.forward_rows(self.norm2.forward_rows(rows)))
        return ops.rows_add(rows, x2)

    def forward(self, x, knn_index=None):
        if knn_index is not None:
            raise NotImplementedError("the knn branch of VN_Block is only used by the PoinTr encoder (out of scope, SURVEY.md 2 #9)")
        B, N, C3 = x.shape
        C = C3 // 3
        rows = x.view(B, N, C, 3).transpose(2, 3).reshape(B * N * 3, C)
        out = self.forward_rows(rows, B, N)
        return out.view(B, N, 3, C).transpose(2, 3).reshape(B, N, C3)
