"""Drop-in VN_PointNet encoder and VN_FoldingNet decoder (reference: models/pcn.py:110-184 and :319-389): same
constructor signature, sub-module names and state_dict keys, same inputs/outputs, executed on the row layout by the
sm_100a kernels.

B200-first restructuring (results equal to the reference's up to fp32 rounding):
  * torch.cat([global.expand(-1,-1,-1,N), local], 1) -> VNLinearLeakyReLU (pcn.py:172-173 and :383-387) never
    materialises the concatenation ([B,2050,3,16384] = 12.9 GB at B=32).  The weight columns that multiply the
    broadcast global feature are applied once per sample ([B*3, C] rows) and enter the per-point GEMM as a
    per-sample bias; only the local channels (512 of 1024 in the encoder, 2 of 2050 in the decoder) form the GEMM K.
  * VNMaxPool's index grids (CPU arange/meshgrid + H2D, vn_layers.py:165) are replaced by an in-kernel arg-max.
  * activations stay channels-last rows between layers; `coarse` and `fine` come out contiguous [B, n, 3] directly.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .vn_layers import VNLinear, VNLinearAndLeakyReLU, VNLinearLeakyReLU, VNMaxPool


class VN_PointNet(nn.Module):
    """models/pcn.py:110-184"""

    def __init__(self, config, num_dense=16384, latent_dim=1024):
        super().__init__()
        self.num_dense = num_dense
        self.latent_dim = latent_dim
        if config.num_coarse == 448:
            self.num_coarse = config.num_coarse // 2
        else:
            self.num_coarse = config.num_coarse
        self.first_conv = nn.Sequential(VNLinearLeakyReLU(1, 128, dim=4), VNLinear(128, 512))
        self.maxpool1 = VNMaxPool(512)
        self.second_conv = nn.Sequential(VNLinearLeakyReLU(1024, 1024, dim=4), VNLinear(1024, self.latent_dim * 2))
        self.maxpool2 = VNMaxPool(self.latent_dim * 2)
        self.mlp = nn.Sequential(
            VNLinearAndLeakyReLU(self.latent_dim * 2, 1024 * 2, dim=4, use_batchnorm='none'),
            VNLinearAndLeakyReLU(1024 * 2, 1024, dim=4, use_batchnorm='none'),
            VNLinear(1024, self.num_coarse))

    def forward(self, xyz):
        B, N, _ = xyz.shape
        # xyz.transpose(2,1).unsqueeze(1) is logical [B,1,3,N]; its rows (b,n,v) x 1 channel are xyz itself
        x0 = xyz.contiguous().view(B * N * 3, 1)
        f0 = self.first_conv[0].forward_rows(x0)                                          # [R,128]
        f1 = ops.linear_rows(f0, self.first_conv[1].map_to_feat.weight)                  # [R,512]
        # f1 feeds the pool and (through the concatenation) second_conv: the pool's sparse gradient is scattered into the dense one
        g1, f1 = self.maxpool1.forward_rows(f1, B, N, tap=True)                           # [B*3,512]
        Cg = g1.shape[1]
        l0 = self.second_conv[0]
        wcat = torch.cat([l0.map_to_feat.weight, l0.map_to_dir.weight], dim=0)            # [2048,1024]
        bias = ops.linear_rows(g1, wcat[:, :Cg])                                          # [B*3,2048]
        f2 = ops.linear_bn_leaky_fused_nograd(f1, wcat[:, Cg:], bias, 3 * N, l0.batchnorm.bn, l0.training, l0.negative_slope)
        if f2 is None:      # GEMM -> (p | d) [R,2048] with the BatchNorm statistics from its epilogue, then one BN + leaky pass -> [R,1024]
            f2 = ops.linear_bn_leaky_rows(f1, wcat[:, Cg:], bias, 3 * N, l0.batchnorm.bn, l0.training, l0.negative_slope)
        # second_conv[1] feeds only maxpool2: fused, its [R,2048] output is transient and its backward is sparse
        fg, idx2 = ops.linear_maxpool_rows(f2, self.second_conv[1].map_to_feat.weight, self.maxpool2.map_to_dir.weight, B, N,
                                           self.maxpool2.forced_idx)                       # [B*3,2048]
        self.maxpool2.last_idx = idx2
        m = self.mlp[0].forward_rows(fg)
        m = self.mlp[1].forward_rows(m)
        m = ops.linear_rows(m, self.mlp[2].map_to_feat.weight)                            # [B*3,num_coarse] rows (b,v)
        # reference: mlp(...)[B,nc,3,1].reshape(-1,nc,3)
        coarse = m.view(B, 3, self.num_coarse).transpose(1, 2).contiguous()
        feature_global = fg.view(B, 3, -1).transpose(1, 2).unsqueeze(-1)                   # logical [B,2048,3,1]
        if self.num_coarse == 224:
            # the 448-coarse variant (pcn.py:179-182): 224 predicted points + 224 furthest-point samples of the input (csrc/graph.cu)
            from . import graph_ops as G
            xyz_c = xyz.contiguous()
            inp_sparse = G.points_gather(xyz_c.view(B * N * 3, 1), G.fps(xyz_c, 224), B, N).view(B, 224, 3)
            coarse_cat = torch.cat([coarse, inp_sparse], dim=1).contiguous()
            return (coarse, coarse_cat), feature_global
        return coarse, feature_global


class VN_FoldingNet(nn.Module):
    """models/pcn.py:319-389"""

    def __init__(self, config, grid_size=4):
        super().__init__()
        self.grid_size = grid_size
        self.latent_dim = config.latent_dim
        self.num_dense = 16384
        if config.num_coarse == 448:
            self.num_coarse = config.num_coarse // 2
            self.num_dense = 14336
            self.grid_size = 8
        else:
            self.num_coarse = config.num_coarse
            self.num_dense = 16384
            self.grid_size = 4
        self.final_conv = nn.Sequential(VNLinearLeakyReLU(self.latent_dim + 1 + 1, 256, dim=4),
                                        VNLinearLeakyReLU(256, 256, dim=4), VNLinear(256, 1))
        gs = self.grid_size
        a = torch.linspace(-0.05, 0.05, steps=gs, dtype=torch.float).view(1, gs).expand(gs, gs).reshape(1, -1)
        b = torch.linspace(-0.05, 0.05, steps=gs, dtype=torch.float).view(gs, 1).expand(gs, gs).reshape(1, -1)
        c = torch.zeros_like(a, dtype=torch.float)
        # a plain attribute (not a buffer) like the reference (pcn.py:362), so state_dict keys match; follows the module's device lazily
        self.folding_seed = torch.cat([a, b, c], dim=0).reshape(1, 1, 3, -1)

    def forward(self, coarse, feature_global, rot=None):
        dev = coarse.device
        if self.folding_seed.device != dev:
            self.folding_seed = self.folding_seed.to(dev)
        B = coarse.shape[0]
        S = self.grid_size ** 2
        nc = self.num_coarse
        nd = nc * S
        seed_pts = self.folding_seed.squeeze(1).transpose(1, 2)                            # [1,S,3]
        if rot is not None:
            seed_pts = rot.transform_points(seed_pts)                                      # [B,S,3]   pcn.py:369-370
        seed_pts = seed_pts.expand(B, S, 3)
        # local channels of the concatenation (pcn.py:375-385): [seed, point_feat] per dense point, rows (b, n, v)
        local = torch.stack([seed_pts[:, None, :, :].expand(B, nc, S, 3), coarse[:, :, None, :].expand(B, nc, S, 3)], dim=-1)
        local = local.reshape(B * nd * 3, 2)
        fg_rows = feature_global.squeeze(-1).transpose(1, 2).reshape(B * 3, -1)            # [B*3,Cg] rows (b,v)
        Cg = fg_rows.shape[1]
        l0, l1, l2 = self.final_conv[0], self.final_conv[1], self.final_conv[2]
        wcat = torch.cat([l0.map_to_feat.weight, l0.map_to_dir.weight], dim=0)            # [512, Cg+2]
        # (a contiguous copy of the Cg broadcast columns when the [512, Cg+2] rows are not 16-byte aligned: 2050 floats per row would push
        # this 96 x 2048 x 512 GEMM and its gradients onto the fp32 SIMT kernel -- 0.25 ms on four CTAs)
        wg = wcat[:, :Cg]
        if (wcat.stride(0) * 4) % 16 != 0:
            wg = wg.contiguous()
        bias = ops.linear_rows(fg_rows, wg)                                                # [B*3,512]
        C0 = l0.map_to_feat.weight.shape[0]
        if ops.smallk_bn_leaky_supported(2, C0, bias) and l0.batchnorm.bn.affine:
            # p = W_feat x + b_p and d = W_dir x + b_d are two FMAs per component: recomputed in every pass, never stored
            # column 0 of `local` is the folding seed: a constant unless the rotation itself is being differentiated
            seed_const = 0 if seed_pts.requires_grad else 1
            h = ops.smallk_bn_leaky(local, wcat[:, Cg:], bias, l0.batchnorm.bn, l0.training, l0.negative_slope, B, nd, seed_const)
        else:
            h = ops.linear_bn_leaky_rows(local, wcat[:, Cg:], bias, 3 * nd, l0.batchnorm.bn, l0.training, l0.negative_slope)
        C1 = l1.map_to_feat.weight.shape[0]
        if torch.is_grad_enabled() and l1.map_to_dir.weight.shape[0] == C1 and ops.bn_leaky_dot_supported(C1):
            # final_conv[1] (BN + leaky) fused with final_conv[2] = VNLinear(256,1) and the residual: its [R,256] output
            # and gradient never touch HBM
            fine = ops.linear_bn_leaky_dot(h, torch.cat([l1.map_to_feat.weight, l1.map_to_dir.weight], dim=0), l1.batchnorm.bn, l1.training,
                                           l1.negative_slope, l2.map_to_feat.weight, local[:, 1])
        else:
            h = l1.forward_rows(h)      # no-grad: BN + leaky fused into the tcgen05 GEMM epilogue (vn_layers.VNLinearLeakyReLU)
            fine = ops.rows_dot(h, l2.map_to_feat.weight, local[:, 1])                     # final VNLinear(256,1) + point_feat
        return fine.view(B, nd, 3)


class Attention_VN_FoldingNet(nn.Module):
    """models/pcn.py:392-520: two VN_Blocks over the 1024 coarse tokens (feature = down-sized global feature +
    the token's own coordinates), then two folding MLPs per token over a 4x4 seed grid.

    B200-first restructuring (same results up to fp32 rounding):
      * the token tensor is built and kept in the row layout; the per-block [B,N,C*3] <-> [B,C,3,N] transposes disappear;
      * cat([seed | fd1, features.expand(16)]) -> VNLinearLeakyReLU(385, 256) (pcn.py:487-491): the 384 broadcast channels are
        applied once per token as a per-token bias ([B*N*3, 512] rows), only the 1 local channel is recomputed per folded
        point, and its (p, d) are never stored (csrc/vn_fused.cu);
      * VNLinearLeakyReLU(256,128) -> VNLinear(128,1) (+ the coarse-point residual, pcn.py:493) is one fused tail.
    `final_conv` is constructed (and stays in the state_dict) but is not used by the reference's forward either."""

    def __init__(self, config, grid_size=4):
        super().__init__()
        from .transformer import VN_Block
        self.grid_size = grid_size
        self.latent_dim = config.latent_dim
        self.num_dense = 16384
        if config.num_coarse == 448:
            self.num_coarse = config.num_coarse // 2
            self.num_dense = 14336
            self.grid_size = 8
        else:
            self.num_coarse = config.num_coarse
            self.num_dense = 16384
            self.grid_size = 4
        self.transformer = nn.ModuleList([VN_Block(dim=384, num_heads=8, mlp_ratio=1, qkv_bias=False, qk_scale=1, drop=0, attn_drop=0)
                                          for _ in range(2)])
        self.final_conv = nn.Sequential(VNLinearLeakyReLU(self.latent_dim + 1 + 1, 256, dim=4), VNLinearLeakyReLU(256, 256, dim=4),
                                        VNLinear(256, 1))
        self.downsize_global = VNLinear(2048, 384)
        gs = self.grid_size
        a = torch.linspace(-1., 1., steps=gs, dtype=torch.float).view(1, gs).expand(gs, gs).reshape(1, -1)
        b = torch.linspace(-1., 1., steps=gs, dtype=torch.float).view(gs, 1).expand(gs, gs).reshape(1, -1)
        c = torch.zeros_like(a, dtype=torch.float)
        self.folding_seed = torch.cat([a, b, c], dim=0)              # [3, S]; plain attribute like the reference (pcn.py:454)
        in_channel, hidden_dim = 384, 256
        self.vn_folding1 = nn.Sequential(VNLinearLeakyReLU(in_channel + 1, hidden_dim, dim=4),
                                         VNLinearLeakyReLU(hidden_dim, hidden_dim // 2, dim=4), VNLinear(hidden_dim // 2, 1))
        self.vn_folding2 = nn.Sequential(VNLinearLeakyReLU(in_channel + 1, hidden_dim, dim=4),
                                         VNLinearLeakyReLU(hidden_dim, hidden_dim // 2, dim=4), VNLinear(hidden_dim // 2, 1))

    @staticmethod
    def _fold(seq, local, feat_rows, T, S, const_local, res=None):
        """one folding MLP: local rows ((token, s), v) x 1, feat_rows (token, v) x 384 -> rows ((token, s), v) [R]"""
        l0, l1, l2 = seq[0], seq[1], seq[2]
        wcat = torch.cat([l0.map_to_feat.weight, l0.map_to_dir.weight], dim=0)           # [512, 385]; column 0 = the local channel
        # (a contiguous copy of the 384 broadcast columns: the [512, 385] view starts one float off 16-byte alignment, which would push this
        # 98304 x 384 x 512 GEMM and its gradients off the tensor-core kernels)
        bias = ops.linear_rows(feat_rows, wcat[:, 1:].contiguous())                       # [T*3, 512]
        C0 = l0.map_to_feat.weight.shape[0]
        if ops.smallk_bn_leaky_supported(1, C0, bias) and l0.batchnorm.bn.affine:
            h = ops.smallk_bn_leaky(local, wcat[:, :1], bias, l0.batchnorm.bn, l0.training, l0.negative_slope, T, S, 1 if const_local else 0)
        else:
            h = ops.linear_bn_leaky_rows(local, wcat[:, :1], bias, 3 * S, l0.batchnorm.bn, l0.training, l0.negative_slope)
        C1 = l1.map_to_feat.weight.shape[0]
        if torch.is_grad_enabled() and ops.bn_leaky_dot_supported(C1):
            return ops.linear_bn_leaky_dot(h, torch.cat([l1.map_to_feat.weight, l1.map_to_dir.weight], dim=0), l1.batchnorm.bn, l1.training,
                                           l1.negative_slope, l2.map_to_feat.weight, res)
        h = l1.forward_rows(h)
        return ops.rows_dot(h, l2.map_to_feat.weight, res)

    def forward(self, coarse, feature_global, rot=None):
        dev = coarse.device
        if self.folding_seed.device != dev:
            self.folding_seed = self.folding_seed.to(dev)
        B, N, _ = coarse.shape
        S = self.grid_size ** 2
        coarse = coarse.contiguous()
        # tokens: rows (b, n, v) x 384 = downsize_global(fg)[(b, v), :] + centers[b, v, n]   (pcn.py:466-470).
        # NOTE the reference builds `repeat_input_centers` with expand(-1,384,-1,-1).reshape(bs,-1,N) on a [B,384,N,3] tensor,
        # which re-interprets each sample's [N,3] coordinate block as [3,N] (no transpose): token n receives the vector
        # (flat[n], flat[N+n], flat[2N+n]) of coarse[b].flatten().  Reproduced as is -- parity with the reference's outputs.
        centers = coarse.reshape(B, 3, N).transpose(1, 2).contiguous().view(B * N * 3, 1)
        fg_rows = feature_global.squeeze(-1).transpose(1, 2).reshape(B * 3, -1)
        dg = ops.linear_rows(fg_rows, self.downsize_global.map_to_feat.weight)            # [B*3, 384]
        ones = torch.ones((dg.shape[1], 1), device=dev, dtype=torch.float32)
        tok = ops.linear_rows(centers, ones, dg, 3 * N)                                   # broadcast add as a K=1 VNLinear + per-sample bias
        for blk in self.transformer:
            tok = blk.forward_rows(tok, B, N)
        # folding (pcn.py:477-493): every token is a "sample" of S points
        T = B * N
        seed = self.folding_seed.t().contiguous()                                          # [S, 3] rows (s, v)
        local1 = seed.unsqueeze(0).expand(T, S, 3).reshape(T * S * 3, 1)
        fd1 = self._fold(self.vn_folding1, local1, tok, T, S, True)
        res = coarse[:, :, None, :].expand(B, N, S, 3).reshape(-1)
        fd2 = self._fold(self.vn_folding2, fd1.view(T * S * 3, 1), tok, T, S, False, res)  # + coarse.unsqueeze(-1)  (pcn.py:493)
        return fd2.view(B, N * S, 3)
