"""Drop-in VN_DGCNN_fps encoder (reference: models/dgcnn.py:164-324): same constructor, sub-module names and state_dict
keys, same input [B, N, 3] and outputs (coarse [B, num_coarse, 3], global_feature [B, 512, 3, 1]), executed on the row
layout by the sm_100a kernels.

The two third-party CUDA packages the reference calls here are replaced by csrc/graph.cu: knn_cuda.KNN(k=16) ->
vnpcc_knn3d, pointnet2 furthest_point_sample / gather_operation -> vnpcc_fps / vnpcc_points_gather.  Every kNN graph on
this path is built on 3-D coordinates (models/dgcnn.py:282 on the input cloud, :289,:296,:303 on the FPS-downsampled
coordinates); the graph of conv5 reuses the neighbour lists of conv4 (same coordinates, models/dgcnn.py:289,296).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import graph_ops as G
from . import ops
from .vn_layers import VNLinear, VNLinearLeakyReLU, VNMaxPool, mean_pool

knn = G.KNN(k=16, transpose_mode=False)          # models/dgcnn.py:11


def fps(pc, num):
    """models/dgcnn.py:14-17: pc [B,N,3] -> FPS-subsampled [B,num,3]"""
    B, N, _ = pc.shape
    idx = G.fps(pc, num)
    return G.points_gather(pc.contiguous().view(B * N * 3, 1), idx, B, N).view(B, num, 3)


class VN_DGCNN_fps(nn.Module):
    def __init__(self, config, only_coarse=False):
        super().__init__()
        if config.num_coarse == 448:
            self.num_coarse = config.num_coarse // 2
        else:
            self.num_coarse = config.num_coarse
        self.conv1 = nn.Sequential(VNLinearLeakyReLU(2, 32))
        self.conv4 = VNLinearLeakyReLU(64, 64)
        self.conv5 = VNLinearLeakyReLU(128, 128)
        self.conv6 = VNLinearLeakyReLU(256, 512)
        self.conv7 = nn.Sequential(VNLinearLeakyReLU(512, 1024, dim=4), VNLinear(1024, self.num_coarse))
        self.pool1 = mean_pool
        self.pool2 = mean_pool
        self.pool3 = mean_pool
        self.pool4 = mean_pool
        self.pool5 = VNMaxPool(512)
        self.k = 16
        self.last_fps_idx = None      # (idx 2048->512, idx 512->128) of the most recent forward, for parity checks
        self.last_knn_idx = None

    @staticmethod
    def _edge_conv(layer, x_rows, idx, B, N, k):
        """graph feature -> VNLinearLeakyReLU(dim=5, BatchNorm2d over (B, N, k)) -> mean over k.

        W [x_j - x_i ; x_i] = W1 x_j + (W2 - W1) x_i: one GEMM over the N points (k-fold fewer FLOPs, always exact fp32 -- the
        difference of two TF32-rounded products would lose the local geometry) and a fused gather-add / BatchNorm / leaky / mean kernel;
        the edge tensor [B, 2C, 3, N, k] and its (p | d) image never reach HBM (csrc/edge_conv.cu)."""
        wf, wd = layer.map_to_feat.weight, layer.map_to_dir.weight
        C, Cin = wf.shape[0], x_rows.shape[1]
        if wd.shape[0] == C and ops.edge_conv_supported(C, layer.batchnorm.bn):
            w1 = torch.cat([wf[:, :Cin], wd[:, :Cin]], dim=0)
            w2 = torch.cat([wf[:, Cin:], wd[:, Cin:]], dim=0)
            uw = ops.linear_rows(x_rows, torch.cat([w1, w2 - w1], dim=0), exact=True)      # rows (b,n,v) x 4C
            return ops.edge_conv(uw, idx, layer.batchnorm.bn, layer.training, layer.negative_slope, B, N)
        e = G.edge_feature(x_rows, idx, B, N)                       # rows ((b,n,j),v) x 2C   (materialised fallback)
        h = layer.forward_rows(e)                                   # rows ((b,n,j),v) x Cout
        return G.group_mean(h, k)                                   # rows (b,n,v) x Cout

    def forward(self, x):
        # x: [B, N, 3]
        B, N, _ = x.shape
        k = self.k
        xyz = x.contiguous()
        rows0 = xyz.view(B * N * 3, 1)                               # logical [B,1,3,N]
        idx0 = G.knn3d(xyz, xyz, k)                                  # dynamic graph of a 1-channel VN feature = the coordinates
        x1 = self._edge_conv(self.conv1[0], rows0, idx0, B, N, k)    # [B,32,3,N]
        n1 = 512
        fi1 = G.fps(xyz, n1)
        coor1 = G.points_gather(rows0, fi1, B, N)                    # rows (b,m,v) x 1
        f_q = G.points_gather(x1, fi1, B, N)                         # rows (b,m,v) x 32
        c1 = coor1.view(B, n1, 3)
        idx1 = G.knn3d(c1, c1, k)
        f = self._edge_conv(self.conv4, f_q, idx1, B, n1, k)         # [B,64,3,512]
        f = self._edge_conv(self.conv5, f, idx1, B, n1, k)           # [B,128,3,512]
        n2 = 128
        fi2 = G.fps(c1, n2)
        coor2 = G.points_gather(coor1, fi2, B, n1)
        f_q = G.points_gather(f, fi2, B, n1)
        c2 = coor2.view(B, n2, 3)
        idx2 = G.knn3d(c2, c2, k)
        f = self._edge_conv(self.conv6, f_q, idx2, B, n2, k)         # [B,512,3,128]
        self.last_fps_idx = (fi1, fi2)
        self.last_knn_idx = (idx0, idx1, idx2)
        g = self.pool5.forward_rows(f, B, n2)                        # rows (b,v) x 512
        h = self.conv7[0].forward_rows(g)
        m = ops.linear_rows(h, self.conv7[1].map_to_feat.weight)     # rows (b,v) x num_coarse
        coarse = m.view(B, 3, self.num_coarse).transpose(1, 2).contiguous()
        global_feature = g.view(B, 3, -1).transpose(1, 2).unsqueeze(-1)       # logical [B,512,3,1]
        if self.num_coarse == 224:
            inp_sparse = fps(xyz, 224)
            coarse_cat = torch.cat([coarse, inp_sparse], dim=1).contiguous()
            return (coarse, coarse_cat), global_feature
        return coarse, global_feature
