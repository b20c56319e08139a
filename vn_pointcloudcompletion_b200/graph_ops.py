"""Host-side point-set graph operators of the VN_DGCNN_fps encoder, each enqueueing kernels of
libvnpcc.so (csrc/graph.cu) on the current stream.  No CPU / ATen fallback.

Drop-ins for the two un-vendored third-party CUDA packages the reference imports on this path:
  KNN                      knn_cuda.KNN(k, transpose_mode)            models/dgcnn.py:6,11  (wheel KNN_CUDA 0.2, README.md:31)
  furthest_point_sample    pointnet2_ops.pointnet2_utils.*            models/dgcnn.py:7,15-16,210,215  (README.md:29)
  gather_operation
and the row-layout forms the network itself uses:
  knn3d, fps, points_gather, edge_feature, group_mean
"""
from __future__ import annotations

import torch

from ._lib import call, ptr, stream
from .ops import _check, _ld, _rows2d


# ---------------------------------------------------------------------------------------------------------------
# index-producing searches (no gradient)
# ---------------------------------------------------------------------------------------------------------------
def knn3d(ref, query, k, want_dist=False):
    """ref [B,Nr,3], query [B,Nq,3] -> idx [B,k,Nq] int64 (and Euclidean dist [B,k,Nq] if want_dist), ordered by
    (distance, index)."""
    _check(ref, "ref")
    _check(query, "query")
    if ref.dim() != 3 or query.dim() != 3 or ref.shape[2] != 3 or query.shape[2] != 3 or ref.shape[0] != query.shape[0]:
        raise ValueError(f"expected [B,Nr,3] and [B,Nq,3], got {tuple(ref.shape)} and {tuple(query.shape)}")
    ref = ref.detach().contiguous()
    query = query.detach().contiguous()
    B, Nr, _ = ref.shape
    Nq = query.shape[1]
    idx = torch.empty((B, k, Nq), device=ref.device, dtype=torch.int64)
    dist = torch.empty((B, k, Nq), device=ref.device, dtype=torch.float32) if want_dist else None
    with torch.cuda.device(ref.device):
        call("vnpcc_knn3d", ptr(ref), ptr(query), B, Nr, Nq, int(k), ptr(idx), ptr(dist), stream())
    return (idx, dist) if want_dist else idx


def fps(xyz, M):
    """xyz [B,N,3] -> idx [B,M] int32 (pointnet2 furthest_point_sample semantics)"""
    _check(xyz, "xyz")
    if xyz.dim() != 3 or xyz.shape[2] != 3:
        raise ValueError(f"expected [B,N,3], got {tuple(xyz.shape)}")
    xyz = xyz.detach().contiguous()
    B, N, _ = xyz.shape
    idx = torch.empty((B, M), device=xyz.device, dtype=torch.int32)
    with torch.cuda.device(xyz.device):
        call("vnpcc_fps", ptr(xyz), B, N, int(M), ptr(idx), stream())
    return idx


# ---------------------------------------------------------------------------------------------------------------
# differentiable gathers on the row layout
# ---------------------------------------------------------------------------------------------------------------
class _PointsGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, idx, B, N):
        x = _rows2d(x, "x")
        C = x.shape[1]
        M = idx.shape[1]
        out = torch.empty((B * M * 3, C), device=x.device, dtype=torch.float32)
        call("vnpcc_points_gather", ptr(x), _ld(x), ptr(idx), B, N, M, C, ptr(out), C, stream())
        ctx.save_for_backward(idx)
        ctx.cfg = (B, N, M, C)
        return out

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        B, N, M, C = ctx.cfg
        g = _rows2d(g, "grad")
        gx = torch.empty((B * N * 3, C), device=g.device, dtype=torch.float32)
        call("vnpcc_points_scatter_add", ptr(g), _ld(g), ptr(idx), B, N, M, C, ptr(gx), C, stream())
        return gx, None, None, None


def points_gather(x, idx, B, N):
    """x rows (b,n,v) x C, idx [B,M] int32 -> rows (b,m,v) x C"""
    if idx.dtype != torch.int32:
        idx = idx.to(torch.int32)
    return _PointsGather.apply(x, idx.contiguous(), B, N)


class _EdgeFeature(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, idx, B, N):
        x = _rows2d(x, "x")
        C = x.shape[1]
        k = idx.shape[1]
        out = torch.empty((B * N * k * 3, 2 * C), device=x.device, dtype=torch.float32)
        call("vnpcc_edge_feature_fwd", ptr(x), _ld(x), ptr(idx), B, N, k, C, ptr(out), 2 * C, stream())
        ctx.save_for_backward(idx)
        ctx.cfg = (B, N, k, C)
        return out

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        B, N, k, C = ctx.cfg
        g = _rows2d(g, "grad")
        gx = torch.empty((B * N * 3, C), device=g.device, dtype=torch.float32)
        call("vnpcc_edge_feature_bwd", ptr(g), _ld(g), ptr(idx), B, N, k, C, ptr(gx), C, stream())
        return gx, None, None, None


def edge_feature(x, idx, B, N):
    """x rows (b,n,v) x C, idx [B,k,N] int64 (knn layout) -> rows ((b,n,j),v) x 2C = (x_j - x_i | x_i)
    (VN_DGCNN_fps.vn_get_graph_feature, models/dgcnn.py:251-278; logical [B, 2C, 3, N, k])"""
    if idx.dtype != torch.int64 or idx.dim() != 3 or idx.shape[0] != B or idx.shape[2] != N:
        raise ValueError(f"idx must be int64 [B,k,N], got {idx.dtype} {tuple(idx.shape)}")
    return _EdgeFeature.apply(x, idx.contiguous(), B, N)


class _GroupMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k):
        x = _rows2d(x, "x")
        R, C = x.shape
        G = R // (3 * k)
        out = torch.empty((G * 3, C), device=x.device, dtype=torch.float32)
        call("vnpcc_rows_group_mean", ptr(x), _ld(x), G, k, C, ptr(out), C, stream())
        ctx.cfg = (G, k, C)
        return out

    @staticmethod
    def backward(ctx, g):
        G, k, C = ctx.cfg
        g = _rows2d(g, "grad")
        gx = torch.empty((G * k * 3, C), device=g.device, dtype=torch.float32)
        call("vnpcc_rows_group_mean_bwd", ptr(g), _ld(g), G, k, C, ptr(gx), C, stream())
        return gx, None


def group_mean(x, k):
    """mean_pool over the last spatial axis of size k: rows ((g,j),v) x C -> rows (g,v) x C"""
    return _GroupMean.apply(x, int(k))


# ---------------------------------------------------------------------------------------------------------------
# drop-ins with the third-party call signatures
# ---------------------------------------------------------------------------------------------------------------
class KNN(torch.nn.Module):
    """knn_cuda.KNN: forward(ref, query) -> (dist, idx).  transpose_mode=False: ref [B,D,Nr], query [B,D,Nq] -> [B,k,Nq];
    transpose_mode=True: ref [B,Nr,D], query [B,Nq,D] -> [B,Nq,k].  Only D == 3 is on the B200 path."""

    def __init__(self, k, transpose_mode=False):
        super().__init__()
        self.k = k
        self._t = transpose_mode

    def forward(self, ref, query):
        if not self._t:
            ref, query = ref.transpose(1, 2), query.transpose(1, 2)
        if ref.shape[2] != 3:
            raise NotImplementedError("the B200 kNN kernel searches 3-D coordinates (every call on the vn_dgcnn_fps path)")
        idx, dist = knn3d(ref, query, self.k, want_dist=True)
        if self._t:
            return dist.transpose(1, 2).contiguous(), idx.transpose(1, 2).contiguous()
        return dist, idx


def furthest_point_sample(xyz, npoint):
    """pointnet2_utils.furthest_point_sample: xyz [B,N,3] -> int32 [B,npoint]"""
    return fps(xyz, npoint)


def gather_operation(features, idx):
    """pointnet2_utils.gather_operation: features [B,C,N], idx [B,M] int32 -> [B,C,M] (differentiable w.r.t. features)"""
    _check(features, "features")
    B, C, N = features.shape
    if C % 3 == 0:
        # [B, C, N] viewed as C/3 VN channels: rows (b,n,v) x C/3
        rows = features.view(B, C // 3, 3, N).permute(0, 3, 2, 1).reshape(B * N * 3, C // 3)
        out = points_gather(rows, idx, B, N)
        return out.view(B, idx.shape[1], 3, C // 3).permute(0, 3, 2, 1).reshape(B, C, idx.shape[1])
    # generic channel count: treat every channel as its own 3-row group by padding the point axis view
    rows = features.permute(0, 2, 1).reshape(B * N, C)
    rows3 = rows.unsqueeze(1).expand(B * N, 3, C).reshape(B * N * 3, C)
    out = points_gather(rows3, idx, B, N)
    return out.view(B, idx.shape[1], 3, C)[:, :, 0, :].permute(0, 2, 1).contiguous()
