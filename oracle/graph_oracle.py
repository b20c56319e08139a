"""oracle/graph_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy (+ C, graph_oracle.c) restatement of the reference's VN_DGCNN_fps encoder, forward and
backward.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

Reference citations (file:line under /root/reference):
  knn = KNN(k=16, transpose_mode=False)     models/dgcnn.py:11      (third-party knn_cuda, restated in graph_oracle.c)
  fps_downsample                            models/dgcnn.py:203-223 (third-party pointnet2_ops, restated in graph_oracle.c)
  vn_get_graph_feature                      models/dgcnn.py:251-278
  VN_DGCNN_fps.forward                      models/dgcnn.py:280-324
  mean_pool                                 models/vn_layers.py:170-171

Parity: the net-level forward / backward is pinned by tests/golden/dgcnn_small.npz (the reference's own VN_DGCNN_fps run
unmodified on CPU with graph_oracle.c's searches plugged into its knn_cuda / pointnet2_ops imports); the two searches
themselves are "parity unpinned" against the third-party binaries (see graph_oracle.c).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import vn_oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _graph_lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle_graph.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/liboracle_graph.so missing: run `make -C oracle` (or __graft_entry__.build())")
        _lib = ctypes.CDLL(path)
        fp = ctypes.POINTER(ctypes.c_float)
        _lib.oracle_knn3d.argtypes = [ctypes.c_int] * 4 + [fp, fp, ctypes.POINTER(ctypes.c_int64), fp]
        _lib.oracle_knn3d.restype = ctypes.c_int
        _lib.oracle_fps.argtypes = [ctypes.c_int] * 3 + [fp, ctypes.POINTER(ctypes.c_int32)]
        _lib.oracle_fps.restype = ctypes.c_int
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def knn3d(ref, query, k):
    """ref [B,Nr,3], query [B,Nq,3] -> (idx [B,k,Nq] int64, dist [B,k,Nq] Euclidean), ordered by (distance, index)"""
    ref = np.ascontiguousarray(ref, np.float32)
    query = np.ascontiguousarray(query, np.float32)
    B, Nr, _ = ref.shape
    Nq = query.shape[1]
    idx = np.zeros((B, k, Nq), np.int64)
    dist = np.zeros((B, k, Nq), np.float32)
    rc = _graph_lib().oracle_knn3d(B, Nr, Nq, k, _fp(ref), _fp(query), idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _fp(dist))
    if rc != 0:
        raise ValueError("oracle_knn3d: bad arguments")
    return idx, dist


def fps(xyz, M):
    """xyz [B,N,3] -> idx [B,M] int32"""
    xyz = np.ascontiguousarray(xyz, np.float32)
    B, N, _ = xyz.shape
    idx = np.zeros((B, M), np.int32)
    rc = _graph_lib().oracle_fps(B, N, M, _fp(xyz), idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    if rc != 0:
        raise ValueError("oracle_fps: bad arguments")
    return idx


def gather_points(x, idx):
    """x [B,C,3,N] (or [B,3,N]), idx [B,M] -> same with the last axis gathered (pointnet2 gather_operation)"""
    ix = idx.astype(np.int64).reshape((idx.shape[0],) + (1,) * (x.ndim - 2) + (idx.shape[1],))
    return np.take_along_axis(x, np.broadcast_to(ix, x.shape[:-1] + (idx.shape[1],)), axis=-1)


def gather_points_bwd(x_shape, idx, g):
    gx = np.zeros(x_shape, g.dtype)
    B = x_shape[0]
    for b in range(B):
        np.add.at(gx[b], (Ellipsis, idx[b].astype(np.int64)), g[b])
    return gx


def graph_feature(x, idx):
    """x [B,C,3,N], idx [B,k,N] -> [B,2C,3,N,k] = cat(x_j - x_i, x_i)   (models/dgcnn.py:251-278)"""
    B, C, _, N = x.shape
    k = idx.shape[1]
    nb = np.swapaxes(idx, 1, 2)                                            # [B,N,k]
    xj = np.stack([x[b][:, :, nb[b]] for b in range(B)], axis=0)           # [B,C,3,N,k]
    xi = np.broadcast_to(x[..., None], xj.shape)
    return np.concatenate([xj - xi, xi], axis=1)


def graph_feature_bwd(x_shape, idx, g):
    B, C, _, N = x_shape
    nb = np.swapaxes(idx, 1, 2)
    g1, g2 = g[:, :C], g[:, C:]
    gx = (g2 - g1).sum(-1, dtype=np.float64)
    for b in range(B):
        acc = np.zeros((C, 3, N), np.float64)
        np.add.at(acc, (slice(None), slice(None), nb[b].reshape(-1)), g1[b].reshape(C, 3, -1).astype(np.float64))
        gx[b] += acc
    return gx.astype(g.dtype)


class VNDGCNNOracle:
    """VN_DGCNN_fps (models/dgcnn.py:164-324); prefix 'encoder.' in the PCNNet state_dict."""

    K = 16

    def __init__(self, P, prefix="encoder.", num_coarse=1024):
        self.P, self.pf, self.num_coarse = P, prefix, num_coarse

    def _vnll(self, name, x, training, update_running):
        P, pf = self.P, self.pf
        bn = O.bn_from_params(P, pf + name + ".batchnorm.bn")
        y, c = O.vn_linear_leaky_relu(x, P[pf + name + ".map_to_feat.weight"], P[pf + name + ".map_to_dir.weight"], bn, training,
                                      update_running=update_running)
        O.bn_to_params(P, pf + name + ".batchnorm.bn", bn)
        return y, c

    def _vnll_bwd(self, name, cache, g, G):
        P, pf = self.P, self.pf
        r = O.vn_linear_leaky_relu_bwd(cache, P[pf + name + ".map_to_feat.weight"], P[pf + name + ".map_to_dir.weight"], g)
        G[pf + name + ".map_to_feat.weight"], G[pf + name + ".map_to_dir.weight"] = r["gWf"], r["gWd"]
        G[pf + name + ".batchnorm.bn.weight"], G[pf + name + ".batchnorm.bn.bias"] = r["gweight"], r["gbias"]
        return r["gx"]

    def forward(self, xyz, training=True, forced_pool_idx=None, update_running=True, forced=None):
        """xyz [B,N,3].  forced = dict(knn=(i0,i1,i2), fps=(f1,f2)) teacher-forces the searches (optional)."""
        P, pf, k = self.P, self.pf, self.K
        B, N, _ = xyz.shape
        coor = np.swapaxes(xyz, 1, 2)                                       # [B,3,N]
        x = coor[:, None]                                                   # [B,1,3,N]
        i0 = forced["knn"][0] if forced else knn3d(xyz, xyz, k)[0]
        e0 = graph_feature(x, i0)
        h0, c0 = self._vnll("conv1.0", e0, training, update_running)
        x1 = h0.mean(-1)
        f1i = forced["fps"][0] if forced else fps(xyz, 512)
        coor1 = gather_points(coor, f1i)                                    # [B,3,512]
        fq1 = gather_points(x1, f1i)
        c1pts = np.ascontiguousarray(np.swapaxes(coor1, 1, 2))
        i1 = forced["knn"][1] if forced else knn3d(c1pts, c1pts, k)[0]
        e1 = graph_feature(fq1, i1)
        h1, c1 = self._vnll("conv4", e1, training, update_running)
        f = h1.mean(-1)
        e2 = graph_feature(f, i1)
        h2, c2 = self._vnll("conv5", e2, training, update_running)
        f2 = h2.mean(-1)
        f2i = forced["fps"][1] if forced else fps(c1pts, 128)
        coor2 = gather_points(coor1, f2i)
        fq2 = gather_points(f2, f2i)
        c2pts = np.ascontiguousarray(np.swapaxes(coor2, 1, 2))
        i2 = forced["knn"][2] if forced else knn3d(c2pts, c2pts, k)[0]
        e3 = graph_feature(fq2, i2)
        h3, c3 = self._vnll("conv6", e3, training, update_running)
        f3 = h3.mean(-1)                                                    # [B,512,3,128]
        g, pidx = O.vn_max_pool(f3, P[pf + "pool5.map_to_dir.weight"], forced_pool_idx)
        gf = g[..., None]                                                   # [B,512,3,1]
        h4, c4 = self._vnll("conv7.0", gf, training, update_running)
        m = O.vn_linear(h4, P[pf + "conv7.1.map_to_feat.weight"])          # [B,nc,3,1]
        coarse = m[..., 0]
        self.cache = dict(x=x, i0=i0, c0=c0, x1=x1, f1i=f1i, fq1=fq1, i1=i1, c1=c1, f=f, c2=c2, f2=f2, f2i=f2i, fq2=fq2, i2=i2, c3=c3,
                          f3=f3, pidx=pidx, c4=c4, h4=h4, k=k)
        self.knn_idx, self.fps_idx, self.pool_idx = (i0, i1, i2), (f1i, f2i), pidx
        return np.ascontiguousarray(coarse), gf

    def backward(self, g_coarse, g_fg=None):
        """returns (grads dict keyed like the state_dict, g_xyz through the feature path)"""
        P, pf, c = self.P, self.pf, self.cache
        k = c["k"]
        G = {}
        gm = g_coarse[..., None]
        gh4, gW = O.vn_linear_bwd(c["h4"], P[pf + "conv7.1.map_to_feat.weight"], gm)
        G[pf + "conv7.1.map_to_feat.weight"] = gW
        ggf = self._vnll_bwd("conv7.0", c["c4"], gh4, G)
        if g_fg is not None:
            ggf = ggf + g_fg
        gf3 = O.vn_max_pool_bwd(c["f3"].shape, c["pidx"], ggf[..., 0], ggf.dtype)
        gh3 = np.broadcast_to(gf3[..., None] / k, gf3.shape + (k,)).astype(gf3.dtype)
        ge3 = self._vnll_bwd("conv6", c["c3"], gh3, G)
        gfq2 = graph_feature_bwd(c["fq2"].shape, c["i2"], ge3)
        gf2 = gather_points_bwd(c["f2"].shape, c["f2i"], gfq2)
        gh2 = np.broadcast_to(gf2[..., None] / k, gf2.shape + (k,)).astype(gf2.dtype)
        ge2 = self._vnll_bwd("conv5", c["c2"], gh2, G)
        gf = graph_feature_bwd(c["f"].shape, c["i1"], ge2)
        gh1 = np.broadcast_to(gf[..., None] / k, gf.shape + (k,)).astype(gf.dtype)
        ge1 = self._vnll_bwd("conv4", c["c1"], gh1, G)
        gfq1 = graph_feature_bwd(c["fq1"].shape, c["i1"], ge1)
        gx1 = gather_points_bwd(c["x1"].shape, c["f1i"], gfq1)
        gh0 = np.broadcast_to(gx1[..., None] / k, gx1.shape + (k,)).astype(gx1.dtype)
        ge0 = self._vnll_bwd("conv1.0", c["c0"], gh0, G)
        gx = graph_feature_bwd(c["x"].shape, c["i0"], ge0)                  # [B,1,3,N]
        return G, np.swapaxes(gx[:, 0], 1, 2)
