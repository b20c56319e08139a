"""Drop-in Vector-Neuron layers: same class names, constructor arguments, sub-module / state_dict names and logical
tensor layout [B, C, 3, N, ...] as the reference's models/vn_layers.py, executed by the sm_100a kernels of libvnpcc.so.

  VNLinear              models/vn_layers.py:12-22
  VNLeakyReLU           models/vn_layers.py:25-43
  VNLinearLeakyReLU     models/vn_layers.py:46-74
  VNLinearAndLeakyReLU  models/vn_layers.py:77-104
  VNBatchNorm           models/vn_layers.py:107-127
  VNMaxPool             models/vn_layers.py:153-167
  mean_pool             models/vn_layers.py:170-171
  VNStdFeature          models/vn_layers.py:174-221

Parameters are held in nn.Linear / nn.BatchNorm sub-modules with the reference's attribute names (map_to_feat,
map_to_dir, batchnorm.bn, linear, leaky_relu, vn1, vn2, vn_lin) purely as containers, so reference checkpoints load
with load_state_dict and the default initialisation is identical; their ATen forward is never called.

Physical layout: outputs are channels-last ([B, *spatial, 3, C] in memory, returned as logical views), which for
dim=4 is exactly what the reference's nn.Linear-on-a-transposed-view produces.  Inputs may have any
strides; one that is not already channels-last is copied once.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops

EPS = 1e-6


def to_rows(x):
    """logical [B, C, 3, *spatial] -> (rows [R, C], B, spatial)"""
    if x.dim() < 3 or x.shape[2] != 3:
        raise ValueError(f"expected a VN tensor [B, C, 3, ...], got {tuple(x.shape)}")
    perm = (0,) + tuple(range(3, x.dim())) + (2, 1)
    xp = x.permute(perm)
    if not xp.is_contiguous():
        xp = xp.contiguous()
    return xp.reshape(-1, x.shape[1]), x.shape[0], tuple(x.shape[3:])


def from_rows(rows, B, spatial):
    """rows [R, C] -> logical [B, C, 3, *spatial] (a view of the channels-last buffer)"""
    C = rows.shape[1]
    y = rows.view((B,) + tuple(spatial) + (3, C))
    k = len(spatial)
    inv = (0, k + 2, k + 1) + tuple(range(1, k + 1))
    return y.permute(inv)


class VNLinear(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.map_to_feat = nn.Linear(in_channels, out_channels, bias=False)

    def forward(self, x):
        rows, B, sp = to_rows(x)
        return from_rows(ops.linear_rows(rows, self.map_to_feat.weight), B, sp)


def _leaky_dir(rows, y_rows, map_to_dir, ns):
    """leaky projection of y_rows along d = map_to_dir(rows); handles share_nonlinearity ([1, C] weight)"""
    d = ops.linear_rows(rows, map_to_dir.weight)
    if d.shape[1] != y_rows.shape[1]:
        d = d.expand(-1, y_rows.shape[1]).contiguous()
    return ops.bn_leaky(y_rows, d, None, False, ns)


class VNLeakyReLU(nn.Module):
    def __init__(self, in_channels, share_nonlinearity=False, negative_slope=0.2):
        super().__init__()
        if share_nonlinearity == True:  # noqa: E712  (mirrors the reference's truthiness test)
            self.map_to_dir = nn.Linear(in_channels, 1, bias=False)
        else:
            self.map_to_dir = nn.Linear(in_channels, in_channels, bias=False)
        self.negative_slope = negative_slope

    def forward(self, x):
        rows, B, sp = to_rows(x)
        return from_rows(_leaky_dir(rows, rows, self.map_to_dir, self.negative_slope), B, sp)


class VNBatchNorm(nn.Module):
    def __init__(self, num_features, dim):
        super().__init__()
        self.dim = dim
        if dim == 3 or dim == 4:
            self.bn = nn.BatchNorm1d(num_features)
        elif dim == 5:
            self.bn = nn.BatchNorm2d(num_features)

    def forward_rows(self, rows):
        return ops.bn_leaky(rows, None, self.bn, self.training, 0.0)

    def forward(self, x):
        rows, B, sp = to_rows(x)
        return from_rows(self.forward_rows(rows), B, sp)


class VNLinearLeakyReLU(nn.Module):
    def __init__(self, in_channels, out_channels, dim=5, share_nonlinearity=False, negative_slope=0.2):
        super().__init__()
        self.dim = dim
        self.negative_slope = negative_slope
        self.map_to_feat = nn.Linear(in_channels, out_channels, bias=False)
        self.batchnorm = VNBatchNorm(out_channels, dim=dim)
        if share_nonlinearity == True:  # noqa: E712
            self.map_to_dir = nn.Linear(in_channels, 1, bias=False)
        else:
            self.map_to_dir = nn.Linear(in_channels, out_channels, bias=False)

    def forward_rows(self, rows, bias_rows=None, rows_per_sample=0, w_slice=None):
        """rows [R, K] -> [R, Cout].  bias_rows/rows_per_sample/w_slice implement the broadcast-channel split used by
        the networks: the GEMM runs on weight columns w_slice only and bias_rows [B*3, 2*Cout] carries the rest."""
        wf, wd = self.map_to_feat.weight, self.map_to_dir.weight
        if wd.shape[0] != wf.shape[0]:       # share_nonlinearity: direction has one channel
            assert bias_rows is None
            p = ops.linear_rows(rows, wf)
            d = ops.linear_rows(rows, wd).expand(-1, wf.shape[0]).contiguous()
            return ops.bn_leaky(p, d, self.batchnorm.bn, self.training, self.negative_slope)
        w = torch.cat([wf, wd], dim=0)       # one GEMM for feat and dir: the input rows are read once
        if w_slice is not None:
            w = w[:, w_slice]
        if self.batchnorm.bn.affine:
            fused = ops.linear_bn_leaky_fused_nograd(rows, w, bias_rows, rows_per_sample, self.batchnorm.bn, self.training,
                                                     self.negative_slope)
            if fused is not None:               # inference: BN + leaky in the GEMM epilogue, p / d never stored
                return fused
        return ops.linear_bn_leaky_rows(rows, w, bias_rows, rows_per_sample, self.batchnorm.bn, self.training, self.negative_slope)

    def forward(self, x):
        rows, B, sp = to_rows(x)
        return from_rows(self.forward_rows(rows), B, sp)


class VNLinearAndLeakyReLU(nn.Module):
    def __init__(self, in_channels, out_channels, dim=5, share_nonlinearity=False, use_batchnorm='norm', negative_slope=0.2):
        super().__init__()
        self.dim = dim
        self.share_nonlinearity = share_nonlinearity
        self.use_batchnorm = use_batchnorm
        self.negative_slope = negative_slope
        self.linear = VNLinear(in_channels, out_channels)
        self.leaky_relu = VNLeakyReLU(out_channels, share_nonlinearity=share_nonlinearity, negative_slope=negative_slope)
        if use_batchnorm != 'none':
            self.batchnorm = VNBatchNorm(out_channels, dim=dim)

    def forward_rows(self, rows):
        y = ops.linear_rows(rows, self.linear.map_to_feat.weight)
        if self.use_batchnorm != 'none':
            y = self.batchnorm.forward_rows(y)
        return _leaky_dir(y, y, self.leaky_relu.map_to_dir, self.negative_slope)

    def forward(self, x):
        rows, B, sp = to_rows(x)
        return from_rows(self.forward_rows(rows), B, sp)


class VNMaxPool(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.map_to_dir = nn.Linear(in_channels, in_channels, bias=False)
        self.last_idx = None        # selections of the most recent forward ([G, C] int64), for parity checks
        self.forced_idx = None      # teacher-forced selections; None = own arg-max

    def forward_rows(self, rows, G, N, tap=False):
        """tap=True returns (pooled, alias of rows): feed the alias to the dense consumer of the same activation (ops._MaxPoolGatherTap)"""
        with torch.no_grad():
            d = ops.gemm_rows(rows.detach(), self.map_to_dir.weight.detach())
        out, idx = ops.maxpool_rows(rows, d, G, N, self.forced_idx, tap and torch.is_grad_enabled() and rows.requires_grad)
        self.last_idx = idx
        if tap and not isinstance(out, tuple):
            return out, rows
        return out

    def forward(self, x):
        rows, B, sp = to_rows(x)
        if len(sp) == 0:
            raise ValueError("VNMaxPool needs a pooled (last) dimension")
        N = sp[-1]
        G = rows.shape[0] // (3 * N)
        out = self.forward_rows(rows, G, N)          # rows (g, v)
        return from_rows(out, B, sp[:-1])


def mean_pool(x, dim=-1, keepdim=False):
    return x.mean(dim=dim, keepdim=keepdim)


class VNStdFeature(nn.Module):
    def __init__(self, in_channels, dim=4, normalize_frame=False, share_nonlinearity=False, negative_slope=0.2):
        super().__init__()
        self.dim = dim
        self.normalize_frame = normalize_frame
        self.vn1 = VNLinearLeakyReLU(in_channels, in_channels // 2, dim=dim, share_nonlinearity=share_nonlinearity,
                                     negative_slope=negative_slope)
        self.vn2 = VNLinearLeakyReLU(in_channels // 2, in_channels // 4, dim=dim, share_nonlinearity=share_nonlinearity,
                                     negative_slope=negative_slope)
        if normalize_frame:
            self.vn_lin = nn.Linear(in_channels // 4, 2, bias=False)
        else:
            self.vn_lin = nn.Linear(in_channels // 4, 3, bias=False)

    def forward(self, x):
        rows, B, sp = to_rows(x)
        z = self.vn1.forward_rows(rows)
        z = self.vn2.forward_rows(z)
        zj = ops.linear_rows(z, self.vn_lin.weight)                  # rows (point, v) x J: the J = 3 (2) frame vectors of every point
        # frame (Gram-Schmidt + cross product when normalize_frame) and the projection of every channel onto it: one kernel each way
        # (csrc/vn_frame.cu); x_std [B, C, 3 (frame axis k), *spatial], frame [B, 3 (component), 3 (k), *spatial] as the reference returns it
        x_std, frame = ops.vn_frame(rows, zj)
        return from_rows(x_std, B, sp), from_rows(frame, B, sp)


class VNLayerNorm(nn.Module):
    """models/vn_layers.py:129-150: LayerNorm over the channel axis of the vector norms, x [B, C, 3, N]"""

    def __init__(self, num_features):
        super().__init__()
        self.layer_norm = nn.LayerNorm(num_features)

    def forward_rows(self, rows):
        return ops.vn_layernorm(rows, self.layer_norm)

    def forward(self, x):
        if x.dim() != 4:
            raise ValueError(f"VNLayerNorm expects [B, C, 3, N] (the reference transposes a 3-D norm tensor), got {tuple(x.shape)}")
        rows, B, sp = to_rows(x)
        return from_rows(self.forward_rows(rows), B, sp)
