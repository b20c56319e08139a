"""Chamfer-based loss variants of the reference on top of the sm_100a nearest-neighbour search:

  calc_cd    utils/loss.py:14-31     per-sample CD-L1 / CD-L2 (+ the `separate` directed means, + F-score at 1e-4, + raw search outputs)
  calc_dcd   utils/loss.py:33-74     density-aware Chamfer distance (Wu et al. 2021): exp(-alpha d) weighted by the inverse number of queries
                                      that share a nearest neighbour
  fscore     extensions/ChamferDistancePytorch/fscore.py:3-16

Every tail -- the per-sample means, the neighbour histogram, the exp-weighted means, the thresholds and all their backward passes -- is a
kernel of csrc/loss_variants.cu on the search's (dist, idx); gradients reach the clouds through chamfer_3DFunction's backward kernel.
Argument names, defaults, return order and shapes follow the reference."""
from __future__ import annotations

import torch

from ._lib import call, ptr, stream
from .chamfer_distance import chamfer_3DDist


class _CDPerSample(torch.autograd.Function):
    """[B,6] = per sample (cd_p, cd_t, mean sqrt d1, mean sqrt d2, mean d1, mean d2)"""

    @staticmethod
    def forward(ctx, dist1, dist2):
        dist1, dist2 = dist1.contiguous(), dist2.contiguous()
        B, N = dist1.shape
        M = dist2.shape[1]
        out = torch.empty((B, 6), device=dist1.device, dtype=torch.float32)
        part = torch.empty((B, 4), device=dist1.device, dtype=torch.float32)
        call("vnpcc_cd_persample_fwd", ptr(dist1), ptr(dist2), B, N, M, ptr(part), ptr(out), stream())
        ctx.save_for_backward(dist1, dist2)
        return out

    @staticmethod
    def backward(ctx, gout):
        dist1, dist2 = ctx.saved_tensors
        B, N = dist1.shape
        M = dist2.shape[1]
        g1, g2 = torch.empty_like(dist1), torch.empty_like(dist2)
        call("vnpcc_cd_persample_bwd", ptr(dist1), ptr(dist2), B, N, M, ptr(gout.contiguous().float()), ptr(g1), ptr(g2), stream())
        return g1, g2


class _DCDTail(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dist1, dist2, idx1, idx2, alpha, n_lambda, frac_21, frac_12):
        dist1, dist2, idx1, idx2 = dist1.contiguous(), dist2.contiguous(), idx1.contiguous(), idx2.contiguous()
        B, N = dist1.shape
        M = dist2.shape[1]
        dev = dist1.device
        cnt1 = torch.empty((B, M), device=dev, dtype=torch.int32)      # idx1 [B,N] points into the other cloud (M points)
        cnt2 = torch.empty((B, N), device=dev, dtype=torch.int32)
        call("vnpcc_nn_counts", ptr(idx1), B, N, M, ptr(cnt1), stream())
        call("vnpcc_nn_counts", ptr(idx2), B, M, N, ptr(cnt2), stream())
        loss = torch.empty(B, device=dev, dtype=torch.float32)
        part = torch.empty((B, 2), device=dev, dtype=torch.float32)
        call("vnpcc_dcd_fwd", ptr(dist1), ptr(dist2), ptr(idx1), ptr(idx2), ptr(cnt1), ptr(cnt2), B, N, M, float(alpha), float(n_lambda),
             float(frac_21), float(frac_12), ptr(part), ptr(loss), stream())
        ctx.save_for_backward(dist1, dist2, idx1, idx2, cnt1, cnt2)
        ctx.cfg = (float(alpha), float(n_lambda), float(frac_21), float(frac_12))
        return loss

    @staticmethod
    def backward(ctx, gloss):
        dist1, dist2, idx1, idx2, cnt1, cnt2 = ctx.saved_tensors
        alpha, n_lambda, frac_21, frac_12 = ctx.cfg
        B, N = dist1.shape
        M = dist2.shape[1]
        g1, g2 = torch.empty_like(dist1), torch.empty_like(dist2)
        call("vnpcc_dcd_bwd", ptr(dist1), ptr(dist2), ptr(idx1), ptr(idx2), ptr(cnt1), ptr(cnt2), B, N, M, alpha, n_lambda, frac_21, frac_12,
             ptr(gloss.contiguous().float()), ptr(g1), ptr(g2), stream())
        return g1, g2, None, None, None, None, None, None


def fscore(dist1, dist2, threshold=0.001):
    """-> (fscore [B], precision_1 [B], precision_2 [B]) on squared distances"""
    dist1, dist2 = dist1.detach().contiguous(), dist2.detach().contiguous()
    B, N = dist1.shape
    out = torch.empty((B, 3), device=dist1.device, dtype=torch.float32)
    call("vnpcc_fscore_sq", ptr(dist1), ptr(dist2), B, N, dist2.shape[1], float(threshold), ptr(out), stream())
    return out[:, 0], out[:, 1], out[:, 2]


def calc_cd(output, gt, calc_f1=False, return_raw=False, normalize=False, separate=False):
    """-> [cd_p [B], cd_t [B]] (or, separate=True, two [2,B] stacks of the directed means) (+ f1) (+ dist1, dist2, idx1, idx2);
    the search runs with gt as the first cloud, so dist1 / idx1 belong to the gt points.  `normalize` is accepted and unused, as upstream."""
    dist1, dist2, idx1, idx2 = chamfer_3DDist()(gt, output)
    o = _CDPerSample.apply(dist1, dist2)
    res = [o[:, 2:4].t(), o[:, 4:6].t()] if separate else [o[:, 0], o[:, 1]]
    if calc_f1:
        res.append(fscore(dist1, dist2, 0.0001)[0])
    if return_raw:
        res.extend([dist1, dist2, idx1, idx2])
    return res


def calc_dcd(x, gt, alpha=1000, n_lambda=1, return_raw=False, non_reg=False):
    """-> [loss [B], cd_p [B], cd_t [B]] (+ dist1, dist2, idx1, idx2)"""
    x, gt = x.float(), gt.float()
    n_x, n_gt = x.shape[1], gt.shape[1]
    assert x.shape[0] == gt.shape[0]
    frac_12, frac_21 = n_x / n_gt, n_gt / n_x
    if non_reg:
        frac_12, frac_21 = max(1, frac_12), max(1, frac_21)
    cd_p, cd_t, dist1, dist2, idx1, idx2 = calc_cd(x, gt, return_raw=True)
    res = [_DCDTail.apply(dist1, dist2, idx1, idx2, alpha, n_lambda, frac_21, frac_12), cd_p, cd_t]
    if return_raw:
        res.extend([dist1, dist2, idx1, idx2])
    return res
