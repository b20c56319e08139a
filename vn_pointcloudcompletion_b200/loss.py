"""Drop-in CD entry points of metrics/loss.py:20-43 (training losses) and metrics/metric.py:12-23 (eval metrics).
The sqrt / mean tails run as one fused reduction kernel (and one backward kernel) instead of 4-6 ATen launches."""
from __future__ import annotations

from . import ops
from .chamfer_distance import ChamferDistance

CD = ChamferDistance()


def cd_loss_L1(pcs1, pcs2):
    """(mean(sqrt(dist1)) + mean(sqrt(dist2))) / 2 over the whole batch  -- metrics/loss.py:20-31"""
    dist1, dist2 = CD(pcs1, pcs2)
    return ops.cd_reduce(dist1, dist2, 0)


def cd_loss_L2(pcs1, pcs2):
    """mean(dist1) + mean(dist2)  -- metrics/loss.py:34-43"""
    dist1, dist2 = CD(pcs1, pcs2)
    return ops.cd_reduce(dist1, dist2, 1)


def l1_cd(pcs1, pcs2):
    """sum_b (mean_n sqrt(dist1) + mean_m sqrt(dist2)) / 2  -- metrics/metric.py:19-23"""
    dist1, dist2 = CD(pcs1, pcs2)
    return ops.cd_reduce(dist1, dist2, 2)


def l2_cd(pcs1, pcs2):
    """sum_b (mean_n dist1 + mean_m dist2)  -- metrics/metric.py:12-16"""
    dist1, dist2 = CD(pcs1, pcs2)
    return ops.cd_reduce(dist1, dist2, 3)


# SURVEY.md 8f row f3 (utils/loss.py:14-74, fscore.py:3-16): kernels in csrc/loss_variants.cu, host side in loss_variants.py
from .loss_variants import calc_cd, calc_dcd, fscore  # noqa: E402,F401
