"""reference: metrics/metric.py:8-48.  l1_cd / l2_cd take CUDA tensors; f_score takes what the reference's test.py hands it (numpy
arrays [N,3], test.py:76) or CUDA tensors, and runs the search on the GPU either way."""
import numpy as np
import torch

from extensions.chamfer_distance.chamfer_distance import ChamferDistance
from vn_pointcloudcompletion_b200 import eval_metrics as _E
from vn_pointcloudcompletion_b200.loss import l1_cd, l2_cd  # noqa: F401

CD = ChamferDistance()
EMD = None


def emd(pcs1, pcs2):
    raise NotImplementedError("emd (extensions/earth_movers_distance) is outside the B200 hot path")


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda() if isinstance(a, np.ndarray) else a


def f_score(pred, gt, th=0.01):
    """metrics/metric.py:31-48 -> Python float like the reference"""
    return float(_E.f_score(_dev(pred), _dev(gt), th))
