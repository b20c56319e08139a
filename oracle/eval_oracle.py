"""oracle/eval_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of the reference's evaluation extras.  Only tests/ may import this.

PARITY UNPINNED: the reference computes these through open3d 0.9 (compute_point_cloud_distance) and pyntcloud (VoxelGrid), neither of
which is installed here, and holds no test or golden vector for them.  Their published algorithms are restated:
  f_score            metrics/metric.py:31-48        Euclidean nearest-neighbour distances both ways, strict `< th`
  points_to_voxels   utils/voxel_util.py:89-105     pyntcloud VoxelGrid(n_x=n_y=n_z=n, regular_bounding_box=True): bounding box of the
                                                     cloud itself grown to a cube, segments = np.linspace(min, max, n + 1) per axis,
                                                     voxel index = clip(searchsorted(segments, x) - 1, 0, n - 1)  (float64 here)
  iou                utils/voxel_util.py:5-14       (plain numpy in the reference: this one IS pinned -- tests/golden/eval_extras.npz holds
                                                     the outputs of the reference's own function on seeded grids, tests/test_eval.py)
"""
from __future__ import annotations

import numpy as np

from . import vn_oracle as O


def f_score(pred, gt, th=0.01):
    """pred [N,3], gt [M,3] -> (precision, recall, F); distances = sqrt of the Chamfer oracle's squared fp32 distances"""
    d1, d2, _, _ = O.chamfer_forward(pred[None], gt[None])
    precision = float((np.sqrt(d1[0]) < np.float32(th)).sum()) / d1.shape[1]
    recall = float((np.sqrt(d2[0]) < np.float32(th)).sum()) / d2.shape[1]
    f = 2 * recall * precision / (recall + precision) if recall + precision else 0.0
    return precision, recall, f


def points_to_voxels(points, size_grid=64):
    n = size_grid
    pts = np.asarray(points, np.float64)
    lo, hi = pts.min(0), pts.max(0)
    margin = (hi - lo).max() - (hi - lo)
    lo, hi = lo - margin / 2, hi + margin / 2
    vox = np.zeros((n, n, n), bool)
    idx = []
    for a in range(3):
        seg = np.linspace(lo[a], hi[a], n + 1)
        idx.append(np.clip(np.searchsorted(seg, pts[:, a]) - 1, 0, n - 1))
    vox[idx[0], idx[1], idx[2]] = True
    return vox


def iou(a, b):
    return np.logical_and(a, b).sum() / np.logical_or(a, b).sum()


def evaluate_iou(pred, gt, size_grid=64):
    return iou(points_to_voxels(pred, size_grid), points_to_voxels(gt, size_grid))
