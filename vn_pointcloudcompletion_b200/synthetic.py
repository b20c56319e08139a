"""Seeded synthetic PCN-shaped data: a complete cloud, a half-space partial view of it and one
SO(3) rotation per sample.  numpy only, so that the build container, the GPU box, the oracle and the golden
generator all see bit-identical inputs for a given seed.

Replaces, for benchmarking/tests only, dataset/shapenet.py:67-68 (2048-pt partial / 16384-pt complete) and the
pytorch3d `random_rotations` + `Rotate` of train.py:131-138 (row-vector convention: transform_points(p) = p @ R).
"""
from __future__ import annotations

import numpy as np


def random_rotations(B, rng):
    """Uniform SO(3) from normalised N(0,1) quaternions -> [B,3,3] float32 (proper rotations, det=+1)."""
    q = rng.standard_normal((B, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    r, i, j, k = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.stack([
        1 - 2 * (j * j + k * k), 2 * (i * j - k * r), 2 * (i * k + j * r),
        2 * (i * j + k * r), 1 - 2 * (i * i + k * k), 2 * (j * k - i * r),
        2 * (i * k - j * r), 2 * (j * k + i * r), 1 - 2 * (i * i + j * j)], axis=1).reshape(B, 3, 3)
    return R.astype(np.float32)


def make_batch(B, n_partial=2048, n_gt=16384, seed=1234, rotate=True):
    """returns (partial [B,n_partial,3], gt [B,n_gt,3], R [B,3,3]) float32, partial and gt already rotated by R."""
    rng = np.random.RandomState(seed)
    gt = rng.uniform(-0.5, 0.5, size=(B, n_gt, 3))
    partial = np.empty((B, n_partial, 3))
    for b in range(B):
        nrm = rng.standard_normal(3)
        nrm /= np.linalg.norm(nrm)
        side = np.nonzero(gt[b] @ nrm > 0)[0]
        pick = rng.choice(side, size=n_partial, replace=len(side) < n_partial)
        partial[b] = gt[b, pick] + rng.standard_normal((n_partial, 3)) * 1e-3
    R = random_rotations(B, rng) if rotate else np.tile(np.eye(3, dtype=np.float32), (B, 1, 1))
    partial = np.matmul(partial, R.astype(np.float64))
    gt = np.matmul(gt, R.astype(np.float64))
    return partial.astype(np.float32), gt.astype(np.float32), R
