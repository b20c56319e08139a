"""pytorch3d.transforms as the reference's loops use it (train.py:20,131-138; test.py:14,58-65): Rotate(R=...).to(device),
RotateAxisAngle(angle=..., axis=..., degrees=...).to(device), random_rotations(n); row-vector convention, transform_points(p) = p @ R."""
import torch

from vn_pointcloudcompletion_b200.eval_metrics import RotateAxisAngle  # noqa: F401
from vn_pointcloudcompletion_b200.model import Rotate as _Rotate
from vn_pointcloudcompletion_b200.model import random_rotations as _random_rotations


class Rotate(_Rotate):
    def __init__(self, R, device=None):
        super().__init__(R if device is None else R.to(device))

    def to(self, device):
        self.R = self.R.to(device)
        return self


def random_rotations(n, dtype=None, device=None):
    return _random_rotations(n, device=device)
