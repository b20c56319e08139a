"""Drop-in for extensions/chamfer_distance/chamfer_distance.py (reference :29-84): chamfer_3DFunction and
ChamferDistance with the same forward/backward signatures; squared distances; int32 indices."""
from __future__ import annotations

from torch import nn

from .ops import chamfer_3DFunction  # noqa: F401  (re-exported: same name as the reference's Function)


class ChamferDistance(nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, input1, input2):
        """input1 (B, N, 3), input2 (B, M, 3) -> dist1 (B, N), dist2 (B, M)"""
        dist1, dist2, _, _ = chamfer_3DFunction.apply(input1, input2)
        return dist1, dist2


class chamfer_3DDist(nn.Module):
    """extensions/ChamferDistancePytorch/chamfer3D/dist_chamfer_3D.py:67-74 (returns the indices too)"""

    def forward(self, input1, input2):
        return chamfer_3DFunction.apply(input1.contiguous(), input2.contiguous())
