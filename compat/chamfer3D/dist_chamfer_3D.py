"""reference: extensions/ChamferDistancePytorch/chamfer3D/dist_chamfer_3D.py:67-74"""
from vn_pointcloudcompletion_b200.chamfer_distance import chamfer_3DDist, chamfer_3DFunction  # noqa: F401
