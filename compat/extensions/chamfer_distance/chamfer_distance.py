"""reference: extensions/chamfer_distance/chamfer_distance.py:29-84 (no JIT build, no chamfer_3D extension module: the kernels are in
libvnpcc.so behind include/vnpcc.h)"""
from vn_pointcloudcompletion_b200.chamfer_distance import ChamferDistance, chamfer_3DFunction  # noqa: F401
