"""reference: extensions/ChamferDistancePytorch/fscore.py:3-16"""
from vn_pointcloudcompletion_b200.loss_variants import fscore  # noqa: F401
