"""reference: utils/loss.py:14-74"""
from vn_pointcloudcompletion_b200.loss_variants import calc_cd, calc_dcd, fscore  # noqa: F401
