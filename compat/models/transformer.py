"""reference: models/transformer.py:25-105"""
from vn_pointcloudcompletion_b200.transformer import Attention, VN_Block  # noqa: F401
