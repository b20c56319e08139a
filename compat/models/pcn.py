"""reference: models/pcn.py -- VN_PointNet :110-184, VN_FoldingNet :319-389, Attention_VN_FoldingNet :392-520"""
from _unsupported import unsupported
from models.vn_layers import *  # noqa: F401,F403  (the reference's module does the same, models/pcn.py:3)
from vn_pointcloudcompletion_b200.pcn import Attention_VN_FoldingNet, VN_FoldingNet, VN_PointNet  # noqa: F401
from vn_pointcloudcompletion_b200.transformer import VN_Block  # noqa: F401

PCN = unsupported("PCN")                    # models/pcn.py:8-107   non-VN baseline
VN_PCN = unsupported("VN_PCN")              # models/pcn.py:187-247
FoldingNet = unsupported("FoldingNet")      # models/pcn.py:250-316 non-VN decoder
