"""reference: models/model.py:9-64"""
from models.dgcnn import *  # noqa: F401,F403
from models.pcn import *  # noqa: F401,F403
from models.vn_layers import *  # noqa: F401,F403
from vn_pointcloudcompletion_b200.model import PCNNet  # noqa: F401
