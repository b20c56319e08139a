"""reference: models/dgcnn.py -- VN_DGCNN_fps :164-324"""
from _unsupported import unsupported
from vn_pointcloudcompletion_b200.dgcnn import VN_DGCNN_fps  # noqa: F401
from vn_pointcloudcompletion_b200.graph_ops import KNN, furthest_point_sample, gather_operation  # noqa: F401

DGCNN = unsupported("DGCNN")                # non-VN baselines
DGCNN_fps = unsupported("DGCNN_fps")
