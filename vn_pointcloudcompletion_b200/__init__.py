"""vn_pointcloudcompletion_b200 -- B200 (sm_100a) implementation of the data-parallel hot path of
ChenBarryHu/VN_PointCloudCompletion: the Vector-Neuron layer stack (vn_pointnet encoder + vn_foldingnet decoder)
and the Chamfer loss/metric, behind the reference's own Python interfaces.  See DESIGN.md / INTEGRATION.md."""
from . import _lib  # noqa: F401
from .chamfer_distance import ChamferDistance, chamfer_3DDist, chamfer_3DFunction  # noqa: F401
from .dgcnn import VN_DGCNN_fps  # noqa: F401
from .eval_metrics import RotateAxisAngle, evaluate_iou, f_score, points_to_voxels, random_sample, read_point_cloud  # noqa: F401
from .graph_ops import KNN, furthest_point_sample, gather_operation  # noqa: F401
from .loss import cd_loss_L1, cd_loss_L2, l1_cd, l2_cd  # noqa: F401
from .model import PCNNet, Rotate, random_rotations  # noqa: F401
from .ops import get_fp32_impl, get_gemm_mode, set_fp32_impl, set_gemm_mode  # noqa: F401
from .pcn import Attention_VN_FoldingNet, VN_FoldingNet, VN_PointNet  # noqa: F401
from .vn_layers import (VNBatchNorm, VNLeakyReLU, VNLinear, VNLinearAndLeakyReLU, VNLinearLeakyReLU,  # noqa: F401
                        VNLayerNorm, VNMaxPool, VNStdFeature, mean_pool)
from .transformer import Attention, VN_Block  # noqa: F401
