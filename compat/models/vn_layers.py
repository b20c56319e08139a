"""reference: models/vn_layers.py -- every class, served by the sm_100a kernels"""
from vn_pointcloudcompletion_b200.vn_layers import (EPS, VNBatchNorm, VNLayerNorm, VNLeakyReLU, VNLinear, VNLinearAndLeakyReLU,  # noqa: F401
                                                    VNLinearLeakyReLU, VNMaxPool, VNStdFeature, mean_pool)
