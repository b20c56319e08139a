"""oracle/ref_chamfer.py -- TEST INFRASTRUCTURE.  Launches the reference's own Chamfer CUDA kernels (device code compiled
unmodified by oracle/build_ref.py into oracle/_ref/ref_chamfer3D.cubin) through the CUDA driver API, with the
reference's launch configuration (extensions/chamfer_distance/chamfer3D.cu:142-143 forward, :184-185 backward).
Used by the -m gpu parity tests (bit-exact dist / idx at full BASELINE sizes) and by tools/chamfer_sweep.py as the
"reference kernel recompiled for sm_100a" comparator.  Never imported by product code."""
import os

import numpy as np
import torch
from cuda.bindings import driver as cu

_HERE = os.path.dirname(os.path.abspath(__file__))
CUBIN = os.path.join(_HERE, "_ref", "ref_chamfer3D.cubin")
_mod = None
_fn = {}


def available():
    return os.path.exists(CUBIN)


def _check(res):
    err = res[0]
    if err != cu.CUresult.CUDA_SUCCESS:
        raise RuntimeError(f"CUDA driver error {err}")
    return res[1] if len(res) == 2 else res[1:]


def _load():
    global _mod
    if _mod is None:
        torch.cuda.init()
        torch.zeros(1, device="cuda")          # make torch's primary context current
        _mod = _check(cu.cuModuleLoad(CUBIN.encode()))
        _fn["fwd"] = _check(cu.cuModuleGetFunction(_mod, b"_Z16NmDistanceKerneliiPKfiS0_PfPi"))
        _fn["bwd"] = _check(cu.cuModuleGetFunction(_mod, b"_Z20NmDistanceGradKerneliiPKfiS0_S0_PKiPfS3_"))
    return _fn


def _launch(fn, grid, block, args, types):
    vals = []
    for a, t in zip(args, types):
        vals.append(np.array([a], dtype=t))
    ptrs = np.array([v.ctypes.data for v in vals], dtype=np.uint64)
    stream = torch.cuda.current_stream().cuda_stream
    _check(cu.cuLaunchKernel(fn, grid[0], grid[1], grid[2], block, 1, 1, 0, stream, ptrs.ctypes.data, 0))


def forward_into(xyz1, xyz2, d1, d2, i1, i2):
    """chamfer_cuda_forward (chamfer3D.cu:136-154) with the pybind `forward` contract (chamfer_cuda.cpp:17-21): caller-allocated outputs,
    raw data pointers (the reference reads .data<float>() and ignores strides)"""
    f = _load()
    B, N, _ = xyz1.shape
    M = xyz2.shape[1]
    T = [np.int32, np.int32, np.uint64, np.int32, np.uint64, np.uint64, np.uint64]
    _launch(f["fwd"], (32, 16, 1), 512, [B, N, xyz1.data_ptr(), M, xyz2.data_ptr(), d1.data_ptr(), i1.data_ptr()], T)
    _launch(f["fwd"], (32, 16, 1), 512, [B, M, xyz2.data_ptr(), N, xyz1.data_ptr(), d2.data_ptr(), i2.data_ptr()], T)


def forward(xyz1, xyz2):
    """returns dist1, dist2, idx1, idx2 (zero-initialised like chamfer_distance.py:40-44)"""
    B, N, _ = xyz1.shape
    M = xyz2.shape[1]
    xyz1, xyz2 = xyz1.contiguous(), xyz2.contiguous()
    d1 = torch.zeros(B, N, device="cuda")
    d2 = torch.zeros(B, M, device="cuda")
    i1 = torch.zeros(B, N, device="cuda", dtype=torch.int32)
    i2 = torch.zeros(B, M, device="cuda", dtype=torch.int32)
    forward_into(xyz1, xyz2, d1, d2, i1, i2)
    return d1, d2, i1, i2


def backward_into(xyz1, xyz2, gd1, gd2, i1, i2, g1, g2):
    """chamfer_cuda_backward (chamfer3D.cu:176-195): accumulates into the caller-zeroed g1, g2"""
    f = _load()
    B, N, _ = xyz1.shape
    M = xyz2.shape[1]
    T = [np.int32, np.int32, np.uint64, np.int32, np.uint64, np.uint64, np.uint64, np.uint64, np.uint64]
    _launch(f["bwd"], (1, 16, 1), 256, [B, N, xyz1.data_ptr(), M, xyz2.data_ptr(), gd1.data_ptr(), i1.data_ptr(),
                                       g1.data_ptr(), g2.data_ptr()], T)
    _launch(f["bwd"], (1, 16, 1), 256, [B, M, xyz2.data_ptr(), N, xyz1.data_ptr(), gd2.data_ptr(), i2.data_ptr(),
                                       g2.data_ptr(), g1.data_ptr()], T)


def backward(xyz1, xyz2, gd1, gd2, i1, i2):
    """returns gradxyz1, gradxyz2 (accumulated into zeros)"""
    g1 = torch.zeros_like(xyz1)
    g2 = torch.zeros_like(xyz2)
    backward_into(xyz1, xyz2, gd1.contiguous(), gd2.contiguous(), i1, i2, g1, g2)
    return g1, g2
