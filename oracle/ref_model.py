"""oracle/ref_model.py -- TEST / BASELINE INFRASTRUCTURE.  Imports the reference's OWN Python implementation (byte-compiled, unmodified,
by oracle/build_ref_py.py into oracle/_ref/refpy.zip) through the import shim of SURVEY.md Appendix F, so that

  * bench.py --impl reference and bench.py's cpu_baseline leg time models.model.PCNNet + metrics.loss.cd_loss_L1 on the box's host
    cores with the reference's CPU-capable Chamfer (chamfer_python.distChamfer) -- BASELINE.md 3;
  * the -m gpu parity tests run the same unmodified reference eagerly on the B200 (its ATen operator chain + its own Chamfer
    kernels from oracle/_ref/ref_chamfer3D.cubin) as the full-size oracle.

Never imported by product code (tests/test_abi.py checks).  Stubs are installed only for the reference's import-time-only
dependencies that are absent from the image (pointnet2_ops, knn_cuda, timm, open3d, emd); none of them is executed on this path.

backend "cpu" : extensions.chamfer_distance.chamfer_distance.ChamferDistance -> chamfer_python.distChamfer (the CUDA extension has no
                CPU path; this is the substitution BASELINE.md 3 prescribes)
backend "cuda": the reference's own extensions/chamfer_distance/chamfer_distance.py wrapper (compiled, unmodified) on top of a
                `chamfer_3D` module whose forward / backward launch the reference's own kernels (oracle/ref_chamfer.py)
"""
import importlib
import importlib.util
import os
import sys
import types
from types import SimpleNamespace

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
PY = os.path.join(HERE, "_ref", "refpy.zip")      # sourceless byte code, imported through zipimport
_loaded = None


def available():
    return os.path.exists(PY)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load(backend="cpu"):
    """returns a namespace with the reference's modules: .model (models.model), .loss (metrics.loss), .vn_layers, .pcn,
    .chamfer_python, .utils_loss.  One backend per process (metrics/loss.py binds CD = ChamferDistance() at import, :16)."""
    global _loaded
    if _loaded is not None:
        if _loaded.backend != backend:
            raise RuntimeError(f"reference already loaded with backend {_loaded.backend}")
        return _loaded
    if not available():
        raise RuntimeError("oracle/_ref/refpy.zip is missing: run `python oracle/build_ref_py.py` in the build container")
    for shadow in ("models", "metrics", "utils", "extensions"):
        if shadow in sys.modules:
            raise RuntimeError(f"a module named {shadow!r} is already imported; the reference's package of that name cannot be loaded")
    sys.path.insert(0, PY)
    sys.path.insert(0, os.path.join(PY, "extensions", "ChamferDistancePytorch"))
    # ---- import-time-only third-party dependencies ----
    pu = _stub("pointnet2_ops.pointnet2_utils")
    _stub("pointnet2_ops", pointnet2_utils=pu)                                       # models/pcn.py:4, models/dgcnn.py:7
    _stub("knn_cuda", KNN=type("KNN", (), {"__init__": lambda self, k, transpose_mode=False: None}))      # models/dgcnn.py:6,11
    tl = _stub("timm.models.layers", DropPath=torch.nn.Identity, trunc_normal_=torch.nn.init.trunc_normal_)
    tm = _stub("timm.models", layers=tl)
    _stub("timm", models=tm)                                                          # models/pointr/vn_pointr.py:4
    _stub("open3d")                                                                   # metrics/metric.py:2
    emd = _stub("extensions.earth_movers_distance.emd",
                EarthMoverDistance=type("EarthMoverDistance", (torch.nn.Module,), {}))      # metrics/loss.py:13 (never called here)
    d3 = _stub("chamfer3D.dist_chamfer_3D")                                           # metrics/loss.py:10 (never called here)
    _stub("chamfer3D", dist_chamfer_3D=d3)
    ext = _stub("extensions")
    ext.__path__ = [os.path.join(PY, "extensions")]
    emdp = _stub("extensions.earth_movers_distance", emd=emd)
    ext.earth_movers_distance = emdp
    if not hasattr(importlib, "find_loader"):                                         # chamfer_distance.py:9 on Python >= 3.12
        importlib.find_loader = lambda name: (sys.modules.get(name) or importlib.util.find_spec(name))
    import chamfer_python                                                             # the reference's CPU Chamfer

    if backend == "cpu":
        class ChamferDistance(torch.nn.Module):
            """stands in for the CUDA-only extension on the host: same return contract as chamfer_distance.py:74-84"""

            def forward(self, input1, input2):
                d1, d2, _, _ = chamfer_python.distChamfer(input1, input2)
                return d1, d2
        cdm = _stub("extensions.chamfer_distance.chamfer_distance", ChamferDistance=ChamferDistance)
        cdp = _stub("extensions.chamfer_distance", chamfer_distance=cdm)
        ext.chamfer_distance = cdp
    elif backend == "cuda":
        from oracle import ref_chamfer as RC

        def _fwd(xyz1, xyz2, dist1, dist2, idx1, idx2):      # chamfer_cuda.cpp:17-21
            RC.forward_into(xyz1, xyz2, dist1, dist2, idx1, idx2)
            return 1

        def _bwd(xyz1, xyz2, gradxyz1, gradxyz2, graddist1, graddist2, idx1, idx2):      # chamfer_cuda.cpp:24-27
            RC.backward_into(xyz1, xyz2, graddist1, graddist2, idx1, idx2, gradxyz1, gradxyz2)
            return 1
        _stub("chamfer_3D", forward=_fwd, backward=_bwd)
        cdp = _stub("extensions.chamfer_distance")
        cdp.__path__ = [os.path.join(PY, "extensions", "chamfer_distance")]
        ext.chamfer_distance = cdp
        import extensions.chamfer_distance.chamfer_distance  # noqa: F401  (the reference's own wrapper, compiled)
    else:
        raise ValueError(backend)

    import metrics.loss as ref_loss
    import models.model as ref_model
    import models.pcn as ref_pcn
    import models.vn_layers as ref_vn
    import utils.loss as ref_uloss
    _loaded = SimpleNamespace(backend=backend, model=ref_model, loss=ref_loss, pcn=ref_pcn, vn_layers=ref_vn,
                              chamfer_python=chamfer_python, utils_loss=ref_uloss)
    return _loaded


class Rotate:
    """pytorch3d.transforms.Rotate stand-in for the decoder's duck-typed `rot` (models/pcn.py:369-370): row-vector convention,
    transform_points(p) = p @ R (third-party arithmetic (2))"""

    def __init__(self, R):
        self.R = R

    def transform_points(self, p):
        return torch.matmul(p, self.R)


def build_pcnnet(device="cpu", enc_type="vn_pointnet", dec_type="vn_foldingnet", seed=0, backend=None):
    """models.model.PCNNet(config, enc_type, dec_type) exactly as train.py:60 builds it; random init under torch.manual_seed(seed)"""
    ref = load(backend or ("cuda" if str(device).startswith("cuda") else "cpu"))
    cfg = SimpleNamespace(num_coarse=1024, latent_dim=2048, only_coarse=False, device=device, enc_pretrained="none")
    torch.manual_seed(seed)
    if str(device) == "cpu":
        # VN_FoldingNet.__init__ calls .cuda() on its folding seed (models/pcn.py:362): identity while constructing on the host
        orig = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self
        try:
            net = ref.model.PCNNet(cfg, enc_type=enc_type, dec_type=dec_type)
        finally:
            torch.Tensor.cuda = orig
    else:
        net = ref.model.PCNNet(cfg, enc_type=enc_type, dec_type=dec_type)
    return net, ref
