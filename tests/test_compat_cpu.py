"""CPU checks of the drop-in layer: compat/ serves the reference's module paths from the package; the reference's OWN unmodified train.py
and test.py import against it (build container only: needs /root/reference); FlatAdam speaks torch.optim.Adam's checkpoint format and
is driven by torch's StepLR (train.py:70,93,262-277)."""
import json
import os
import subprocess
import sys

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def _py(code, extra_path=()):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(REPO, "compat"), os.path.join(REPO, "compat_shims"), REPO] + list(extra_path))
    env["OUTPUT_DIR"] = "/tmp"
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_compat_serves_reference_module_paths():
    o = _py("""
import json
from models.model import PCNNet
from models.vn_layers import VNLinear, VNLeakyReLU, VNLinearLeakyReLU, VNLinearAndLeakyReLU, VNBatchNorm, VNMaxPool, VNStdFeature, mean_pool
from models.pcn import VN_PointNet, VN_FoldingNet, Attention_VN_FoldingNet, PCN
from models import PCN as PCN2, VN_PCN, DGCNN
from models.dgcnn import VN_DGCNN_fps, DGCNN_fps
from models.transformer import VN_Block
from metrics.loss import cd_loss_L1, cd_loss_L2, emd_loss
from metrics.metric import l1_cd, l2_cd, emd, f_score
from extensions.chamfer_distance.chamfer_distance import ChamferDistance, chamfer_3DFunction
from extensions.earth_movers_distance.emd import EarthMoverDistance
from chamfer3D import dist_chamfer_3D
from fscore import fscore
from utils.loss import calc_cd, calc_dcd
from utils.voxel_util import points_to_voxels, evaluate_iou
from pytorch3d.transforms import Rotate, RotateAxisAngle, random_rotations
mods = {f.__name__: f.__module__ for f in (PCNNet, VNLinear, VNStdFeature, VN_PointNet, VN_FoldingNet, cd_loss_L1, l1_cd, ChamferDistance,
                                           chamfer_3DFunction, calc_dcd, fscore, VN_DGCNN_fps, VN_Block)}
try:
    PCN(16384, 1024, 4)
    mods['PCN'] = 'constructed'
except NotImplementedError:
    mods['PCN'] = 'NotImplementedError'
EarthMoverDistance()
print(json.dumps(mods))
""")
    assert o.pop("PCN") == "NotImplementedError"
    assert all(m.startswith("vn_pointcloudcompletion_b200") for m in o.values()), o


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs /root/reference (build container)")
def test_reference_train_and_test_scripts_import_unmodified_against_compat():
    """sys.path = [compat, compat_shims, repo, test shims for tensorboardX / open3d / matplotlib, the reference checkout]: the reference's
    own train.py and test.py import with zero edits; hot-path names resolve to the package, everything else to the reference's files"""
    o = _py("""
import json, os, sys
os.chdir('/tmp')
import train, test
import utils.experiments, dataset
print(json.dumps({'PCNNet': train.PCNNet.__module__, 'cd_loss_L1': train.cd_loss_L1.__module__, 'l1_cd': train.l1_cd.__module__,
                  'calc_dcd': train.calc_dcd.__module__, 'test_PCNNet': test.PCNNet.__module__, 'f_score': test.f_score.__module__,
                  'evaluate_iou': test.evaluate_iou.__module__, 'Rotate': train.Rotate.__module__,
                  'experiments': utils.experiments.__file__, 'dataset': dataset.__file__, 'train': train.__file__}))
""", extra_path=[os.path.join(REPO, "tests", "shims"), REF])
    assert o["PCNNet"] == "vn_pointcloudcompletion_b200.model" and o["test_PCNNet"] == "vn_pointcloudcompletion_b200.model"
    assert o["cd_loss_L1"] == "vn_pointcloudcompletion_b200.loss" and o["l1_cd"] == "vn_pointcloudcompletion_b200.loss"
    assert o["calc_dcd"] == "vn_pointcloudcompletion_b200.loss_variants"
    assert o["f_score"] == "metrics.metric" and o["evaluate_iou"] == "utils.voxel_util"
    assert o["Rotate"] == "pytorch3d.transforms"
    assert o["experiments"].startswith(REF) and o["dataset"].startswith(REF) and o["train"].startswith(REF)


def test_flat_adam_speaks_torch_adam_checkpoints_and_steplr():
    from vn_pointcloudcompletion_b200.trainer import FlatAdam
    shapes = [(3, 4), (5,), (2,)]
    ref = torch.optim.Adam([torch.nn.Parameter(torch.randn(*s)) for s in shapes], lr=5e-4)
    for p in ref.param_groups[0]["params"][:2]:      # the third parameter never receives a gradient (like VNMaxPool.map_to_dir)
        p.grad = torch.ones_like(p)
    ref.step()
    ref.step()
    sd = ref.state_dict()
    mine = FlatAdam([torch.nn.Parameter(torch.randn(*s)) for s in shapes])
    assert isinstance(mine, torch.optim.Optimizer)
    mine.load_state_dict(sd)
    assert mine.step_count == 2 and mine.lr == 5e-4
    assert torch.equal(mine.m[:12].view(3, 4), sd["state"][0]["exp_avg"]) and torch.equal(mine.v[12:17], sd["state"][1]["exp_avg_sq"])
    assert float(mine.m[17:].abs().sum()) == 0.0
    back = mine.state_dict()
    assert sorted(back["state"].keys()) == [0, 1] and back["param_groups"][0]["params"] == [0, 1, 2]
    ref2 = torch.optim.Adam([torch.nn.Parameter(torch.randn(*s)) for s in shapes])
    ref2.load_state_dict(back)      # torch's own loader accepts it
    assert torch.equal(ref2.state_dict()["state"][1]["exp_avg"], sd["state"][1]["exp_avg"])
    sched = torch.optim.lr_scheduler.StepLR(mine, step_size=50, gamma=0.8)      # train.py:93
    for _ in range(50):
        sched.step()
    assert mine.lr == pytest.approx(5e-4 * 0.8)
    # legacy (round-1) flat format still loads
    mine.load_state_dict({"step": 7, "exp_avg": torch.zeros(19), "exp_avg_sq": torch.ones(19), "lr": 1e-3})
    assert mine.step_count == 7 and mine.lr == 1e-3
    # a parameter that lost its alias of the flat buffer is detected
    mine.params[0].data = torch.zeros(3, 4)
    with pytest.raises(RuntimeError):
        mine._check_bindings()
