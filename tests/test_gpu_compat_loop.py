"""The drop-in claim end to end: a training / validation loop written against the REFERENCE'S module paths (models.model.PCNNet,
metrics.loss.cd_loss_L1, metrics.metric.l1_cd / f_score, utils.loss.calc_dcd, utils.voxel_util.evaluate_iou, pytorch3d.transforms) runs on
compat/ with plain torch.optim.Adam + StepLR exactly as train.py:60-93,127-186 writes it, matches the package's own DataParallelTrainer
(flat buffers + fused Adam) step for step, and the two optimizers read each other's checkpoints (train.py:72-80, 262-277)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(mode):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(REPO, "compat"), os.path.join(REPO, "compat_shims"), REPO])
    r = subprocess.run([sys.executable, os.path.join(REPO, "tests", "compat_loop_script.py"), mode], env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_reference_shaped_loop_matches_trainer_fp32():
    o = _run("fp32")
    print(o)
    assert o["losses"][0] == pytest.approx(o["trainer_losses"][0], rel=1e-6)
    # the two optimizers took the same step: identical BatchNorm buffers, Adam moments equal to the fp32 atomic-add noise of the backward
    # kernels, parameter updates equal except on elements whose gradient IS that noise (tests/test_gpu_trainer.py checks the fused Adam
    # kernel against torch.optim.Adam on identical gradients to 1e-5)
    assert o["step1_buffer_diff"] <= 1e-6 and o["step1_state_rel_diff"] <= 1e-4 and o["step1_update_rel_l2"] <= 5e-2, o
    assert o["step1_param_diff"] <= 2.5e-4, o
    assert o["losses"][1] == pytest.approx(o["trainer_losses"][1], rel=1e-3)
    assert o["lr"] == pytest.approx(o["trainer_lr"]) and o["lr"] == pytest.approx(1e-4 * 0.8)
    assert o["state_keys_equal"] and o["n_state"] == o["n_params"] - 2      # the two VNMaxPool direction weights never get Adam state
    assert o["resumed_step"] == 3 and o["ckpt_keys"] == ["best_epoch", "best_metrics", "epoch", "optim_state_dict"]
    assert 0.0 <= o["f_score"] <= 1.0 and 0.0 <= o["iou"] <= 1.0 and o["l1_cd"] > 0 and o["l2_cd"] > 0 and o["dcd"] > 0
    assert o["trainer_eval"][0] > 0 and o["trainer_eval"][2] == pytest.approx(o["trainer_eval"][0] + o["trainer_eval"][1])


def test_reference_shaped_loop_runs_in_tf32_mode():
    o = _run("tf32")
    assert all(x == x and x > 0 for x in o["losses"] + o["trainer_losses"])
    assert o["step1_state_rel_diff"] <= 1e-3 and o["step1_update_rel_l2"] <= 1e-1, o
