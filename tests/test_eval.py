"""Evaluation extras / input helpers: CUDA path vs the numpy oracle (oracle/eval_oracle.py).
F-score counts and voxel occupancy are integer work: exact.  ("parity unpinned" against open3d / pyntcloud, see the oracle header.)"""
import os
import struct

import numpy as np
import pytest

from oracle import eval_oracle as EO


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_voxel_oracle_properties():
    rng = np.random.RandomState(0)
    pts = rng.uniform(-0.3, 0.4, (500, 3)).astype(np.float32) * np.array([1.0, 0.5, 0.25], np.float32)
    v = EO.points_to_voxels(pts, 16)
    assert v.shape == (16, 16, 16) and 0 < v.sum() <= 500
    assert v[:, :3].sum() == 0 and v[:, 13:].sum() == 0          # regular (cubic) bounding box: the short axes are centred
    assert EO.iou(v, v) == 1.0
    assert abs(EO.evaluate_iou(pts, pts + 10.0, 16) - 1.0) < 1e-9   # translation invariant: each cloud uses its own box


def test_read_ply_roundtrip(tmp_path):
    from vn_pointcloudcompletion_b200.eval_metrics import read_point_cloud
    rng = np.random.RandomState(1)
    pts = rng.standard_normal((37, 3)).astype(np.float32)
    a = tmp_path / "a.ply"
    with open(a, "w") as f:
        f.write("ply\nformat ascii 1.0\ncomment test\nelement vertex 37\nproperty float x\nproperty float y\nproperty float z\nend_header\n")
        for p in pts:
            f.write("%.9g %.9g %.9g\n" % tuple(p))
    np.testing.assert_array_equal(read_point_cloud(str(a)), pts)
    b = tmp_path / "b.ply"
    with open(b, "wb") as f:
        f.write(b"ply\nformat binary_little_endian 1.0\nelement vertex 37\nproperty double x\nproperty double y\nproperty double z\n"
                b"property uchar red\nelement face 0\nproperty list uchar int vertex_indices\nend_header\n")
        for p in pts:
            f.write(struct.pack("<dddB", float(p[0]), float(p[1]), float(p[2]), 7))
    np.testing.assert_array_equal(read_point_cloud(str(b)), pts)


@pytest.mark.gpu
@pytest.mark.parametrize("N,M,th", [(300, 500, 0.05), (2048, 16384, 0.01), (1, 7, 0.2)])
def test_f_score_vs_oracle(N, M, th):
    from vn_pointcloudcompletion_b200 import eval_metrics as E
    rng = np.random.RandomState(N)
    pred = rng.uniform(-0.5, 0.5, (3, N, 3)).astype(np.float32)
    gt = rng.uniform(-0.5, 0.5, (3, M, 3)).astype(np.float32)
    got = E.f_score_batch(_dev(pred), _dev(gt), th).cpu().numpy()
    for b in range(3):
        want = EO.f_score(pred[b], gt[b], th)
        np.testing.assert_allclose(got[b], want, rtol=1e-6, atol=1e-7)
    one = E.f_score(_dev(pred[0]), _dev(gt[0]), th)
    assert one.dim() == 0 and abs(one.item() - EO.f_score(pred[0], gt[0], th)[2]) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("N,n", [(500, 16), (16384, 64), (3, 8)])
def test_voxels_vs_oracle(N, n):
    from vn_pointcloudcompletion_b200 import eval_metrics as E
    rng = np.random.RandomState(N + n)
    a = (rng.uniform(-0.4, 0.5, (2, N, 3)) * np.array([1.0, 0.7, 0.3])).astype(np.float32)
    b = (a + rng.standard_normal(a.shape) * 0.02).astype(np.float32)
    va = E.points_to_voxels(_dev(a), n).cpu().numpy()
    for s in range(2):
        assert np.array_equal(va[s], EO.points_to_voxels(a[s], n))
    got = E.evaluate_iou(_dev(a), _dev(b), n).cpu().numpy()
    for s in range(2):
        assert abs(got[s] - EO.evaluate_iou(a[s], b[s], n)) < 1e-6
    assert abs(E.iou(E.points_to_voxels(_dev(a[0]), n), E.points_to_voxels(_dev(b[0]), n)).item() - got[0]) < 1e-6


@pytest.mark.gpu
def test_input_helpers():
    import torch

    from vn_pointcloudcompletion_b200 import eval_metrics as E
    g = torch.Generator(device="cuda").manual_seed(0)
    pc = torch.arange(30, device="cuda", dtype=torch.float32).view(10, 3)
    s = E.random_sample(pc, 4, g)
    assert s.shape == (4, 3) and len({tuple(r) for r in s.cpu().tolist()}) == 4
    s = E.random_sample(pc, 25, g)
    assert s.shape == (25, 3) and {tuple(r) for r in s[:10].cpu().tolist()} == {tuple(r) for r in pc.cpu().tolist()}
    rot = E.RotateAxisAngle(torch.tensor([90.0, 0.0]), axis="Z", degrees=True).to("cuda")
    p = torch.tensor([[[1.0, 0.0, 0.0]], [[1.0, 2.0, 3.0]]], device="cuda")
    out = rot.transform_points(p).cpu().numpy()
    np.testing.assert_allclose(out[0, 0], [0.0, 1.0, 0.0], atol=1e-6)      # +90 degrees about Z takes x to y
    np.testing.assert_allclose(out[1, 0], [1.0, 2.0, 3.0], atol=1e-6)
    # orthonormal, det +1
    R = rot.R.cpu().numpy()
    np.testing.assert_allclose(R @ R.transpose(0, 2, 1), np.tile(np.eye(3), (2, 1, 1)), atol=1e-6)


# ---- the pure-numpy part of the reference's eval extras, pinned to the reference run (tests/golden/make_golden.py gen_eval_extras) ----
def _extras():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eval_extras.npz"))


def _compat_voxel_util():
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "compat", "utils", "voxel_util.py")
    spec = importlib.util.spec_from_file_location("compat_voxel_util", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_iou_vs_reference_golden():
    """utils/voxel_util.py:5-13 run unmodified on seeded grids: the oracle, the compat mirror (numpy path) and torch path agree with it"""
    import torch
    from vn_pointcloudcompletion_b200 import eval_metrics as E
    g = _extras()
    vu = _compat_voxel_util()
    a, b, want = g["iou.a"], g["iou.b"], g["iou.out"]
    for i in range(4):
        assert EO.iou(a[i], b[i]) == want[i]
        assert vu.iou(a[i], b[i]) == want[i]
        assert abs(E.iou(torch.from_numpy(a[i]), torch.from_numpy(b[i])).item() - want[i]) < 1e-7
    assert vu.iou(a, b) == want[4]


def test_voxel_mesh_export_vs_reference_golden(tmp_path):
    """utils/voxel_util.py:22-62 (voxel2mesh / write_obj / voxel2obj): same vertices, faces and OBJ bytes as the reference's loops"""
    g = _extras()
    vu = _compat_voxel_util()
    vox = g["mesh.vox"]
    for name, sv in (("surface", True), ("all", False)):
        verts, faces = vu.voxel2mesh(vox.copy(), sv)
        np.testing.assert_allclose(verts, g[f"mesh.{name}.verts"], rtol=0, atol=1e-15)
        np.testing.assert_array_equal(faces, g[f"mesh.{name}.faces"])
    assert len(g["mesh.surface.verts"]) < len(g["mesh.all.verts"])          # the solid block's interior is hidden in surface view
    path = tmp_path / "m.obj"
    vu.voxel2obj(str(path), vox.copy(), True)
    assert open(path, "rb").read() == g["mesh.obj_text"].tobytes()
