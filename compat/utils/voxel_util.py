"""reference: utils/voxel_util.py:5-19,89-105 -- numpy in / numpy out like the reference (test.py:73-78), computed on the GPU"""
import numpy as np
import torch

from vn_pointcloudcompletion_b200 import eval_metrics as _E


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda() if isinstance(a, np.ndarray) else a


def iou(preds, gt):
    if isinstance(preds, np.ndarray):
        return np.sum(np.logical_and(preds, gt)) / np.sum(np.logical_or(preds, gt))
    return _E.iou(preds, gt)


def points_to_voxels(points, size_grid=64):
    out = _E.points_to_voxels(_dev(points), size_grid)
    return out.cpu().numpy() if isinstance(points, np.ndarray) else out


def evaluate_iou(preds_pc, gt_pc, size_grid=64):
    out = _E.evaluate_iou(_dev(preds_pc), _dev(gt_pc), size_grid)
    return float(out) if isinstance(preds_pc, np.ndarray) else out


def ply_to_voxels(filename=None, size_grid=64):
    """utils/voxel_util.py:69-87: occupancy grid of a PLY file's vertices"""
    if filename is None:
        return None
    return points_to_voxels(_E.read_point_cloud(filename), size_grid)


# unit cube: corner (a, b, c) has index 4a + 2b + c; 12 triangles (1-based OBJ indices are added when writing)
_CORNERS = np.array([[a, b, c] for a in (0, 1) for b in (0, 1) for c in (0, 1)], np.float64)
_TRIS = np.array([[0, 1, 2], [1, 3, 2], [2, 3, 6], [3, 7, 6], [0, 2, 6], [0, 6, 4], [0, 5, 1], [0, 4, 5], [6, 7, 5], [6, 5, 4], [1, 7, 3], [1, 5, 7]])


def voxel2mesh(voxels, surface_view):
    """utils/voxel_util.py:22-47 (an export helper, not on the hot path): one cube of edge 0.01 per occupied voxel (> 0.3), cubes spaced
    1.1 apart; with surface_view only voxels whose 3x3x3 neighbourhood is not completely occupied are kept.  Vectorised."""
    occ = np.asarray(voxels) > 0.3
    keep = occ.copy()
    if surface_view:
        pad = np.pad(occ, 1).astype(np.int32)
        n = occ.shape
        full = sum(pad[1 + a:1 + a + n[0], 1 + b:1 + b + n[1], 1 + c:1 + c + n[2]] for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1))
        keep &= full < 27
    ijk = np.argwhere(keep).astype(np.float64)                              # [V, 3]
    verts = 0.01 * (_CORNERS[None] + 1.1 * ijk[:, None, :]).reshape(-1, 3)
    faces = (_TRIS[None] + 1 + 8 * np.arange(len(ijk))[:, None, None]).reshape(-1, 3)
    return verts, faces


def write_obj(filename, verts, faces):
    with open(filename, "w") as f:
        f.write("g\n# %d vertex\n" % len(verts))
        f.writelines("v %f %f %f\n" % tuple(v) for v in verts)
        f.write("# %d faces\n" % len(faces))
        f.writelines("f %d %d %d\n" % tuple(t) for t in faces)


def voxel2obj(filename, pred, surface_view=True):
    write_obj(filename, *voxel2mesh(pred, surface_view))
