import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(REPO, "tests", "golden", name + ".npz"))
    return load


def assert_grad_close(actual, ref, name="", l2=5e-3, mx=2e-2):
    """Gradient comparison that is robust to the few discontinuities of this network (leaky mask at <p,d>~0, nearest
    neighbour ties, 1/(2 sqrt(d)) at tiny Chamfer distances): relative L2 error of the whole tensor and max-abs error
    relative to the largest reference entry."""
    import numpy as np
    a = np.asarray(actual, np.float64).ravel()
    r = np.asarray(ref, np.float64).ravel()
    assert a.shape == r.shape, f"{name}: shape {a.shape} vs {r.shape}"
    nr = np.linalg.norm(r) + 1e-300
    e2 = np.linalg.norm(a - r) / nr
    em = np.abs(a - r).max() / (np.abs(r).max() + 1e-300)
    assert e2 <= l2 and em <= mx, f"{name}: rel-L2 error {e2:.3e} (limit {l2}), max error / max|ref| {em:.3e} (limit {mx})"
