"""SO(3)-equivariance, the property the reference is built around (README.md:1-7; SURVEY.md 4 iii): f(x R) = f(x) R layer by layer, checked on
the two CPU restatements of the reference that every parity test leans on -- the numpy oracle (oracle/vn_oracle.py) and the plain-PyTorch
port (tests/eager_port.py) -- in float64, so that the property is pinned to 1e-10 and any indexing / layout slip shows up at once.  (End to
end the property only holds up to VNMaxPool near-ties; here the pools are checked through their selections.)"""
import numpy as np
import pytest
import torch

import eager_port as EP
from oracle import vn_oracle as O


def _rotation(seed):
    q = np.random.RandomState(seed).standard_normal(4)
    q /= np.linalg.norm(q)
    r, i, j, k = q
    return np.array([[1 - 2 * (j * j + k * k), 2 * (i * j - k * r), 2 * (i * k + j * r)],
                     [2 * (i * j + k * r), 1 - 2 * (i * i + k * k), 2 * (j * k - i * r)],
                     [2 * (i * k - j * r), 2 * (j * k + i * r), 1 - 2 * (i * i + j * j)]])


def _rot(x, R):
    """rotate the vector axis (axis 2) of [B, C, 3, ...] by the row-vector convention v -> v R"""
    return np.moveaxis(np.moveaxis(x, 2, -1) @ R, -1, 2)


@pytest.mark.parametrize("seed", [0, 1])
def test_oracle_layers_are_equivariant(seed):
    rng = np.random.RandomState(10 + seed)
    R = _rotation(seed)
    B, Cin, Cout, N = 3, 6, 10, 17
    x = rng.standard_normal((B, Cin, 3, N))
    Wf, Wd = rng.standard_normal((Cout, Cin)), rng.standard_normal((Cout, Cin))
    np.testing.assert_allclose(O.vn_linear(_rot(x, R), Wf), _rot(O.vn_linear(x, Wf), R), atol=1e-10)
    bn = O.BNState(Cout, np.float64)
    bn.weight[:] = rng.uniform(0.5, 1.5, Cout)
    bn.bias[:] = rng.standard_normal(Cout) * 0.2
    y0 = O.vn_linear_leaky_relu(x, Wf, Wd, bn, training=True, update_running=False)
    y1 = O.vn_linear_leaky_relu(_rot(x, R), Wf, Wd, bn, training=True, update_running=False)
    y0, y1 = (y[0] if isinstance(y, tuple) else y for y in (y0, y1))
    np.testing.assert_allclose(y1, _rot(y0, R), atol=1e-10)
    Wp = rng.standard_normal((Cin, Cin))
    p0 = O.vn_max_pool(x, Wp)
    p1 = O.vn_max_pool(_rot(x, R), Wp)
    (v0, i0), (v1, i1) = p0[:2], p1[:2]
    assert np.array_equal(np.asarray(i0), np.asarray(i1))                 # the scores <x, W x> are invariant
    np.testing.assert_allclose(v1, np.moveaxis(np.moveaxis(v0, 2, -1) @ R, -1, 2), atol=1e-10)


def test_eager_port_network_is_equivariant_with_forced_selections():
    """VN_PointNet + VN_FoldingNet(rot=R): rotating the input and the folding seed rotates coarse and fine (the pools' selections of the
    unrotated run are forced, which is what equivariance up to near-ties means)"""
    from types import SimpleNamespace

    import vn_pointcloudcompletion_b200 as V
    cfg = SimpleNamespace(num_coarse=1024, latent_dim=2048, only_coarse=False, device="cpu", enc_pretrained="none")
    torch.manual_seed(0)
    P = {k: (v.double() if v.is_floating_point() else v) for k, v in EP.params_from_module(V.PCNNet(cfg)).items()}
    g = torch.Generator().manual_seed(3)
    xyz = torch.rand(2, 96, 3, generator=g, dtype=torch.float64) - 0.5
    R = torch.from_numpy(_rotation(5))
    with torch.no_grad():
        c0, f0, idx = EP.pcn_forward(P, xyz, None, True)
        c1, f1, _ = EP.pcn_forward(P, xyz @ R, R.expand(2, 3, 3), True, idx)
    np.testing.assert_allclose(c1.numpy(), (c0 @ R).numpy(), atol=1e-9)
    np.testing.assert_allclose(f1.numpy(), (f0 @ R).numpy(), atol=1e-9)
