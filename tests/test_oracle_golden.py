"""Pins the CPU oracle (oracle/) against the golden fixtures produced by the reference itself
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import vn_oracle as O
from conftest import assert_grad_close

CH_CASES = ["unit", "ragged", "tiny", "one2one", "big"]


@pytest.mark.parametrize("case", CH_CASES)
def test_chamfer_oracle_vs_reference_distchamfer(golden, case):
    g = golden("chamfer_unit")
    p1, p2 = g[f"ch_{case}_p1"], g[f"ch_{case}_p2"]
    d1, d2, i1, i2 = O.chamfer_forward(p1, p2)
    # the reference's own acceptance test: ChamferDistancePytorch/unit_test.py:23-33
    assert np.mean((d1 - g[f"ch_{case}_d1"]) ** 2) + np.mean((d2 - g[f"ch_{case}_d2"]) ** 2) < 1e-8
    assert np.array_equal(i1, g[f"ch_{case}_i1"]) and np.array_equal(i2, g[f"ch_{case}_i2"])
    # tighter: difference-form fp32 vs float64 expansion form
    np.testing.assert_allclose(d1, g[f"ch_{case}_d1"], rtol=2e-5, atol=1e-9)
    np.testing.assert_allclose(d2, g[f"ch_{case}_d2"], rtol=2e-5, atol=1e-9)
    g1, g2 = O.chamfer_backward(p1, p2, g[f"ch_{case}_w1"], g[f"ch_{case}_w2"], i1, i2)
    np.testing.assert_allclose(g1, g[f"ch_{case}_g1"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(g2, g[f"ch_{case}_g2"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("case", CH_CASES)
def test_cd_entry_points(golden, case):
    g = golden("chamfer_unit")
    p1, p2 = g[f"ch_{case}_p1"], g[f"ch_{case}_p2"]
    l, cache = O.cd_loss_L1(p1, p2)
    np.testing.assert_allclose(l, g[f"ch_{case}_l1"], rtol=1e-5)
    g1, g2 = O.cd_loss_L1_bwd(cache)
    np.testing.assert_allclose(g1, g[f"ch_{case}_l1_g1"], rtol=2e-4, atol=1e-7)
    np.testing.assert_allclose(g2, g[f"ch_{case}_l1_g2"], rtol=2e-4, atol=1e-7)
    np.testing.assert_allclose(O.cd_loss_L2(p1, p2)[0], g[f"ch_{case}_l2"], rtol=1e-5)
    np.testing.assert_allclose(O.l2_cd(p1, p2), g[f"ch_{case}_l2cd"], rtol=1e-5)
    np.testing.assert_allclose(O.l1_cd(p1, p2), g[f"ch_{case}_l1cd"], rtol=1e-5)


def test_chamfer_tie_breaks_to_lowest_index():
    p1 = np.zeros((1, 3, 3), np.float32)
    p2 = np.tile(np.array([[1, 0, 0]], np.float32), (1, 1030, 1))   # all candidates equidistant, spans 3 tiles
    p2[0, 700] = [0.5, 0, 0]
    p2[0, 900] = [0.5, 0, 0]
    d1, d2, i1, i2 = O.chamfer_forward(p1, p2)
    assert (i1 == 700).all() and np.allclose(d1, 0.25)
    assert (i2 == 0).all()


def _bn(g, key, when="pre", prefix="batchnorm.bn"):
    bn = O.BNState(g[f"{key}.{when}.sd.{prefix}.weight"].shape[0])
    bn.weight = g[f"{key}.{when}.sd.{prefix}.weight"]
    bn.bias = g[f"{key}.{when}.sd.{prefix}.bias"]
    bn.running_mean = g[f"{key}.{when}.sd.{prefix}.running_mean"]
    bn.running_var = g[f"{key}.{when}.sd.{prefix}.running_var"]
    bn.num_batches_tracked = int(g[f"{key}.{when}.sd.{prefix}.num_batches_tracked"])
    return bn


def _check_bn_post(g, key, bn, prefix="batchnorm.bn"):
    np.testing.assert_allclose(bn.running_mean, g[f"{key}.post.sd.{prefix}.running_mean"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(bn.running_var, g[f"{key}.post.sd.{prefix}.running_var"], rtol=1e-5, atol=1e-6)
    assert bn.num_batches_tracked == int(g[f"{key}.post.sd.{prefix}.num_batches_tracked"])


TOL = dict(rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize("key", ["VNLinear", "VNLinear_dim3"])
def test_vn_linear(golden, key):
    g = golden("vn_layers")
    W = g[f"{key}.pre.sd.map_to_feat.weight"]
    np.testing.assert_allclose(O.vn_linear(g[f"{key}.x"], W), g[f"{key}.y0"], **TOL)
    gx, gW = O.vn_linear_bwd(g[f"{key}.x"], W, g[f"{key}.gy0"])
    np.testing.assert_allclose(gx, g[f"{key}.gx"], **TOL)
    np.testing.assert_allclose(gW, g[f"{key}.grad.map_to_feat.weight"], rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("key,ns", [("VNLeakyReLU", 0.2), ("VNLeakyReLU_shared", 0.2), ("VNLeakyReLU_ns", 0.0)])
def test_vn_leaky_relu(golden, key, ns):
    g = golden("vn_layers")
    Wd = g[f"{key}.pre.sd.map_to_dir.weight"]
    y, cache = O.vn_leaky_relu(g[f"{key}.x"], Wd, ns)
    np.testing.assert_allclose(y, g[f"{key}.y0"], **TOL)
    gx, gWd = O.vn_leaky_relu_bwd(cache, Wd, g[f"{key}.gy0"], ns)
    np.testing.assert_allclose(gx, g[f"{key}.gx"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(gWd, g[f"{key}.grad.map_to_dir.weight"], rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("key,train", [("VNLinearLeakyReLU", True), ("VNLinearLeakyReLU_eval", False),
                                       ("VNLinearLeakyReLU_k1", True), ("VNLinearLeakyReLU_dim5", True),
                                       ("VNLinearLeakyReLU_shared", True)])
def test_vn_linear_leaky_relu(golden, key, train):
    g = golden("vn_layers")
    Wf, Wd = g[f"{key}.pre.sd.map_to_feat.weight"], g[f"{key}.pre.sd.map_to_dir.weight"]
    bn = _bn(g, key)
    y, cache = O.vn_linear_leaky_relu(g[f"{key}.x"], Wf, Wd, bn, training=train)
    np.testing.assert_allclose(y, g[f"{key}.y0"], **TOL)
    _check_bn_post(g, key, bn)
    r = O.vn_linear_leaky_relu_bwd(cache, Wf, Wd, g[f"{key}.gy0"])
    np.testing.assert_allclose(r["gx"], g[f"{key}.gx"], rtol=1e-3, atol=2e-4)
    np.testing.assert_allclose(r["gWf"], g[f"{key}.grad.map_to_feat.weight"], rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(r["gWd"], g[f"{key}.grad.map_to_dir.weight"], rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(r["gweight"], g[f"{key}.grad.batchnorm.bn.weight"], rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(r["gbias"], g[f"{key}.grad.batchnorm.bn.bias"], rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("key,has_bn", [("VNLinearAndLeakyReLU_none", False), ("VNLinearAndLeakyReLU_norm", True)])
def test_vn_linear_and_leaky_relu(golden, key, has_bn):
    g = golden("vn_layers")
    W, Wd = g[f"{key}.pre.sd.linear.map_to_feat.weight"], g[f"{key}.pre.sd.leaky_relu.map_to_dir.weight"]
    bn = _bn(g, key) if has_bn else None
    y, cache = O.vn_linear_and_leaky_relu(g[f"{key}.x"], W, Wd, bn)
    np.testing.assert_allclose(y, g[f"{key}.y0"], **TOL)
    r = O.vn_linear_and_leaky_relu_bwd(cache, W, Wd, g[f"{key}.gy0"])
    np.testing.assert_allclose(r["gx"], g[f"{key}.gx"], rtol=1e-3, atol=2e-4)
    np.testing.assert_allclose(r["gW"], g[f"{key}.grad.linear.map_to_feat.weight"], rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(r["gWd"], g[f"{key}.grad.leaky_relu.map_to_dir.weight"], rtol=1e-3, atol=1e-3)
    if has_bn:
        _check_bn_post(g, key, bn)
        np.testing.assert_allclose(r["gweight"], g[f"{key}.grad.batchnorm.bn.weight"], rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("key,train", [("VNBatchNorm", True), ("VNBatchNorm_eval", False), ("VNBatchNorm_dim3", True)])
def test_vn_batchnorm(golden, key, train):
    g = golden("vn_layers")
    bn = _bn(g, key, prefix="bn")
    y, cache = O.vn_batchnorm(g[f"{key}.x"], bn, training=train)
    np.testing.assert_allclose(y, g[f"{key}.y0"], **TOL)
    _check_bn_post(g, key, bn, prefix="bn")
    gx, gw, gb = O.vn_batchnorm_bwd(cache, g[f"{key}.gy0"])
    np.testing.assert_allclose(gx, g[f"{key}.gx"], rtol=1e-3, atol=2e-4)
    np.testing.assert_allclose(gw, g[f"{key}.grad.bn.weight"], rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(gb, g[f"{key}.grad.bn.bias"], rtol=1e-3, atol=1e-3)


def test_vn_max_pool(golden):
    g = golden("vn_layers")
    Wd = g["VNMaxPool.pre.sd.map_to_dir.weight"]
    y, idx = O.vn_max_pool(g["VNMaxPool.x"], Wd)
    assert np.array_equal(idx, g["VNMaxPool.idx"])            # selections: exact
    np.testing.assert_array_equal(y, g["VNMaxPool.y0"])       # a gather: exact
    gx = O.vn_max_pool_bwd(g["VNMaxPool.x"].shape, idx, g["VNMaxPool.gy0"])
    np.testing.assert_array_equal(gx, g["VNMaxPool.gx"])
    assert g["VNMaxPool.grad.map_to_dir.weight"].size == 0    # no gradient reaches map_to_dir


@pytest.mark.parametrize("key,frame", [("VNStdFeature", False), ("VNStdFeature_frame", True)])
def test_vn_std_feature(golden, key, frame):
    g = golden("vn_layers")
    vn = []
    for n in ("vn1", "vn2"):
        vn.append((g[f"{key}.pre.sd.{n}.map_to_feat.weight"], g[f"{key}.pre.sd.{n}.map_to_dir.weight"],
                   _bn(g, key, prefix=f"{n}.batchnorm.bn")))
    x_std, z0 = O.vn_std_feature(g[f"{key}.x"], vn[0], vn[1], g[f"{key}.pre.sd.vn_lin.weight"], normalize_frame=frame)
    np.testing.assert_allclose(x_std, g[f"{key}.y0"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(z0, g[f"{key}.y1"], rtol=1e-3, atol=1e-4)


def test_mean_pool(golden):
    g = golden("vn_layers")
    np.testing.assert_allclose(O.mean_pool(g["mean_pool.x"]), g["mean_pool.y"], rtol=1e-5, atol=1e-6)


# pcn_small (B=2) has a decoder BatchNorm channel whose norm variance is 7.6e-7 at mean 0.66 (the two samples' global
# features give almost the same norm): fp32 rounding of the norms (5e-7) is amplified ~1e3x there, so any two fp32
# implementations differ by ~1e-4 in `fine`.  pcn_b6 (B=6, condition number <= 16) is compared at the north-star 1e-4.
PCN_FIXTURES = [("pcn_small", 5e-4), ("pcn_b6", 1e-4)]


@pytest.mark.parametrize("fixture,ftol", PCN_FIXTURES)
def test_pcn_oracle_vs_reference_golden(golden, fixture, ftol):
    """pins oracle.PCNNetOracle (encoder + decoder + CD-L1 train loss, forward and backward) against the reference's own
    modules run by tests/golden/make_golden.py (B=2, 256-pt partial, 2048-pt GT, torch.manual_seed(0) weights whose
    digests are stored in the fixture).  VNMaxPool selections: own arg-max must match except at near-ties; values are
    compared with the reference's selections forced."""
    from types import SimpleNamespace

    import torch

    import vn_pointcloudcompletion_b200 as V
    g = golden(fixture)
    cfg = SimpleNamespace(num_coarse=1024, latent_dim=2048, only_coarse=False, device="cpu", enc_pretrained="none")
    torch.manual_seed(0)
    net = V.PCNNet(cfg)             # parameter container only (weights == reference's, checked by digest in test_abi)
    P = {k: v.detach().numpy().copy() for k, v in net.state_dict().items()}
    orc = O.PCNNetOracle(P)
    orc.enc.forward(g["p"], training=True, update_running=False)
    bad = orc.enc.idx[0] != g["idx1"]
    assert (np.abs(g["gap1"]).reshape(bad.shape)[bad] < 1e-4).all() and bad.mean() < 0.1
    coarse, fine = orc.forward(g["p"], g["R"], training=True, forced_idx=(g["idx1"], g["idx2"]))
    np.testing.assert_allclose(coarse, g["coarse"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(fine, g["fine"], rtol=ftol, atol=ftol * 0.25)
    loss, l1, l2, G = orc.loss_and_grads(g["c"])
    np.testing.assert_allclose(l1, g["loss1"], rtol=1e-4)
    np.testing.assert_allclose(l2, g["loss2"], rtol=1e-4)
    for k in g.files:
        if k.startswith("grad."):
            assert_grad_close(G[k[5:]], g[k], k)
        elif k.startswith("grad_head."):
            assert_grad_close(G[k[10:]].ravel()[:256], g[k], k)
        elif k.startswith("grad_none."):
            assert k[10:] not in G
        elif k.startswith("buf_post.") and not k.endswith("num_batches_tracked"):
            np.testing.assert_allclose(P[k[9:]], g[k], rtol=1e-4, atol=1e-6, err_msg=k)


@pytest.mark.parametrize("fixture,ftol", PCN_FIXTURES)
def test_eager_port_matches_reference_golden(golden, fixture, ftol):
    """pins tests/eager_port.py (the plain-PyTorch restatement used as the full-size GPU reference and as the eager yardstick) against the
    reference's own modules: same weights, same inputs, the reference's VNMaxPool selections forced; outputs, both losses (computed with
    the C oracle's Chamfer search) and the autograd gradients of every trained parameter."""
    from types import SimpleNamespace

    import torch

    import vn_pointcloudcompletion_b200 as V
    import eager_port as EP          # tests/ is on sys.path (rootdir conftest)
    g = golden(fixture)
    cfg = SimpleNamespace(num_coarse=1024, latent_dim=2048, only_coarse=False, device="cpu", enc_pretrained="none")
    torch.manual_seed(0)
    P = EP.params_from_module(V.PCNNet(cfg), requires_grad=True)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))      # noqa: E731
    forced = (t(g["idx1"]).reshape(g["p"].shape[0], -1).long(), t(g["idx2"]).reshape(g["p"].shape[0], -1).long())
    coarse, fine, _ = EP.pcn_forward(P, t(g["p"]), t(g["R"]), True, forced)
    np.testing.assert_allclose(coarse.detach().numpy(), g["coarse"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(fine.detach().numpy(), g["fine"], rtol=ftol, atol=ftol * 0.25)

    class _Chamfer(torch.autograd.Function):      # the C oracle's search (reference arithmetic) under autograd
        @staticmethod
        def forward(ctx, a, b):
            d1, d2, i1, i2 = O.chamfer_forward(a.detach().numpy(), b.detach().numpy())
            ctx.save_for_backward(a, b)
            ctx.idx = (i1, i2)
            return torch.from_numpy(d1), torch.from_numpy(d2)

        @staticmethod
        def backward(ctx, g1, g2):
            a, b = ctx.saved_tensors
            ga, gb = O.chamfer_backward(a.detach().numpy(), b.detach().numpy(), g1.contiguous().numpy(), g2.contiguous().numpy(), *ctx.idx)
            return torch.from_numpy(ga), torch.from_numpy(gb)

    c = t(g["c"])
    l1 = EP.cd_loss_l1(_Chamfer.apply, coarse, c)
    l2 = EP.cd_loss_l1(_Chamfer.apply, fine, c)
    np.testing.assert_allclose(l1.item(), g["loss1"], rtol=1e-4)
    np.testing.assert_allclose(l2.item(), g["loss2"], rtol=1e-4)
    (l1 + l2).backward()
    for k in g.files:
        if k.startswith("grad."):
            assert_grad_close(P[k[5:]].grad.numpy(), g[k], k)
        elif k.startswith("grad_none."):
            assert P[k[10:]].grad is None
