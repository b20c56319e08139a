"""GPU parity of every drop-in VN layer against the golden fixtures produced by the reference's own classes
(tests/golden/vn_layers.npz): outputs, input gradients, parameter gradients, BatchNorm buffers.
fp32 GEMM mode: tolerance 1e-4 relative (north star); arg-max selections and gathers exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = dict(rtol=2e-4, atol=2e-5)
GTOL = dict(rtol=1e-3, atol=2e-4)
WTOL = dict(rtol=1e-3, atol=1e-3)


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _load(mod, g, key, when="pre"):
    sd = {}
    pref = f"{key}.{when}.sd."
    for k in g.files:
        if k.startswith(pref):
            sd[k[len(pref):]] = torch.from_numpy(g[k])
    mod.load_state_dict(sd, strict=True)
    return mod.cuda()


def _run(mod, g, key, train=True, tuple_out=False):
    mod.train(train)
    x = _dev(g[key + ".x"]).requires_grad_(True)
    y = mod(x)
    ys = y if tuple_out else (y,)
    sum((t * _dev(g[f"{key}.gy{i}"])).sum() for i, t in enumerate(ys)).backward()
    return x, ys


def _check_common(mod, g, key, x, ys, gtol=GTOL, wtol=WTOL, check_post=True):
    for i, t in enumerate(ys):
        np.testing.assert_allclose(t.detach().cpu().numpy(), g[f"{key}.y{i}"], err_msg=f"{key}.y{i}", **TOL)
    np.testing.assert_allclose(x.grad.cpu().numpy(), g[key + ".gx"], err_msg=key + ".gx", **gtol)
    for n_, p in mod.named_parameters():
        ref = g[f"{key}.grad.{n_}"]
        if ref.size == 0:
            assert p.grad is None, f"{key}: {n_} must not receive a gradient"
        else:
            np.testing.assert_allclose(p.grad.cpu().numpy(), ref, err_msg=f"{key}.grad.{n_}", **wtol)
    if check_post:
        for k, v in mod.state_dict().items():
            ref = g[f"{key}.post.sd.{k}"]
            np.testing.assert_allclose(v.cpu().numpy(), ref, rtol=1e-5, atol=1e-6, err_msg=f"{key}.post.{k}")


@pytest.fixture(autouse=True)
def _fp32_mode():
    import vn_pointcloudcompletion_b200 as V
    V.set_gemm_mode("fp32")
    yield


@pytest.mark.parametrize("key,args", [("VNLinear", (12, 20)), ("VNLinear_dim3", (12, 20))])
def test_vn_linear(golden, key, args):
    import vn_pointcloudcompletion_b200 as V
    g = golden("vn_layers")
    mod = _load(V.VNLinear(*args), g, key)
    x, ys = _run(mod, g, key)
    _check_common(mod, g, key, x, ys)
    # physical layout contract: channels-last output, logical view
    if ys[0].dim() == 4:
        B, C, _, N = ys[0].shape
        assert ys[0].stride() == (3 * N * C, 1, C, 3 * C)


@pytest.mark.parametrize("key,kw", [("VNLeakyReLU", {}), ("VNLeakyReLU_shared", dict(share_nonlinearity=True)),
                                    ("VNLeakyReLU_ns", dict(negative_slope=0.0))])
def test_vn_leaky_relu(golden, key, kw):
    import vn_pointcloudcompletion_b200 as V
    g = golden("vn_layers")
    mod = _load(V.VNLeakyReLU(16, **kw), g, key)
    x, ys = _run(mod, g, key)
    _check_common(mod, g, key, x, ys)


@pytest.mark.parametrize("key,ctor,train", [
    ("VNLinearLeakyReLU", lambda V: V.VNLinearLeakyReLU(12, 24, dim=4), True),
    ("VNLinearLeakyReLU_eval", lambda V: V.VNLinearLeakyReLU(12, 24, dim=4), False),
    ("VNLinearLeakyReLU_k1", lambda V: V.VNLinearLeakyReLU(1, 16, dim=4), True),
    ("VNLinearLeakyReLU_dim5", lambda V: V.VNLinearLeakyReLU(6, 10), True),
    ("VNLinearLeakyReLU_shared", lambda V: V.VNLinearLeakyReLU(12, 24, dim=4, share_nonlinearity=True), True),
    ("VNLinearAndLeakyReLU_none", lambda V: V.VNLinearAndLeakyReLU(12, 24, dim=4, use_batchnorm="none"), True),
    ("VNLinearAndLeakyReLU_norm", lambda V: V.VNLinearAndLeakyReLU(12, 24, dim=4), True),
    ("VNBatchNorm", lambda V: V.VNBatchNorm(16, dim=4), True),
    ("VNBatchNorm_eval", lambda V: V.VNBatchNorm(16, dim=4), False),
    ("VNBatchNorm_dim3", lambda V: V.VNBatchNorm(16, dim=3), True),
])
def test_vn_bn_leaky_family(golden, key, ctor, train):
    import vn_pointcloudcompletion_b200 as V
    g = golden("vn_layers")
    mod = _load(ctor(V), g, key)
    x, ys = _run(mod, g, key, train=train)
    _check_common(mod, g, key, x, ys)


def test_vn_max_pool(golden):
    import vn_pointcloudcompletion_b200 as V
    g = golden("vn_layers")
    mod = _load(V.VNMaxPool(16), g, "VNMaxPool")
    x, ys = _run(mod, g, "VNMaxPool")
    assert np.array_equal(mod.last_idx.cpu().numpy().reshape(g["VNMaxPool.idx"].shape), g["VNMaxPool.idx"])
    np.testing.assert_array_equal(ys[0].detach().cpu().numpy(), g["VNMaxPool.y0"])
    np.testing.assert_array_equal(x.grad.cpu().numpy(), g["VNMaxPool.gx"])
    assert mod.map_to_dir.weight.grad is None


def test_maxpool_kernel_bit_exact_selection_given_x_and_d():
    """selections are asserted at the pool-kernel boundary with identical (x, d); score = three rounded
    products summed left to right; first maximum wins (incl. exact ties)."""
    from vn_pointcloudcompletion_b200 import ops
    rng = np.random.RandomState(3)
    G, N, C = 5, 777, 70
    x = rng.standard_normal((G * N * 3, C)).astype(np.float32)
    d = rng.standard_normal((G * N * 3, C)).astype(np.float32)
    x[3 * 10:3 * 11] = x[3 * 500:3 * 501]     # exact tie inside group 0
    d[3 * 10:3 * 11] = d[3 * 500:3 * 501]
    idx = ops.maxpool_select(_dev(x), _dev(d), G, N).cpu().numpy()
    xv = x.reshape(G, N, 3, C)
    dv = d.reshape(G, N, 3, C)
    score = (xv[:, :, 0] * dv[:, :, 0] + xv[:, :, 1] * dv[:, :, 1]) + xv[:, :, 2] * dv[:, :, 2]
    assert np.array_equal(idx, score.argmax(axis=1))


@pytest.mark.parametrize("key,frame", [("VNStdFeature", False), ("VNStdFeature_frame", True)])
def test_vn_std_feature(golden, key, frame):
    import vn_pointcloudcompletion_b200 as V
    g = golden("vn_layers")
    mod = _load(V.VNStdFeature(16, dim=4, normalize_frame=frame), g, key)
    x, ys = _run(mod, g, key, tuple_out=True)
    for i, t in enumerate(ys):
        np.testing.assert_allclose(t.detach().cpu().numpy(), g[f"{key}.y{i}"], rtol=1e-3, atol=1e-4)
    ref = g[key + ".gx"]
    np.testing.assert_allclose(x.grad.cpu().numpy(), ref, rtol=2e-3, atol=1e-4 * np.abs(ref).max())


def test_mean_pool(golden):
    import vn_pointcloudcompletion_b200 as V
    g = golden("vn_layers")
    np.testing.assert_allclose(V.mean_pool(_dev(g["mean_pool.x"])).cpu().numpy(), g["mean_pool.y"], rtol=1e-5, atol=1e-6)


def test_arbitrary_input_strides_accepted():
    import vn_pointcloudcompletion_b200 as V
    torch.manual_seed(0)
    lin = V.VNLinear(8, 16).cuda()
    xc = torch.randn(2, 8, 3, 33, device="cuda")                       # contiguous [B,C,3,N]
    xl = xc.permute(0, 3, 2, 1).contiguous().permute(0, 3, 2, 1)       # channels-last physical, same logical values
    assert torch.allclose(lin(xc), lin(xl), rtol=0, atol=0)


@pytest.mark.parametrize("R,K,Cout", [(300, 1, 7), (1000, 2, 512), (257, 130, 129), (4096, 256, 256), (96, 2048, 1024)])
def test_gemm_fp32_vs_float64(R, K, Cout):
    from vn_pointcloudcompletion_b200 import ops
    rng = np.random.RandomState(R + K)
    x = rng.standard_normal((R, K)).astype(np.float32)
    w = rng.standard_normal((Cout, K)).astype(np.float32)
    gy = rng.standard_normal((R, Cout)).astype(np.float32)
    y = ops.gemm_rows(_dev(x), _dev(w)).cpu().numpy()
    np.testing.assert_allclose(y, x.astype(np.float64) @ w.astype(np.float64).T, rtol=1e-4, atol=1e-4)
    gx = ops.gemm_rows(_dev(gy), _dev(w), True).cpu().numpy()
    np.testing.assert_allclose(gx, gy.astype(np.float64) @ w.astype(np.float64), rtol=1e-4, atol=1e-4)
    gw = ops.gemm_wgrad(_dev(gy), _dev(x)).cpu().numpy()
    np.testing.assert_allclose(gw, gy.astype(np.float64).T @ x.astype(np.float64), rtol=1e-4, atol=2e-3)


def test_layer_runs_on_the_tensors_device_not_the_current_one():
    """the reference lets a model live on config.device without torch.cuda.set_device (models/model.py:14-20): every launch must go to
    the device (and that device's current stream) its tensors live on"""
    import vn_pointcloudcompletion_b200 as V
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    torch.cuda.set_device(0)
    torch.manual_seed(0)
    lin = V.VNLinearLeakyReLU(8, 16, dim=4)
    x = torch.randn(2, 8, 3, 33)
    y0 = lin.cuda(0)(x.cuda(0)).cpu()
    lin1 = lin.to("cuda:1")
    lin1.batchnorm.bn.reset_running_stats()
    assert torch.cuda.current_device() == 0
    x1 = x.to("cuda:1").requires_grad_(True)
    y1 = lin1(x1)
    y1.sum().backward()
    torch.cuda.synchronize(1)
    assert y1.device.index == 1 and x1.grad.device.index == 1 and torch.cuda.current_device() == 0
    assert torch.allclose(y1.cpu(), y0, rtol=1e-6, atol=1e-6)
    d1, d2, i1, i2 = V.chamfer_3DFunction.apply(torch.rand(2, 50, 3, device="cuda:1"), torch.rand(2, 70, 3, device="cuda:1"))
    assert d1.device.index == 1 and i2.device.index == 1


def test_mixed_device_operands_are_rejected():
    from vn_pointcloudcompletion_b200 import _lib, ops
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with pytest.raises(_lib.VnpccError):
        ops.gemm_rows(torch.randn(64, 8, device="cuda:0"), torch.randn(16, 8, device="cuda:1"))
