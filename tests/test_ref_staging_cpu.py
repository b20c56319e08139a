"""The byte-compiled reference under oracle/_ref/refpy.zip (oracle/build_ref_py.py) -- what bench.py --impl reference times and what the -m gpu
full-size tests use as their oracle -- IS the reference: on CPU it reproduces the committed goldens that tests/golden/make_golden.py made
from /root/reference directly (same seeded weights, inputs, outputs, losses).  Skipped when oracle/_ref/refpy.zip was not built."""
import numpy as np
import pytest
import torch


def test_staged_reference_reproduces_pcn_small_golden(golden):
    from oracle import ref_model as RM
    if not RM.available():
        pytest.skip("oracle/_ref/refpy.zip not built (needs /root/reference in the build container)")
    try:
        net, ref = RM.build_pcnnet("cpu", seed=0)
    except RuntimeError as e:      # the other backend was loaded earlier in this process (a combined CPU + GPU pytest run)
        pytest.skip(str(e))
    g = golden("pcn_small")
    net.train()
    sd = net.state_dict()
    assert len(sd) == 38
    for k, v in sd.items():          # the seeded init is the one the goldens were generated with
        a = v.double().numpy().ravel()
        dg = np.array([a.sum(), np.abs(a).sum(), (a * a).sum(), a[:: max(1, a.size // 97)][:64].sum()])
        np.testing.assert_allclose(dg, g["sd_digest." + k], rtol=1e-6, atol=1e-9, err_msg=k)
    p, c, R = (torch.from_numpy(g[k]) for k in ("p", "c", "R"))
    coarse, fine = net(p, RM.Rotate(R))
    np.testing.assert_allclose(coarse.detach().numpy(), g["coarse"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(fine.detach().numpy(), g["fine"], rtol=1e-4, atol=2e-5)
    l1 = ref.loss.cd_loss_L1(coarse, c)           # metrics/loss.py:20-31, unmodified, over chamfer_python.distChamfer
    l2 = ref.loss.cd_loss_L1(fine, c)
    np.testing.assert_allclose(l1.item(), g["loss1"], rtol=1e-5)
    np.testing.assert_allclose(l2.item(), g["loss2"], rtol=1e-4)


def test_bench_reference_arm_uses_the_staged_reference():
    """bench.py's CPU arm reports kind 'reference' exactly when oracle/_ref/refpy.zip exists (the port is a fallback only)"""
    import inspect

    import bench
    from oracle import ref_model as RM
    src = inspect.getsource(bench.cpu_reference_step)
    assert "RM.available()" in src and '"kind": "reference"' in src and '"kind": "port"' in src
    assert callable(RM.build_pcnnet)
