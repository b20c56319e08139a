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


def test_bench_reference_arm_prints_the_contract_line():
    """`python bench.py --impl reference --steps 1 --warmup 0` on the host cores: one JSON line with this arm's metric / config / unit and the
    reference-arm keys the measurement contract names (impl, cpu_baseline.kind, e2e with zero copy bytes)"""
    import json
    import os
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(repo, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=900, cwd=repo)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "vn_pcn_train_samples_per_s" and line["unit"] == "samples/s"
    assert line["steps"] == 1 and line["higher_is_better"] is True and line["value"] > 0
    assert "workload" in line["config"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
