"""GPU parity at the FULL BASELINE shapes (2048-point partial input -> 1024 coarse / 16384 dense points against a 16384-point ground truth),
where the numpy oracle takes minutes: the CUDA path (parity mode, fp32) against tests/eager_port.py, the plain-PyTorch restatement of the
reference's operator chain that tests/test_oracle_golden.py pins to the reference's own outputs and gradients.  Same weights
(torch.manual_seed(0) default init = the reference's), same synthetic SO(3)-rotated batch, VNMaxPool selections of the CUDA path forced
into the port (near-ties flip between any two fp32 evaluations) after checking that the port's own selections agree except at
near-ties.  Tolerances: values 1e-4 relative (BASELINE north star), gradients rel-L2 5e-3 (conftest.assert_grad_close)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import eager_port as EP
from conftest import assert_grad_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B", [6])
def test_full_size_train_step_vs_eager_port(B):
    import vn_pointcloudcompletion_b200 as V
    from vn_pointcloudcompletion_b200.synthetic import make_batch
    V.set_gemm_mode("fp32")
    cfg = SimpleNamespace(num_coarse=1024, latent_dim=2048, only_coarse=False, device="cuda", enc_pretrained="none")
    torch.manual_seed(0)
    net = V.PCNNet(cfg).train()
    P = EP.params_from_module(net, requires_grad=True)
    p, c, R = (torch.from_numpy(a).cuda() for a in make_batch(B, 2048, 16384, seed=4321))

    coarse, fine = net(p, V.Rotate(R))
    idx = (net.encoder.maxpool1.last_idx.reshape(B, -1), net.encoder.maxpool2.last_idx.reshape(B, -1))
    # the port's own selections: equal except where its top-2 score gap is within fp32 noise
    with torch.no_grad():
        _, _, own = EP.encoder(P, p, True)
    assert (own[0] != idx[0]).float().mean().item() < 0.02, "maxpool1 selections differ on more than 2 % of the channels"

    rc, rf, _ = EP.pcn_forward(P, p, R, True, idx)
    assert coarse.shape == (B, 1024, 3) and fine.shape == (B, 16384, 3)
    np.testing.assert_allclose(coarse.detach().cpu().numpy(), rc.detach().cpu().numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(fine.detach().cpu().numpy(), rf.detach().cpu().numpy(), rtol=1e-4, atol=1e-5)

    chamfer = V.chamfer_3DFunction.apply          # bit-identical to the reference kernel (tests/test_gpu_chamfer.py)
    loss = V.cd_loss_L1(coarse, c) + V.cd_loss_L1(fine, c)
    rloss = EP.cd_loss_l1(chamfer, rc, c) + EP.cd_loss_l1(chamfer, rf, c)
    np.testing.assert_allclose(loss.item(), rloss.item(), rtol=1e-4)
    loss.backward()
    rloss.backward()
    checked = 0
    for name, prm in net.named_parameters():
        ref = P[name].grad
        if ref is None:
            assert prm.grad is None or float(prm.grad.abs().max()) == 0.0, name
            continue
        assert_grad_close(prm.grad.cpu().numpy(), ref.cpu().numpy(), name)
        checked += 1
    assert checked >= 20
