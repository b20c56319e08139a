"""GPU tests of the tcgen05/TMEM/TMA GEMMs (TF32 operands, fp32 accumulate) through the C-ABI.
With operands pre-truncated to TF32 (10-bit mantissa) every product is exact, so the kernel must match float64 to
fp32-accumulation accuracy -- this pins tile addressing, swizzle/descriptor layout and the epilogue mapping.  With raw
fp32 operands the stated TF32 tolerance applies: |err| <= 2^-9 * sqrt(K) * rms(x) * rms(w) * 4."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _tf32(a):
    return (a.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


@pytest.fixture()
def tf32_mode():
    import vn_pointcloudcompletion_b200 as V
    V.set_gemm_mode("tf32")
    yield
    V.set_gemm_mode("fp32")


def _with_knob(knob, value, fn):
    from vn_pointcloudcompletion_b200 import _lib
    _lib.raw("vnpcc_set_tuning", knob, value)
    try:
        return fn()
    finally:
        _lib.raw("vnpcc_set_tuning", knob, 0)


SHAPES = [(512, 128, 512), (768, 512, 2048), (1000, 96, 200), (96, 2048, 1024), (4096, 256, 256), (300, 1024, 130), (70000, 64, 128),
          # few rows: the contraction is split over CTAs and accumulated with red.add (K not a multiple of the split, one-tile outputs)
          (96, 1024, 1024), (96, 1024, 3072), (64, 256, 512), (128, 1000, 384), (100, 4096, 64)]


@pytest.mark.parametrize("R,K,Cout", SHAPES)
def test_rows_gemm_tf32_exact_on_truncated_operands(tf32_mode, R, K, Cout):
    from vn_pointcloudcompletion_b200 import _lib, ops
    rng = np.random.RandomState(R + K + Cout)
    x = _tf32(rng.standard_normal((R, K)).astype(np.float32))
    w = _tf32(rng.standard_normal((Cout, K)).astype(np.float32))
    y = torch.empty((R, Cout), device="cuda")
    rc = _lib.raw("vnpcc_gemm_rows_tf32", _dev(x).data_ptr(), K, _dev(w).data_ptr(), K, y.data_ptr(), Cout, R, K, Cout, None, 0, 0,
                  torch.cuda.current_stream().cuda_stream)
    xd, wd = _dev(x), _dev(w)
    rc = _lib.raw("vnpcc_gemm_rows_tf32", xd.data_ptr(), K, wd.data_ptr(), K, y.data_ptr(), Cout, R, K, Cout, None, 0, 0,
                  torch.cuda.current_stream().cuda_stream)
    assert rc == 0, f"tensor-core kernel refused the shape (rc={rc})"
    torch.cuda.synchronize()
    ref = x.astype(np.float64) @ w.astype(np.float64).T
    np.testing.assert_allclose(y.cpu().numpy(), ref, rtol=1e-5, atol=1e-4 * np.sqrt(K / 64))


@pytest.mark.parametrize("B,N", [(3, 100), (40, 5), (6, 16), (2, 700), (5, 86)])
def test_rows_gemm_tf32_bias_and_strides(tf32_mode, B, N):
    """per-sample bias rows in the epilogue: tiles inside one sample, tiles over two samples (rows_per_sample >= 256), chunks that straddle
    a sample boundary, and samples shorter than a 32-row chunk (row-by-row walk)"""
    from vn_pointcloudcompletion_b200 import ops
    rng = np.random.RandomState(1)
    K, Cout = 64, 256
    R = B * N * 3
    xfull = _tf32(rng.standard_normal((R, K + 32)).astype(np.float32))
    wfull = _tf32(rng.standard_normal((Cout, 2 * K)).astype(np.float32))
    bias = rng.standard_normal((B * 3, Cout)).astype(np.float32)
    xd, wd = _dev(xfull), _dev(wfull)
    y = ops.gemm_rows(xd[:, 32:], wd[:, K:], False, _dev(bias), 3 * N).cpu().numpy()     # strided views: ld != K
    ref = xfull[:, 32:].astype(np.float64) @ wfull[:, K:].astype(np.float64).T
    ref = ref + np.repeat(bias.reshape(B, 1, 3, Cout), N, axis=1).reshape(R, Cout)
    np.testing.assert_allclose(y, ref, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("R,K,Cout", [(4096, 256, 256), (768, 512, 1024)])
def test_rows_gemm_tf32_tolerance_on_raw_operands(tf32_mode, R, K, Cout):
    from vn_pointcloudcompletion_b200 import ops
    rng = np.random.RandomState(7)
    x = rng.standard_normal((R, K)).astype(np.float32)
    w = rng.standard_normal((Cout, K)).astype(np.float32)
    y = ops.gemm_rows(_dev(x), _dev(w)).cpu().numpy()
    ref = x.astype(np.float64) @ w.astype(np.float64).T
    assert np.abs(y - ref).max() <= 2.0 ** -9 * np.sqrt(K) * 4
    gx = ops.gemm_rows(_dev(ref.astype(np.float32)), _dev(w), True).cpu().numpy()       # dgrad form (weight transposed on the fly)
    ref2 = ref.astype(np.float32).astype(np.float64) @ w.astype(np.float64)
    assert np.abs(gx - ref2).max() <= 2.0 ** -9 * np.sqrt(Cout) * np.abs(ref).std() * 8


@pytest.mark.parametrize("R,K,Cout", [(4096, 256, 256), (6000, 128, 512), (3000, 1024, 2048), (1536, 96, 160), (100000, 256, 512)])
def test_wgrad_tf32_exact_on_truncated_operands(tf32_mode, R, K, Cout):
    from vn_pointcloudcompletion_b200 import _lib
    rng = np.random.RandomState(R + K)
    x = _tf32(rng.standard_normal((R, K)).astype(np.float32))
    gy = _tf32(rng.standard_normal((R, Cout)).astype(np.float32))
    xd, gd = _dev(x), _dev(gy)
    g = torch.empty((Cout, K), device="cuda")
    rc = _lib.raw("vnpcc_gemm_wgrad_tf32", gd.data_ptr(), Cout, xd.data_ptr(), K, g.data_ptr(), K, R, Cout, K, None, 0,
                  torch.cuda.current_stream().cuda_stream)
    assert rc == 0, f"tensor-core wgrad refused the shape (rc={rc})"
    torch.cuda.synchronize()
    ref = gy.astype(np.float64).T @ x.astype(np.float64)
    np.testing.assert_allclose(g.cpu().numpy(), ref, rtol=1e-4, atol=2e-4 * np.sqrt(R / 64))


def test_pcn_train_step_tf32_close_to_fp32(tf32_mode):
    """whole train step in TF32 mode vs fp32 mode on the same seeded input with teacher-forced selections: stated
    tolerance for the tensor-core path = 2e-2 relative on coarse/fine, loss within 1e-2 relative"""
    from types import SimpleNamespace
    import vn_pointcloudcompletion_b200 as V
    from vn_pointcloudcompletion_b200.synthetic import make_batch
    cfg = SimpleNamespace(num_coarse=1024, latent_dim=2048, only_coarse=False, device="cuda", enc_pretrained="none")
    p, c, R = (torch.from_numpy(a).cuda() for a in make_batch(4, 256, 2048, seed=11))
    outs = {}
    idx = None
    for mode in ("fp32", "tf32"):
        V.set_gemm_mode(mode)
        torch.manual_seed(0)
        net = V.PCNNet(cfg).train()
        if idx is not None:
            net.encoder.maxpool1.forced_idx, net.encoder.maxpool2.forced_idx = idx
        coarse, fine = net(p, V.Rotate(R))
        loss = V.cd_loss_L1(coarse, c) + V.cd_loss_L1(fine, c)
        loss.backward()
        if idx is None:
            idx = (net.encoder.maxpool1.last_idx.clone(), net.encoder.maxpool2.last_idx.clone())
        outs[mode] = (coarse.detach(), fine.detach(), loss.item(), net.decoder.final_conv[1].map_to_feat.weight.grad.clone())
    a, b = outs["fp32"], outs["tf32"]
    assert (a[0] - b[0]).abs().max() <= 2e-2 * a[0].abs().max()
    assert (a[1] - b[1]).abs().max() <= 2e-2 * a[1].abs().max()
    assert abs(a[2] - b[2]) <= 1e-2 * abs(a[2])
    assert (a[3] - b[3]).norm() <= 5e-2 * a[3].norm()


@pytest.mark.parametrize("train", [False, True])
def test_fused_epilogue_vn_gemm_matches_unfused(tf32_mode, train):
    """VNLinearLeakyReLU no-grad forward with BN + leaky fused into the tcgen05 epilogue vs the unfused kernels on the same
    TF32 GEMM: same MMA sequence -> the only differences are fp32 rounding of the elementwise tail"""
    import vn_pointcloudcompletion_b200 as V
    from vn_pointcloudcompletion_b200 import ops
    torch.manual_seed(3)
    B, N, K, C = 3, 100, 64, 256            # 300 points: exercises a partial last tile and sample boundaries inside tiles
    m = V.VNLinearLeakyReLU(K, C, dim=4).cuda()
    with torch.no_grad():
        m.batchnorm.bn.weight.copy_(torch.randn(C))
        m.batchnorm.bn.bias.copy_(torch.randn(C) * 0.3)
        m.batchnorm.bn.running_mean.copy_(torch.rand(C) * 5)
        m.batchnorm.bn.running_var.copy_(torch.rand(C) + 0.5)
    m.train(train)
    rows = torch.randn(B * N * 3, K, device="cuda")
    bias = torch.randn(B * 3, 2 * C, device="cuda")
    for b_rows, rps in ((None, 0), (bias, 3 * N)):
        rm0 = m.batchnorm.bn.running_mean.clone()
        with torch.enable_grad():
            ref = m.forward_rows(rows, b_rows, rps).detach()
        rm_ref = m.batchnorm.bn.running_mean.clone()
        m.batchnorm.bn.running_mean.copy_(rm0)
        with torch.no_grad():
            w = torch.cat([m.map_to_feat.weight, m.map_to_dir.weight], 0)
            fused = ops.linear_bn_leaky_fused_nograd(rows, w, b_rows, rps, m.batchnorm.bn, m.training, m.negative_slope)
        assert fused is not None, "fused kernel refused the shape"
        assert (fused - ref).abs().max() <= 2e-4 * ref.abs().max()
        assert torch.allclose(m.batchnorm.bn.running_mean, rm_ref, rtol=1e-5, atol=1e-6)
        m.batchnorm.bn.running_mean.copy_(rm0)


def test_pcn_eval_forward_fused_vs_unfused(tf32_mode):
    from types import SimpleNamespace
    import vn_pointcloudcompletion_b200 as V
    from vn_pointcloudcompletion_b200.synthetic import make_batch
    cfg = SimpleNamespace(num_coarse=1024, latent_dim=2048, only_coarse=False, device="cuda", enc_pretrained="none")
    torch.manual_seed(0)
    net = V.PCNNet(cfg).eval()
    p, c, R = (torch.from_numpy(a).cuda() for a in make_batch(3, 256, 2048, seed=5))
    with torch.enable_grad():
        c0, f0 = net(p, V.Rotate(R))
        idx = (net.encoder.maxpool1.last_idx.clone(), net.encoder.maxpool2.last_idx.clone())
    net.encoder.maxpool1.forced_idx, net.encoder.maxpool2.forced_idx = idx
    with torch.no_grad():
        c1, f1 = net(p, V.Rotate(R))
    assert (c0.detach() - c1).abs().max() <= 1e-3 * c0.abs().max()
    assert (f0.detach() - f1).abs().max() <= 1e-3 * f0.abs().max()


def test_fused_pool_gemm_selects_like_unfused(tf32_mode):
    """VNLinear -> VNMaxPool with the arg-max inside the tcgen05 epilogue: same TF32 MMA sequence as the unfused GEMMs, so
    the selections must be identical; pooled rows are recomputed in exact fp32 from the selected inputs"""
    from vn_pointcloudcompletion_b200 import ops
    torch.manual_seed(5)
    G, N, K, C = 3, 100, 64, 256             # N not a multiple of 32: tiles straddle groups
    x = torch.randn(G * N * 3, K, device="cuda")
    w = torch.randn(C, K, device="cuda") / 8
    wdir = torch.randn(C, C, device="cuda") / 16
    out, idx = ops.linear_maxpool_rows(x, w, wdir, G, N)
    wc = ops.gemm_rows(wdir, w, True)
    f = ops.gemm_rows(x, w)
    d = ops.gemm_rows(x, wc)
    ref_idx = ops.maxpool_select(f, d, G, N)
    assert torch.equal(idx, ref_idx)
    xs = x.view(G, N, 3, K)
    sel = torch.gather(xs, 1, idx.view(G, C, 1, 1).expand(G, C, 3, K).permute(0, 1, 2, 3).contiguous().view(G, C, 3, K)[:, :, :, :].permute(0, 1, 2, 3))
    want = torch.einsum("gcvk,ck->gvc", sel.double(), w.double()).reshape(G * 3, C)
    assert torch.allclose(out.double(), want, rtol=1e-4, atol=1e-4)
    # gradient path (sparse backward) still works on the fused forward
    x.requires_grad_(True)
    w.requires_grad_(True)
    o2, _ = ops.linear_maxpool_rows(x, w, wdir, G, N)
    o2.sum().backward()
    assert x.grad.abs().sum() > 0 and w.grad.abs().sum() > 0


@pytest.mark.parametrize("R,K,Cout,Cs,with_bias", [(3 * 1000, 64, 256, 128, False), (3 * 4321, 256, 512, 256, False), (3 * 2048 * 3, 512, 2048, 1024, True),
                                                   (3 * 80, 32, 64, 32, False), (3 * 1110, 128, 256, 256, True),
                                                   # K <= 256 and enough row blocks: the resident-weight variant (one channel tile per CTA)
                                                   (3 * 8001, 128, 256, 256, True), (3 * 5000, 200, 384, 128, False), (3 * 40002, 256, 512, 256, True),
                                                   (3 * 3000, 256, 1024, 256, False)])
def test_gemm_rows_stats_epilogue(tf32_mode, R, K, Cout, Cs, with_bias):
    """vnpcc_gemm_rows_tf32_stats: the output equals the plain tcgen05 GEMM's bit for bit (same MMA sequence per element; only the row tile
    differs: 240 = 80 whole points instead of 256), and the BatchNorm-on-norm statistics accumulated in its epilogue equal a separate
    fp64 pass over that output (vnpcc_vn_norm_stats) to 2e-6 relative (fp32 partial sums over 16 points, MUFU rsqrt)"""
    from vn_pointcloudcompletion_b200 import _lib, ops
    torch.manual_seed(R + K)
    x = torch.randn(R, K, device="cuda")
    w = torch.randn(Cout, K, device="cuda") / K ** 0.5
    nsamp = 3
    bias = torch.randn(nsamp * 3, Cout, device="cuda") if with_bias else None
    rps = R // nsamp if with_bias else 0
    if with_bias:
        assert rps % 3 == 0
    y0 = ops.gemm_rows(x, w, False, bias, rps)
    assert ops._LAST_KERNEL[0] == "gemm_rows_tf32"
    sums = torch.full((2 * Cs,), 7.0, device="cuda", dtype=torch.float64)      # garbage: the entry point zeroes it
    y1 = torch.empty_like(y0)
    rc = _lib.raw("vnpcc_gemm_rows_tf32_stats", x, K, w, K, y1, Cout, R, K, Cout, bias, Cout if with_bias else 0, rps, sums, Cs, _lib.stream())
    assert rc == 0, rc
    assert torch.equal(y0, y1)
    ref = torch.empty(2 * Cs, device="cuda", dtype=torch.float64)
    _lib.call("vnpcc_vn_norm_stats", y0, Cout, R // 3, Cs, ref, _lib.stream())
    n = (y0.view(R // 3, 3, Cout)[:, :, :Cs].double().pow(2).sum(1).sqrt().float() + 1e-6).double()
    assert torch.allclose(ref[:Cs], n.sum(0), rtol=1e-6) and torch.allclose(ref[Cs:], (n * n).sum(0), rtol=1e-6)
    # epilogue: MUFU rsqrt norms (2 ulp) and fp32 pairwise partial sums of 16 points, fp64 across passes
    assert torch.allclose(sums, ref, rtol=2e-6, atol=0), ((sums - ref).abs() / ref.abs()).max()
    # and through the operator: ops.gemm_rows(stats=...) routes to the same kernel
    s2 = torch.empty(2 * Cs, device="cuda", dtype=torch.float64)
    y2 = ops.gemm_rows(x, w, False, bias, rps, stats=(s2, Cs))
    assert torch.equal(y2, y0) and torch.allclose(s2, ref, rtol=2e-6, atol=0)


@pytest.mark.parametrize("R,K,Cout", [(4096, 256, 256), (96 * 8, 2048, 1024), (30000, 512, 2048), (3000, 128, 512)])
def test_3xtf32_is_fp32_accurate(R, K, Cout):
    """'fp32x3' mode: operands split into TF32 hi + lo (vnpcc_split_tf32), three products in one tcgen05 GEMM.  Against float64 the error
    must be ~1e-5 of the result scale (the tensor core's fp32 accumulator truncates once per K=8 instruction, so it is a little coarser than
    sequential fp32 FMAs), two orders below plain TF32, for forward, dgrad and wgrad"""
    import vn_pointcloudcompletion_b200 as V
    from vn_pointcloudcompletion_b200 import ops
    rng = np.random.RandomState(R + K)
    x = rng.standard_normal((R, K)).astype(np.float32)
    w = (rng.standard_normal((Cout, K)) / np.sqrt(K)).astype(np.float32)
    gy = rng.standard_normal((R, Cout)).astype(np.float32)
    xd, wd, gd = _dev(x), _dev(w), _dev(gy)
    ref_y = x.astype(np.float64) @ w.astype(np.float64).T
    ref_gx = gy.astype(np.float64) @ w.astype(np.float64)
    ref_gw = gy.astype(np.float64).T @ x.astype(np.float64)
    errs = {}
    try:
        for mode in ("fp32x3", "fp32", "tf32"):
            V.set_gemm_mode(mode)
            y = ops.gemm_rows(xd, wd).cpu().numpy()
            kern = ops._LAST_KERNEL[0]
            gx = ops.gemm_rows(gd, wd, True).cpu().numpy()
            gw = ops.gemm_wgrad(gd, xd).cpu().numpy()
            errs[mode] = (kern, np.abs(y - ref_y).max() / np.abs(ref_y).max(), np.abs(gx - ref_gx).max() / np.abs(ref_gx).max(),
                          np.abs(gw - ref_gw).max() / np.abs(ref_gw).max())
    finally:
        V.set_gemm_mode("fp32")
    print(errs)
    k3, e3y, e3x, e3w = errs["fp32x3"]
    assert k3 == "gemm_rows_tf32x3" and errs["fp32"][0] == "sgemm_fp32" and errs["tf32"][0] == "gemm_rows_tf32"
    assert e3y < 3e-5 and e3x < 3e-5 and e3w < 3e-5, errs
    assert errs["tf32"][1] > 20 * e3y          # plain TF32 is far coarser


@pytest.mark.parametrize("P,Cin,C,with_res", [(32 * 40 + 13, 256, 256, True), (5000, 128, 128, False), (7, 256, 128, True), (96 * 33, 256, 256, False),
                                              # enough points for the CTA-pair form of tail_dgrad (Cin = 256): ragged last pair tile, one CTA
                                              # of the last pair without rows
                                              (64 * 74 + 21, 256, 256, True), (64 * 150 + 40, 256, 128, False)])
def test_tail_fused_backward_matches_unfused(tf32_mode, P, Cin, C, with_res):
    """the decoder tail VNLinearLeakyReLU(Cin -> C) -> VNLinear(C, 1) (+ residual), models/pcn.py:340-345,387, as one autograd node: the fused
    TF32 backward (sums pre-pass + tail_dgrad_tf32_kernel, which forms the final gradient of (p | d) inside the dgrad GEMM) against the
    unfused kernel sequence (bwd1 -> bwd2 -> dgrad GEMM) on the same saved tensors; ragged point counts included"""
    import torch.nn as nn
    from vn_pointcloudcompletion_b200 import ops
    torch.manual_seed(P + C)
    R = 3 * P
    h = torch.randn(R, Cin, device="cuda")
    wcat = torch.randn(2 * C, Cin, device="cuda") / Cin ** 0.5
    w2 = torch.randn(1, C, device="cuda") / C ** 0.5
    res = torch.randn(R, device="cuda") if with_res else None
    gy = torch.randn(R, device="cuda")
    bw, bb = torch.rand(C) + 0.5, torch.randn(C) * 0.3
    out = {}
    for fused in (True, "no-wgrad", False):
        ops._TAIL_FUSED_BWD = bool(fused)
        ops._TAIL_FUSED_WGRAD = fused is True
        try:
            bn = nn.BatchNorm1d(C).cuda().train()
            with torch.no_grad():
                bn.weight.copy_(bw)
                bn.bias.copy_(bb)
            hh, ww, w22 = h.clone().requires_grad_(True), wcat.clone().requires_grad_(True), w2.clone().requires_grad_(True)
            rr = res.clone().requires_grad_(True) if with_res else None
            y = ops.linear_bn_leaky_dot(hh, ww, bn, True, 0.2, w22, rr)
            y.backward(gy)
            kern = ops._LAST_KERNEL[0]
            out[fused] = (y.detach(), hh.grad, ww.grad, w22.grad, bn.weight.grad, bn.bias.grad, rr.grad if with_res else None, kern)
        finally:
            ops._TAIL_FUSED_BWD = True
            ops._TAIL_FUSED_WGRAD = False
    # the CTA-pair and the one-SM form of tail_dgrad issue the same MMA sequence per element: identical input gradients
    ops._TAIL_FUSED_BWD, ops._TAIL_FUSED_WGRAD = True, False

    def one_sm():
        bn = nn.BatchNorm1d(C).cuda().train()
        with torch.no_grad():
            bn.weight.copy_(bw)
            bn.bias.copy_(bb)
        hh = h.clone().requires_grad_(True)
        ops.linear_bn_leaky_dot(hh, wcat.clone().requires_grad_(True), bn, True, 0.2, w2.clone().requires_grad_(True), None).backward(gy)
        return hh.grad
    gh_default, gh_one = one_sm(), _with_knob(7, 3, one_sm)
    assert torch.equal(gh_default, gh_one)
    b = out[False]
    names = ["y", "gh", "gw", "gw2", "ggamma", "gbeta", "gres"]
    for key in (True, "no-wgrad"):
      a = out[key]
      assert torch.equal(a[0], b[0])
      for i in range(1, 7):
        if a[i] is None:
            assert b[i] is None
            continue
        rel = float((a[i] - b[i]).norm() / (b[i].norm() + 1e-30))
        print(key, names[i], rel)
        # (R < 64: the unfused dgrad falls back to the exact SIMT kernel, so the difference is TF32 operand rounding itself)
        assert rel < (2e-4 if R >= 64 else 3e-3), (names[i], rel)


@pytest.mark.parametrize("R,K,Cout,Cs,nsamp", [(3 * 4000, 256, 512, 0, 0), (3 * 4000, 256, 512, 256, 0), (3 * 3003, 512, 1024, 512, 3), (3 * 9001, 320, 256, 0, 0),
                                                (3 * 2750, 1024, 768, 0, 0)])
def test_cta_pair_rows_gemm_equals_one_sm_kernel(tf32_mode, R, K, Cout, Cs, nsamp):
    """tcgen05 cta_group::2 (two SMs per 256-channel tile, each staging half of the row block) against the one-SM kernel: the same MMA
    sequence per output element -> identical bits, with and without the per-sample bias and the statistics epilogue; row counts that
    leave partial tiles and an odd number of tiles per pair"""
    from vn_pointcloudcompletion_b200 import ops
    torch.manual_seed(R + K)
    x = torch.randn(R, K, device="cuda")
    w = torch.randn(Cout, K, device="cuda") / K ** 0.5
    bias = torch.randn(nsamp * 3, Cout, device="cuda") if nsamp else None
    rps = R // nsamp if nsamp else 0

    def run():
        sums = torch.zeros(2 * Cs, device="cuda", dtype=torch.float64) if Cs else None
        y = ops.gemm_rows(x, w, False, bias, rps, stats=(sums, Cs)) if Cs else ops.gemm_rows(x, w, False, bias, rps)
        return y, sums

    y1, s1 = _with_knob(2, 1, run)      # knob 2 = 1: one SM per tile
    y2, s2 = run()                      # default: CTA pairs where eligible (Cout % 256 == 0, K >= 256, R >= 8192)
    assert torch.equal(y1, y2)
    if Cs:
        assert torch.allclose(s1, s2, rtol=1e-12, atol=0)      # fp64 atomics in a different order
    ref = x.double() @ w.double().t()
    if bias is not None:
        ref = ref + bias.double().view(nsamp, 1, 3, Cout).expand(nsamp, R // (3 * nsamp), 3, Cout).reshape(R, Cout)
    assert (y2.double() - ref).abs().max() <= 2.0 ** -9 * K ** 0.5 * 4


def test_cta_pair_fused_vn_kernels_equal_one_sm_kernels(tf32_mode):
    """the fused VN GEMM (BN + leaky epilogue, statistics-only, conv -> max-pool arg-max) on CTA pairs against one SM per tile"""
    import vn_pointcloudcompletion_b200 as V
    from vn_pointcloudcompletion_b200 import ops
    torch.manual_seed(11)
    B, N, K, C = 3, 1000, 256, 256
    m = V.VNLinearLeakyReLU(K, C, dim=4).cuda().train()
    rows = torch.randn(B * N * 3, K, device="cuda")
    bias = torch.randn(B * 3, 2 * C, device="cuda")
    w = torch.cat([m.map_to_feat.weight, m.map_to_dir.weight], 0).detach()
    rm0 = m.batchnorm.bn.running_mean.clone()

    def fused():
        m.batchnorm.bn.running_mean.copy_(rm0)
        with torch.no_grad():
            return ops.linear_bn_leaky_fused_nograd(rows, w, bias, 3 * N, m.batchnorm.bn, True, 0.2)

    a, b = fused(), _with_knob(2, 4, fused)      # knob 2 = 4: CTA pairs for the fused kernels (default: one SM per tile)
    assert a is not None and b is not None
    assert (a - b).abs().max() <= 1e-6 * a.abs().max()          # batch statistics: fp64 atomics in a different order
    wl = torch.randn(2 * C, K, device="cuda") / K ** 0.5
    wdir = torch.randn(2 * C, 2 * C, device="cuda") / (2 * C) ** 0.5

    def pool():
        with torch.no_grad():
            return ops.linear_maxpool_rows(rows, wl, wdir, B, N)

    (o1, i1), (o2, i2) = pool(), _with_knob(2, 4, pool)
    assert torch.equal(i1, i2) and torch.equal(o1, o2)
