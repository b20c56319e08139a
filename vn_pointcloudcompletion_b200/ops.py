"""Host-side operators over the channels-last ROW layout, each a torch.autograd.Function whose forward and backward
enqueue kernels of libvnpcc.so on the current CUDA stream (include/vnpcc.h).  PyTorch is used for device memory,
streams and autograd bookkeeping only; no arithmetic of the hot path runs in ATen, and there is no CPU fallback.

Row layout: a logical VN tensor [B, C, 3, *spatial] is a matrix [R, C] with R = B*prod(spatial)*3 rows, row
r = (point)*3 + v, channels contiguous (the physical layout the reference's nn.Linear calls produce).

Reference semantics (paths under /root/reference):
  linear_rows            nn.Linear(bias=False) of VNLinear & friends   models/vn_layers.py:21,38,65,69,162,194
  bn_leaky               VNBatchNorm + leaky projection                models/vn_layers.py:116-127, 39-42, 70-73
  maxpool_rows           VNMaxPool                                     models/vn_layers.py:158-167
  rows_dot               VNLinear(C, 1) (+ residual)                   models/pcn.py:345,387
  chamfer_3DFunction     extensions/chamfer_distance/chamfer_distance.py:29-71
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call, ptr, stream

EPS = 1e-6  # models/vn_layers.py:10

_GEMM_MODE = "fp32"      # "fp32": SIMT fp32-exact ; "tf32": tcgen05 tensor cores (TF32 operands, fp32 accumulate)


def set_gemm_mode(mode):
    """'fp32'   parity mode: exact fp32 FMAs (SIMT GEMMs), IEEE sqrt / division -- the mode the 1e-6-level parity tests run in;
    'fp32x3' parity mode on the tensor cores: every GEMM the tcgen05 kernels take runs as 3xTF32 (operands split into TF32 hi + lo parts,
             three products accumulated in fp32: ~1e-5 of the result scale, inside the north star's 1e-4), 2.4x the SIMT mode's step rate;
    'tf32'   throughput mode: tcgen05 TF32 GEMMs + MUFU elementwise kernels (what bench.py times; SURVEY 8d tolerance applies)."""
    global _GEMM_MODE, _FP32_IMPL
    if mode not in ("fp32", "fp32x3", "tf32"):
        raise ValueError(mode)
    _GEMM_MODE = "tf32" if mode == "tf32" else "fp32"
    _FP32_IMPL = "tf32x3" if mode != "fp32" else "simt"
    _lib.load().vnpcc_set_fast_math(1 if mode == "tf32" else 0)


def get_gemm_mode():
    return "fp32x3" if (_GEMM_MODE == "fp32" and _FP32_IMPL == "tf32x3") else _GEMM_MODE


# how fp32-accurate contractions run: "tf32x3" = on the tensor cores (operands split into TF32 hi + lo parts, three products accumulated
# in fp32; csrc/misc.cu split_tf32_kernel + the ordinary tcgen05 GEMM kernels over a tripled contraction axis) wherever the tensor-core
# kernels take the shape, "simt" = fp32 FMAs (csrc/gemm_simt.cu).  Set by set_gemm_mode; `exact=True` GEMMs of the throughput mode (the
# edge-convolution point GEMM) use tf32x3.
_FP32_IMPL = "simt"


def set_fp32_impl(impl):
    global _FP32_IMPL
    if impl not in ("tf32x3", "simt"):
        raise ValueError(impl)
    _FP32_IMPL = impl


def get_fp32_impl():
    return _FP32_IMPL


def _split3(t, layout):
    """tripled TF32 operand of a [R, K] matrix (see vnpcc_split_tf32)"""
    R, K = t.shape
    out = torch.empty((R, 3 * K) if layout < 2 else (3 * R, K), device=t.device, dtype=torch.float32)
    call("vnpcc_split_tf32", ptr(t), _ld(t), R, K, ptr(out), _ld(out), layout, stream())
    return out


# optional kernel-class timer (bench.py): an object with .start(cls, work) -> token and .stop(token); CUDA events on
# the launching stream, so it adds no synchronisation
_TIMER = None


def set_timer(timer):
    global _TIMER
    _TIMER = timer


_LAST_KERNEL = [None]     # name of the kernel family the most recent GEMM launch helper actually used


class _Timed:
    """work: algorithmic FLOPs (GEMMs, attention) or directed pairs (Chamfer) of the launch; nbytes: its algorithmic HBM bytes
    (operands read once, result written once) -- together they place the launch on the roofline min(tensor, AI x HBM)"""
    __slots__ = ("tok", "cls")

    def __init__(self, cls, work, nbytes=0.0):
        self.cls = cls
        _LAST_KERNEL[0] = None
        self.tok = _TIMER.start(cls, (work, nbytes)) if _TIMER is not None else None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        if self.tok is not None:
            _TIMER.stop(self.tok, _LAST_KERNEL[0] or self.cls)
        return False


def _check(t, name="tensor"):
    if t is None:
        return
    if not t.is_cuda:
        raise _lib.VnpccError(f"{name} must be a CUDA tensor: the B200 kernels have no CPU fallback")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")


def _rows2d(t, name="rows"):
    """accept a 2-D tensor whose last stride is 1 (row stride = leading dimension); make it so otherwise"""
    _check(t, name)
    if t.dim() != 2:
        raise ValueError(f"{name} must be 2-D [rows, channels]")
    if t.stride(1) != 1 and t.shape[1] != 1:
        t = t.contiguous()
    if t.shape[1] == 1 and t.stride(1) != 1:
        t = t.contiguous()
    return t


def _ld(t):
    return t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))


# ---------------------------------------------------------------------------------------------------------------
# raw launch helpers (no autograd)
# ---------------------------------------------------------------------------------------------------------------
def gemm_rows(x, w, trans_w=False, bias=None, rows_per_sample=0, out=None, accumulate=False, exact=False, stats=None):
    """y[r,o] = sum_k x[r,k] * (w[o,k] | w[k,o] if trans_w) (+ bias[(r // rows_per_sample)*3 + r%3, o]); exact=True forces the
    fp32 SIMT kernel even in TF32 mode (operands whose differences matter, e.g. the edge-convolution point GEMM).
    stats = (sums [2*Cs] float64, Cs): also produce the BatchNorm-on-norm batch statistics of the first Cs output channels
    (sum ||y|| + 1e-6, and squared) -- in the tcgen05 kernel's epilogue when it takes the shape, else by a pass over y."""
    R, K = x.shape
    Cout = w.shape[1] if trans_w else w.shape[0]
    assert (w.shape[0] if trans_w else w.shape[1]) == K, (x.shape, w.shape, trans_w)
    if out is None:
        out = torch.empty((R, Cout), device=x.device, dtype=torch.float32)
    if R == 0 or Cout == 0:
        if stats is not None:
            stats[0].zero_()
        return out
    if stats is not None:
        sums, Cs = stats
        if _GEMM_MODE == "tf32" and not accumulate and not exact and not trans_w:
            with _Timed("gemm", 2.0 * R * K * Cout, 4.0 * (R * K + R * Cout + K * Cout)):
                rc = _lib.raw("vnpcc_gemm_rows_tf32_stats", ptr(x), _ld(x), ptr(w), _ld(w), ptr(out), _ld(out), R, K, Cout, ptr(bias),
                              _ld(bias) if bias is not None else 0, rows_per_sample, ptr(sums), Cs, stream())
                if rc == 0:
                    _LAST_KERNEL[0] = "gemm_rows_tf32"
            if rc == 0:
                return out
            if rc != 10003:
                raise _lib.VnpccError(f"vnpcc_gemm_rows_tf32_stats failed with code {rc}")
    with _Timed("gemm", 2.0 * R * K * Cout, 4.0 * (R * K + R * Cout + K * Cout)):
        _gemm_rows_launch(x, w, trans_w, bias, rows_per_sample, out, accumulate, R, K, Cout, exact)
    if stats is not None:
        call("vnpcc_vn_norm_stats", ptr(out), _ld(out), R // 3, stats[1], ptr(stats[0]), stream())
    return out


def _gemm_rows_launch(x, w, trans_w, bias, rows_per_sample, out, accumulate, R, K, Cout, exact=False):
    # <= 4 input (or, for the dgrad form, output) channels: HBM-bound streaming kernels, exact fp32 in both modes
    if not accumulate and not trans_w and K <= 4:
        rc = _lib.raw("vnpcc_smallk_fwd", ptr(x), _ld(x), ptr(w), _ld(w), ptr(bias), _ld(bias) if bias is not None else 0,
                      rows_per_sample, ptr(out), _ld(out), R, K, Cout, stream())
        if rc == 0:
            _LAST_KERNEL[0] = "smallk"
            return out
        if rc != 10003:
            raise _lib.VnpccError(f"vnpcc_smallk_fwd failed with code {rc}")
    if not accumulate and trans_w and Cout <= 4 and bias is None:
        rc = _lib.raw("vnpcc_smallk_dgrad", ptr(x), _ld(x), ptr(w), _ld(w), ptr(out), _ld(out), R, Cout, K, stream())
        if rc == 0:
            _LAST_KERNEL[0] = "smallk"
            return out
        if rc != 10003:
            raise _lib.VnpccError(f"vnpcc_smallk_dgrad failed with code {rc}")
    if _GEMM_MODE == "tf32" and not accumulate and not exact:
        wt = w
        if trans_w:   # the tensor-core kernel wants K-contiguous weights; weights are small, transpose them
            wt = torch.empty((Cout, K), device=w.device, dtype=torch.float32)
            call("vnpcc_transpose", ptr(w), _ld(w), ptr(wt), K, K, Cout, stream())
        rc = _lib.raw("vnpcc_gemm_rows_tf32", ptr(x), _ld(x), ptr(wt), _ld(wt), ptr(out), _ld(out), R, K, Cout, ptr(bias),
                      _ld(bias) if bias is not None else 0, rows_per_sample, stream())
        if rc == 0:
            _LAST_KERNEL[0] = "gemm_rows_tf32"
            return out
        if rc != 10003:   # VNPCC_ERR_UNSUPPORTED -> shape not taken by the tensor-core kernel
            raise _lib.VnpccError(f"vnpcc_gemm_rows_tf32 failed with code {rc}")
    if _GEMM_MODE == "fp32" and _FP32_IMPL == "tf32x3" and not accumulate and K >= 32 and K % 4 == 0 and Cout >= 64 and R >= 64:
        wt = w
        if trans_w:
            wt = torch.empty((Cout, K), device=w.device, dtype=torch.float32)
            call("vnpcc_transpose", ptr(w), _ld(w), ptr(wt), K, K, Cout, stream())
        if _ld(x) % 4 == 0 and x.data_ptr() % 16 == 0:
            x3, w3 = _split3(x, 0), _split3(wt, 1)
            rc = _lib.raw("vnpcc_gemm_rows_tf32", ptr(x3), 3 * K, ptr(w3), 3 * K, ptr(out), _ld(out), R, 3 * K, Cout, ptr(bias),
                          _ld(bias) if bias is not None else 0, rows_per_sample, stream())
            if rc == 0:
                _LAST_KERNEL[0] = "gemm_rows_tf32x3"
                return out
            if rc != 10003:
                raise _lib.VnpccError(f"vnpcc_gemm_rows_tf32 (3xTF32) failed with code {rc}")
    _LAST_KERNEL[0] = "sgemm_fp32"
    call("vnpcc_gemm_rows_fp32", ptr(x), _ld(x), ptr(w), _ld(w), 1 if trans_w else 0, ptr(out), _ld(out), R, K, Cout,
         ptr(bias), _ld(bias) if bias is not None else 0, rows_per_sample, 1 if accumulate else 0, stream())
    return out


_WS = {}


def _workspace(nbytes, device, key="ws"):
    """scratch buffer cached per (purpose, device, stream): two streams running the same op concurrently (validation on a side stream
    during training) must not share scratch, and a buffer is only ever reused in the order of the stream it was handed to"""
    k = (key, device, torch.cuda.current_stream(device).cuda_stream)
    buf = _WS.get(k)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), device=device, dtype=torch.uint8)
        _WS[k] = buf
    return buf


def gemm_wgrad(dy, x, out=None, accumulate=False):
    """g[o,k] (+)= sum_r dy[r,o] * x[r,k]"""
    R, Cout = dy.shape
    K = x.shape[1]
    if out is None:
        out = torch.empty((Cout, K), device=x.device, dtype=torch.float32)
        accumulate = False
    with _Timed("gemm", 2.0 * R * K * Cout, 4.0 * (R * K + R * Cout + K * Cout)):
        return _gemm_wgrad_launch(dy, x, out, accumulate, R, Cout, K)


def smallk_wgrad(dy, x, B, N, want_bias):
    """K <= 4: gW [Cout, K] and (optionally) the per-sample bias gradient [B*3, Cout] in ONE pass over dy"""
    R, Cout = dy.shape
    K = x.shape[1]
    gw = torch.empty((Cout, K), device=x.device, dtype=torch.float32)
    gb = torch.empty((B * 3, Cout), device=x.device, dtype=torch.float32) if want_bias else None
    call("vnpcc_smallk_wgrad", ptr(dy), _ld(dy), ptr(x), _ld(x), B, N, K, Cout, ptr(gw), K, ptr(gb), Cout, stream())
    return gw, gb


def _gemm_wgrad_launch(dy, x, out, accumulate, R, Cout, K):
    if not accumulate and K <= 4 and R % 3 == 0 and out.stride(0) == K:
        rc = _lib.raw("vnpcc_smallk_wgrad", ptr(dy), _ld(dy), ptr(x), _ld(x), 1, R // 3, K, Cout, ptr(out), K, None, 0, stream())
        if rc == 0:
            _LAST_KERNEL[0] = "smallk"
            return out
        if rc != 10003:
            raise _lib.VnpccError(f"vnpcc_smallk_wgrad failed with code {rc}")
    if _GEMM_MODE == "tf32" and not accumulate and R > 0:
        nb = _lib.raw("vnpcc_gemm_wgrad_tf32_workspace_bytes", R, Cout, K)
        ws = _workspace(nb, x.device, "wgrad")
        rc = _lib.raw("vnpcc_gemm_wgrad_tf32", ptr(dy), _ld(dy), ptr(x), _ld(x), ptr(out), _ld(out), R, Cout, K, ptr(ws),
                      ws.numel(), stream())
        if rc == 0:
            _LAST_KERNEL[0] = "gemm_wgrad_tf32"
            return out
        if rc != 10003:
            raise _lib.VnpccError(f"vnpcc_gemm_wgrad_tf32 failed with code {rc}")
    if (_GEMM_MODE == "fp32" and _FP32_IMPL == "tf32x3" and not accumulate and R >= 256 and Cout >= 32 and K >= 32 and Cout % 4 == 0
            and K % 4 == 0 and 3 * R < (1 << 31)):
        dy3, x3 = _split3(dy, 2), _split3(x, 3)
        rc = _lib.raw("vnpcc_gemm_wgrad_tf32", ptr(dy3), Cout, ptr(x3), K, ptr(out), _ld(out), 3 * R, Cout, K, None, 0, stream())
        if rc == 0:
            _LAST_KERNEL[0] = "gemm_wgrad_tf32x3"
            return out
        if rc != 10003:
            raise _lib.VnpccError(f"vnpcc_gemm_wgrad_tf32 (3xTF32) failed with code {rc}")
    _LAST_KERNEL[0] = "sgemm_fp32"
    call("vnpcc_gemm_wgrad_fp32", ptr(dy), _ld(dy), ptr(x), _ld(x), ptr(out), _ld(out), R, Cout, K, 1 if accumulate else 0,
         stream())
    return out


def rows_sample_sum(g, B, N):
    """out[(b,v), c] = sum_n g[(b,n,v), c]"""
    C = g.shape[1]
    out = torch.empty((B * 3, C), device=g.device, dtype=torch.float32)
    call("vnpcc_rows_sample_sum", ptr(g), _ld(g), B, N, C, ptr(out), C, stream())
    return out


# ---------------------------------------------------------------------------------------------------------------
# VNLinear on rows
# ---------------------------------------------------------------------------------------------------------------
class _LinearRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, rows_per_sample, exact=False, stats=None):
        x = _rows2d(x, "x")
        _check(w, "weight")
        if w.stride(1) != 1:
            w = w.contiguous()
        if bias is not None:
            bias = _rows2d(bias, "bias")
        y = gemm_rows(x, w, False, bias, rows_per_sample, exact=exact, stats=stats)
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        ctx.rps = rows_per_sample
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gy = _rows2d(gy, "grad")
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = gemm_rows(gy, w, True)
        if (ctx.has_bias and x.shape[1] <= 4 and ctx.needs_input_grad[1] and ctx.needs_input_grad[2]
                and gy.shape[1] % 4 == 0 and _ld(gy) % 4 == 0):
            R = gy.shape[0]
            gw, gb = smallk_wgrad(gy, x, R // ctx.rps, ctx.rps // 3, True)
            return gx, gw, gb, None, None, None
        if ctx.needs_input_grad[1]:
            gw = gemm_wgrad(gy, x)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            R = gy.shape[0]
            B = R // ctx.rps
            gb = rows_sample_sum(gy, B, ctx.rps // 3)
        return gx, gw, gb, None, None, None


def linear_rows(x, w, bias=None, rows_per_sample=0, exact=False, stats=None):
    """x [R,K], w [Cout,K] -> [R,Cout]; optional per-sample bias rows [B*3, Cout] (row (b,v)) added to every point of
    sample b (the broadcast half of torch.cat([global.expand(N), local]) folded out of the GEMM, models/pcn.py:172,385).
    stats: see gemm_rows (BatchNorm-on-norm batch statistics from the GEMM epilogue)."""
    return _LinearRows.apply(x, w, bias, rows_per_sample, exact, stats)


def bn_needs_batch_stats(bn, training):
    return bn is not None and (training or bn.running_mean is None)


class _LinearBNLeaky(torch.autograd.Function):
    """out = leaky(BN(x Wf^T + b_p), x Wd^T + b_d) for stacked weights wcat = (Wf ; Wd) as ONE autograd node: the forward is GEMM (+ batch
    statistics in its epilogue) -> BN + leaky pass; the backward is bwd1 -> bwd2 -> dgrad -> wgrad, and because the node owns all four it
    can take the per-sample bias gradient out of bwd2 (vnpcc_vn_bn_bwd2_sbias) instead of re-reading the stacked gradient."""

    @staticmethod
    def forward(ctx, x, wcat, bias, gamma, beta, rows_per_sample, ns, bn, training):
        x = _rows2d(x, "x")
        _check(wcat, "weight")
        if wcat.stride(1) != 1:
            wcat = wcat.contiguous()
        if bias is not None:
            bias = _rows2d(bias, "bias")
        R = x.shape[0]
        C = wcat.shape[0] // 2
        sums = torch.empty(2 * C, device=x.device, dtype=torch.float64) if bn_needs_batch_stats(bn, training) else None
        pd = gemm_rows(x, wcat, False, bias, rows_per_sample, stats=(sums, C) if sums is not None else None)
        stat, use_batch = _bn_prepare(pd[:, :C], C, bn, training, R // 3, sums=sums)
        out = torch.empty((R, C), device=x.device, dtype=torch.float32)
        if R > 0:
            call("vnpcc_vn_bn_leaky_fwd", ptr(pd), _ld(pd), ptr(pd[:, C:]), _ld(pd), ptr(out), C, R // 3, C, ptr(stat), ptr(gamma), ptr(beta),
                 float(ns), stream())
        ctx.save_for_backward(x, wcat, pd, gamma, beta, stat)
        ctx.cfg = (C, float(ns), bool(use_batch), int(rows_per_sample), bias is not None)
        return out

    @staticmethod
    def backward(ctx, g):
        x, wcat, pd, gamma, beta, stat = ctx.saved_tensors
        C, ns, use_batch, rps, has_bias = ctx.cfg
        g = _rows2d(g, "grad")
        R = pd.shape[0]
        P = R // 3
        dev = pd.device
        gpd = torch.empty((R, 2 * C), device=dev, dtype=torch.float32)
        sums = torch.empty(2 * C, device=dev, dtype=torch.float64)
        ggamma = torch.empty(C, device=dev, dtype=torch.float32)
        gbeta = torch.empty(C, device=dev, dtype=torch.float32)
        gb = None
        if R > 0:
            call("vnpcc_vn_bn_leaky_bwd1", ptr(g), _ld(g), ptr(pd), _ld(pd), ptr(pd[:, C:]), _ld(pd), ptr(gpd), 2 * C, ptr(gpd[:, C:]), 2 * C, P, C,
                 ptr(stat), ptr(gamma), ptr(beta), ns, ptr(sums), stream())
            want_gb = has_bias and ctx.needs_input_grad[2]
            if want_gb:
                B = R // rps
                gb = torch.empty((B * 3, 2 * C), device=dev, dtype=torch.float32)
                rc = _lib.raw("vnpcc_vn_bn_bwd2_sbias", ptr(gpd), 2 * C, ptr(pd), _ld(pd), P, C, ptr(stat), ptr(gamma), ptr(beta), ptr(sums),
                              float(P), 1 if use_batch else 0, ptr(ggamma), ptr(gbeta), ptr(gpd[:, C:]), 2 * C, ptr(gb), 2 * C, rps // 3,
                              stream())
                if rc == 10003:
                    gb = None
                elif rc != 0:
                    raise _lib.VnpccError(f"vnpcc_vn_bn_bwd2_sbias failed with code {rc}")
            if gb is None:
                call("vnpcc_vn_bn_bwd2", ptr(gpd), 2 * C, ptr(pd), _ld(pd), P, C, ptr(stat), ptr(gamma), ptr(beta), ptr(sums), float(P),
                     1 if use_batch else 0, ptr(ggamma), ptr(gbeta), stream())
                if want_gb:
                    gb = rows_sample_sum(gpd, R // rps, rps // 3)
        gx = gemm_rows(gpd, wcat, True) if ctx.needs_input_grad[0] else None
        gw = gemm_wgrad(gpd, x) if ctx.needs_input_grad[1] else None
        return gx, gw, gb, ggamma, gbeta, None, None, None, None


def linear_bn_leaky_rows(x, wcat, bias, rows_per_sample, bn, training, ns):
    """training-capable VNLinearLeakyReLU on rows with stacked weights wcat [2C, K] = (W_feat ; W_dir) (models/vn_layers.py:60-74):
    ONE GEMM writes (p | d) and, in its epilogue, accumulates the BatchNorm-on-norm batch statistics of p; one streaming pass applies
    BatchNorm + the leaky projection."""
    C = wcat.shape[0] // 2
    if bn is not None and bn.affine and x.shape[1] > 4 and x.shape[0] > 0:
        return _LinearBNLeaky.apply(x, wcat, bias, bn.weight, bn.bias, rows_per_sample, ns, bn, training)
    sums = None
    if bn_needs_batch_stats(bn, training):
        sums = torch.empty(2 * C, device=x.device, dtype=torch.float64)
    pd = linear_rows(x, wcat, bias, rows_per_sample, stats=(sums, C) if sums is not None else None)
    return bn_leaky(pd, None, bn, training, ns, stacked=True, sums=sums)


def _bn_prepare(p, C, bn, training, count, stats_fn=None, sums=None):
    """returns stat [2C] (mean | invstd) and updates the running buffers in training mode.  stats_fn(sums) may supply
    the per-channel sums (sum n | sum n^2, fp64) itself, or `sums` may already hold them (GEMM epilogue); by default they are
    reduced from the rows p."""
    dev = bn.weight.device if p is None else p.device
    stat = torch.empty(2 * C, device=dev, dtype=torch.float32)
    use_batch = training or bn.running_mean is None
    have = sums is not None
    if not have:
        sums = torch.empty(2 * C, device=dev, dtype=torch.float64)
    if use_batch and not have:
        if stats_fn is not None:
            stats_fn(sums)
        else:
            call("vnpcc_vn_norm_stats", ptr(p), _ld(p), count, C, ptr(sums), stream())
    momentum = bn.momentum
    upd = training and bn.track_running_stats and bn.running_mean is not None
    if upd:
        bn.num_batches_tracked += 1
        if momentum is None:
            momentum = 1.0 / float(bn.num_batches_tracked)
    call("vnpcc_bn_finalize", ptr(sums), float(count), C, 1 if use_batch else 0,
         ptr(bn.running_mean) if (upd or not use_batch) else None, ptr(bn.running_var) if (upd or not use_batch) else None,
         float(momentum if momentum is not None else 0.0), float(bn.eps), ptr(stat), stream())
    return stat, use_batch


class _BNLeaky(torch.autograd.Function):
    """out = leaky(BN(p), d).  `pd` is either the stacked [R, 2C] buffer (p | d) or p alone with d passed separately
    (d may be None: BatchNorm only; stat may be None: leaky only)."""

    @staticmethod
    def forward(ctx, p, d, gamma, beta, stat, use_batch, ns, stacked):
        p = _rows2d(p, "p")
        if stacked:
            C = p.shape[1] // 2
            pv, dv = p[:, :C], p[:, C:]
        else:
            C = p.shape[1]
            pv, dv = p, (_rows2d(d, "d") if d is not None else None)
        R = p.shape[0]
        P = R // 3
        out = torch.empty((R, C), device=p.device, dtype=torch.float32)
        if R > 0:
            call("vnpcc_vn_bn_leaky_fwd", ptr(pv), _ld(p), ptr(dv), _ld(dv) if dv is not None else 0, ptr(out), C, P, C,
                 ptr(stat), ptr(gamma), ptr(beta), float(ns), stream())
        ctx.save_for_backward(p, d if not stacked else None, gamma, beta, stat)
        ctx.cfg = (C, P, float(ns), bool(stacked), bool(use_batch))
        return out

    @staticmethod
    def backward(ctx, g):
        p, d, gamma, beta, stat = ctx.saved_tensors
        C, P, ns, stacked, use_batch = ctx.cfg
        g = _rows2d(g, "grad")
        R = p.shape[0]
        dev = p.device
        if stacked:
            pv, dv = p[:, :C], p[:, C:]
            gpd = torch.empty((R, 2 * C), device=dev, dtype=torch.float32)
            gp, gd = gpd[:, :C], gpd[:, C:]
            ldg_ = 2 * C
        else:
            pv, dv = p, d
            gp = torch.empty((R, C), device=dev, dtype=torch.float32)
            gd = torch.empty((R, C), device=dev, dtype=torch.float32) if d is not None else None
            ldg_ = C
        sums = torch.empty(2 * C, device=dev, dtype=torch.float64) if stat is not None else None
        ggamma = gbeta = None
        if R > 0:
            call("vnpcc_vn_bn_leaky_bwd1", ptr(g), _ld(g), ptr(pv), _ld(p), ptr(dv), _ld(dv) if dv is not None else 0, ptr(gp),
                 ldg_, ptr(gd), ldg_ if gd is not None else 0, P, C, ptr(stat), ptr(gamma), ptr(beta), ns, ptr(sums), stream())
            if stat is not None:
                ggamma = torch.empty(C, device=dev, dtype=torch.float32)
                gbeta = torch.empty(C, device=dev, dtype=torch.float32)
                call("vnpcc_vn_bn_bwd2", ptr(gp), ldg_, ptr(pv), _ld(p), P, C, ptr(stat), ptr(gamma), ptr(beta), ptr(sums),
                     float(P), 1 if use_batch else 0, ptr(ggamma), ptr(gbeta), stream())
        if stacked:
            return gpd, None, ggamma, gbeta, None, None, None, None
        return gp, gd, ggamma, gbeta, None, None, None, None


def bn_leaky(p, d, bn, training, ns, stacked=False, sums=None):
    """p (and d) rows; bn: an nn.BatchNorm1d/2d module or None; returns leaky(BN(p), d) rows [R, C].  sums: batch statistics of p
    already accumulated by the producing GEMM (gemm_rows(stats=...))."""
    p = _rows2d(p, "p")
    C = p.shape[1] // 2 if stacked else p.shape[1]
    stat, use_batch, gamma, beta = None, False, None, None
    if bn is not None:
        stat, use_batch = _bn_prepare(p[:, :C] if stacked else p, C, bn, training, p.shape[0] // 3, sums=sums)
        gamma = bn.weight if bn.weight is not None else torch.ones(C, device=p.device)
        beta = bn.bias if bn.bias is not None else torch.zeros(C, device=p.device)
    return _BNLeaky.apply(p, d, gamma, beta, stat, use_batch, ns, stacked)


class _BNLeakyDot(torch.autograd.Function):
    """y[r] = sum_c leaky(BN(p), d)[r,c] * w2[c] (+ res[r]) on the stacked buffer pd = (p | d): the last
    VNLinearLeakyReLU of the decoder fused with VNLinear(C,1) and the residual (models/pcn.py:340-345,387)."""

    @staticmethod
    def forward(ctx, pd, gamma, beta, stat, use_batch, ns, w2, res):
        pd = _rows2d(pd, "pd")
        R = pd.shape[0]
        C = pd.shape[1] // 2
        P = R // 3
        w2 = w2.reshape(-1).contiguous()
        if res is not None:
            res = res.reshape(-1).contiguous()
        y = torch.empty(R, device=pd.device, dtype=torch.float32)
        call("vnpcc_bn_leaky_dot_fwd", ptr(pd), _ld(pd), ptr(pd[:, C:]), _ld(pd), P, C, ptr(stat), ptr(gamma), ptr(beta), float(ns),
             ptr(w2), ptr(res), ptr(y), stream())
        ctx.save_for_backward(pd, gamma, beta, stat, w2)
        ctx.cfg = (C, P, float(ns), bool(use_batch), res is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        pd, gamma, beta, stat, w2 = ctx.saved_tensors
        C, P, ns, use_batch, has_res = ctx.cfg
        gy = gy.contiguous()
        R = pd.shape[0]
        dev = pd.device
        gpd = torch.empty((R, 2 * C), device=dev, dtype=torch.float32)
        sums = torch.empty(2 * C, device=dev, dtype=torch.float64) if stat is not None else None
        gw2d = torch.empty(C, device=dev, dtype=torch.float64)
        call("vnpcc_bn_leaky_dot_bwd1", ptr(gy), ptr(pd), _ld(pd), ptr(pd[:, C:]), _ld(pd), ptr(gpd), 2 * C, ptr(gpd[:, C:]), 2 * C, P, C,
             ptr(stat), ptr(gamma), ptr(beta), ns, ptr(sums), ptr(w2), ptr(gw2d), stream())
        ggamma = gbeta = None
        if stat is not None:
            ggamma = torch.empty(C, device=dev, dtype=torch.float32)
            gbeta = torch.empty(C, device=dev, dtype=torch.float32)
            call("vnpcc_vn_bn_bwd2", ptr(gpd), 2 * C, ptr(pd), _ld(pd), P, C, ptr(stat), ptr(gamma), ptr(beta), ptr(sums), float(P),
                 1 if use_batch else 0, ptr(ggamma), ptr(gbeta), stream())
        gw2 = torch.empty(C, device=dev, dtype=torch.float32)
        call("vnpcc_double_to_float", ptr(gw2d), ptr(gw2), C, stream())
        return gpd, ggamma, gbeta, None, None, None, gw2.view(1, C), (gy if has_res else None)


_TAIL_FUSED_BWD = True      # tests switch these off to compare the fused tail backward with the unfused kernel sequence
# weight gradient with the gradient formed on the fly as well (tail_wgrad_tf32_kernel: gpd never stored).  Correct (tests) but MEASURED
# SLOWER at the decoder shape -- 3.74 ms against 3.08 ms for "fused dgrad writes gpd once + ordinary weight-gradient GEMM": without the gpd
# store the dgrad kernel only drops from 1.76 to 1.55 ms (it re-streams the 512 KB of weights from L2 for every 96-row tile: bound by the
# per-SM L2 bandwidth, not by HBM) while the fused weight gradient takes 1.38 ms against 0.80 ms (profiles/r2_tail_bench.md).  Off.
_TAIL_FUSED_WGRAD = False


class _LinearBNLeakyDot(torch.autograd.Function):
    """y[r] = sum_c leaky(BN(h Wf^T), h Wd^T)[r, c] w2[c] (+ res[r]): VNLinearLeakyReLU(Cin -> C) fused with VNLinear(C, 1) and the residual
    (the decoder tail, models/pcn.py:340-345,387) as ONE autograd node: GEMM (+ batch statistics in its epilogue) -> fused BN + leaky + dot
    forward; the backward, in TF32 mode, is the sums pre-pass + the fused tcgen05 dgrad kernel (vnpcc_tail_bwd_tf32: the final gradient of
    (p | d) is formed inside the GEMM and written once) + the weight-gradient GEMM."""

    @staticmethod
    def forward(ctx, h, wcat, gamma, beta, w2, res, ns, bn, training):
        h = _rows2d(h, "h")
        if wcat.stride(1) != 1:
            wcat = wcat.contiguous()
        R = h.shape[0]
        C = wcat.shape[0] // 2
        sums = torch.empty(2 * C, device=h.device, dtype=torch.float64) if bn_needs_batch_stats(bn, training) else None
        pd = gemm_rows(h, wcat, stats=(sums, C) if sums is not None else None)
        stat, use_batch = _bn_prepare(pd[:, :C], C, bn, training, R // 3, sums=sums)
        w2 = w2.reshape(-1).contiguous()
        if res is not None:
            res = res.reshape(-1).contiguous()
        y = torch.empty(R, device=h.device, dtype=torch.float32)
        call("vnpcc_bn_leaky_dot_fwd", ptr(pd), _ld(pd), ptr(pd[:, C:]), _ld(pd), R // 3, C, ptr(stat), ptr(gamma), ptr(beta), float(ns),
             ptr(w2), ptr(res), ptr(y), stream())
        ctx.save_for_backward(h, wcat, pd, gamma, beta, stat, w2)
        ctx.cfg = (C, float(ns), bool(use_batch), res is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        h, wcat, pd, gamma, beta, stat, w2 = ctx.saved_tensors
        C, ns, use_batch, has_res = ctx.cfg
        gy = gy.contiguous()
        R = pd.shape[0]
        P = R // 3
        Cin = h.shape[1]
        dev = pd.device
        gpd = None
        sums = torch.empty(2 * C, device=dev, dtype=torch.float64)
        gw2d = torch.empty(C, device=dev, dtype=torch.float64)
        ggamma = torch.empty(C, device=dev, dtype=torch.float32)
        gbeta = torch.empty(C, device=dev, dtype=torch.float32)
        gh = gw = None
        if _GEMM_MODE == "tf32" and _TAIL_FUSED_BWD and ctx.needs_input_grad[0] and P > 0:
            wt = torch.empty((Cin, 2 * C), device=dev, dtype=torch.float32)
            call("vnpcc_transpose", ptr(wcat), _ld(wcat), ptr(wt), 2 * C, 2 * C, Cin, stream())
            gh = torch.empty((R, Cin), device=dev, dtype=torch.float32)
            fused_w = _TAIL_FUSED_WGRAD and ctx.needs_input_grad[1] and C % 128 == 0 and _ld(h) % 4 == 0 and h.data_ptr() % 16 == 0
            if fused_w:
                gw = torch.empty((2 * C, Cin), device=dev, dtype=torch.float32)
            else:
                gpd = torch.empty((R, 2 * C), device=dev, dtype=torch.float32)
            flops = 2.0 * R * 2 * C * Cin * (2 if fused_w else 1)
            nbytes = 4.0 * (2.0 * R * 2 * C + R * Cin + (0 if fused_w else R * 2 * C) + (2 * R * Cin if fused_w else 0))
            with _Timed("gemm", flops, nbytes):
                rc = _lib.raw("vnpcc_tail_bwd_tf32", ptr(gy), ptr(pd), _ld(pd), P, C, ptr(stat), ptr(gamma), ptr(beta), ns, ptr(w2), ptr(wt),
                              2 * C, Cin, 1 if use_batch else 0, ptr(sums), ptr(gw2d), ptr(gpd), 2 * C if gpd is not None else 0, ptr(gh), Cin,
                              ptr(h) if fused_w else None, _ld(h) if fused_w else 0, ptr(gw), Cin if fused_w else 0, stream())
                if rc == 0:
                    _LAST_KERNEL[0] = "tail_bwd_tf32"
            if rc == 10003:
                gh = gw = gpd = None
            elif rc != 0:
                raise _lib.VnpccError(f"vnpcc_tail_bwd_tf32 failed with code {rc}")
            else:
                call("vnpcc_double_to_float", ptr(sums), ptr(gbeta), C, stream())
                call("vnpcc_double_to_float", ptr(sums[C:]), ptr(ggamma), C, stream())
        if gh is None:
            gpd = torch.empty((R, 2 * C), device=dev, dtype=torch.float32)
            call("vnpcc_bn_leaky_dot_bwd1", ptr(gy), ptr(pd), _ld(pd), ptr(pd[:, C:]), _ld(pd), ptr(gpd), 2 * C, ptr(gpd[:, C:]), 2 * C, P, C,
                 ptr(stat), ptr(gamma), ptr(beta), ns, ptr(sums), ptr(w2), ptr(gw2d), stream())
            call("vnpcc_vn_bn_bwd2", ptr(gpd), 2 * C, ptr(pd), _ld(pd), P, C, ptr(stat), ptr(gamma), ptr(beta), ptr(sums), float(P),
                 1 if use_batch else 0, ptr(ggamma), ptr(gbeta), stream())
            if ctx.needs_input_grad[0]:
                gh = gemm_rows(gpd, wcat, True)
        if gw is None and ctx.needs_input_grad[1]:
            gw = gemm_wgrad(gpd, h)
        gw2 = torch.empty(C, device=dev, dtype=torch.float32)
        call("vnpcc_double_to_float", ptr(gw2d), ptr(gw2), C, stream())
        return gh, gw, ggamma, gbeta, gw2.view(1, C), (gy if has_res else None), None, None, None


def linear_bn_leaky_dot(h, wcat, bn, training, ns, w2, res=None):
    """the decoder tail on rows h [R, Cin] with stacked weights wcat [2C, Cin]; w2 is the [1, C] weight of VNLinear(C, 1)"""
    return _LinearBNLeakyDot.apply(h, wcat, bn.weight, bn.bias, w2, res, ns, bn, training)


def linear_bn_leaky_fused_nograd(x, wcat, bias, rows_per_sample, bn, training, ns):
    """No-grad forward of VNLinearLeakyReLU with BatchNorm-on-norm + leaky projection fused into the tcgen05 GEMM epilogue
    (csrc/gemm_tcgen05.cu, gemm_vn_fused_kernel): p and d never reach HBM.  Returns None when the shape is not taken (the
    caller then uses the unfused kernels).  Only used with gradients disabled (validation / test loops, train.py:199-226,
    test.py:54-72): the training backward needs p and d."""
    if torch.is_grad_enabled() or _GEMM_MODE != "tf32":
        return None
    x = _rows2d(x, "x")
    R, K = x.shape
    C = wcat.shape[0] // 2
    if wcat.stride(1) != 1:
        wcat = wcat.contiguous()
    if C % 128 != 0 or K < 32 or K % 4 != 0 or R % 3 != 0 or _ld(x) % 4 != 0 or _ld(wcat) % 4 != 0 or x.data_ptr() % 16 or wcat.data_ptr() % 16:
        return None
    ldb = _ld(bias) if bias is not None else 0
    stat, gamma, beta = None, None, None
    if bn is not None:
        def stats_fn(sums):
            call("vnpcc_gemm_vn_stats", ptr(x), _ld(x), ptr(wcat), _ld(wcat), R, K, C, ptr(bias), ldb, rows_per_sample, ptr(sums), stream())
        stat, _ = _bn_prepare(None, C, bn, training, R // 3, stats_fn)
        gamma, beta = bn.weight, bn.bias
    out = torch.empty((R, C), device=x.device, dtype=torch.float32)
    with _Timed("gemm_vn_fused", 2.0 * R * K * 2 * C, 4.0 * (R * K + R * C + 2 * C * K)):
        call("vnpcc_gemm_vn_apply", ptr(x), _ld(x), ptr(wcat), _ld(wcat), ptr(out), C, R, K, C, ptr(bias), ldb, rows_per_sample,
             ptr(stat), ptr(gamma), ptr(beta), float(ns), stream())
    return out


class _SmallKBNLeaky(torch.autograd.Function):
    """out = leaky(BN(Wf x + b_p), Wd x + b_d) for x with <= 4 channels and per-sample bias rows, p / d never stored
    (csrc/vn_fused.cu)."""

    @staticmethod
    def forward(ctx, x, w, bias, gamma, beta, stat, use_batch, ns, B, N, gx_first_col):
        C = w.shape[0] // 2
        K = x.shape[1]
        out = torch.empty((x.shape[0], C), device=x.device, dtype=torch.float32)
        call("vnpcc_fold_fwd", ptr(x), _ld(x), ptr(w), _ld(w), ptr(bias), _ld(bias) if bias is not None else 0, B, N, K, C, ptr(stat),
             ptr(gamma), ptr(beta), float(ns), ptr(out), C, stream())
        ctx.save_for_backward(x, w, bias, gamma, beta, stat)
        ctx.cfg = (B, N, K, C, float(ns), bool(use_batch), int(gx_first_col))
        return out

    @staticmethod
    def backward(ctx, g):
        x, w, bias, gamma, beta, stat = ctx.saved_tensors
        B, N, K, C, ns, use_batch, gx_first_col = ctx.cfg
        g = _rows2d(g, "grad")
        dev = x.device
        gx = torch.empty((x.shape[0], K), device=dev, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        gw = torch.empty((2 * C, K), device=dev, dtype=torch.float32)
        gb = torch.empty((B * 3, 2 * C), device=dev, dtype=torch.float32) if bias is not None else None
        sums = torch.empty(2 * C, device=dev, dtype=torch.float64)
        ggamma = torch.empty(C, device=dev, dtype=torch.float32) if stat is not None else None
        gbeta = torch.empty(C, device=dev, dtype=torch.float32) if stat is not None else None
        call("vnpcc_fold_bwd", ptr(g), _ld(g), ptr(x), _ld(x), ptr(w), _ld(w), ptr(bias), _ld(bias) if bias is not None else 0, B, N, K, C,
             ptr(stat), ptr(gamma), ptr(beta), ns, 1 if use_batch else 0, ptr(sums), ptr(gx), K, gx_first_col, ptr(gw), K, ptr(gb),
             2 * C, ptr(ggamma), ptr(gbeta), stream())
        return gx, gw, gb, ggamma, gbeta, None, None, None, None, None, None


def smallk_bn_leaky_supported(K, C, bias):
    return 1 <= K <= 4 and C % 128 == 0 and C <= 1024 and (bias is None or (bias.stride(0) % 4 == 0 and bias.stride(1) == 1))


def smallk_bn_leaky(x, w, bias, bn, training, ns, B, N, gx_first_col=0):
    """VNLinearLeakyReLU on rows x [B*N*3, K<=4] with stacked weights w [2C, K] and per-sample bias rows [B*3, 2C].
    Input columns < gx_first_col are treated as constants (their gradient is returned as zero and not computed)."""
    x = _rows2d(x, "x")
    _check(w, "weight")
    if w.stride(1) != 1:
        w = w.contiguous()
    C = w.shape[0] // 2
    K = x.shape[1]
    stat, use_batch, gamma, beta = None, False, None, None
    if bn is not None:
        def stats_fn(sums):
            call("vnpcc_fold_stats", ptr(x), _ld(x), ptr(w), _ld(w), ptr(bias), _ld(bias) if bias is not None else 0, B, N, K, C,
                 ptr(sums), stream())
        stat, use_batch = _bn_prepare(None, C, bn, training, B * N, stats_fn)
        gamma, beta = bn.weight, bn.bias
    return _SmallKBNLeaky.apply(x, w, bias, gamma, beta, stat, use_batch, ns, B, N, gx_first_col)


def bn_leaky_dot_supported(C):
    return C % 128 == 0 and C <= 1024


def bn_leaky_dot(pd, bn, training, ns, w2, res=None, sums=None):
    """fused  leaky(BN(p), d) . w2 (+ res)  on the stacked (p | d) rows; w2 is the [1, C] weight of VNLinear(C, 1)"""
    pd = _rows2d(pd, "pd")
    C = pd.shape[1] // 2
    stat, use_batch, gamma, beta = None, False, None, None
    if bn is not None:
        stat, use_batch = _bn_prepare(pd[:, :C], C, bn, training, pd.shape[0] // 3, sums=sums)
        gamma = bn.weight if bn.weight is not None else torch.ones(C, device=pd.device)
        beta = bn.bias if bn.bias is not None else torch.zeros(C, device=pd.device)
    return _BNLeakyDot.apply(pd, gamma, beta, stat, use_batch, ns, w2, res)


# ---------------------------------------------------------------------------------------------------------------
# VNMaxPool on rows: groups of N consecutive points
# ---------------------------------------------------------------------------------------------------------------
def maxpool_select(x, d, G, N):
    """idx [G, C] int64 = argmax_n <x, d> within each group of N consecutive points (first maximum wins)"""
    C = x.shape[1]
    ws = torch.empty(G * C, device=x.device, dtype=torch.int64)
    idx = torch.empty((G, C), device=x.device, dtype=torch.int64)
    call("vnpcc_vn_maxpool_argmax", ptr(x), _ld(x), ptr(d), _ld(d), G, N, C, ptr(ws), ptr(idx), stream())
    return idx


class _MaxPoolGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, idx, G, N):
        x = _rows2d(x, "x")
        C = x.shape[1]
        out = torch.empty((G * 3, C), device=x.device, dtype=torch.float32)
        call("vnpcc_vn_maxpool_gather", ptr(x), _ld(x), ptr(idx), G, N, C, ptr(out), C, stream())
        ctx.save_for_backward(idx)
        ctx.cfg = (G, N, C)
        return out

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        G, N, C = ctx.cfg
        g = _rows2d(g, "grad")
        gx = torch.zeros((G * N * 3, C), device=g.device, dtype=torch.float32)
        call("vnpcc_vn_maxpool_scatter_add", ptr(g), _ld(g), ptr(idx), G, N, C, ptr(gx), C, stream())
        return gx, None, None, None


class _MaxPoolGatherTap(torch.autograd.Function):
    """_MaxPoolGather that also hands its input through (an alias): for an activation that feeds the pool AND a dense consumer
    (models/pcn.py:168-173: feature -> maxpool1 and -> cat -> second_conv).  The pooled gradient touches one point per (group, channel);
    the backward scatters it IN PLACE into the dense consumer's gradient instead of materialising a second dense tensor of zeros and
    summing the two (a 400 MB fill + a 1.2 GB add per step at B = 32).
    CONTRACT: the alias must be consumed by exactly one Function whose backward returns a freshly allocated gradient (here: the linear
    layers' dgrad output, pcn.py).  A consumer that forwards its incoming gradient unchanged (an add, a view, a hook that keeps a
    reference) would see the pooled gradient scattered into ITS tensor too -- route such a consumer through maxpool_rows(tap=False)."""

    @staticmethod
    def forward(ctx, x, idx, G, N):
        x = _rows2d(x, "x")
        C = x.shape[1]
        out = torch.empty((G * 3, C), device=x.device, dtype=torch.float32)
        call("vnpcc_vn_maxpool_gather", ptr(x), _ld(x), ptr(idx), G, N, C, ptr(out), C, stream())
        ctx.save_for_backward(idx)
        ctx.cfg = (G, N, C)
        return out, x.view_as(x)

    @staticmethod
    def backward(ctx, g, gpass):
        (idx,) = ctx.saved_tensors
        G, N, C = ctx.cfg
        if gpass is None:
            gx = torch.zeros((G * N * 3, C), device=idx.device, dtype=torch.float32)
        else:
            gx = _rows2d(gpass, "grad")      # the dense consumer's gradient: a fresh tensor owned by this backward pass
            if gx.stride(1) != 1 or gx.stride(0) < C:      # a broadcast / strided gradient is not ours to write into
                gx = gx.contiguous()
        if g is not None:
            g = _rows2d(g, "grad")
            call("vnpcc_vn_maxpool_scatter_add", ptr(g), _ld(g), ptr(idx), G, N, C, ptr(gx), _ld(gx), stream())
        return gx, None, None, None


def maxpool_rows(x, d, G, N, forced_idx=None, tap=False):
    """x, d rows [G*N*3, C] -> (pooled rows [G*3, C], idx [G, C]).  d carries no gradient (argmax), SURVEY B.3.
    tap=True: the pooled rows come with a pass-through alias of x for the dense consumer of the same activation, (pooled, x_alias)."""
    x = _rows2d(x, "x")
    with torch.no_grad():
        idx = maxpool_select(x, _rows2d(d.detach(), "d"), G, N) if forced_idx is None else forced_idx.reshape(G, -1).contiguous()
    if tap:
        return _MaxPoolGatherTap.apply(x, idx, G, N), idx
    return _MaxPoolGather.apply(x, idx, G, N), idx


class _LinearMaxPool(torch.autograd.Function):
    """out = VNMaxPool(VNLinear(x)) with the pooled layer's [R, C] activation kept only transiently and NO dense
    gradient: the backward scatters / gathers the one selected point per (sample, channel)."""

    @staticmethod
    def forward(ctx, x, w, wdir, G, N, forced_idx):
        x = _rows2d(x, "x")
        if w.stride(1) != 1:
            w = w.contiguous()
        C, K = w.shape
        idx = None
        if forced_idx is not None:
            idx = forced_idx.reshape(G, -1).contiguous()
        elif _GEMM_MODE == "tf32":
            # d = Wdir (W x) = (Wdir W) x: contract over K (the layer's input width) instead of C >= K, and run both
            # GEMMs as ONE tcgen05 kernel whose epilogue does the arg-max: neither f nor d is written to HBM
            wc = gemm_rows(wdir, w, True)
            wcat = torch.cat([w, wc], dim=0)
            best = torch.empty(G * C, device=x.device, dtype=torch.int64)
            with _Timed("gemm_vn_fused", 2.0 * x.shape[0] * K * 2 * C, 4.0 * (x.shape[0] * K + 2 * C * K)):
                rc = _lib.raw("vnpcc_gemm_vn_pool", ptr(x), _ld(x), ptr(wcat), _ld(wcat), x.shape[0], K, C, N, ptr(best), stream())
            if rc == 0:
                idx = torch.empty((G, C), device=x.device, dtype=torch.int64)
                call("vnpcc_vn_maxpool_decode", ptr(best), G * C, ptr(idx), stream())
            elif rc != 10003:
                raise _lib.VnpccError(f"vnpcc_gemm_vn_pool failed with code {rc}")
        out = torch.empty((G * 3, C), device=x.device, dtype=torch.float32)
        if idx is not None and _GEMM_MODE == "tf32" and K % 4 == 0:
            # pooled rows recomputed from the selected inputs (exact fp32 dot products)
            call("vnpcc_pool_linear_gather", ptr(x), _ld(x), ptr(w), _ld(w), ptr(idx), G, N, C, K, ptr(out), C, stream())
        else:
            f = gemm_rows(x, w)
            if idx is None:
                d = gemm_rows(f, wdir)       # reference order of operations (parity mode)
                idx = maxpool_select(f, d, G, N)
                del d
            call("vnpcc_vn_maxpool_gather", ptr(f), _ld(f), ptr(idx), G, N, C, ptr(out), C, stream())
        ctx.save_for_backward(x, w, idx)
        ctx.cfg = (G, N)
        ctx.mark_non_differentiable(idx)
        return out, idx

    @staticmethod
    def backward(ctx, g, _gidx):
        x, w, idx = ctx.saved_tensors
        G, N = ctx.cfg
        g = _rows2d(g, "grad")
        C, K = w.shape
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gw = torch.empty((C, K), device=x.device, dtype=torch.float32) if ctx.needs_input_grad[1] else None
        call("vnpcc_pool_linear_bwd", ptr(g), _ld(g), ptr(idx), ptr(x), _ld(x), ptr(w), _ld(w), G, N, C, K, ptr(gx),
             _ld(gx) if gx is not None else 0, ptr(gw), K, stream())
        return gx, gw, None, None, None, None


def linear_maxpool_rows(x, w, wdir, G, N, forced_idx=None):
    """VNLinear(K -> C) followed by VNMaxPool(C) over groups of N points: (pooled rows [G*3, C], idx [G, C])"""
    return _LinearMaxPool.apply(x, w, wdir.detach(), G, N, forced_idx)


# ---------------------------------------------------------------------------------------------------------------
# VNStdFeature: frame construction + projection of every channel onto the frame (csrc/vn_frame.cu)
# ---------------------------------------------------------------------------------------------------------------
class _VNFrame(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, z):
        x = _rows2d(x, "x")
        z = _rows2d(z, "z")
        R, C = x.shape
        P = R // 3
        J = z.shape[1]
        out = torch.empty((R, C), device=x.device, dtype=torch.float32)
        zout = torch.empty((R, 3), device=x.device, dtype=torch.float32)
        call("vnpcc_vn_frame_fwd", ptr(x), _ld(x), ptr(z), _ld(z), P, C, J, ptr(out), C, ptr(zout), stream())
        ctx.save_for_backward(x, z)
        return out, zout

    @staticmethod
    def backward(ctx, gout, gzout):
        x, z = ctx.saved_tensors
        R, C = x.shape
        J = z.shape[1]
        gout = _rows2d(gout, "grad") if gout is not None else torch.zeros((R, C), device=x.device, dtype=torch.float32)
        if gzout is not None:
            gzout = gzout.contiguous()
        gx = torch.empty((R, C), device=x.device, dtype=torch.float32)
        gz = torch.empty((R, J), device=x.device, dtype=torch.float32)
        call("vnpcc_vn_frame_bwd", ptr(gout), _ld(gout), ptr(gzout), ptr(x), _ld(x), ptr(z), _ld(z), R // 3, C, J, ptr(gx), C, ptr(gz), J,
             stream())
        return gx, gz


def vn_frame(x, z):
    """x rows (point, v) x C, z rows (point, v) x J (J = 3, or 2 = normalised frame) -> (x_std rows (point, k) x C, frame rows (point, k) x 3)"""
    return _VNFrame.apply(x, z)


# ---------------------------------------------------------------------------------------------------------------
# VNLinear(C -> 1) (+ residual)
# ---------------------------------------------------------------------------------------------------------------
class _RowsDot(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, res):
        x = _rows2d(x, "x")
        R, C = x.shape
        w = w.reshape(-1).contiguous()
        y = torch.empty(R, device=x.device, dtype=torch.float32)
        if res is not None:
            res = res.reshape(-1).contiguous()
        call("vnpcc_rows_dot", ptr(x), _ld(x), ptr(w), R, C, ptr(res), ptr(y), stream())
        ctx.save_for_backward(x, w)
        ctx.has_res = res is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gy = gy.contiguous()
        R, C = x.shape
        gx = torch.empty((R, C), device=x.device, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        gw = torch.empty(C, device=x.device, dtype=torch.float32) if ctx.needs_input_grad[1] else None
        call("vnpcc_rows_dot_bwd", ptr(gy), ptr(x), _ld(x), ptr(w), R, C, ptr(gx), C, ptr(gw), stream())
        return gx, (gw.view(1, C) if gw is not None else None), (gy if ctx.has_res else None)


def rows_dot(x, w, res=None):
    """y[r] = sum_c x[r,c] w[0,c] (+ res[r]);  w is the [1, C] weight of VNLinear(C, 1)"""
    return _RowsDot.apply(x, w, res)


# ---------------------------------------------------------------------------------------------------------------
# Chamfer
# ---------------------------------------------------------------------------------------------------------------
class chamfer_3DFunction(torch.autograd.Function):
    """Same contract as the reference's chamfer_3DFunction (extensions/chamfer_distance/chamfer_distance.py:29-71):
    forward(xyz1 [B,N,3], xyz2 [B,M,3]) -> (dist1 [B,N], dist2 [B,M], idx1, idx2 int32), squared distances."""

    @staticmethod
    def forward(ctx, xyz1, xyz2):
        _check(xyz1, "xyz1")
        _check(xyz2, "xyz2")
        if xyz1.dim() != 3 or xyz2.dim() != 3 or xyz1.shape[2] != 3 or xyz2.shape[2] != 3 or xyz1.shape[0] != xyz2.shape[0]:
            raise ValueError(f"expected [B,N,3] and [B,M,3], got {tuple(xyz1.shape)} and {tuple(xyz2.shape)}")
        xyz1 = xyz1.contiguous()
        xyz2 = xyz2.contiguous()
        B, N, _ = xyz1.shape
        M = xyz2.shape[1]
        dev = xyz1.device
        # zero-initialised like the reference only when a cloud is empty (outputs stay 0 then); otherwise the kernels overwrite every
        # element and four fill launches per call would be most of the cost of a small search
        alloc = torch.empty if (B > 0 and N > 0 and M > 0) else torch.zeros
        dist1 = alloc((B, N), device=dev, dtype=torch.float32)
        dist2 = alloc((B, M), device=dev, dtype=torch.float32)
        idx1 = alloc((B, N), device=dev, dtype=torch.int32)
        idx2 = alloc((B, M), device=dev, dtype=torch.int32)
        if B > 0 and N > 0 and M > 0:
            nb = _lib.raw("vnpcc_chamfer_workspace_bytes", B, N, M)
            ws = _workspace(nb, dev, "chamfer")
            with torch.cuda.device(dev), _Timed("chamfer_fwd", 2.0 * B * N * M):
                call("vnpcc_chamfer_forward", ptr(xyz1), ptr(xyz2), B, N, M, ptr(dist1), ptr(dist2), ptr(idx1), ptr(idx2),
                     ptr(ws), ws.numel(), stream())
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, dist2, idx1, idx2

    @staticmethod
    def backward(ctx, graddist1, graddist2, gradidx1=None, gradidx2=None):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        B, N, _ = xyz1.shape
        M = xyz2.shape[1]
        dev = xyz1.device
        graddist1 = graddist1.contiguous()
        graddist2 = graddist2.contiguous()
        need1, need2 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g1 = torch.empty_like(xyz1) if need1 else None
        g2 = torch.empty_like(xyz2) if need2 else None
        if B > 0 and (need1 or need2):
            with torch.cuda.device(dev):
                call("vnpcc_chamfer_backward", ptr(xyz1), ptr(xyz2), B, N, M, ptr(graddist1), ptr(graddist2), ptr(idx1),
                     ptr(idx2), ptr(g1), ptr(g2), stream())
        return g1, g2


class _CDReduce(torch.autograd.Function):
    """the sqrt/mean tails of cd_loss_L1/L2, l1_cd/l2_cd fused into one reduction (+ its backward)"""

    @staticmethod
    def forward(ctx, dist1, dist2, mode):
        dist1 = dist1.contiguous()
        dist2 = dist2.contiguous()
        B, N = dist1.shape
        M = dist2.shape[1]
        out = torch.empty(1, device=dist1.device, dtype=torch.float32)
        scratch = torch.empty(2, device=dist1.device, dtype=torch.float64)
        call("vnpcc_cd_reduce", ptr(dist1), ptr(dist2), B, N, M, mode, ptr(scratch), ptr(out), stream())
        ctx.save_for_backward(dist1, dist2)
        ctx.mode = mode
        return out.view(())

    @staticmethod
    def backward(ctx, gout):
        dist1, dist2 = ctx.saved_tensors
        B, N = dist1.shape
        M = dist2.shape[1]
        gout = gout.reshape(1).contiguous().float()
        g1 = torch.empty_like(dist1)
        g2 = torch.empty_like(dist2)
        call("vnpcc_cd_reduce_bwd", ptr(dist1), ptr(dist2), B, N, M, ctx.mode, ptr(gout), ptr(g1), ptr(g2), stream())
        return g1, g2, None


def cd_reduce(dist1, dist2, mode):
    return _CDReduce.apply(dist1, dist2, mode)


# ---------------------------------------------------------------------------------------------------------------
# fused Adam over flat buffers
# ---------------------------------------------------------------------------------------------------------------
def adam_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    call("vnpcc_adam_step", ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), float(lr), float(beta1), float(beta2), float(eps),
         float(weight_decay), int(step), float(grad_scale), stream())


def adam_step_dev(p, g, m, v, state, beta1, beta2, eps, weight_decay, grad_scale=1.0):
    """adam_step with lr / step / bias corrections in the 4-float device tensor `state` (graph-replayable)"""
    call("vnpcc_adam_step_dev", ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), ptr(state), float(beta1), float(beta2), float(eps),
         float(weight_decay), float(grad_scale), stream())


# ---------------------------------------------------------------------------------------------------------------
# transformer-refined decoder: VNLayerNorm, residual add, VN multi-head attention core
# ---------------------------------------------------------------------------------------------------------------
class _VNLayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        x = _rows2d(x, "x")
        R, C = x.shape
        P = R // 3
        y = torch.empty((R, C), device=x.device, dtype=torch.float32)
        need = any(ctx.needs_input_grad[:3])
        stats = torch.empty((P, 2), device=x.device, dtype=torch.float32) if need else None
        call("vnpcc_vn_layernorm_fwd", ptr(x), _ld(x), P, C, ptr(weight), ptr(bias), float(eps), ptr(y), C, ptr(stats), stream())
        ctx.save_for_backward(x, weight, bias, stats)
        return y

    @staticmethod
    def backward(ctx, g):
        x, weight, bias, stats = ctx.saved_tensors
        g = _rows2d(g, "grad")
        R, C = x.shape
        gx = torch.empty((R, C), device=x.device, dtype=torch.float32)
        gw = torch.empty(C, device=x.device, dtype=torch.float32)
        gb = torch.empty(C, device=x.device, dtype=torch.float32)
        call("vnpcc_vn_layernorm_bwd", ptr(g), _ld(g), ptr(x), _ld(x), R // 3, C, ptr(weight), ptr(bias), ptr(stats), ptr(gx), C, ptr(gw), ptr(gb),
             stream())
        return gx, gw, gb, None


def vn_layernorm(x, ln):
    """x rows [P*3, C]; ln: an nn.LayerNorm(C) module (parameter container) -> VNLayerNorm rows (models/vn_layers.py:140-150)"""
    if not ln.elementwise_affine:
        raise NotImplementedError("VNLayerNorm without affine parameters")
    return _VNLayerNorm.apply(x, ln.weight, ln.bias, ln.eps)


class _RowsAdd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a = _rows2d(a, "a")
        b = _rows2d(b, "b")
        out = torch.empty((a.shape[0], a.shape[1]), device=a.device, dtype=torch.float32)
        call("vnpcc_rows_add", ptr(a), _ld(a), ptr(b), _ld(b), ptr(out), a.shape[1], a.shape[0], a.shape[1], stream())
        return out

    @staticmethod
    def backward(ctx, g):
        return g, g


def rows_add(a, b):
    """a + b on rows (residual connections of VN_Block, models/transformer.py:60,68)"""
    return _RowsAdd.apply(a, b)


_ATTN_DS_WORKSPACE = True      # tests switch these off to exercise the other backward routes
_ATTN_STORE_P = True


class _VNAttention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, B, N, H, scale):
        qkv = _rows2d(qkv, "qkv")
        C = qkv.shape[1] // 3
        D = C // H
        out = torch.empty((qkv.shape[0], C), device=qkv.device, dtype=torch.float32)
        lse = torch.empty((B, H, N), device=qkv.device, dtype=torch.float32)
        pbuf = None
        with _Timed("attention_fwd", 4.0 * B * H * N * N * 3 * D):
            rc = 10003
            if _GEMM_MODE == "tf32":      # tcgen05 / TMEM forward (csrc/attention_tc.cu); shapes it does not take fall through
                if _ATTN_STORE_P and ctx.needs_input_grad[0] and N % 32 == 0 and D == 48:
                    # the attention weights P (B*H*N*N floats) are kept for the backward, which then consists of streaming GEMMs
                    pbuf = torch.empty(B * H * N * N, device=qkv.device, dtype=torch.float32)
                rc = _lib.raw("vnpcc_vn_attention_fwd_tf32", ptr(qkv), _ld(qkv), B, N, H, D, float(scale), ptr(out), C, ptr(lse), ptr(pbuf), stream())
                if rc not in (0, 10003):
                    raise _lib.VnpccError(f"vnpcc_vn_attention_fwd_tf32 failed with code {rc}")
                if rc == 0:
                    _LAST_KERNEL[0] = "attention_fwd_tf32"
                else:
                    pbuf = None
            if rc != 0:
                call("vnpcc_vn_attention_fwd", ptr(qkv), _ld(qkv), B, N, H, D, float(scale), ptr(out), C, ptr(lse), stream())
        ctx.save_for_backward(qkv, out, lse)
        ctx.pbuf = pbuf
        ctx.cfg = (B, N, H, D, float(scale))
        return out

    @staticmethod
    def backward(ctx, g):
        qkv, out, lse = ctx.saved_tensors
        B, N, H, D, scale = ctx.cfg
        g = _rows2d(g, "grad")
        if _ld(g) % 4 != 0 or g.data_ptr() % 16:
            g = g.contiguous()
        dqkv = torch.empty_like(qkv)
        delta = torch.empty((B, H, N), device=qkv.device, dtype=torch.float32)
        with _Timed("attention_bwd", 10.0 * B * H * N * N * 3 * D):
            rc = 10003
            if _GEMM_MODE == "tf32":      # tcgen05 / TMEM backward (csrc/attention_tc.cu)
                pbuf, ctx.pbuf = ctx.pbuf, None      # consumed: the backward turns P into dS in place
                ws = None
                if pbuf is None and N % 32 == 0 and _ATTN_DS_WORKSPACE:      # dS [B*H*N, N]: lets dK run as a plain streaming GEMM
                    ws = _workspace(4 * B * H * N * N, qkv.device, "attn_ds")
                rc = _lib.raw("vnpcc_vn_attention_bwd_tf32", ptr(qkv), _ld(qkv), ptr(g), _ld(g), ptr(out), _ld(out), ptr(lse), B, N, H, D, scale,
                              ptr(dqkv), _ld(dqkv), ptr(delta), ptr(ws), ws.numel() if ws is not None else 0, ptr(pbuf), stream())
                if rc not in (0, 10003):
                    raise _lib.VnpccError(f"vnpcc_vn_attention_bwd_tf32 failed with code {rc}")
                if rc == 0:
                    _LAST_KERNEL[0] = "attention_bwd_tf32"
            if rc != 0:
                call("vnpcc_vn_attention_bwd", ptr(qkv), _ld(qkv), ptr(g), _ld(g), ptr(out), _ld(out), ptr(lse), B, N, H, D, scale, ptr(dqkv),
                     _ld(dqkv), ptr(delta), stream())
        return dqkv, None, None, None, None


def vn_attention(qkv, B, N, H, scale):
    """qkv rows [B*N*3, 3C] = (q | k | v) -> attention output rows [B*N*3, C] (models/transformer.py:89-100)"""
    return _VNAttention.apply(qkv, B, N, H, scale)


# ---------------------------------------------------------------------------------------------------------------
# edge convolution without the edge tensor (csrc/edge_conv.cu): VN_DGCNN_fps graph feature -> VNLinearLeakyReLU(dim=5) -> mean over k
# ---------------------------------------------------------------------------------------------------------------
class _EdgeConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, uw, idx, gamma, beta, stat, use_batch, ns, B, N):
        uw = _rows2d(uw, "uw")
        C = uw.shape[1] // 4
        k = idx.shape[1]
        out = torch.empty((B * N * 3, C), device=uw.device, dtype=torch.float32)
        call("vnpcc_edge_conv_fwd", ptr(uw), _ld(uw), ptr(idx), B, N, k, C, ptr(stat), ptr(gamma), ptr(beta), float(ns), ptr(out), C, stream())
        ctx.save_for_backward(uw, idx, gamma, beta, stat)
        ctx.cfg = (B, N, k, C, float(ns), bool(use_batch))
        return out

    @staticmethod
    def backward(ctx, g):
        uw, idx, gamma, beta, stat = ctx.saved_tensors
        B, N, k, C, ns, use_batch = ctx.cfg
        g = _rows2d(g, "grad")
        if _ld(g) % 4 != 0 or g.data_ptr() % 16:
            g = g.contiguous()
        dev = uw.device
        guw = torch.empty((B * N * 3, 4 * C), device=dev, dtype=torch.float32)
        sums = torch.empty(2 * C, device=dev, dtype=torch.float64)
        ggamma = torch.empty(C, device=dev, dtype=torch.float32)
        gbeta = torch.empty(C, device=dev, dtype=torch.float32)
        call("vnpcc_edge_conv_bwd", ptr(g), _ld(g), ptr(uw), _ld(uw), ptr(idx), B, N, k, C, ptr(stat), ptr(gamma), ptr(beta), ns,
             1 if use_batch else 0, ptr(sums), ptr(guw), 4 * C, ptr(ggamma), ptr(gbeta), stream())
        return guw, None, ggamma, gbeta, None, None, None, None, None


def edge_conv_supported(C, bn):
    return C % 4 == 0 and 4 <= C <= 1024 and 256 % (C // 4) == 0 and bn is not None and bn.affine


def edge_conv(uw, idx, bn, training, ns, B, N):
    """uw rows (b,n,v) x 4C = (U_p | U_d | W_p | W_d) of the point GEMM, idx [B,k,N] int64 -> rows (b,n,v) x C =
    mean_j leaky(BN(U_p[j] + W_p[i]), U_d[j] + W_d[i]); BatchNorm statistics over the B*N*k edges"""
    uw = _rows2d(uw, "uw")
    C = uw.shape[1] // 4
    k = idx.shape[1]

    def stats_fn(sums):
        call("vnpcc_edge_conv_stats", ptr(uw), _ld(uw), ptr(idx), B, N, k, C, ptr(sums), stream())
    stat, use_batch = _bn_prepare(None, C, bn, training, B * N * k, stats_fn)
    return _EdgeConv.apply(uw, idx, bn.weight, bn.bias, stat, use_batch, ns, B, N)
