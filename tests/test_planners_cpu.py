"""Host logic of the launch planners, without a GPU: the C-ABI's introspection entry points return the work-item splits / chunk lengths /
grids the launchers would use for a problem size.  Invariants: the plan covers the whole problem, respects the kernels' granularities, and
wastes at most a bounded part of its last wave (the property the planners exist for)."""
import ctypes as C
import math

import pytest

from vn_pointcloudcompletion_b200 import _lib

SMS = 148


def _ints(n, ctype=C.c_int):
    return (ctype * n)()


@pytest.mark.parametrize("B,N,M", [(32, 16384, 16384), (32, 1024, 16384), (32, 16384, 1024), (32, 1024, 1024), (32, 2048, 2048), (4, 100, 200),
                                   (1, 1, 1), (2, 300, 5), (32, 65536, 65536), (7, 5000, 333)])
def test_chamfer_plan(B, N, M):
    out = _ints(4)
    _lib.raw("vnpcc_debug_chamfer_plan", B, N, M, out)
    nq, ns, sl, qb = list(out)
    assert qb == 1024 and nq == math.ceil(N / qb)
    assert ns >= 1 and sl % 256 == 0 and sl >= 256            # whole 32-candidate chunks, the kernel's tile granularity
    assert ns * sl >= M and (ns - 1) * sl < M                  # the splits cover every candidate, none is empty
    assert ns <= 64
    slots = SMS * 8          # the planner sizes the grid for 8 CTAs per SM (chamfer_ctas_per_sm): late CTAs back-fill the tail wave
    items = B * nq * ns
    waves = math.ceil(items / slots)
    # never worse than the unsplit plan by the planner's own cost model
    unsplit = math.ceil(B * nq / slots) * (math.ceil(M / 256) * 256 + 256)
    assert waves * (sl + 256) <= unsplit
    # the workspace the ABI asks for holds (best, second, chunk) per (query, split) in both directions
    assert _lib.raw("vnpcc_chamfer_workspace_bytes", B, N, M) >= 3 * 4 * B * N * ns


@pytest.mark.parametrize("B,N,C,resident,lanes", [(32, 16384, 256, 1, 4), (32, 16384, 256, 2, 2), (32, 16384, 256, 3, 4), (2, 70, 128, 2, 4),
                                                  (6, 16384, 512, 2, 2), (32768, 16, 256, 2, 2), (32768, 16, 128, 1, 4), (1, 5, 128, 2, 4)])
def test_fold_geometry(B, N, C, resident, lanes):
    out = _ints(6)
    _lib.raw("vnpcc_debug_fold_geometry", B, N, C, resident, lanes, out)
    gx, gy, bx, by, n_chunk, row_mode = list(out)
    assert bx == C // lanes and bx % 32 == 0 and bx * by <= 256 and by >= 1
    if row_mode:      # every block row owns whole samples and loops over them
        assert N <= 64 and gx == 1 and n_chunk == N and 1 <= gy <= SMS * resident and gy * by <= B + by - 1
    else:
        assert gy == B and gx * n_chunk >= N and (gx - 1) * n_chunk < N
        slots = SMS * resident
        waves = math.ceil(gx * gy / slots)
        if N >= 64 * by:      # enough points to choose from: at most 15 % of the launched block slots idle
            assert gx * gy / (waves * slots) >= 0.85, (gx, gy, waves)


@pytest.mark.parametrize("R,Cout,K", [(196608, 2048, 512), (1572864, 512, 256), (196608, 512, 128), (98304 * 3, 1152, 384), (4096, 256, 256),
                                      (300, 64, 64)])
def test_wgrad_plan(R, Cout, K):
    out = _ints(4, C.c_longlong)
    _lib.raw("vnpcc_debug_wgrad_plan", R, Cout, K, SMS, out)
    gm, gn, splits, rps = list(out)
    assert gm == math.ceil(Cout / 128) and gn == math.ceil(K / 256)
    assert splits >= 1 and rps % 32 == 0 and splits * rps >= R and (splits - 1) * rps < R
    ctas = gm * gn * splits
    waves = math.ceil(ctas / SMS)
    if R >= 65536:
        assert ctas / (waves * SMS) >= 0.85, (ctas, waves)


@pytest.mark.parametrize("groups,N,slots,lanes,min_chunk", [(512, 2048, 1184, 8, 64), (2048, 2048, 1184, 8, 64), (4, 64, 1184, 8, 64),
                                                            (96, 16384, 592, 8, 64), (1, 1, 148, 8, 64)])
def test_plan_chunk_len(groups, N, slots, lanes, min_chunk):
    n_chunk = _lib.raw("vnpcc_debug_plan_chunk_len", groups, N, slots, lanes, min_chunk)
    assert n_chunk >= min(N, min_chunk) or n_chunk >= N
    chunks = math.ceil(N / n_chunk)
    assert chunks * n_chunk >= N
    # no chunk count the planner may consider is cheaper under its cost model
    def cost(c):
        ln = math.ceil(N / c)
        return math.ceil(groups * math.ceil(N / ln) / slots) * (math.ceil(ln / lanes) + 4)
    hi = max(1, min(max(1, N // min_chunk), math.ceil(8 * slots / groups)))
    assert cost(chunks) <= min(cost(c) for c in range(1, hi + 1))


def _rows_plan(R, K, Cout, bias=0, stats=0, sms=SMS):
    out = _ints(5, C.c_longlong)
    _lib.raw("vnpcc_debug_rows_plan", R, K, Cout, bias, stats, sms, out)
    return dict(zip(("variant", "ksplit", "bn", "grid", "tiles"), list(out)))


def test_rows_gemm_plan_at_the_train_step_shapes():
    """which form of the tcgen05 rows GEMM the BASELINE step's launches get: CTA pairs for the big K >= 256 layers,
    one SM per tile for K = 128, split-K for the 96-row heads"""
    dec_fwd = _rows_plan(1572864, 256, 512, stats=1)                    # decoder final_conv[1] forward with statistics
    assert dec_fwd["variant"] == 1 and dec_fwd["bn"] == 240 and dec_fwd["grid"] == SMS
    enc_fwd = _rows_plan(196608, 512, 2048, bias=1, stats=1)           # encoder second_conv[0] forward (per-sample bias)
    assert enc_fwd["variant"] == 1 and enc_fwd["tiles"] == 8 * math.ceil(196608 / 240)
    assert _rows_plan(1572864, 512, 256)["variant"] == 1                # decoder dgrad
    assert _rows_plan(196608, 128, 512)["variant"] == 0                 # K = 128: measured slower on pairs
    assert _rows_plan(196608, 512, 128)["variant"] == 0                 # no 256-channel tile
    head = _rows_plan(96, 1024, 1024)                                   # per-sample heads: 8 CTAs -> 64
    assert head["variant"] == 2 and head["ksplit"] == 8 and head["grid"] == 64 and head["bn"] == 128
    assert _rows_plan(96, 1024, 1024, bias=1)["variant"] == 0           # partial products cannot carry the bias


@pytest.mark.parametrize("R,K,Cout,bias,stats", [(96, 1024, 3072, 0, 0), (64, 256, 512, 0, 0), (128, 1000, 384, 0, 0), (100, 4096, 64, 0, 0),
                                                  (96, 96, 1024, 0, 0), (8192, 256, 256, 0, 0), (8191, 256, 256, 0, 0), (30000, 320, 768, 1, 0),
                                                  (3 * 5000, 200, 384, 0, 1), (10 ** 6, 2048, 4096, 0, 0), (300, 64, 128, 0, 0)])
def test_rows_gemm_plan_invariants(R, K, Cout, bias, stats):
    p = _rows_plan(R, K, Cout, bias, stats)
    num_m, num_kb = math.ceil(Cout / 128), math.ceil(K / 32)
    assert 1 <= p["grid"] <= SMS and p["grid"] <= max(p["tiles"], 1) * (2 if p["variant"] == 1 else 1)
    if p["variant"] == 2:       # split-K: few rows, no bias, every split non-empty and at least ~4 K blocks long, about one CTA per SM at most
        assert R <= 128 and not bias and not stats
        kb_per = math.ceil(num_kb / p["ksplit"])
        assert p["ksplit"] >= 2 and kb_per * (p["ksplit"] - 1) < num_kb and kb_per >= 4
        assert p["tiles"] == num_m * p["ksplit"] and p["tiles"] <= SMS + num_m
    elif p["variant"] == 1:     # CTA pairs: whole 256-channel tiles, long contraction, many row blocks; clusters of two CTAs
        assert Cout % 256 == 0 and K >= 256 and R >= 8192
        assert p["grid"] % 2 == 0 and p["tiles"] == (Cout // 256) * math.ceil(R / p["bn"])
    else:
        assert p["tiles"] == num_m * math.ceil(R / p["bn"]) and p["ksplit"] == 1
    assert p["bn"] == (128 if (R <= 128 and not stats) else (240 if stats else 256))
