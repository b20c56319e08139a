"""development aid: rows GEMM on CTA pairs (tcgen05 cta_group::2, the default where eligible) against the one-SM kernel: equality and time"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vn_pointcloudcompletion_b200 as V
from vn_pointcloudcompletion_b200 import _lib, ops
from stream_bench import timed

V.set_gemm_mode("tf32")
for (R, K, Cout, Cs, nb) in [(3 * 5000, 64, 256, 0, 0), (1572864, 256, 512, 256, 0), (196608, 512, 2048, 1024, 32), (196608, 128, 512, 0, 0), (1572864, 512, 256, 0, 0),
                             (196608, 2048, 512, 0, 0)]:
    x = torch.randn(R, K, device="cuda")
    w = torch.randn(Cout, K, device="cuda") / K ** 0.5
    bias = torch.randn(nb * 3, Cout, device="cuda") if nb else None
    rps = R // nb if nb else 0
    out, tms, sm = [], [], []
    for knob in (1, 0):      # knob 2: 1 = one SM per tile (4 stages), 0 = default (CTA pairs where eligible)
        _lib.raw("vnpcc_set_tuning", 2, knob)
        y = torch.empty(R, Cout, device="cuda")
        sums = torch.zeros(2 * Cs, device="cuda", dtype=torch.float64) if Cs else None
        fn = (lambda: ops.gemm_rows(x, w, False, bias, rps, out=y, stats=(sums, Cs))) if Cs else (lambda: ops.gemm_rows(x, w, False, bias, rps, out=y))
        fn()
        torch.cuda.synchronize()
        out.append(y.clone())
        sm.append(sums.clone() if Cs else None)
        tms.append(timed(fn, 10))
    _lib.raw("vnpcc_set_tuning", 2, 0)
    same = torch.equal(out[0], out[1])
    md = float((out[0] - out[1]).abs().max())
    sd = float(((sm[0] - sm[1]).abs() / sm[0].abs()).max()) if Cs else 0.0
    gb = 4.0 * (R * K + R * Cout) / 1e9
    print(f"R={R} K={K} Cout={Cout} stats={Cs} bias={nb > 0}: one SM {tms[0]:.3f} ms, pair {tms[1]:.3f} ms ({gb / tms[1]:.2f} TB/s, {2e-9 * R * K * Cout / tms[1]:.0f} TF/s)"
          f"  equal={same} maxdiff={md:.2e} stats rel diff={sd:.1e}", flush=True)
