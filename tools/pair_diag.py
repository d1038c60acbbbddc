"""Diagnostic for the cta_group::2 row-streaming convolution (rowconv2.cu): parity against fp32 PyTorch and against the
single-CTA kernel, then timing at the BASELINE size.  Usage: python tools/pair_diag.py [swap] [time_only]"""
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crfr_b200 import _lib as L, ops                        # noqa: E402
from tests.util import bf16_round, nhwc_from, rel_err, to_nchw  # noqa: E402


def opt(name, v):
    L.call("crfr_set_option", name.encode(), v)


def run(x, w, b, pair, dgrad=False):
    opt("rowconv_pair", pair)
    n, c, h, w_ = x.shape
    if dgrad:
        y = ops.conv_dgrad(nhwc_from(x), ops.pack_conv_weight(w.cuda(), for_dgrad=True), (n, h, w_, c), c, c, 3, 1, 1,
                           engine=L.ENGINE_TCGEN05)
        st = None
    else:
        y, _, st = ops.conv_fwd(nhwc_from(x), ops.pack_conv_weight(w.cuda()), c, c, 3, 1, 1, bias=b.cuda(),
                                engine=L.ENGINE_TCGEN05, want_stats=True)
    torch.cuda.synchronize()
    return y, st


def main():
    swap = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    opt("pair_swap", swap)
    if len(sys.argv) <= 2:
        for n, h in ((2, 128), (4, 16), (2, 1), (2, 2), (6, 3), (300, 1), (40, 16), (8, 50), (128, 128)):
            g = torch.Generator().manual_seed(100 + n)
            x = bf16_round(torch.randn(n, 64, h, 128, generator=g))
            w = bf16_round(torch.randn(64, 64, 3, 3, generator=g) * 0.05)
            b = torch.randn(64, generator=g)
            ref = F.conv2d(x, w, b, 1, 1)
            y1, s1 = run(x, w, b, 0)
            y2, s2 = run(x, w, b, 1)
            e1, e2 = rel_err(to_nchw(y1), ref), rel_err(to_nchw(y2), ref)
            same = torch.equal(y1, y2)
            print("fwd   n=%3d h=%3d swap=%d: single %.2e pair %.2e identical=%s stats %.2e" % (
                n, h, swap, e1, e2, same, rel_err(s2, s1)), flush=True)
            xr = x.clone().requires_grad_(True)
            F.conv2d(xr, w, None, 1, 1).backward(x)          # dy := x (same shape)
            d1, _ = run(x, w, b, 0, dgrad=True)
            d2, _ = run(x, w, b, 1, dgrad=True)
            print("dgrad n=%3d h=%3d swap=%d: single %.2e pair %.2e identical=%s" % (
                n, h, swap, rel_err(to_nchw(d1), xr.grad), rel_err(to_nchw(d2), xr.grad), torch.equal(d1, d2)), flush=True)
    # timing at the BASELINE size
    n, h = 128, 128
    x = torch.randn(n, h, 128, 64, device="cuda").to(torch.bfloat16)
    w = ops.pack_conv_weight((torch.randn(64, 64, 3, 3) * 0.05).cuda())

    def timed(stats):
        for _ in range(3):
            ops.conv_fwd(x, w, 64, 64, 3, 1, 1, engine=L.ENGINE_TCGEN05, want_stats=stats)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.conv_fwd(x, w, 64, 64, 3, 1, 1, engine=L.ENGINE_TCGEN05, want_stats=stats)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 20

    for dbg in ([int(a) for a in sys.argv[3:]] or [0]):
        opt("pair_debug", dbg)
        for pair in ((0, 1) if dbg == 0 else (1,)):
            for stats in ((True, False) if dbg == 0 else (True,)):
                opt("rowconv_pair", pair)
                ms = timed(stats)
                if dbg & 32:
                    import ctypes as C
                    import numpy as np
                    buf = (C.c_longlong * (74 * 16))()
                    L.call("crfr_debug_pair_profile", buf, 74 * 16)
                    a = np.array(list(buf), dtype=np.float64).reshape(74, 16)[:, :10].mean(0)
                    names = ["mma total", "wait acc_empty", "wait full", "wait peer_full", "issue+commit", "epi total",
                             "epi wait acc_full", "epi tmem ld/st/arrive", "epi pack+store", "epi stats"]
                    rows_per = 128 * 64 / 74.0
                    for k, (nm, v) in enumerate(zip(names, a)):
                        own = rows_per / 2 if k >= 6 else rows_per
                        print("    %-24s %9.0f cycles = %6.0f per %s" % (nm, v, v / own, "own row" if k >= 6 else "row"))
                print("time pair=%d stats=%d dbg=%2d: %.1f us per launch (+finalize) = %.0f TFLOP/s" % (
                    pair, stats, dbg, ms * 1e3, 2 * n * h * 128 * 64 * 576 / ms / 1e9))


if __name__ == "__main__":
    main()
