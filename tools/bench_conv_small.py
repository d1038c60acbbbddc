"""CUDA-event timing of the tile-engine convolutions at the low-resolution FSRNet shapes (encoder 64 ch @ 32x32, prior /
hourglass 128 ch @ 32x32 .. 8x8).  Usage: python tools/bench_conv_small.py [images]"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from crfr_b200 import _lib as L, ops   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
g = torch.Generator(device="cuda").manual_seed(3)


def run(c, h, reps=20):
    xs = [torch.randn(n, h, h, c, generator=g, device="cuda").to(torch.bfloat16) for _ in range(3)]
    w = ops.pack_conv_weight(torch.randn(c, c, 3, 3, generator=g, device="cuda") * 0.05)
    d = ops.conv_desc(xs[0], c, c, 3, 1, 1)
    y = torch.empty((n, h, h, c), dtype=torch.bfloat16, device="cuda")
    stats = torch.empty((n, c, 2), dtype=torch.float32, device="cuda")
    ws = ops.workspace(L.lib().crfr_conv_workspace_bytes(C.byref(d)))

    def call(i, st):
        L.call("crfr_conv_fwd", L.ENGINE_TCGEN05, C.byref(d), xs[i % 3].data_ptr(), w.data_ptr(), c, None, y.data_ptr(),
               None, st, 1e-5, ws.data_ptr(), ws.numel(), ops.stream())
    for with_stats in (False, True):
        for i in range(3):
            call(i, stats.data_ptr() if with_stats else None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            call(i, stats.data_ptr() if with_stats else None)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        fl = 2.0 * n * h * h * c * c * 9
        print("conv3x3 %3d->%3d @%3dx%-3d x%d %s  %7.1f us  %7.1f TFLOP/s" % (c, c, h, h, n, "+stats" if with_stats else "      ",
                                                                             us, fl / us / 1e6), flush=True)


for c, h in ((64, 32), (128, 32), (128, 16), (128, 8)):
    run(c, h)
