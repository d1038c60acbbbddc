"""CUDA-event timing of the tile-engine weight gradients at the low-resolution shapes (64 ch @ 32x32 / 56x56, 128 ch @ 32x32 /
16x16).  Usage: python tools/bench_wgrad_small.py [images]"""
import sys

import torch

sys.path.insert(0, ".")
from crfr_b200 import _lib as L, ops   # noqa: E402
from tests.util import rel_err          # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
g = torch.Generator(device="cuda").manual_seed(3)
for c, h in ((64, 32), (64, 56), (128, 32), (128, 16)):
    xs = [torch.randn(n, h, h, c, generator=g, device="cuda").to(torch.bfloat16) for _ in range(3)]
    dy = torch.randn(n, h, h, c, generator=g, device="cuda").to(torch.bfloat16)
    dw, _ = ops.conv_wgrad(xs[0], dy, c, c, 3, 1, 1, engine=L.ENGINE_TCGEN05)
    ref, _ = ops.conv_wgrad(xs[0][:16], dy[:16], c, c, 3, 1, 1, engine=L.ENGINE_DIRECT)
    chk, _ = ops.conv_wgrad(xs[0][:16], dy[:16], c, c, 3, 1, 1, engine=L.ENGINE_TCGEN05)
    for i in range(3):
        ops.conv_wgrad(xs[i], dy, c, c, 3, 1, 1, engine=L.ENGINE_TCGEN05)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(12):
        ops.conv_wgrad(xs[i % 3], dy, c, c, 3, 1, 1, engine=L.ENGINE_TCGEN05)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 12 * 1e3
    print("wgrad3x3 %3d->%3d @%3dx%-3d x%d  %7.1f us (incl. memset + unpack)  %7.1f TFLOP/s   vs direct engine %.2e"
          % (c, c, h, h, n, us, 2.0 * n * h * h * c * c * 9 / us / 1e6, rel_err(chk, ref)), flush=True)
