"""Row-streaming weight gradient (rowwgrad.cu): parity against fp32 PyTorch on ragged shapes, then timing at the BASELINE
size.  Usage: python tools/wgrad_diag.py"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crfr_b200 import _lib as L, ops                        # noqa: E402
from tests.util import bf16_round, nhwc_from, rel_err       # noqa: E402


def main():
    for n, h in ((2, 128), (3, 128), (7, 50), (300, 1), (200, 2), (40, 16), (1, 5), (64, 128)):
        g = torch.Generator().manual_seed(100 + n)
        x = bf16_round(torch.randn(n, 64, h, 128, generator=g))
        dy = bf16_round(torch.randn(n, 64, h, 128, generator=g))
        w = torch.zeros(64, 64, 3, 3, requires_grad=True)
        F.conv2d(x, w, None, 1, 1).backward(dy)
        dw, _ = ops.conv_wgrad(nhwc_from(x), nhwc_from(dy), 64, 64, 3, 1, 1, engine=L.ENGINE_TCGEN05)
        dw2, _ = ops.conv_wgrad(nhwc_from(x), nhwc_from(dy), 64, 64, 3, 1, 1, engine=L.ENGINE_TCGEN05)
        torch.cuda.synchronize()
        print("wgrad n=%3d h=%3d: rel err %.2e deterministic=%s" % (n, h, rel_err(dw, w.grad), torch.equal(dw, dw2)), flush=True)
    n, h = 128, 128
    x = torch.randn(n, h, 128, 64, device="cuda").to(torch.bfloat16)
    dy = torch.randn(n, h, 128, 64, device="cuda").to(torch.bfloat16)
    for _ in range(3):
        ops.conv_wgrad(x, dy, 64, 64, 3, 1, 1, engine=L.ENGINE_TCGEN05)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.conv_wgrad(x, dy, 64, 64, 3, 1, 1, engine=L.ENGINE_TCGEN05)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("time: %.1f us per wgrad (kernel + slab reduce + zero/alloc) = %.0f TFLOP/s" % (ms * 1e3, 2 * n * h * 128 * 64 * 576 / ms / 1e9))


if __name__ == "__main__":
    main()
