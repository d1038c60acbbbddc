"""GPU-side timing of the fused cosine top-k matcher. Usage: python tools/bench_match.py [probes] [gallery]"""
import sys

import torch

sys.path.insert(0, ".")
from crfr_b200 import ops   # noqa: E402

p = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
g = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
gen = torch.Generator(device="cuda").manual_seed(11)
gal = ops.l2norm_bf16(torch.randn(g, 512, generator=gen, device="cuda"))
pr = ops.l2norm_bf16(torch.randn(p, 512, generator=gen, device="cuda"))
ops.cosine_topk(pr, gal, 5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    ops.cosine_topk(pr, gal, 5)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print("%d x %d x 512: %.2f ms  %.1f TFLOP/s  %.0f probes/s" % (p, g, ms, 2.0 * p * g * 512 / ms / 1e9, p / ms * 1e3))
