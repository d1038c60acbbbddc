"""CUDA-event timing of conv(PReLU(InstanceNorm(y) (+ res))) forward at 64 x 128 x 128: crfr_norm_act_conv_fwd with the
normalisation in the convolution's producer warps (fuse_norm_fwd = 1) or as two kernels (0)."""
import sys

import torch

sys.path.insert(0, ".")
from crfr_b200 import _lib as L, ops   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
c, h = 64, 128
g = torch.Generator(device="cuda").manual_seed(3)
mk = lambda s=1.0: (torch.randn(n, h, h, c, generator=g, device="cuda") * s).to(torch.bfloat16)
ys = [mk(1.3) for _ in range(2)]
res = mk()
w = (torch.randn(c, c, 3, 3, generator=g, device="cuda") * 0.05).to(torch.bfloat16).float()
wp = ops.pack_conv_weight(w)
gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.3
alpha = torch.rand(c, device="cuda") * 0.5
stats = [ops.norm_stats(y) for y in ys]
_big = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)


def timeit(fn, reps=10):
    for i in range(2):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(4):
        torch.mm(_big, _big)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for name, r in (("plain", None), ("residual", res)):
    for mode in (0, 1):
        ops.set_option("fuse_norm_fwd", mode)
        t = timeit(lambda i: ops.norm_act_conv_fwd(ys[i % 2], stats[i % 2], wp, c, c, 3, 1, 1, gamma, beta, alpha, res=r,
                                                   engine=L.ENGINE_TCGEN05))
        print("norm + conv forward (+ statistics), %-8s fuse=%d   %7.1f us" % (name, mode, t))
    ops.set_option("fuse_norm_fwd", 1)
