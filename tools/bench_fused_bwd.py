"""CUDA-event timing of the backward across conv(PReLU(InstanceNorm(y) (+ res))) at 64 x 128 x 128: dgrad alone, the
normalisation backward alone, and crfr_conv_dgrad_norm_bwd with the first pass fused into the dgrad epilogue or not."""
import sys

import torch

sys.path.insert(0, ".")
from crfr_b200 import _lib as L, ops   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
c, h = 64, 128
g = torch.Generator(device="cuda").manual_seed(3)
mk = lambda s=1.0: (torch.randn(n, h, h, c, generator=g, device="cuda") * s).to(torch.bfloat16)
ys = [mk(1.3) for _ in range(2)]
douts = [mk() for _ in range(2)]
res, dxb = mk(), mk(0.5)
w = (torch.randn(c, c, 3, 3, generator=g, device="cuda") * 0.05).to(torch.bfloat16).float()
wt = ops.pack_conv_weight(w, for_dgrad=True)
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


t = timeit(lambda i: ops.conv_dgrad(douts[i % 2], wt, (n, h, h, c), c, c, 3, 1, 1, engine=L.ENGINE_TCGEN05))
print("dgrad alone                         %7.1f us" % t)
for name, r, b in (("plain", None, None), ("second gradient", None, dxb), ("residual", res, None),
                   ("residual + second gradient", res, dxb)):
    t = timeit(lambda i: ops.norm_act_bwd(douts[i % 2], ys[i % 2], stats[i % 2], gamma, beta, alpha, res=r, dout_b=b))
    print("norm backward alone, %-26s %7.1f us" % (name, t))
    for mode in (0, 1):
        ops.set_option("fuse_norm_bwd", mode)
        t = timeit(lambda i: ops.conv_dgrad_norm_bwd(douts[i % 2], wt, ys[i % 2], stats[i % 2], c, c, 3, 1, 1, gamma, beta,
                                                     alpha, res=r, dx_b=b, engine=L.ENGINE_TCGEN05))
        print("  dgrad + norm backward, fuse=%d                 %7.1f us" % (mode, t))
    ops.set_option("fuse_norm_bwd", 1)
