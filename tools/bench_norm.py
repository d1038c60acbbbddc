"""GPU-side timing of the HBM-bound normalise/activation kernels at the full-resolution FSRNet shape."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from crfr_b200 import _lib as L, ops   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
c, h = 64, 128
g = torch.Generator(device="cuda").manual_seed(3)
mk = lambda: torch.randn(n, h, h, c, generator=g, device="cuda").to(torch.bfloat16)
ys = [mk() for _ in range(4)]
res, da, db = mk(), mk(), mk()
gamma = torch.rand(c, device="cuda") + 0.5
beta = torch.randn(c, device="cuda")
alpha = torch.rand(c, device="cuda") * 0.5
_big = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
MiB = n * h * h * c * 2 / 2**20


def timeit(fn, reps=20):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(6):
        torch.mm(_big, _big)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


stats = ops.norm_stats(ys[0])
t = timeit(lambda i: ops.norm_stats(ys[i % 4]))
print("stats (partial+finalize)      %6.1f us  %5.2f TB/s (1 map read)" % (t, MiB * 2**20 / t / 1e6))
t = timeit(lambda i: ops.norm_act_fwd(ys[i % 4], stats, gamma, beta, alpha, res=res))
print("fwd apply (+res)              %6.1f us  %5.2f TB/s (2 reads + 1 write)" % (t, 3 * MiB * 2**20 / t / 1e6))
t = timeit(lambda i: ops.norm_act_bwd(da, ys[i % 4], stats, gamma, beta, alpha, res=res, dout_b=db))
print("bwd (reduce+fold+param+apply) %6.1f us  %5.2f TB/s (6 reads + 2 writes)" % (t, 8 * MiB * 2**20 / t / 1e6))
