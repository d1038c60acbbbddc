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
MiB = n * h * h * c * 2 / 2**20
K = max(4, int(600 / MiB) + 1)          # rotate enough buffer sets that nothing survives in the 126 MB L2 between reps
ys = [mk() for _ in range(K)]
das = [mk() for _ in range(K)]
res, da, db = mk(), das[0], mk()
gamma = torch.rand(c, device="cuda") + 0.5
beta = torch.randn(c, device="cuda")
alpha = torch.rand(c, device="cuda") * 0.5
_big = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)


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
t = timeit(lambda i: ops.norm_stats(ys[i % K]))
print("stats (partial+finalize)      %6.1f us  %5.2f TB/s (1 map read)" % (t, MiB * 2**20 / t / 1e6))
for fs in (0, 1):
    ops.set_option("norm_fwd_stream", fs)
    t = timeit(lambda i: ops.norm_act_fwd(ys[i % K], stats, gamma, beta, alpha, res=res))
    print("fwd %s apply (+res)      %6.1f us  %5.2f TB/s (2 reads + 1 write)" % ("stream" if fs else "regs  ", t, 3 * MiB * 2**20 / t / 1e6))
    t = timeit(lambda i: ops.norm_act_fwd(ys[i % K], stats, gamma, beta, alpha))
    print("fwd %s apply             %6.1f us  %5.2f TB/s (1 read + 1 write)" % ("stream" if fs else "regs  ", t, 2 * MiB * 2**20 / t / 1e6))
ops.set_option("norm_fwd_stream", 1)
import os
MODES = [(0, "regs    "), (1, "stream  ")]
if os.environ.get("BENCH_MODES"):
    MODES = [m for m in MODES if str(m[0]) in os.environ["BENCH_MODES"]]
for mode, name in MODES:
    ops.set_option("norm_bwd_impl", mode)
    t = timeit(lambda i: ops.norm_act_bwd(das[i % K], ys[i % K], stats, gamma, beta, alpha, res=res, dout_b=db))
    print("bwd %s res + 2nd grad     %6.1f us  %5.2f TB/s (6 reads + 2 writes)"
          % (name, t, 8 * MiB * 2**20 / t / 1e6))
    t = timeit(lambda i: ops.norm_act_bwd(das[i % K], ys[i % K], stats, gamma, beta, alpha, res=res))
    print("bwd %s res                %6.1f us  %5.2f TB/s (5 reads + 2 writes)"
          % (name, t, 7 * MiB * 2**20 / t / 1e6))
    t = timeit(lambda i: ops.norm_act_bwd(das[i % K], ys[i % K], stats, gamma, beta, alpha))
    print("bwd %s plain              %6.1f us  %5.2f TB/s (4 reads + 1 write)"
          % (name, t, 5 * MiB * 2**20 / t / 1e6))
ops.set_option("norm_bwd_impl", -1)
