"""Correctness + CUDA-event timing of the main conv shapes (3x3 64->64 @128x128, chunk of images)."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from crfr_b200 import _lib as L, ops   # noqa: E402
from tests.util import rel_err          # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
c, h = 64, 128
g = torch.Generator(device="cuda").manual_seed(3)
xs = [torch.randn(n, h, h, c, generator=g, device="cuda").to(torch.bfloat16) for _ in range(4)]
dy = torch.randn(n, h, h, c, generator=g, device="cuda").to(torch.bfloat16)
w = (torch.randn(c, c, 3, 3, generator=g, device="cuda") * 0.05).to(torch.bfloat16).float()
b = torch.randn(c, generator=g, device="cuda")
wp, wt = ops.pack_conv_weight(w), ops.pack_conv_weight(w, for_dgrad=True)

y_tc, _, _ = ops.conv_fwd(xs[0], wp, c, c, 3, 1, 1, bias=b, engine=L.ENGINE_TCGEN05)
y_d, _, _ = ops.conv_fwd(xs[0], wp, c, c, 3, 1, 1, bias=b, engine=L.ENGINE_DIRECT)
torch.cuda.synchronize()
print("fwd   tc vs direct rel_err %.3e" % rel_err(y_tc.float(), y_d.float()), flush=True)
dx_tc = ops.conv_dgrad(dy, wt, (n, h, h, c), c, c, 3, 1, 1, engine=L.ENGINE_TCGEN05)
dx_d = ops.conv_dgrad(dy, wt, (n, h, h, c), c, c, 3, 1, 1, engine=L.ENGINE_DIRECT)
torch.cuda.synchronize()
print("dgrad tc vs direct rel_err %.3e" % rel_err(dx_tc.float(), dx_d.float()), flush=True)


_big = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)


def timeit(fn, reps=20):
    """GPU-side time per call: the calls are enqueued behind a long matmul so host launch cost is hidden."""
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


d = ops.conv_desc(xs[0], c, c, 3, 1, 1)
y = torch.empty((n, h, h, c), dtype=torch.bfloat16, device="cuda")
ws = ops.workspace(L.lib().crfr_conv_workspace_bytes(C.byref(d)))
dw = torch.zeros((c, c, 3, 3), device="cuda")
flops = 2.0 * n * h * h * c * c * 9
t = timeit(lambda i: L.call("crfr_conv_fwd", L.ENGINE_TCGEN05, C.byref(d), xs[i % 4].data_ptr(), wp.data_ptr(), c, None,
                            y.data_ptr(), None, None, 1e-5, ws.data_ptr(), ws.numel(), ops.stream()))
print("fwd   %.1f us  %.1f TFLOP/s" % (t, flops / t / 1e6))
t = timeit(lambda i: L.call("crfr_conv_dgrad", L.ENGINE_TCGEN05, C.byref(d), xs[i % 4].data_ptr(), wt.data_ptr(), c,
                            y.data_ptr(), ws.data_ptr(), ws.numel(), ops.stream()))
print("dgrad %.1f us  %.1f TFLOP/s" % (t, flops / t / 1e6))
t = timeit(lambda i: L.call("crfr_conv_wgrad", L.ENGINE_TCGEN05, C.byref(d), xs[i % 4].data_ptr(), dy.data_ptr(),
                            dw.data_ptr(), None, ws.data_ptr(), ws.numel(), ops.stream()))
print("wgrad %.1f us  %.1f TFLOP/s (incl. memset + unpack)" % (t, flops / t / 1e6))
