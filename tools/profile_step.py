"""One-chunk FSRNet train step for profiling (ncu launch list / full capture). Usage: python tools/profile_step.py [B] [reps]"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
import crfr_b200                                           # noqa: E402
from crfr_b200 import _lib as L, ops                       # noqa: E402
from crfr_b200.model import FSRnet as M                    # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(1234)
net = M.OverallNetwork()
net.apply(M.weights_init)
net = net.cuda()
params = net.ordered_parameters()
grads = [torch.zeros_like(p) for p in params]
g = torch.Generator().manual_seed(1)
x = torch.randn(B, 3, 128, 128, generator=g).cuda()
hr = torch.randn(B, 3, 128, 128, generator=g).cuda()
hm = torch.rand(B, 32, 32, generator=g).cuda()
lbl = torch.randint(0, 11, (B, 1, 32, 32), generator=g).cuda()
outs = M.alloc_outputs(x)
io = M._io(x, outs, (hr, hm, lbl), loss_div=2.0 * B, w_pix=5.0)
ws = torch.empty(L.lib().crfr_fsrnet_workspace_bytes(B, 128, 1), dtype=torch.uint8, device="cuda")
losses = torch.zeros(5, device="cuda")
pt, gt = M._ParamTable([p.detach() for p in params]), M._ParamTable(grads)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
import time
host = []
ev[0].record()
for i in range(reps):
    t0 = time.perf_counter()
    L.call("crfr_fsrnet_train_step", L.ENGINE_AUTO, pt.arr, gt.arr, C.byref(io), losses.data_ptr(), ws.data_ptr(),
           ws.numel(), ops.stream())
    host.append((time.perf_counter() - t0) * 1e3)
    ev[i + 1].record()
torch.cuda.synchronize()
print("host enqueue ms per chunk step:", host)
print("losses", losses.tolist())
print("ms per chunk step:", [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)], "launches", L.lib().crfr_launch_count())
