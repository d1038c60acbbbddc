"""A few launches of the row-streaming pair kernel at 128 images for ncu.  argv: fwd | bwd_plain | bwd_res2"""
import sys

import torch

sys.path.insert(0, ".")
from crfr_b200 import _lib as L, ops   # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"
n, c, h = 128, 64, 128
g = torch.Generator(device="cuda").manual_seed(3)
mk = lambda s=1.0: (torch.randn(n, h, h, c, generator=g, device="cuda") * s).to(torch.bfloat16)
y, dout, res, dxb = mk(1.3), mk(), mk(), mk(0.5)
w = (torch.randn(c, c, 3, 3, generator=g, device="cuda") * 0.05).to(torch.bfloat16).float()
wp, wt = ops.pack_conv_weight(w), ops.pack_conv_weight(w, for_dgrad=True)
gamma, beta, alpha = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.3, torch.rand(c, device="cuda") * 0.5
stats = ops.norm_stats(y)
for _ in range(4):
    if mode == "fwd":
        ops.conv_fwd(y, wp, c, c, 3, 1, 1, engine=L.ENGINE_TCGEN05, want_stats=True)
    else:
        ops.conv_dgrad_norm_bwd(dout, wt, y, stats, c, c, 3, 1, 1, gamma, beta, alpha, res=res if mode == "bwd_res2" else None,
                                dx_b=dxb if mode == "bwd_res2" else None, engine=L.ENGINE_TCGEN05)
torch.cuda.synchronize()
print("ok")
