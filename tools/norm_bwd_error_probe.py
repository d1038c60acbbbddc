"""Measured error of the InstanceNorm + PReLU (+ residual) backward kernels against fp32 autograd (tolerances of
tests/test_kernels_gpu.py / tests/test_fused_bwd_gpu.py)."""
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from crfr_b200 import ops                                  # noqa: E402
from tests.util import bf16_round, nhwc_from, rel_err, to_nchw   # noqa: E402


def case(n, c, h, w, res, two, seed):
    g = torch.Generator().manual_seed(seed)
    y = bf16_round(torch.randn(n, c, h, w, generator=g) * 1.7 + 0.8)
    r = bf16_round(torch.randn(n, c, h, w, generator=g)) if res else None
    gamma, beta, alpha = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g), torch.rand(c, generator=g) * 0.5
    yr, gr, br, ar = (t.clone().requires_grad_(True) for t in (y, gamma, beta, alpha))
    rr = r.clone().requires_grad_(True) if res else None
    z = F.instance_norm(yr, weight=gr, bias=br, eps=1e-5)
    if res:
        z = z + rr
    out = F.prelu(z, ar)
    d1 = bf16_round(torch.randn(out.shape, generator=g))
    d2 = bf16_round(torch.randn(out.shape, generator=g)) if two else None
    out.backward(d1 + d2 if two else d1)
    yg = nhwc_from(y)
    st = ops.norm_stats(yg)
    for mode in (0, 1):
        ops.set_option("norm_bwd_impl", mode)
        dz, dy, dg, db, da = ops.norm_act_bwd(nhwc_from(d1), yg, st, gamma.cuda(), beta.cuda(), alpha.cuda(),
                                              res=None if r is None else nhwc_from(r), dout_b=None if d2 is None else nhwc_from(d2))
        print("n=%d c=%d %dx%d res=%d two=%d impl=%d: dy %.2e dz %s dgamma %.2e dbeta %.2e dalpha %.2e" % (
            n, c, h, w, res, two, mode, rel_err(to_nchw(dy), yr.grad), "%.2e" % rel_err(to_nchw(dz), rr.grad) if res else "-",
            rel_err(dg, gr.grad), rel_err(db, br.grad), rel_err(da, ar.grad)))
    ops.set_option("norm_bwd_impl", -1)


for a in [(3, 64, 32, 32, True, True), (3, 128, 8, 8, True, True), (3, 128, 16, 16, False, False), (40, 64, 32, 32, True, True),
          (40, 64, 32, 32, False, False), (24, 128, 16, 16, True, False), (300, 64, 8, 8, False, True), (5, 64, 40, 40, True, True),
          (8, 64, 128, 128, True, True)]:
    case(*a, seed=sum(a[:4]))
