"""Diagnostics for the tcgen05 engine on a real B200: localises descriptor / TMA-coordinate / layout bugs.
Usage (on the GPU box): python tools/tc_diag.py [fwd|dgrad|wgrad|match]"""
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from crfr_b200 import _lib as L, ops                      # noqa: E402
from tests.util import bf16_round, nhwc_from, rel_err, to_nchw  # noqa: E402


def summarize(name, got, ref):
    e = rel_err(got, ref)
    d = (got.double().cpu() - ref.double().cpu()).abs()
    print("%-44s rel_err %.3e  max_abs %.3e  ref_norm %.3e  got_norm %.3e  nan %d" % (
        name, e, d.max().item(), ref.double().norm().item(), got.double().norm().item(),
        int(torch.isnan(got).sum())), flush=True)
    return e


def diag_fwd(n=2, cin=64, cout=64, h=32):
    g = torch.Generator().manual_seed(1)
    x = bf16_round(torch.randn(n, cin, h, h, generator=g))
    xg = nhwc_from(x)
    # 1. centre-tap identity: y must equal x
    w = torch.zeros(cout, cin, 3, 3)
    for c in range(min(cin, cout)):
        w[c, c, 1, 1] = 1.0
    y, _, _ = ops.conv_fwd(xg, ops.pack_conv_weight(w.cuda()), cin, cout, 3, 1, 1, engine=L.ENGINE_TCGEN05)
    torch.cuda.synchronize()
    e = summarize("fwd centre-tap identity", to_nchw(y), F.conv2d(x, w, None, 1, 1))
    if e > 1e-3:
        yy, rr = to_nchw(y), F.conv2d(x, w, None, 1, 1)
        bad = (yy - rr).abs() > 1e-3
        print("  bad fraction %.4f; per-channel bad count (first 16): %s" % (bad.float().mean().item(), bad.sum((0, 2, 3))[:16].tolist()))
        print("  per-row bad count (image 0, first 8 rows): %s" % bad[0].sum((0, 2))[:8].tolist())
        print("  sample got[0,:8,0,0] %s\n  sample ref[0,:8,0,0] %s" % (yy[0, :8, 0, 0].tolist(), rr[0, :8, 0, 0].tolist()))
    # 2. one tap at a time with a random channel-mixing matrix
    m = bf16_round(torch.randn(cout, cin, generator=g) * 0.1)
    for ky in range(3):
        for kx in range(3):
            w = torch.zeros(cout, cin, 3, 3)
            w[:, :, ky, kx] = m
            y, _, _ = ops.conv_fwd(xg, ops.pack_conv_weight(w.cuda()), cin, cout, 3, 1, 1, engine=L.ENGINE_TCGEN05)
            torch.cuda.synchronize()
            summarize("fwd single tap (%d,%d)" % (ky, kx), to_nchw(y), F.conv2d(x, w, None, 1, 1))
    w = bf16_round(torch.randn(cout, cin, 3, 3, generator=g) * 0.05)
    b = torch.randn(cout, generator=g)
    y, _, _ = ops.conv_fwd(xg, ops.pack_conv_weight(w.cuda()), cin, cout, 3, 1, 1, bias=b.cuda(), engine=L.ENGINE_TCGEN05)
    torch.cuda.synchronize()
    summarize("fwd full random + bias", to_nchw(y), F.conv2d(x, w, b, 1, 1))


def diag_dgrad(n=2, cin=64, cout=64, h=32):
    g = torch.Generator().manual_seed(2)
    w = bf16_round(torch.randn(cout, cin, 3, 3, generator=g) * 0.05)
    dy = bf16_round(torch.randn(n, cout, h, h, generator=g))
    xr = torch.zeros(n, cin, h, h, requires_grad=True)
    F.conv2d(xr, w, None, 1, 1).backward(dy)
    dx = ops.conv_dgrad(nhwc_from(dy), ops.pack_conv_weight(w.cuda(), for_dgrad=True), (n, h, h, cin), cin, cout, 3, 1, 1,
                        engine=L.ENGINE_TCGEN05)
    torch.cuda.synchronize()
    summarize("dgrad random", to_nchw(dx), xr.grad)


def diag_wgrad(n=2, cin=64, cout=64, h=32):
    g = torch.Generator().manual_seed(3)
    x = bf16_round(torch.randn(n, cin, h, h, generator=g))
    dy = bf16_round(torch.randn(n, cout, h, h, generator=g))
    wr = torch.zeros(cout, cin, 3, 3, requires_grad=True)
    F.conv2d(x, wr, None, 1, 1).backward(dy)
    dw, _ = ops.conv_wgrad(nhwc_from(x), nhwc_from(dy), cin, cout, 3, 1, 1, engine=L.ENGINE_TCGEN05)
    torch.cuda.synchronize()
    e = summarize("wgrad random %dx%d->%d" % (h, cin, cout), dw, wr.grad)
    if e > 1e-3:
        for ky in range(3):
            for kx in range(3):
                summarize("  wgrad tap (%d,%d)" % (ky, kx), dw[:, :, ky, kx], wr.grad[:, :, ky, kx])
        # hypothesis: transposed channel roles
        summarize("  wgrad vs ref^T (co<->ci)", dw, wr.grad.permute(1, 0, 2, 3).contiguous() if cin == cout else wr.grad)


def diag_match():
    import numpy as np
    from oracle import eval_oracle as EO
    gal, pr, ids = EO.synthetic_gallery(3000, 200, dim=512, seed=4)
    gb = ops.l2norm_bf16(torch.from_numpy(gal).cuda())
    pb = ops.l2norm_bf16(torch.from_numpy(pr).cuda())
    val, idx = ops.cosine_topk(pb, gb, 5)
    torch.cuda.synchronize()
    oval, oidx = EO.cosine_topk(pb.float().cpu().numpy(), gb.float().cpu().numpy(), 5)
    print("matcher: idx equal %s  rank-1 hits %d/%d  max |dval| %.3e" % (
        np.array_equal(idx.cpu().numpy(), oidx), int((idx[:, 0].cpu().numpy() == ids).sum()), len(ids),
        float(np.abs(val.cpu().numpy() - oval).max())), flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "fwd"
    print("device:", torch.cuda.get_device_name(0), flush=True)
    {"fwd": diag_fwd, "dgrad": diag_dgrad, "wgrad": diag_wgrad, "match": diag_match}[what]()
    if what == "fwd":
        diag_fwd(2, 128, 128, 16)
    if what == "wgrad":
        diag_wgrad(2, 128, 128, 16)
