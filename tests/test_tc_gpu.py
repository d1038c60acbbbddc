"""tcgen05 / TMEM / TMA engine parity: implicit-GEMM conv forward, dgrad, wgrad and the fused cosine top-k matcher
against fp32 PyTorch / the numpy oracle on identical bf16-representable inputs."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.util import bf16_round, nhwc_from, rel_err, to_nchw

pytestmark = pytest.mark.gpu

BF16_TOL = 5e-3
F32_TOL = 1e-4

# (n, cin, cout, h): every 3x3 s1 p1 shape of FSRNet at 128x128 input (and the small maps of a 64x64 input)
TC_SHAPES = [
    (2, 64, 64, 128),     # coarse / decoder residual convs (FSRnet.py:79,85) - 87 % of the MACs
    (3, 64, 64, 32),      # encoder residual convs
    (2, 128, 128, 32),    # prior residual convs, hourglass level 2
    (3, 128, 128, 16),    # hourglass level 1
    (5, 128, 128, 8),     # hourglass level 0 (two images per tile, ragged last tile)
    (4, 128, 128, 8),     # the same with whole tiles: statistics fused in the epilogue, two images per tile
    (9, 128, 128, 4),     # 64x64-input hourglass floor (eight images per tile)
    (2, 192, 64, 32),     # decoder conv_input on cat(prior, encoder)
    (1, 64, 64, 64),      # two rows per tile
    (2, 64, 64, 56),      # ResNet_34 layer1 (model/resnet.py:161): 112-pixel tiles (partial, zero-filled for wgrad)
    (2, 128, 128, 28),    # layer2
    (3, 256, 256, 14),    # layer3: 126-pixel tiles, ragged last tile
    (5, 512, 512, 7),     # layer4: two images per tile, two 256-wide column tiles
    (2, 64, 128, 24),     # cin != cout with 128 output channels (weight gradient: dY as the M operand, one K chunk)
    (3, 128, 64, 12),     # ... and with 64 (M = 64 accumulator, two K chunks)
]


def _mk(n, cin, cout, h, seed):
    g = torch.Generator().manual_seed(seed)
    x = bf16_round(torch.randn(n, cin, h, h, generator=g))
    w = bf16_round(torch.randn(cout, cin, 3, 3, generator=g) * 0.05)
    b = torch.randn(cout, generator=g)
    dy = bf16_round(torch.randn(n, cout, h, h, generator=g))
    return x, w, b, dy


@pytest.mark.parametrize("n,cin,cout,h", TC_SHAPES)
def test_tc_conv_fwd(cuda, n, cin, cout, h):
    from crfr_b200 import _lib as L, ops
    assert ops.engine_supported(L.ENGINE_TCGEN05, 0, h, h, cin, cout, 3, 1, 1)
    x, w, b, _ = _mk(n, cin, cout, h, 11)
    ref = F.conv2d(x, w, b, 1, 1)
    y, _, st = ops.conv_fwd(nhwc_from(x), ops.pack_conv_weight(w.cuda()), cin, cout, 3, 1, 1, bias=b.cuda(),
                            engine=L.ENGINE_TCGEN05, want_stats=True)
    torch.cuda.synchronize()
    assert rel_err(to_nchw(y), ref) < BF16_TOL
    yb = to_nchw(y)
    assert rel_err(st[..., 0], yb.mean((2, 3))) < 1e-3
    assert rel_err(st[..., 1], 1.0 / torch.sqrt(yb.var((2, 3), unbiased=False) + 1e-5)) < 1e-3
    # the CUDA-core engine computes the same thing (cross-check used at full size below)
    y2, _, _ = ops.conv_fwd(nhwc_from(x), ops.pack_conv_weight(w.cuda()), cin, cout, 3, 1, 1, bias=b.cuda(),
                            engine=L.ENGINE_DIRECT)
    assert rel_err(y.float(), y2.float()) < BF16_TOL


@pytest.mark.parametrize("n,cin,cout,h", TC_SHAPES)
def test_tc_conv_dgrad(cuda, n, cin, cout, h):
    from crfr_b200 import _lib as L, ops
    assert ops.engine_supported(L.ENGINE_TCGEN05, 1, h, h, cin, cout, 3, 1, 1)
    x, w, _, dy = _mk(n, cin, cout, h, 12)
    xr = x.clone().requires_grad_(True)
    F.conv2d(xr, w, None, 1, 1).backward(dy)
    dx = ops.conv_dgrad(nhwc_from(dy), ops.pack_conv_weight(w.cuda(), for_dgrad=True), (n, h, h, cin), cin, cout, 3, 1, 1,
                        engine=L.ENGINE_TCGEN05)
    torch.cuda.synchronize()
    assert rel_err(to_nchw(dx), xr.grad) < BF16_TOL


@pytest.mark.parametrize("n,cin,cout,h", TC_SHAPES)
def test_tc_conv_wgrad(cuda, n, cin, cout, h):
    from crfr_b200 import _lib as L, ops
    assert ops.engine_supported(L.ENGINE_TCGEN05, 2, h, h, cin, cout, 3, 1, 1)
    x, w, _, dy = _mk(n, cin, cout, h, 13)
    wr = w.clone().requires_grad_(True)
    F.conv2d(x, wr, None, 1, 1).backward(dy)
    dw, _ = ops.conv_wgrad(nhwc_from(x), nhwc_from(dy), cin, cout, 3, 1, 1, engine=L.ENGINE_TCGEN05)
    torch.cuda.synchronize()
    assert rel_err(dw, wr.grad) < F32_TOL


def test_tc_conv_views_and_accumulation(cuda):
    """Channel-slice views (ld > c, the concat-free decoder input) and += semantics of the weight gradient."""
    from crfr_b200 import _lib as L, ops
    n, h = 2, 32
    x, w, b, dy = _mk(n, 64, 64, h, 14)
    wide = torch.randn(n, h, h, 192).to(torch.bfloat16).cuda()
    wide[..., 128:192] = nhwc_from(x)
    view = wide[..., 128:]           # pointer offset 128, ld 192 (ops read shape[3] as ld -> build desc by hand)
    import ctypes as C
    d = L.ConvDesc(n, h, h, 64, 64, 3, 1, 1, h, h, 192, 64, 0)
    y = torch.empty((n, h, h, 64), dtype=torch.bfloat16, device="cuda")
    wp = ops.pack_conv_weight(w.cuda())
    ws = ops.workspace(L.lib().crfr_conv_workspace_bytes(C.byref(d)))
    L.call("crfr_conv_fwd", L.ENGINE_TCGEN05, C.byref(d), view.data_ptr(), wp.data_ptr(), 64, None, y.data_ptr(), None,
           None, 1e-5, ws.data_ptr(), ws.numel(), ops.stream())
    assert rel_err(to_nchw(y), F.conv2d(x, w, None, 1, 1)) < BF16_TOL
    wr = w.clone().requires_grad_(True)
    F.conv2d(x, wr, None, 1, 1).backward(dy)
    dw = torch.ones((64, 64, 3, 3), dtype=torch.float32, device="cuda")
    dyg = nhwc_from(dy)
    for _ in range(2):
        L.call("crfr_conv_wgrad", L.ENGINE_TCGEN05, C.byref(d), view.data_ptr(), dyg.data_ptr(), dw.data_ptr(), None,
               ws.data_ptr(), ws.numel(), ops.stream())
    assert rel_err(dw, 1.0 + 2.0 * wr.grad) < F32_TOL


def test_tc_conv_full_size_against_direct_engine(cuda):
    """BASELINE size (batch 16 of the 64ch 128x128 layer): the two engines agree, and the conv is linear."""
    from crfr_b200 import _lib as L, ops
    g = torch.Generator().manual_seed(15)
    n, c, h = 16, 64, 128
    x = torch.randn(n, h, h, c, generator=g).to(torch.bfloat16).cuda()
    w = bf16_round(torch.randn(c, c, 3, 3, generator=g) * 0.05).cuda()
    wp = ops.pack_conv_weight(w)
    y_tc, _, _ = ops.conv_fwd(x, wp, c, c, 3, 1, 1, engine=L.ENGINE_TCGEN05)
    y_d, _, _ = ops.conv_fwd(x, wp, c, c, 3, 1, 1, engine=L.ENGINE_DIRECT)
    assert rel_err(y_tc.float(), y_d.float()) < BF16_TOL
    y2, _, _ = ops.conv_fwd((x.float() * 2).to(torch.bfloat16), wp, c, c, 3, 1, 1, engine=L.ENGINE_TCGEN05)
    assert rel_err(y2.float(), 2 * y_tc.float()) < BF16_TOL
    dy = torch.randn(n, h, h, c, generator=g).to(torch.bfloat16).cuda()
    dw_tc, _ = ops.conv_wgrad(x, dy, c, c, 3, 1, 1, engine=L.ENGINE_TCGEN05)
    dw_d, _ = ops.conv_wgrad(x, dy, c, c, 3, 1, 1, engine=L.ENGINE_DIRECT)
    assert rel_err(dw_tc, dw_d) < F32_TOL
    dx_tc = ops.conv_dgrad(dy, ops.pack_conv_weight(w, for_dgrad=True), (n, h, h, c), c, c, 3, 1, 1, engine=L.ENGINE_TCGEN05)
    dx_d = ops.conv_dgrad(dy, ops.pack_conv_weight(w, for_dgrad=True), (n, h, h, c), c, c, 3, 1, 1, engine=L.ENGINE_DIRECT)
    assert rel_err(dx_tc.float(), dx_d.float()) < BF16_TOL


@pytest.mark.parametrize("p,g,dim,k", [(64, 1000, 128, 5), (200, 5000, 512, 5), (130, 777, 512, 8), (1, 3, 64, 5)])
def test_matcher_small_against_oracle(cuda, p, g, dim, k):
    """Bit-exact top-k indices vs the fp32 oracle on bf16-rounded unit vectors (planted identities, tie-free)."""
    from crfr_b200 import ops
    from oracle import eval_oracle as EO
    gal, pr, ids = EO.synthetic_gallery(g, p, dim=dim, seed=p + g)
    gb = ops.l2norm_bf16(torch.from_numpy(gal).cuda())
    pb = ops.l2norm_bf16(torch.from_numpy(pr).cuda())
    val, idx = ops.cosine_topk(pb, gb, k)
    torch.cuda.synchronize()
    oval, oidx = EO.cosine_topk(pb.float().cpu().numpy(), gb.float().cpu().numpy(), min(k, g))
    kk = min(k, g)
    assert np.array_equal(idx.cpu().numpy()[:, :kk], oidx)
    assert np.allclose(val.cpu().numpy()[:, :kk], oval, rtol=1e-5, atol=1e-5)
    if g > k:
        assert np.array_equal(idx.cpu().numpy()[:, 0], ids)
    else:
        assert (idx.cpu().numpy()[:, kk:] == -1).all()


def test_matcher_ties_and_sharding(cuda):
    """Duplicate gallery rows tie exactly -> lowest index first; gallery-sharded top-k merges to the unsharded one."""
    from crfr_b200 import ops
    from oracle import eval_oracle as EO
    gal, pr, ids = EO.synthetic_gallery(4096, 96, dim=256, seed=3)
    gal[2000:2010] = gal[5]            # ten exact duplicates of entry 5
    gb = ops.l2norm_bf16(torch.from_numpy(gal).cuda())
    pb = ops.l2norm_bf16(torch.from_numpy(gal[[5]] * 1.0).cuda())
    val, idx = ops.cosine_topk(pb, gb, 5)
    assert idx[0].tolist() == [5, 2000, 2001, 2002, 2003]
    pb = ops.l2norm_bf16(torch.from_numpy(pr).cuda())
    full_v, full_i = ops.cosine_topk(pb, gb, 5)
    parts_v, parts_i = [], []
    for s in range(4):
        v, i = ops.cosine_topk(pb, gb[s * 1024:(s + 1) * 1024], 5, index_base=s * 1024)
        parts_v.append(v); parts_i.append(i)
    mv, mi = ops.topk_merge(torch.stack(parts_v), torch.stack(parts_i), 5)
    assert torch.equal(mi, full_i) and torch.equal(mv, full_v)


@pytest.mark.parametrize("n,h", [(200, 2), (7, 50), (300, 1), (3, 128), (40, 16)])
def test_rowconv_ragged_row_ranges(cuda, n, h):
    """Row-streaming kernels (rowconv.cu / rowwgrad.cu, width 128, 64 -> 64): CTAs that own single rows, whole
    images or several images; forward + fused InstanceNorm statistics, dgrad and wgrad against fp32 PyTorch."""
    from crfr_b200 import _lib as L, ops
    g = torch.Generator().manual_seed(100 + n)
    c, w_ = 64, 128
    x = bf16_round(torch.randn(n, c, h, w_, generator=g))
    w = bf16_round(torch.randn(c, c, 3, 3, generator=g) * 0.05)
    dy = bf16_round(torch.randn(n, c, h, w_, generator=g))
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    ref = F.conv2d(xr, wr, None, 1, 1)
    ref.backward(dy)
    y, _, st = ops.conv_fwd(nhwc_from(x), ops.pack_conv_weight(w.cuda()), c, c, 3, 1, 1, engine=L.ENGINE_TCGEN05,
                            want_stats=True)
    torch.cuda.synchronize()
    assert rel_err(to_nchw(y), ref) < BF16_TOL
    yb = to_nchw(y)
    assert rel_err(st[..., 0], yb.mean((2, 3))) < 1e-3
    assert rel_err(st[..., 1], 1.0 / torch.sqrt(yb.var((2, 3), unbiased=False) + 1e-5)) < 1e-3
    dx = ops.conv_dgrad(nhwc_from(dy), ops.pack_conv_weight(w.cuda(), for_dgrad=True), (n, h, w_, c), c, c, 3, 1, 1,
                        engine=L.ENGINE_TCGEN05)
    assert rel_err(to_nchw(dx), xr.grad) < BF16_TOL
    dw, _ = ops.conv_wgrad(nhwc_from(x), nhwc_from(dy), c, c, 3, 1, 1, engine=L.ENGINE_TCGEN05)
    assert rel_err(dw, wr.grad) < F32_TOL
    # run-to-run determinism of the reductions (fixed-order partials, no float atomics)
    dw2, _ = ops.conv_wgrad(nhwc_from(x), nhwc_from(dy), c, c, 3, 1, 1, engine=L.ENGINE_TCGEN05)
    _, _, st2 = ops.conv_fwd(nhwc_from(x), ops.pack_conv_weight(w.cuda()), c, c, 3, 1, 1, engine=L.ENGINE_TCGEN05,
                             want_stats=True)
    assert torch.equal(dw, dw2) and torch.equal(st, st2)


@pytest.mark.parametrize("n,h", [(2, 1), (2, 2), (6, 3), (300, 1), (40, 16), (8, 50), (2, 128), (34, 128)])
def test_rowconv_pair_kernel_is_bit_identical(cuda, n, h):
    """The cta_group::2 row-streaming kernel (rowconv2.cu: two CTAs walk the same rows of two images, one M = 256 MMA
    per K step for the pair) against the single-CTA kernel on the paired row-block partitioning: 1-row segments, ranges
    that start / end inside an image, TMEM-ring wraps, clusters that span several image pairs.  Same MMA order per
    output element, so forward and dgrad must agree BIT FOR BIT; the statistics differ only in fp32 summation order."""
    from crfr_b200 import _lib as L, ops
    g = torch.Generator().manual_seed(500 + n + h)
    c, w_ = 64, 128
    x = bf16_round(torch.randn(n, c, h, w_, generator=g))
    w = bf16_round(torch.randn(c, c, 3, 3, generator=g) * 0.05)
    b = torch.randn(c, generator=g)
    ref = F.conv2d(x, w, b, 1, 1)
    res = {}
    try:
        for pair in (0, 1):
            L.call("crfr_set_option", b"rowconv_pair", pair)
            y, _, st = ops.conv_fwd(nhwc_from(x), ops.pack_conv_weight(w.cuda()), c, c, 3, 1, 1, bias=b.cuda(),
                                    engine=L.ENGINE_TCGEN05, want_stats=True)
            dx = ops.conv_dgrad(nhwc_from(x), ops.pack_conv_weight(w.cuda(), for_dgrad=True), (n, h, w_, c), c, c, 3, 1, 1,
                                engine=L.ENGINE_TCGEN05)
            torch.cuda.synchronize()
            res[pair] = (y, st, dx)
    finally:
        L.call("crfr_set_option", b"rowconv_pair", 1)
    assert rel_err(to_nchw(res[1][0]), ref) < BF16_TOL
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][2], res[1][2])
    assert rel_err(res[1][1], res[0][1]) < 1e-5


def test_matcher_full_scale_probe_subset(cuda):
    """BASELINE.json configs[3] at full size (10 000 probes x 1 000 000 x 512 gallery, top-5): the scores and indices of a
    probe subset against an fp32 evaluation of the same bf16 vectors chunked over the gallery (utils/eval.py:11: p @ g.T,
    topk).  A random gallery has near-ties at 1e-6, where the fp32 summation order decides: an index may differ from the
    checker's only where the two scores agree to 2e-6; rank 1 (the planted identity, separated by ~0.5) must be exact."""
    from crfr_b200 import ops
    P, G, D, K, SUB = 10000, 1000000, 512, 5, 192
    g = torch.Generator(device="cuda").manual_seed(11)
    gal = torch.randn(G, D, generator=g, device="cuda")
    ids = torch.randint(0, G, (P,), generator=g, device="cuda")
    pr = gal[ids] + 0.6 * torch.randn(P, D, generator=g, device="cuda")
    gb, pb = ops.l2norm_bf16(gal), ops.l2norm_bf16(pr)
    del gal, pr
    val, idx = ops.cosine_topk(pb, gb, K)
    torch.cuda.synchronize()
    sel = torch.linspace(0, P - 1, SUB, device="cuda").long()
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        q = pb[sel].float()
        best_v = torch.full((SUB, K), -2.0, device="cuda")
        best_i = torch.full((SUB, K), -1, dtype=torch.long, device="cuda")
        for lo in range(0, G, 125000):
            s = q @ gb[lo:lo + 125000].float().T
            v, i = s.topk(K, dim=1)
            cv, ci = torch.cat([best_v, v], 1), torch.cat([best_i, i + lo], 1)
            o = cv.argsort(dim=1, descending=True, stable=True)[:, :K]
            best_v, best_i = cv.gather(1, o), ci.gather(1, o)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    ours_v, ours_i = val[sel], idx[sel].long()
    assert torch.equal(ours_i[:, 0], ids[sel]) and torch.equal(best_i[:, 0], ids[sel])
    assert (ours_v - best_v).abs().max().item() < 2e-6
    diff = ours_i != best_i
    assert not diff[:, 0].any()
    # a differing index must be a near-tie: our score for it equals the checker's score at that rank to 2e-6
    assert ((ours_v - best_v).abs()[diff] < 2e-6).all()
    assert diff.float().mean().item() < 0.02
