"""Module-level parity of the native FSRNet program against the reference's own outputs (golden fixture) and the
CPU oracle.

Why the end-to-end bound is statistical.  With bf16 storage this 40-layer InstanceNorm/PReLU network is chaotic at
random init: changing one weight by 1e-7 flips a few bf16 roundings, those flip PReLU signs downstream, and the
result moves by the full quantisation-noise level (outputs ~1-5 %, gradients ~20 %; measured with the oracle's own
bf16 storage model, DESIGN.md "bf16 contract").  No two bf16 implementations - including the reference under
autocast - can agree better than that end to end.  So:
  * every kernel is held to the north_star tolerance on identical inputs in tests/test_kernels_gpu.py and
    tests/test_tc_gpu.py (1e-2 relative for bf16 results - measured ~2e-3 -, 1e-4 for fp32 results);
  * here the whole program must deviate from the fp32 reference no more than the reference algorithm itself does
    when evaluated under the same storage contract (oracle Precision("bf16")), tensor by tensor;
  * paths that share the forward pass (train step vs autograd, where backward is linear) must agree to 2e-2.
"""
import os

import numpy as np
import pytest
import torch

from tests.util import rel_err

pytestmark = pytest.mark.gpu

SLACK = 1.6             # allowed ratio between our deviation and the bf16-emulated oracle's deviation
LOSS_TOL = 1.5e-2       # north_star 1e-2 + the measured 1 % quantisation shift of the dominant landmark term


def _net(engine):
    from crfr_b200.model.FSRnet import OverallNetwork, weights_init
    torch.manual_seed(1234)
    net = OverallNetwork()
    net.apply(weights_init)
    net.engine = engine
    return net.cuda().train()


def _loss(outs, hr, hm, lbl, b):
    from crfr_b200.loss import CrossEntropyLoss2d, MSELoss_Landmark, MSELossFunc
    mse, lmk, ce = MSELossFunc(), MSELoss_Landmark(), CrossEntropyLoss2d()
    parts = (mse(outs[1], hr), mse(outs[0], hr), lmk(outs[2], hm), ce(outs[3], lbl))
    return (5. * parts[0] + 5. * parts[1] + parts[2] + parts[3]) / (2.0 * b), parts


def _flat(grads, names):
    return torch.cat([grads[k].double().reshape(-1).cpu() for k in names])


@pytest.mark.parametrize("engine_name", ["auto", "direct"])
def test_overall_network_against_reference(cuda, golden_dir, engine_name):
    from crfr_b200 import _lib as L
    from oracle import fsrnet_oracle as FO
    g = np.load(os.path.join(golden_dir, "fsrnet_small.npz"))       # produced by the reference's own modules
    net = _net(L.ENGINE_AUTO if engine_name == "auto" else L.ENGINE_DIRECT)
    x, hr, lbl, hm = FO.synthetic_batch(2, 64)
    sd = FO.build_fsrnet_state_dict(1234)
    ref = FO.fsrnet_loss_and_grads(sd, x, hr, hm, lbl)                       # fp32 restatement (== golden)
    emu = FO.fsrnet_loss_and_grads(sd, x, hr, hm, lbl, precision="bf16")     # same algorithm, bf16 storage contract

    outs = net(x.cuda())
    for i, name in enumerate(("coarse", "out", "landmark", "parsing")):
        o = outs[i]
        assert o.dtype == torch.float32 and tuple(o.shape) == g[name].shape and torch.isfinite(o).all()
        ours = rel_err(o, torch.from_numpy(g[name]))
        bound = SLACK * rel_err(emu[0][i], ref[0][i]) + 1e-3
        assert ours < bound, (name, ours, bound)
    total, parts = _loss(outs, hr.cuda(), hm.cuda(), lbl.cuda(), 2)
    np.testing.assert_allclose(total.item(), g["total"], rtol=LOSS_TOL)
    np.testing.assert_allclose([p.item() for p in parts], g["parts"], rtol=5e-2)
    total.backward()

    gnorm = float(g["global_grad_norm"])
    ours_g, names = {}, []
    for k, p in net.named_parameters():
        if FO.fsrnet_dead_param(k):
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
        if k in FO.FSRNET_NULL_GRAD:            # conv bias in front of an InstanceNorm: mathematically zero
            assert p.grad.norm().item() < 1e-3 * gnorm, k
            continue
        ours_g[k] = p.grad
        names.append(k)
        e_ours, e_emu = rel_err(p.grad, ref[3][k]), rel_err(emu[3][k], ref[3][k])
        assert e_ours < SLACK * e_emu + 2e-2, (k, e_ours, e_emu)
    a, b, c = _flat(ours_g, names), _flat(ref[3], names), _flat(emu[3], names)
    cos_ours = float(a @ b / (a.norm() * b.norm()))
    cos_emu = float(c @ b / (c.norm() * b.norm()))
    assert cos_ours > 0.95 and cos_ours > cos_emu - 0.02, (cos_ours, cos_emu)
    assert abs(a.norm().item() - gnorm) < 5e-2 * gnorm


def test_train_step_matches_module_path(cuda):
    """crfr_fsrnet_train_step (fused losses + backward) == forward + drop-in loss modules + autograd: the forward is
    shared bit for bit and the backward is linear in the loss gradients, so this comparison is well conditioned."""
    import ctypes as C
    from crfr_b200 import _lib as L, ops
    from crfr_b200.model import FSRnet as M
    from oracle import fsrnet_oracle as FO
    net = _net(L.ENGINE_AUTO)
    x, hr, lbl, hm = (t.cuda() for t in FO.synthetic_batch(2, 64, seed=5))
    outs = net(x)
    total, parts = _loss(outs, hr, hm, lbl, 2)
    total.backward()
    params = net.ordered_parameters()
    grads = [torch.zeros_like(p) for p in params]
    outs2 = M.alloc_outputs(x)
    io = M._io(x, outs2, (hr, hm, lbl.contiguous()), loss_div=4.0, w_pix=5.0)
    ws = torch.empty(L.lib().crfr_fsrnet_workspace_bytes(2, 64, 1), dtype=torch.uint8, device="cuda")
    losses = torch.zeros(5, device="cuda")
    pt, gt = M._ParamTable([p.detach() for p in params]), M._ParamTable(grads)
    L.call("crfr_fsrnet_train_step", L.ENGINE_AUTO, pt.arr, gt.arr, C.byref(io), losses.data_ptr(), ws.data_ptr(),
           ws.numel(), ops.stream())
    torch.cuda.synchronize()
    assert abs(losses[0].item() - total.item()) < 1e-4 * abs(total.item())
    np.testing.assert_allclose(losses[1:].cpu().numpy(), [p.item() for p in parts], rtol=1e-4)
    for a, b in zip(outs, outs2):
        assert torch.equal(a, b)                     # deterministic forward
    for (k, p), gr in zip(net.named_parameters(), grads):
        if p.grad is None or FO.fsrnet_dead_param(k) or k in FO.FSRNET_NULL_GRAD:
            continue
        # (the 3-element image-conv bias gradients are summed in fp32 from the fp32 loss gradient on both paths;
        # conv_mid.bias also sums the two stems' bf16 input gradients, a sign-cancelling sum that amplifies their rounding
        # noise - 1.7e-2 in the oracle's own bf16 evaluation - and has its own bound as in test_fsrnet_forced_gpu.py)
        # The two paths hand the backward loss gradients that differ in their last bf16 bit (fp32 autograd of the loss
        # modules against the fused loss kernels): two bf16 evaluations of the same backward, bounded like the other
        # implementation-equivalence comparisons (REASSOC_TOL_DEEP = 2.5e-2 in test_fsrnet_forced_gpu.py; measured worst
        # 2.03e-2 on an InstanceNorm bias of the coarse network, depending on the random initialisation)
        assert rel_err(gr, p.grad) < (1e-1 if k == "_coarse_sr_network.conv_mid.bias" else 2.5e-2), k


def test_chunked_accumulation_equals_full_batch(cuda):
    """InstanceNorm only -> samples are independent: two chunks of 2 with loss_div = 2*G*G/c accumulate the same
    gradient as one batch of 4 (the trainer's micro-batching and the DP loss scaling rely on this)."""
    import ctypes as C
    from crfr_b200 import _lib as L, ops
    from crfr_b200.model import FSRnet as M
    from crfr_b200.trainer import chunk_loss_div
    from oracle import fsrnet_oracle as FO
    net = _net(L.ENGINE_AUTO)
    params = net.ordered_parameters()
    x, hr, lbl, hm = (t.cuda() for t in FO.synthetic_batch(4, 64, seed=8))
    pt = M._ParamTable([p.detach() for p in params])

    def run(slices, G):
        grads = [torch.zeros_like(p) for p in params]
        gt = M._ParamTable(grads)
        tot = 0.0
        for s in slices:
            xs = x[s].contiguous()
            outs = M.alloc_outputs(xs)
            io = M._io(xs, outs, (hr[s].contiguous(), hm[s].contiguous(), lbl[s].contiguous()),
                       loss_div=chunk_loss_div(G, xs.shape[0]), w_pix=5.0)
            ws = torch.empty(L.lib().crfr_fsrnet_workspace_bytes(xs.shape[0], 64, 1), dtype=torch.uint8, device="cuda")
            losses = torch.zeros(5, device="cuda")
            L.call("crfr_fsrnet_train_step", L.ENGINE_AUTO, pt.arr, gt.arr, C.byref(io), losses.data_ptr(),
                   ws.data_ptr(), ws.numel(), ops.stream())
            tot += losses[0].item()
        return tot, grads

    t_full, g_full = run([slice(0, 4)], 4)
    t_two, g_two = run([slice(0, 2), slice(2, 4)], 4)
    assert abs(t_full - t_two) < 1e-4 * abs(t_full)
    for (k, _), a, b in zip(net.named_parameters(), g_two, g_full):
        if FO.fsrnet_dead_param(k) or k in FO.FSRNET_NULL_GRAD:
            continue
        assert rel_err(a, b) < 2e-2, k


def test_rejects_cpu_and_bad_shapes(cuda):
    net = _net(0)
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 3, 64, 64))
    with pytest.raises(ValueError):
        net(torch.zeros(1, 3, 60, 60, device="cuda"))


def test_trainer_lanes_match_single_stream(cuda):
    """FSRNetTrainer with concurrent lanes (chunks on independent streams / workspaces / gradient arenas) produces the
    same gradients, losses and parameter update as the single-stream trainer."""
    from crfr_b200 import _lib as L
    from crfr_b200.trainer import FSRNetTrainer
    from oracle import fsrnet_oracle as FO
    x, hr, lbl, hm = (t.cuda() for t in FO.synthetic_batch(4, 64, seed=9))
    results = []
    for lanes, chunk, graph in ((1, 4, False), (2, 2, False), (3, 1, False), (1, 4, True), (2, 2, True)):
        tr = FSRNetTrainer(_net(L.ENGINE_AUTO), lr=1e-3, chunk=chunk, lanes=lanes, use_graph=graph)
        losses = tr.step(x, hr, hm, lbl).clone()
        if graph:                                 # second call = pure replay from the static buffers
            tr2 = FSRNetTrainer(_net(L.ENGINE_AUTO), lr=0.0, chunk=chunk, lanes=lanes, use_graph=True)
            tr2.step(x * 0.5, hr, hm, lbl)        # captured on other data (lr = 0: the parameters stay put) ...
            l2 = tr2.step(x, hr, hm, lbl).clone()  # ... and replayed on ours
            torch.cuda.synchronize()
            assert rel_err(tr2.flat_g, results[0][0]) < 2e-2 and rel_err(l2, results[0][2]) < 1e-4
        torch.cuda.synchronize()
        results.append((tr.flat_g.clone(), tr.flat_p.clone(), losses))
    g0, p0, l0 = results[0]
    for g, p, l in results[1:]:
        assert rel_err(l, l0) < 1e-4
        assert rel_err(g, g0) < 2e-2            # chunked bf16 rounding differs slightly from the full batch
        assert rel_err(p, p0) < 1e-3
