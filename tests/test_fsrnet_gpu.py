"""Module-level parity of the native FSRNet program: outputs, losses and every parameter gradient against the
golden fixture produced by the reference's own modules, and against the CPU oracle on bf16-rounded weights."""
import os

import numpy as np
import pytest
import torch

from tests.util import bf16_round, rel_err

pytestmark = pytest.mark.gpu

OUT_TOL = 1e-2          # north_star: 1e-2 relative for bf16 outputs / losses / gradients (norm-wise per tensor)
GRAD_TOL = 5e-2         # per-tensor gradient bound against the fp32 reference (depth ~40 conv+IN layers in bf16)


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


@pytest.mark.parametrize("engine_name", ["auto", "direct"])
def test_overall_network_against_golden(cuda, golden_dir, engine_name):
    from crfr_b200 import _lib as L
    from oracle import fsrnet_oracle as FO
    g = np.load(os.path.join(golden_dir, "fsrnet_small.npz"))
    net = _net(L.ENGINE_AUTO if engine_name == "auto" else L.ENGINE_DIRECT)
    x, hr, lbl, hm = FO.synthetic_batch(2, 64)
    outs = net(x.cuda())
    for name, o in zip(("coarse", "out", "landmark", "parsing"), outs):
        assert o.dtype == torch.float32 and tuple(o.shape) == g[name].shape
        assert rel_err(o, torch.from_numpy(g[name])) < OUT_TOL, name
    total, parts = _loss(outs, hr.cuda(), hm.cuda(), lbl.cuda(), 2)
    np.testing.assert_allclose([p.item() for p in parts], g["parts"], rtol=OUT_TOL)
    np.testing.assert_allclose(total.item(), g["total"], rtol=OUT_TOL)
    total.backward()
    norms = dict(zip([str(n) for n in g["grad_names"]], g["grad_norms"]))
    gnorm = float(g["global_grad_norm"])
    worst = 0.0
    sq = 0.0
    for k, p in net.named_parameters():
        if FO.fsrnet_dead_param(k):
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
        if k in FO.FSRNET_NULL_GRAD:
            assert p.grad.norm().item() < 1e-3 * gnorm, k
            continue
        sq += p.grad.double().norm().item() ** 2
        assert abs(p.grad.norm().item() - norms[k]) < GRAD_TOL * norms[k] + 1e-7 * gnorm, (k, p.grad.norm().item(), norms[k])
        if "grad:" + k in g.files:
            e = rel_err(p.grad, torch.from_numpy(g["grad:" + k]))
            worst = max(worst, e)
            assert e < GRAD_TOL, (k, e)
    assert abs(np.sqrt(sq) - gnorm) < OUT_TOL * gnorm


def test_overall_network_against_bf16_oracle(cuda):
    """Same comparison with the oracle run on bf16-rounded conv weights (isolates activation rounding)."""
    from crfr_b200 import _lib as L
    from oracle import fsrnet_oracle as FO
    net = _net(L.ENGINE_AUTO)
    sd = {k: (bf16_round(v) if v.dim() == 4 else v.clone()) for k, v in FO.build_fsrnet_state_dict(1234).items()}
    x, hr, lbl, hm = FO.synthetic_batch(2, 64, seed=99)
    x = bf16_round(x)
    o_outs, o_total, o_parts, gd = FO.fsrnet_loss_and_grads(sd, x, hr, hm, lbl)
    outs = net(x.cuda())
    for a, b in zip(outs, o_outs):
        assert rel_err(a, b) < OUT_TOL
    total, _ = _loss(outs, hr.cuda(), hm.cuda(), lbl.cuda(), 2)
    assert abs(total.item() - o_total.item()) < OUT_TOL * abs(o_total.item())
    total.backward()
    for k, p in net.named_parameters():
        if gd[k] is None or k in FO.FSRNET_NULL_GRAD:
            continue
        assert rel_err(p.grad, gd[k]) < GRAD_TOL, k


def test_train_step_matches_module_path(cuda):
    """crfr_fsrnet_train_step (fused losses + backward) == forward + drop-in loss modules + autograd."""
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
    for a, b in zip(outs, outs2):
        assert rel_err(b, a) < 1e-5
    for (k, p), gr in zip(net.named_parameters(), grads):
        if p.grad is None or FO.fsrnet_dead_param(k) or k in FO.FSRNET_NULL_GRAD:
            continue
        assert rel_err(gr, p.grad) < 2e-2, k      # loss gradients enter in bf16 on one path, via fp32 torch on the other


def test_rejects_cpu_and_bad_shapes(cuda):
    net = _net(0)
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 3, 64, 64))
    with pytest.raises(ValueError):
        net(torch.zeros(1, 3, 60, 60, device="cuda"))
