"""SUPER_RESOLUTION FSRNet variant (SUPER_RESOLUTION/model/FSRnet.py:251-416 of the reference): the four
sub-networks composed from native ops (crfr_b200.functional) against fixtures the reference's own modules produced
(tests/golden/sr_variant.npz, oracle pinned in oracle/make_golden.py) and against fp32 autograd through the oracle."""
import numpy as np
import pytest
import torch

from tests.util import rel_err

pytestmark = pytest.mark.gpu

SLACK = 1.6
CASES = [("coarse", "Coarse_SR_Network", 700), ("encoder", "Fine_SR_Encoder", 701),
         ("prior", "Prior_Estimation_Network", 702), ("decoder", "Fine_SR_Decoder", 703)]


def _build(cls_name, seed):
    from crfr_b200.SUPER_RESOLUTION.model import FSRnet as M
    torch.manual_seed(seed)
    net = getattr(M, cls_name)()
    with torch.no_grad():                        # the same draws as oracle/make_golden.py:golden_sr
        for _, p in net.named_parameters():
            if p.dim() == 1:
                p.add_(0.2 * torch.randn_like(p))
    return net


def _inputs():
    g = torch.Generator().manual_seed(31)
    return torch.randn(2, 3, 64, 64, generator=g), torch.randn(2, 128, 64, 64, generator=g)


@pytest.mark.parametrize("name,cls_name,seed", CASES)
def test_sr_subnetwork_against_reference_golden(cuda, golden_dir, name, cls_name, seed):
    gd = np.load(golden_dir + "/sr_variant.npz")
    net = _build(cls_name, seed).cuda().train()
    x, xd = _inputs()
    with torch.no_grad():
        outs = net((xd if name == "decoder" else x).cuda())
    outs = outs if isinstance(outs, tuple) else (outs,)
    for i, o in enumerate(outs):
        ref = torch.from_numpy(gd["%s_%d_sample" % (name, i)])
        tol = SLACK * float(gd["%s_%d_emu_rel" % (name, i)]) + 1e-2
        assert o.dtype == torch.float32 and rel_err(o[:, :, ::8, ::8], ref) < tol, (name, i, rel_err(o[:, :, ::8, ::8], ref), tol)
        mean, std, norm = gd["%s_%d_mean_std_norm" % (name, i)]
        assert abs(o.norm().item() - norm) < 3e-2 * norm and abs(o.std().item() - std) < 3e-2 * std


def test_sr_encoder_gradients_against_oracle(cuda):
    from oracle import fsrnet_oracle as FO
    from oracle import sr_oracle as SO
    net = _build("Fine_SR_Encoder", 701)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net = net.cuda().train()
    g = torch.Generator().manual_seed(8)
    x = torch.randn(2, 3, 32, 32, generator=g).to(torch.bfloat16).float()
    dy = torch.randn(2, 64, 32, 32, generator=g)
    xg = x.cuda().requires_grad_(True)
    net(xg).backward(dy.cuda())
    res = {}
    for tag, pr in (("ref", FO.FP32), ("emu", FO.Precision("bf16"))):
        leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        xr = x.clone().requires_grad_(True)
        SO.encoder_forward(leaves, xr, "", pr).backward(dy)
        res[tag] = (xr.grad, {k: v.grad for k, v in leaves.items()})
    assert rel_err(xg.grad, res["ref"][0]) < SLACK * rel_err(res["emu"][0], res["ref"][0]) + 2e-2
    for k, p in net.named_parameters():
        r, e = res["ref"][1][k], res["emu"][1][k]
        assert p.grad is not None and rel_err(p.grad, r) < SLACK * rel_err(e, r) + 2e-2, (k, rel_err(p.grad, r), rel_err(e, r))


def test_sr_network_wiring_and_backward(cuda):
    """coarse -> (encoder, prior) -> cat -> decoder end to end: shapes, and a gradient for every parameter."""
    from crfr_b200.SUPER_RESOLUTION.model.FSRnet import SRNetwork
    torch.manual_seed(5)
    net = SRNetwork().cuda().train()
    x = torch.randn(2, 3, 32, 32, device="cuda")
    coarse, out, lm, ps = net(x)
    assert tuple(coarse.shape) == (2, 3, 32, 32) == tuple(out.shape)
    assert tuple(lm.shape) == (2, 68, 32, 32) and tuple(ps.shape) == (2, 13, 32, 32)
    assert float(out.abs().max()) <= 1.0 and float(coarse.abs().max()) <= 1.0          # Tanh heads
    (out.pow(2).mean() + coarse.mean() + lm.mean() + ps.mean()).backward()
    for k, p in net.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), k


def test_reflect_pad_and_tanh_ops(cuda):
    from crfr_b200 import functional as Fn
    g = torch.Generator().manual_seed(1)
    for c, pad in ((64, 1), (64, 3), (3, 3), (3, 1)):
        x = torch.randn(2, c, 9, 11, generator=g).to(torch.bfloat16).float()
        xr = x.clone().requires_grad_(True)
        ref = torch.nn.functional.pad(xr, (pad,) * 4, mode="reflect")
        dy = torch.randn(ref.shape, generator=g).to(torch.bfloat16).float()
        ref.backward(dy)
        xg = x.cuda().requires_grad_(True)
        out = Fn.to_nchw(Fn.reflect_pad(Fn.to_nhwc(xg), pad, c), c)
        out.backward(dy.cuda())
        assert torch.equal(out.detach().cpu(), ref.detach())
        assert rel_err(xg.grad, xr.grad) < 4e-3            # sums of up to four bf16 values, rounded once
    t = torch.randn(1000, generator=g)
    tg = t.cuda().requires_grad_(True)
    y = Fn.tanh(tg)
    y.sum().backward()
    assert rel_err(y, torch.tanh(t)) < 1e-6 and rel_err(tg.grad, 1 - torch.tanh(t) ** 2) < 1e-5
