"""Discriminator and OverallNetwork_GAN (model/FSRnet.py:461-486, 512-545 of the reference): the discriminator is composed
from single native ops (crfr_b200.functional: tcgen05 conv / linear GEMMs, fused BatchNorm + PReLU passes), the four
sub-networks are the native sub-programs, chained through autograd as the reference chains its modules."""
import numpy as np
import pytest
import torch

from tests.util import rel_err

pytestmark = pytest.mark.gpu

SLACK = 1.6


def test_discriminator_forward_backward_against_oracle(cuda):
    from crfr_b200.model.FSRnet import Discriminator, weights_init
    from oracle import fsrnet_oracle as FO
    torch.manual_seed(5)
    d = Discriminator(spatial=16)
    d.apply(weights_init)
    with torch.no_grad():                                   # non-trivial affine parameters
        for p in (d.bn_mid.weight, d.bn_mid.bias, d.bn_end.weight, d.bn_end.bias, d.relu.weight, d.conv_input.bias):
            p.add_(0.3 * torch.randn_like(p))
    sd = {k: v.detach().clone() for k, v in d.state_dict().items() if "running_" not in k and "num_batches" not in k}
    assert [k for k, _ in FO.discriminator_param_shapes(16)] == list(sd.keys())
    d = d.cuda().train()
    g = torch.Generator().manual_seed(6)
    x = torch.randn(8, 192, 16, 16, generator=g).to(torch.bfloat16).float()
    dy = torch.randn(8, 512, generator=g)
    xg = x.cuda().requires_grad_(True)
    out = d(xg)
    out.backward(dy.cuda())
    res = {}
    for tag, pr in (("ref", FO.FP32), ("emu", FO.Precision("bf16"))):
        leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        xr = x.clone().requires_grad_(True)
        o = FO.discriminator_forward(leaves, xr, "", pr)
        o.backward(dy)
        res[tag] = (o.detach(), xr.grad, {k: v.grad for k, v in leaves.items()})
    ref, emu = res["ref"], res["emu"]
    assert out.shape == (8, 512) and rel_err(out, ref[0]) < SLACK * rel_err(emu[0], ref[0]) + 1e-2
    assert rel_err(xg.grad, ref[1]) < SLACK * rel_err(emu[1], ref[1]) + 2e-2
    for k, p in d.named_parameters():
        if k.startswith("residual."):                       # constructed, never called (:481-482)
            assert p.grad is None, k
            continue
        if k in ("conv_input.bias", "fc.bias"):             # cancelled by the BatchNorm that follows: ~0
            continue
        assert rel_err(p.grad, ref[2][k]) < SLACK * rel_err(emu[2][k], ref[2][k]) + 2e-2, k
    # running statistics: bn_mid is applied twice per forward, bn_end once (nn.BatchNorm semantics)
    assert int(d.bn_mid.num_batches_tracked) == 2 and int(d.bn_end.num_batches_tracked) == 1
    d.eval()
    with torch.no_grad():
        assert torch.isfinite(d(x.cuda())).all()


def test_overall_network_gan_against_reference_golden(cuda, golden_dir):
    """forward(lr, hr) at the reference's own size (224 x 224, fc 64*56*56) against the embeddings the reference's modules
    produced (tests/golden/gan224.npz, oracle pinned in oracle/make_golden.py); the bound is the deviation of the oracle's
    own bf16-storage evaluation stored with the fixture."""
    from crfr_b200.model.FSRnet import OverallNetwork_GAN, weights_init
    gd = np.load(golden_dir + "/gan224.npz")
    torch.manual_seed(4242)
    net = OverallNetwork_GAN()
    net.apply(weights_init)
    assert [k for k in net.state_dict().keys() if "running_" not in k and "num_batches" not in k] == gd["keys"].tolist()
    net = net.cuda().train()
    g = torch.Generator().manual_seed(99)
    lr, hr = torch.randn(4, 3, 224, 224, generator=g), torch.randn(4, 3, 224, 224, generator=g)
    sr, coarse, lm, ps, e1, e2 = net(lr.cuda(), hr.cuda())
    emu = gd["emu_rel"]
    assert tuple(sr.shape) == (4, 3, 224, 224) and tuple(lm.shape) == (4, 97, 56, 56) and tuple(e1.shape) == (4, 512)
    assert rel_err(e1, torch.from_numpy(gd["embedding1"])) < SLACK * emu[4] + 2e-2
    assert rel_err(e2, torch.from_numpy(gd["embedding2"])) < SLACK * emu[5] + 2e-2
    assert abs(sr.mean().item() - float(gd["sr_mean"])) < 3e-2 and abs(sr.std().item() - float(gd["sr_std"])) < 5e-2 * float(gd["sr_std"])
    # one backward through all five modules: every live parameter receives a finite gradient
    ((e1 - e2.detach()).pow(2).mean() + sr.pow(2).mean() + lm.mean() + ps.mean()).backward()
    for k, p in net.named_parameters():
        dead = (".bn_end." in k and "_discriminator" not in k) or ".residual_next." in k or \
            k.startswith("_fine_sr_encoder.conv_mid") or ".instance_norm." in k or k.startswith("_discriminator.residual.")
        if dead:
            assert p.grad is None, k
        else:
            assert p.grad is not None and torch.isfinite(p.grad).all(), k
