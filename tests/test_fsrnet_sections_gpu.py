"""The four FSRNet sub-networks on their own (Course_SR_Network.forward(x) -> (out, out_coarse),
Fine_SR_Encoder.forward(x) -> out, Prior_Estimation_Network.forward(x) -> (out, landmark_out, parsing_out),
Fine_SR_Decoder.forward(x) -> out; ref model/FSRnet.py:328-340, 359-379, 408-426, 448-459): each is a native
sub-program with its own autograd node, checked against the oracle's restatement of the same sub-network and - chained
through PyTorch autograd exactly as the reference's OverallNetwork_GAN.forward chains them (:538-541) - against the
fused whole-network program."""
import pytest
import torch

from tests.util import rel_err

pytestmark = pytest.mark.gpu


def _net():
    from crfr_b200.model.FSRnet import OverallNetwork, weights_init
    torch.manual_seed(1234)
    net = OverallNetwork()
    net.apply(weights_init)
    return net.cuda().train()


def _sub_sd(sd, prefix):
    return {k: v for k, v in sd.items() if k.startswith(prefix)}


@pytest.mark.parametrize("which", ["coarse", "encoder", "prior", "decoder"])
def test_subnetwork_forward_against_oracle(cuda, which):
    """Standalone module (own state_dict, reference key names without the OverallNetwork prefix) against the oracle's
    restatement, free running at 64 x 64: the deviation is bounded by what the oracle's own bf16-storage evaluation
    deviates from fp32 (tests/test_fsrnet_gpu.py explains the yardstick)."""
    from crfr_b200.model import FSRnet as M
    from oracle import fsrnet_oracle as FO
    sd = FO.build_fsrnet_state_dict(1234)
    cls, prefix, fn = {"coarse": (M.Course_SR_Network, "_coarse_sr_network.", FO.coarse_forward),
                       "encoder": (M.Fine_SR_Encoder, "_fine_sr_encoder.", FO.encoder_forward),
                       "prior": (M.Prior_Estimation_Network, "_prior_estimation_network.", FO.prior_forward),
                       "decoder": (M.Fine_SR_Decoder, "_fine_sr_decoder.", FO.decoder_forward)}[which]
    mod = cls()
    mod.load_state_dict({k[len(prefix):]: v for k, v in _sub_sd(sd, prefix).items()})
    mod = mod.cuda()
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 192, 16, 16, generator=g) if which == "decoder" else torch.randn(2, 3, 64, 64, generator=g)
    with torch.no_grad():
        ours = mod(x.cuda())
    ref = fn(sd, x, prefix)
    emu = fn(sd, x, prefix, FO.Precision("bf16"))
    ours = ours if isinstance(ours, tuple) else (ours,)
    ref = ref if isinstance(ref, tuple) else (ref,)
    emu = emu if isinstance(emu, tuple) else (emu,)
    assert len(ours) == len(ref)
    for a, b, e in zip(ours, ref, emu):
        assert a.dtype == torch.float32 and tuple(a.shape) == tuple(b.shape)
        assert rel_err(a, b) < 1.6 * rel_err(e, b) + 2e-3, (which, rel_err(a, b), rel_err(e, b))


@pytest.mark.parametrize("size,batch", [(64, 2), (128, 2)])
def test_chained_subnetworks_equal_fused_network(cuda, size, batch):
    """coarse -> (encoder, prior) -> cat -> decoder through four autograd nodes == the fused crfr_fsrnet_forward /
    _backward: the same kernels on the same data, so the forward is bit-identical; the backward differs only in where the
    three gradients of the coarse image are summed (fp32 by autograd here, bf16 slots there)."""
    from crfr_b200.loss import CrossEntropyLoss2d, MSELoss_Landmark, MSELossFunc
    from oracle import fsrnet_oracle as FO
    x, hr, lbl, hm = (t.cuda() for t in FO.synthetic_batch(batch, size, seed=21))

    def loss_of(outs):
        coarse, out, lm, ps = outs
        return (5. * MSELossFunc()(out, hr) + 5. * MSELossFunc()(coarse, hr) + MSELoss_Landmark()(lm, hm)
                + CrossEntropyLoss2d()(ps, lbl)) / (2.0 * batch)

    net = _net()
    fused = net(x)
    loss_of(fused).backward()
    g_fused = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
    net.zero_grad(set_to_none=True)

    _, coarse = net._coarse_sr_network(x)
    enc = net._fine_sr_encoder(coarse)
    pe, lm, ps = net._prior_estimation_network(coarse)
    out = net._fine_sr_decoder(torch.cat((pe, enc), 1))
    chained = (coarse, out, lm, ps)
    for a, b, name in zip(chained, fused, ("coarse", "out", "landmark", "parsing")):
        assert torch.equal(a, b), name
    loss_of(chained).backward()
    for k, p in net.named_parameters():
        if FO.fsrnet_dead_param(k):
            assert p.grad is None, k
            continue
        assert p.grad is not None, k
        if k in FO.FSRNET_NULL_GRAD:
            continue
        tol = 1e-1 if k == "_coarse_sr_network.conv_mid.bias" else 2e-2
        assert rel_err(p.grad, g_fused[k]) < tol, (k, rel_err(p.grad, g_fused[k]))


def test_subnetwork_input_gradient(cuda):
    """dx of a sub-network (needed when sub-networks are chained, FSR_main.py:146,158-159 addresses them by name):
    the encoder's input gradient against fp32 autograd through the oracle's restatement on bf16-representable input."""
    from crfr_b200.model import FSRnet as M
    from oracle import fsrnet_oracle as FO
    sd = FO.build_fsrnet_state_dict(1234)
    prefix = "_fine_sr_encoder."
    mod = M.Fine_SR_Encoder()
    mod.load_state_dict({k[len(prefix):]: v for k, v in _sub_sd(sd, prefix).items()})
    mod = mod.cuda()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 64, 64, generator=g).to(torch.bfloat16).float()
    dy = torch.randn(2, 64, 16, 16, generator=g)
    xg = x.cuda().requires_grad_(True)
    mod(xg).backward(dy.cuda())
    xr = x.clone().requires_grad_(True)
    emu_x = x.clone().requires_grad_(True)
    FO.encoder_forward(sd, xr, prefix).backward(dy)
    FO.encoder_forward(sd, emu_x, prefix, FO.Precision("bf16")).backward(dy)
    assert xg.grad is not None and tuple(xg.grad.shape) == tuple(x.shape)
    assert rel_err(xg.grad, xr.grad) < 1.6 * rel_err(emu_x.grad, xr.grad) + 2e-2
