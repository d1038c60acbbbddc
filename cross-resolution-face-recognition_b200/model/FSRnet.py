"""Drop-in FSRNet modules backed by the native sm_100a network program.

Mirror of the reference's operator interface for this path (model/FSRnet.py of the reference): same class names,
constructor arguments, sub-module / parameter names (202 identical ``state_dict`` keys, including the dead ones:
``bn_end``, ``residual_next.*``, the encoder's inherited ``conv_mid`` and the decoder's ``instance_norm``), same
construction order (so ``torch.manual_seed(s); OverallNetwork(); apply(weights_init)`` draws identical weights) and
the same ``forward`` signature.  The torch layer objects below are parameter containers only: ``forward`` never
calls them.  ``OverallNetwork.forward`` runs the whole network in ``crfr_fsrnet_forward`` and registers one autograd
node whose backward is ``crfr_fsrnet_backward``; there is no eager / CPU fallback.

Wiring: the reference's ``OverallNetwork.forward`` (model/FSRnet.py:497-508) feeds the 64-channel coarse feature into
3-channel stems and raises; the wiring that runs is ``OverallNetwork_GAN``'s (:538-541): the encoder and the prior
network consume the 3-channel coarse image.  That is what is implemented.
"""
import ctypes as C

import torch
import torch.nn as nn

from .. import _lib as L
from .. import ops


def _conv(cin, cout, k, stride=1, pad=0, bias=True):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=pad, bias=bias)


class _Residual_Block(nn.Module):
    """ref: model/FSRnet.py:75-98 (conv3x3 -> IN -> PReLU -> conv3x3 -> IN -> +x -> PReLU)."""

    def __init__(self, out_channels, in_channels=64):
        super().__init__()
        self.conv1 = _conv(in_channels, out_channels, 3, 1, 1, bias=False)
        self.in1 = nn.InstanceNorm2d(out_channels, affine=True)
        self.relu = nn.PReLU(out_channels)
        self.conv2 = _conv(out_channels, out_channels, 3, 1, 1, bias=False)
        self.in2 = nn.InstanceNorm2d(out_channels, affine=True)
        self.relu_out = nn.PReLU(out_channels)


class BasicBlock(nn.Module):
    """ref: model/FSRnet.py:105-135 (hourglass flavour: 128 channels hard-coded, non-affine IN, one shared PReLU)."""
    expansion = 2

    def __init__(self, inplanes=128, planes=128, stride=1, downsample=None):
        super().__init__()
        self.conv1 = _conv(128, 128, 3, stride, 1, bias=False)
        self.bn1 = nn.InstanceNorm2d(planes * 2)
        self.relu = nn.PReLU(128)
        self.conv2 = _conv(planes * 2, planes * 2, 3, 1, 1, bias=False)
        self.bn2 = nn.InstanceNorm2d(planes * 2)
        self.downsample = downsample
        self.stride = stride


class Hourglass(nn.Module):
    """ref: model/FSRnet.py:176-215: hg[d] holds 3 (4 at d == 0) Sequentials of ``num_blocks`` blocks."""

    def __init__(self, block, num_blocks, planes, depth):
        super().__init__()
        self.depth = depth
        self.block = block
        self.hg = nn.ModuleList(
            nn.ModuleList(nn.Sequential(*[block(planes * block.expansion, planes) for _ in range(num_blocks)])
                          for _ in range(4 if d == 0 else 3))
            for d in range(depth))


def _stack(n, channels, in_channels=None):
    return nn.Sequential(*[_Residual_Block(channels, channels if in_channels is None else in_channels)
                           for _ in range(n)])


def _subnet_forward_unavailable(name):
    raise NotImplementedError(
        "%s.forward on its own is not part of the native hot path yet; call OverallNetwork.forward (the sub-networks "
        "run fused inside crfr_fsrnet_forward)" % name)


class Course_SR_Network(nn.Module):
    """ref: model/FSRnet.py:308-340; forward(x) -> (feat64, coarse3)."""

    def __init__(self):
        super().__init__()
        self.conv_input = _conv(3, 64, 3, 1, 1)
        self.relu = nn.PReLU(64)
        self.residual = _stack(3, 64)
        self.dropout = nn.Dropout2d(p=0.5, inplace=True)
        self.conv_mid = _conv(64, 3, 3, 1, 1)
        self.bn_mid = nn.InstanceNorm2d(64, affine=True)
        self.bn_end = nn.InstanceNorm2d(3, affine=True)

    def forward(self, x):
        _subnet_forward_unavailable(type(self).__name__)


class Fine_SR_Encoder(Course_SR_Network):
    """ref: model/FSRnet.py:342-379; forward(x) -> feat64 at 1/4 resolution."""

    def __init__(self):
        super().__init__()
        self.conv_input = _conv(3, 64, 7, 4, 3)
        self.relu = nn.PReLU(64)
        self.bn_mid = nn.InstanceNorm2d(64, affine=True)
        self.residual = _stack(3, 64)
        self.conv_end = _conv(64, 64, 3, 1, 1)


class Prior_Estimation_Network(nn.Module):
    """ref: model/FSRnet.py:381-426; forward(x) -> (feat128, landmark97, parsing11) at 1/4 resolution."""

    def __init__(self):
        super().__init__()
        self.conv = _conv(3, 128, 7, 4, 3)
        self.bn = nn.InstanceNorm2d(128, affine=True)
        self.relu = nn.PReLU(128)
        self.residual = _stack(3, 128, 128)
        self.residual_next = _stack(3, 128, 128)
        self.hg = Hourglass(planes=64, depth=2, block=BasicBlock, num_blocks=2)
        self.dropout = nn.Dropout2d(p=0.5, inplace=True)
        self.fc = _conv(128, 11, 1)
        self.fc_landmark = _conv(128, 97, 1)

    def forward(self, x):
        _subnet_forward_unavailable(type(self).__name__)


class Fine_SR_Decoder(nn.Module):
    """ref: model/FSRnet.py:428-459; forward(x192) -> sr3 at 4x resolution."""

    def __init__(self):
        super().__init__()
        self.conv_input = _conv(192, 64, 3, 1, 1)
        self.relu = nn.PReLU(64)
        self.bn_mid = nn.InstanceNorm2d(64, affine=True)
        self.deconv = nn.ConvTranspose2d(64, 64, kernel_size=7, stride=4, bias=True, padding=2, output_padding=1)
        self.residual = _stack(3, 64)
        self.dropout = nn.Dropout2d(p=0.5, inplace=True)
        self.conv_out = _conv(64, 3, 3, 1, 1)
        self.instance_norm = nn.InstanceNorm2d(3, affine=True)

    def forward(self, x):
        _subnet_forward_unavailable(type(self).__name__)


def weights_init(m):
    """ref: FSR_main.py:38-58 (xavier-uniform Conv2d weights, zero conv biases); usable with ``model.apply``."""
    for each in m.modules():
        if isinstance(each, nn.Conv2d):
            nn.init.xavier_uniform_(each.weight.data)
            if each.bias is not None:
                each.bias.data.zero_()
        elif isinstance(each, nn.BatchNorm2d):
            each.weight.data.fill_(1)
            each.bias.data.zero_()
        elif isinstance(each, nn.Linear):
            nn.init.xavier_uniform_(each.weight.data)
            each.bias.data.zero_()


class _ParamTable:
    """The 202 parameter (and gradient) device pointers in state_dict order, as a C array of void*."""

    def __init__(self, tensors):
        self.arr = (C.c_void_p * L.FSRNET_NPARAMS)(*[None if t is None else t.data_ptr() for t in tensors])
        self.keep = tensors


def _io(x, outs, targets=None, loss_div=1.0, w_pix=5.0):
    b, _, h, _ = x.shape
    io = L.FsrnetIO()
    io.batch, io.size = b, h
    io.x = x.data_ptr()
    io.coarse, io.out, io.landmark, io.parsing = (t.data_ptr() for t in outs)
    if targets is not None:
        io.hr, io.heatmap, io.labels = (t.data_ptr() for t in targets)
    io.loss_div, io.w_pix = loss_div, w_pix
    return io


def alloc_outputs(x):
    b, _, h, w = x.shape
    dev = x.device
    return (torch.empty((b, 3, h, w), dtype=torch.float32, device=dev),
            torch.empty((b, 3, h, w), dtype=torch.float32, device=dev),
            torch.empty((b, 97, h // 4, w // 4), dtype=torch.float32, device=dev),
            torch.empty((b, 11, h // 4, w // 4), dtype=torch.float32, device=dev))


def check_input(x):
    if not x.is_cuda:
        raise RuntimeError("crfr_b200 FSRNet needs a CUDA tensor: the hot path has no CPU fallback")
    if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != x.shape[3] or x.shape[2] % 16 or x.shape[2] < 32:
        raise ValueError("expected [B,3,S,S] with S a multiple of 16 (>= 32), got %s" % (tuple(x.shape),))


class _FSRNetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, engine, *params):
        x = x.contiguous().float()
        outs = alloc_outputs(x)
        need_grad = any(ctx.needs_input_grad[2:])
        b, s = x.shape[0], x.shape[2]
        nbytes = L.lib().crfr_fsrnet_workspace_bytes(b, s, 1 if need_grad else 0)
        # a private workspace per call when a backward may follow: it holds the saved activations
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device) if need_grad else ops.workspace(nbytes)
        table = _ParamTable([p.detach() for p in params])
        io = _io(x, outs)
        L.call("crfr_fsrnet_forward", engine, table.arr, C.byref(io), 1 if need_grad else 0, ws.data_ptr(),
               ws.numel(), ops.stream())
        ctx.engine, ctx.ws, ctx.x, ctx.outs = engine, ws, x, outs
        ctx.save_for_backward(*params)
        return outs

    @staticmethod
    def backward(ctx, d_coarse, d_out, d_landmark, d_parsing):
        params = ctx.saved_tensors
        sizes = [p.numel() for p in params]
        offs, tot = [], 0
        for n in sizes:
            offs.append(tot)
            tot += (n + 3) // 4 * 4
        flat = torch.zeros(tot, dtype=torch.float32, device=ctx.x.device)
        grads = [flat[o:o + n].view(p.shape) for o, n, p in zip(offs, sizes, params)]
        ptable = _ParamTable([p.detach() for p in params])
        gtable = _ParamTable(grads)
        io = _io(ctx.x, ctx.outs)

        def g(t):
            return None if t is None else t.contiguous().float()
        dc, do, dl, dp = g(d_coarse), g(d_out), g(d_landmark), g(d_parsing)
        L.call("crfr_fsrnet_backward", ctx.engine, ptable.arr, gtable.arr, C.byref(io), ops.ptr(dc), ops.ptr(do),
               ops.ptr(dl), ops.ptr(dp), ctx.ws.data_ptr(), ctx.ws.numel(), ops.stream())
        return (None, None) + tuple(grads)


class OverallNetwork(nn.Module):
    """ref: model/FSRnet.py:488-508; forward(x) -> (coarse_out, out, landmark_out, parsing_out)."""

    def __init__(self):
        super().__init__()
        self._coarse_sr_network = Course_SR_Network()
        self._prior_estimation_network = Prior_Estimation_Network()
        self._fine_sr_encoder = Fine_SR_Encoder()
        self._fine_sr_decoder = Fine_SR_Decoder()
        self.softmax = nn.Softmax()
        self.engine = L.ENGINE_AUTO

    def ordered_parameters(self):
        """Parameters in state_dict order (the order the native program indexes them in)."""
        sd = dict(self.named_parameters())
        return [sd[k] for k in self.state_dict().keys()]

    def forward(self, x):
        check_input(x)
        params = self.ordered_parameters()
        if len(params) != L.FSRNET_NPARAMS:
            raise RuntimeError("parameter table has %d entries, expected %d" % (len(params), L.FSRNET_NPARAMS))
        return _FSRNetFunction.apply(x, self.engine, *params)
