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


# first index of each sub-network's parameters in the 202-entry table (state_dict order of OverallNetwork)
_SECTION_BASE = {L.FSRNET_COARSE: 0, L.FSRNET_PRIOR: 33, L.FSRNET_ENCODER: 132, L.FSRNET_DECODER: 167}
_SECTION_COUNT = {L.FSRNET_COARSE: 33, L.FSRNET_PRIOR: 99, L.FSRNET_ENCODER: 35, L.FSRNET_DECODER: 35}


def _section_outputs(section, x):
    b, dev = x.shape[0], x.device
    if section == L.FSRNET_DECODER:
        s = x.shape[2] * 4
        return s, (torch.empty((b, 3, s, s), dtype=torch.float32, device=dev),)
    s, q = x.shape[2], x.shape[2] // 4
    e = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    if section == L.FSRNET_COARSE:
        return s, (e(b, 64, s, s), e(b, 3, s, s))
    if section == L.FSRNET_ENCODER:
        return s, (e(b, 64, q, q),)
    return s, (e(b, 128, q, q), e(b, 97, q, q), e(b, 11, q, q))


def _section_io(section, x, size, outs):
    io = L.FsrnetSectionIO()
    io.section, io.batch, io.size = section, x.shape[0], size
    io.x = x.data_ptr()
    for i, t in enumerate(outs):
        io.out[i] = t.data_ptr()
    return io


def _section_table(section, tensors):
    full = [None] * L.FSRNET_NPARAMS
    base = _SECTION_BASE[section]
    full[base:base + len(tensors)] = tensors
    return _ParamTable(full)


class _SectionFunction(torch.autograd.Function):
    """One sub-network as a native program (crfr_fsrnet_section_forward) with its own autograd node
    (crfr_fsrnet_section_backward): parameter gradients and the gradient w.r.t. the section input."""

    @staticmethod
    def forward(ctx, x, engine, section, *params):
        x = x.contiguous().float()
        size, outs = _section_outputs(section, x)
        need_grad = any(ctx.needs_input_grad)
        nbytes = L.lib().crfr_fsrnet_section_workspace_bytes(section, x.shape[0], size, 1 if need_grad else 0)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device) if need_grad else ops.workspace(nbytes)
        table = _section_table(section, [p.detach() for p in params])
        io = _section_io(section, x, size, outs)
        L.call("crfr_fsrnet_section_forward", engine, table.arr, C.byref(io), 1 if need_grad else 0, ws.data_ptr(),
               ws.numel(), ops.stream())
        ctx.engine, ctx.section, ctx.size, ctx.ws = engine, section, size, ws
        ctx.xshape, ctx.oshapes, ctx.dev = tuple(x.shape), [tuple(o.shape) for o in outs], x.device
        ctx.save_for_backward(*params)
        return outs

    @staticmethod
    def backward(ctx, *d_outs):
        params = ctx.saved_tensors
        flat, grads = _flat_grads(params, ctx.dev)
        ptable = _section_table(ctx.section, [p.detach() for p in params])
        gtable = _section_table(ctx.section, grads)
        # the backward pass reads neither the input nor the outputs: only their shapes travel (no reference cycle
        # output -> grad_fn -> ctx -> output keeps the workspace alive)
        io = L.FsrnetSectionIO()
        io.section, io.batch, io.size = ctx.section, ctx.xshape[0], ctx.size
        dummy = torch.empty(1, device=ctx.dev)
        io.x = dummy.data_ptr()
        for i in range(len(ctx.oshapes)):
            io.out[i] = dummy.data_ptr()
        keep = [None if d is None else d.contiguous().float() for d in d_outs]
        darr = (C.c_void_p * 3)(*[None if d is None else d.data_ptr() for d in keep] + [None] * (3 - len(keep)))
        dx = torch.empty(ctx.xshape, dtype=torch.float32, device=ctx.dev) if ctx.needs_input_grad[0] else None
        L.call("crfr_fsrnet_section_backward", ctx.engine, ptable.arr, gtable.arr, C.byref(io), darr, ops.ptr(dx),
               ctx.ws.data_ptr(), ctx.ws.numel(), ops.stream())
        ctx.ws = None
        base = _SECTION_BASE[ctx.section]
        return (dx, None, None) + tuple(None if base + i in _DEAD_PARAMS else t for i, t in enumerate(grads))


class _SubNet(nn.Module):
    """Shared forward of the four sub-networks: own parameters in state_dict order -> the section's slice of the table."""
    _section = None
    engine = L.ENGINE_AUTO

    def _run(self, x):
        check_section_input(self._section, x)
        sd = dict(self.named_parameters())
        params = [sd[k] for k in self.state_dict().keys()]
        if len(params) != _SECTION_COUNT[self._section]:
            raise RuntimeError("%s has %d parameters, expected %d" % (type(self).__name__, len(params),
                                                                      _SECTION_COUNT[self._section]))
        return _SectionFunction.apply(x, self.engine, self._section, *params)


class Course_SR_Network(_SubNet):
    """ref: model/FSRnet.py:308-340; forward(x) -> (feat64, coarse3)."""
    _section = L.FSRNET_COARSE

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
        out, out_coarse = self._run(x)
        return out, out_coarse


class Fine_SR_Encoder(Course_SR_Network):
    """ref: model/FSRnet.py:342-379; forward(x) -> feat64 at 1/4 resolution."""
    _section = L.FSRNET_ENCODER

    def __init__(self):
        super().__init__()
        self.conv_input = _conv(3, 64, 7, 4, 3)
        self.relu = nn.PReLU(64)
        self.bn_mid = nn.InstanceNorm2d(64, affine=True)
        self.residual = _stack(3, 64)
        self.conv_end = _conv(64, 64, 3, 1, 1)

    def forward(self, x):
        return self._run(x)[0]


class Prior_Estimation_Network(_SubNet):
    """ref: model/FSRnet.py:381-426; forward(x) -> (feat128, landmark97, parsing11) at 1/4 resolution."""
    _section = L.FSRNET_PRIOR

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
        out, landmark_out, parsing_out = self._run(x)
        return out, landmark_out, parsing_out


class Fine_SR_Decoder(_SubNet):
    """ref: model/FSRnet.py:428-459; forward(x192) -> sr3 at 4x resolution."""
    _section = L.FSRNET_DECODER

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
        return self._run(x)[0]


class Discriminator(nn.Module):
    """ref: model/FSRnet.py:461-486: conv3x3 192 -> 64 -> BatchNorm2d -> PReLU -> the SAME BatchNorm2d again (the residual
    stack is constructed but commented out of forward, :481-482) -> flatten -> Linear(64 * 56 * 56, 512) -> BatchNorm1d.
    The reference hard-codes the 56 x 56 map of a 224 x 224 input; ``spatial`` (default 56: identical ``state_dict``
    shapes) lets the 128 x 128 BASELINE size (32 x 32 maps) use the same module.  Composed from single native ops
    (crfr_b200.functional): tcgen05 conv and linear GEMMs, fused BatchNorm + PReLU passes; train / eval as nn.BatchNorm."""

    def __init__(self, spatial=56):
        super().__init__()
        self.conv_input = _conv(192, 64, 3, 1, 1)
        self.relu = nn.PReLU(64)
        self.bn_mid = nn.BatchNorm2d(64, affine=True)
        self.residual = _stack(3, 64, 64)
        self.fc = nn.Linear(64 * spatial * spatial, 512)
        self.bn_end = nn.BatchNorm1d(512)
        self.spatial = spatial

    def forward(self, x):
        from .. import functional as Fn
        if not x.is_cuda:
            raise RuntimeError("crfr_b200 Discriminator needs a CUDA tensor: the hot path has no CPU fallback")
        if x.dim() != 4 or x.shape[1] != 192 or x.shape[2] != self.spatial or x.shape[3] != self.spatial:
            raise ValueError("expected [B,192,%d,%d], got %s" % (self.spatial, self.spatial, tuple(x.shape)))
        h = Fn.to_nhwc(x.float())
        y = Fn.conv2d(h, self.conv_input.weight, self.conv_input.bias, 1, 1)
        a = Fn.batch_norm(y, self.bn_mid, alpha=self.relu.weight)      # self.relu(self.bn_mid(self.conv_input(x)))  :480
        a = Fn.batch_norm(a, self.bn_mid)                               # out = self.bn_mid(out)                      :483
        o = Fn.linear(a, self.fc.weight, self.fc.bias)                  # view(B, -1) -> fc                          :484-485
        o = Fn.batch_norm(o, self.bn_end)                               # :486
        return Fn.to_nchw(o).view(x.shape[0], 512)


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


# table indices of the parameters that never influence the outputs (888 789 dead parameters):
# coarse bn_end (31, 32), prior residual_next.* (62..85), encoder conv_mid / bn_end (159, 160, 163, 164), decoder
# instance_norm (200, 201)
_DEAD_PARAMS = frozenset([31, 32] + list(range(62, 86)) + [159, 160, 163, 164, 200, 201])


class _ParamTable:
    """The 202 parameter (and gradient) device pointers in state_dict order, as a C array of void*."""

    def __init__(self, tensors):
        self.arr = (C.c_void_p * L.FSRNET_NPARAMS)(*[None if t is None else t.data_ptr() for t in tensors])
        self.keep = tensors


def _check_native_tensor(t, dtype, name):
    """The native side reads raw device pointers: dtype and contiguity must be exactly what include/crfr.h states."""
    if t.dtype != dtype or not t.is_contiguous():
        raise TypeError("%s must be a contiguous %s tensor (got %s, contiguous=%s): coerce with .contiguous().to(...)"
                        % (name, dtype, t.dtype, t.is_contiguous()))


def _io(x, outs, targets=None, loss_div=1.0, w_pix=5.0):
    b, _, h, _ = x.shape
    _check_native_tensor(x, torch.float32, "x")
    if targets is not None:
        hr_, hm_, lb_ = targets
        _check_native_tensor(hr_, torch.float32, "hr")
        _check_native_tensor(hm_, torch.float32, "heatmap")
        _check_native_tensor(lb_, torch.int64, "labels")
        q = h // 4
        if tuple(hr_.shape) != tuple(x.shape) or tuple(hm_.shape) != (b, q, q) or lb_.numel() != b * q * q:
            raise ValueError("target shapes %s %s %s do not match input %s" % (tuple(hr_.shape), tuple(hm_.shape),
                                                                               tuple(lb_.shape), tuple(x.shape)))
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


def check_section_input(section, x):
    if not x.is_cuda:
        raise RuntimeError("crfr_b200 FSRNet needs a CUDA tensor: the hot path has no CPU fallback")
    if section == L.FSRNET_DECODER:
        ok = x.dim() == 4 and x.shape[1] == 192 and x.shape[2] == x.shape[3] and x.shape[2] % 4 == 0 and x.shape[2] >= 8
        if not ok:
            raise ValueError("expected [B,192,Q,Q] with Q a multiple of 4 (>= 8), got %s" % (tuple(x.shape),))
    else:
        check_input(x)


def _flat_grads(params, dev):
    """One zeroed flat fp32 buffer with a 16-byte aligned view per parameter (the native backward accumulates)."""
    sizes = [p.numel() for p in params]
    offs, tot = [], 0
    for n in sizes:
        offs.append(tot)
        tot += (n + 3) // 4 * 4
    flat = torch.zeros(tot, dtype=torch.float32, device=dev)
    return flat, [flat[o:o + n].view(p.shape) for o, n, p in zip(offs, sizes, params)]


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
        # only shapes are kept: storing the outputs on ctx would close the cycle output -> grad_fn -> ctx -> output and
        # keep the multi-GB workspace alive until the cyclic collector runs
        ctx.engine, ctx.ws, ctx.xshape, ctx.dev = engine, ws, tuple(x.shape), x.device
        ctx.save_for_backward(*params)
        return outs

    @staticmethod
    def backward(ctx, d_coarse, d_out, d_landmark, d_parsing):
        params = ctx.saved_tensors
        flat, grads = _flat_grads(params, ctx.dev)
        ptable = _ParamTable([p.detach() for p in params])
        gtable = _ParamTable(grads)
        io = L.FsrnetIO()                      # the backward pass reads neither x nor the outputs
        io.batch, io.size = ctx.xshape[0], ctx.xshape[2]

        def g(t):
            return None if t is None else t.contiguous().float()
        dc, do, dl, dp = g(d_coarse), g(d_out), g(d_landmark), g(d_parsing)
        L.call("crfr_fsrnet_backward", ctx.engine, ptable.arr, gtable.arr, C.byref(io), ops.ptr(dc), ops.ptr(do),
               ops.ptr(dl), ops.ptr(dp), ctx.ws.data_ptr(), ctx.ws.numel(), ops.stream())
        ctx.ws = None                          # the saved activations are not needed again
        # parameters that do not influence the outputs (bn_end, residual_next.*, the encoder's conv_mid, the decoder's
        # instance_norm) get no gradient, as under the reference's autograd (p.grad stays None)
        return (None, None) + tuple(None if i in _DEAD_PARAMS else t for i, t in enumerate(grads))


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


class OverallNetwork_GAN(nn.Module):
    """ref: model/FSRnet.py:512-545; forward(lr, hr) -> (sr, coarse, landmark_out1, parsing_out1, embedding1, embedding2).
    The four sub-networks are native sub-programs chained through autograd exactly as the reference chains its modules;
    the discriminator embeds the concatenated prior / encoder features of the coarse image and of the HR image."""

    def __init__(self, spatial=56):
        super().__init__()
        self._coarse_sr_network = Course_SR_Network()
        self._prior_estimation_network = Prior_Estimation_Network()
        self._fine_sr_encoder = Fine_SR_Encoder()
        self._fine_sr_decoder = Fine_SR_Decoder()
        self._discriminator = Discriminator(spatial)

    def forward_once(self, x):
        out_sr = self._fine_sr_encoder(x)
        out_pe, landmark_out, parsing_out = self._prior_estimation_network(x)
        out = torch.cat((out_pe, out_sr), 1)
        criterion_out = self._discriminator(out)
        return out, landmark_out, parsing_out, criterion_out

    def forward(self, lr, hr):
        out, coarse = self._coarse_sr_network(lr)
        out1, landmark_out1, parsing_out1, embedding1 = self.forward_once(coarse)
        out2, landmark_out2, parsing_out2, embedding2 = self.forward_once(hr)
        sr = self._fine_sr_decoder(out1)
        return sr, coarse, landmark_out1, parsing_out1, embedding1, embedding2
