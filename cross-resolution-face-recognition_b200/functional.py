"""Autograd nodes over the single-op entry points of the C-ABI (conv, normalise + activation, linear, layout).

The hot-path networks (``OverallNetwork``, its four sub-networks, ``ResNet_34``, ``IR_50``) run as whole native programs;
the modules around them - the discriminator of ``OverallNetwork_GAN`` (model/FSRnet.py:461-486 of the reference), the
SUPER_RESOLUTION variant - are composed here op by op.  Every node's forward and backward is one or two native
kernels; activations between nodes are NHWC bf16 tensors, parameters stay fp32 ``nn.Parameter``s in the reference's
layouts.  There is no eager / CPU path: the nodes raise on CPU tensors.
"""
import ctypes as C

import torch

from . import _lib as L
from . import ops


class _ToNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ld):
        ctx.c = x.shape[1]
        return ops.nchw_to_nhwc(x, ld=ld)

    @staticmethod
    def backward(ctx, g):
        return ops.nhwc_to_nchw(g.contiguous(), c=ctx.c), None


class _ToNCHW(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, c):
        ctx.ld = x.shape[3]
        return ops.nhwc_to_nchw(x.contiguous(), c=c)

    @staticmethod
    def backward(ctx, g):
        return ops.nchw_to_nhwc(g, ld=ctx.ld), None


def to_nhwc(x, ld=None):
    """fp32 NCHW -> bf16 NHWC (channel count padded to ``ld``; 3-channel images use ld = 4)."""
    c = x.shape[1]
    return _ToNHWC.apply(x, (4 if c < 8 else c) if ld is None else ld)


def to_nchw(x, c=None):
    """bf16 NHWC -> fp32 NCHW."""
    return _ToNCHW.apply(x, x.shape[3] if c is None else c)


class _Conv2d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad, cin, engine):
        cout, k = weight.shape[0], weight.shape[2]
        wp = ops.pack_conv_weight(weight.detach(), s_pad=x.shape[3] if cin < 8 else None)
        out_ld = 4 if cout < 8 else (cout + 7) // 8 * 8          # 16-byte aligned pixels (13 parsing / 68 landmark channels)
        y, _, _ = ops.conv_fwd(x, wp, cin, cout, k, stride, pad, bias=None if bias is None else bias.detach(), engine=engine,
                               out_ld=out_ld)
        if out_ld != cout:
            y[..., cout:] = 0                                     # padding channels are never written by the kernels
        ctx.save_for_backward(x, weight)
        ctx.cfg = (stride, pad, cin, cout, k, engine, bias is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        stride, pad, cin, cout, k, engine, has_bias = ctx.cfg
        dy = dy.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            wt = ops.pack_conv_weight(weight.detach(), for_dgrad=True)
            dx = ops.conv_dgrad(dy, wt, tuple(x.shape), cin, cout, k, stride, pad, engine=engine)
        if ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2]):
            dw, db = ops.conv_wgrad(x, dy, cin, cout, k, stride, pad, engine=engine, want_bias=has_bias)
        return dx, dw, db, None, None, None, None


def conv2d(x, weight, bias=None, stride=1, pad=0, engine=L.ENGINE_AUTO):
    """nn.Conv2d on an NHWC bf16 activation; ``weight`` fp32 [cout, cin, k, k] (the reference layout)."""
    return _Conv2d.apply(x, weight, bias, stride, pad, weight.shape[1], engine)


class _NormAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, gamma, beta, alpha, res, relu, batch_norm, stats, eps):
        if stats is None:
            stats = ops.norm_stats(y, eps=eps, groups_as_batch=batch_norm)
        out = ops.norm_act_fwd(y, stats, None if gamma is None else gamma.detach(), None if beta is None else beta.detach(),
                               None if alpha is None else alpha.detach(), relu, res, batch_norm=batch_norm)
        ctx.save_for_backward(y, stats, gamma, beta, alpha, res)
        ctx.cfg = (relu, batch_norm)
        ctx.mark_non_differentiable(stats)
        return out, stats

    @staticmethod
    def backward(ctx, dout, _):
        y, stats, gamma, beta, alpha, res = ctx.saved_tensors
        relu, batch_norm = ctx.cfg
        det = lambda t: None if t is None else t.detach()
        dz, dy, dg, db, da = ops.norm_act_bwd(dout.contiguous(), y, stats, det(gamma), det(beta), det(alpha), relu, res,
                                              batch_norm=batch_norm)
        return dy, dg, db, da, dz, None, None, None, None


def norm_act(y, gamma=None, beta=None, alpha=None, res=None, relu=False, batch_norm=False, stats=None, eps=1e-5):
    """out = act(gamma * (y - mean) * rstd + beta (+ res)): InstanceNorm2d (per image) or train-mode BatchNorm (whole
    batch), followed by PReLU (``alpha``), ReLU (``relu``) or nothing.  ``stats`` overrides the batch statistics (eval-mode
    BatchNorm: crfr_bn_running_to_stats).  Returns (out, stats); stats [n or 1, c, 2] = (mean, rstd) fp32."""
    return _NormAct.apply(y, gamma, beta, alpha, res, relu, batch_norm, stats, eps)


def batch_norm(y, bn, alpha=None, relu=False, res=None):
    """nn.BatchNorm2d / BatchNorm1d module semantics on an NHWC bf16 activation: batch statistics and the running-buffer
    update in training mode, running statistics in eval mode."""
    c = bn.num_features
    if bn.training:
        out, stats = norm_act(y, bn.weight, bn.bias, alpha, res, relu, True, None, bn.eps)
        count = y.shape[0] * y.shape[1] * y.shape[2]
        mom = 0.1 if bn.momentum is None else bn.momentum
        L.call("crfr_bn_update_running", ops.ptr(stats), ops.ptr(bn.running_mean), ops.ptr(bn.running_var),
               ops.ptr(bn.num_batches_tracked), c, count, mom, bn.eps, ops.stream())
        return out
    stats = torch.empty((1, c, 2), dtype=torch.float32, device=y.device)
    L.call("crfr_bn_running_to_stats", ops.ptr(bn.running_mean), ops.ptr(bn.running_var), c, bn.eps, ops.ptr(stats),
           ops.stream())
    return norm_act(y, bn.weight, bn.bias, alpha, res, relu, True, stats, bn.eps)[0]


class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        b, h, w, c = x.shape
        o = weight.shape[0]
        y = torch.empty((b, 1, 1, o), dtype=torch.bfloat16, device=x.device)
        ws = ops.workspace(L.lib().crfr_linear_workspace_bytes(b, h * w, c, o))
        L.call("crfr_linear_fwd", ops.ptr(x), b, h * w, c, ops.ptr(weight.detach()), ops.ptr(None if bias is None else bias.detach()),
               o, ops.ptr(y), ops.ptr(ws), ws.numel(), ops.stream())
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        b, h, w, c = x.shape
        o = weight.shape[0]
        dy = dy.contiguous()
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = torch.zeros_like(weight) if ctx.needs_input_grad[1] else None
        db = torch.zeros((o,), dtype=torch.float32, device=x.device) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        ws = ops.workspace(L.lib().crfr_linear_workspace_bytes(b, h * w, c, o))
        L.call("crfr_linear_bwd", ops.ptr(x), ops.ptr(dy), b, h * w, c, ops.ptr(weight.detach()), o, ops.ptr(dx), ops.ptr(dw),
               ops.ptr(db), ops.ptr(ws), ws.numel(), ops.stream())
        return dx, dw, db


def linear(x, weight, bias=None):
    """nn.Linear applied to ``x.view(B, -1)`` of the NCHW tensor whose NHWC bf16 form is ``x`` [B, h, w, c];
    ``weight`` fp32 [out, c*h*w] in the reference's flatten order.  Returns NHWC bf16 [B, 1, 1, out]."""
    return _Linear.apply(x.contiguous(), weight, bias)


class _ReflectPad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pad, c):
        n, h, w, ld = x.shape
        out = torch.empty((n, h + 2 * pad, w + 2 * pad, ld), dtype=torch.bfloat16, device=x.device)
        L.call("crfr_reflect_pad_fwd", ops.ptr(x), ld, ops.ptr(out), ld, n, h, w, c, pad, ops.stream())
        ctx.cfg = (n, h, w, ld, c, pad)
        return out

    @staticmethod
    def backward(ctx, dout):
        n, h, w, ld, c, pad = ctx.cfg
        dout = dout.contiguous()
        dx = torch.zeros((n, h, w, ld), dtype=torch.bfloat16, device=dout.device) if ld != c else \
            torch.empty((n, h, w, ld), dtype=torch.bfloat16, device=dout.device)
        L.call("crfr_reflect_pad_bwd", ops.ptr(dout), ld, ops.ptr(dx), ld, n, h, w, c, pad, ops.stream())
        return dx, None, None


def reflect_pad(x, pad, c=None):
    """nn.ReflectionPad2d(pad) on an NHWC bf16 activation (``c`` real channels; 3-channel images are stored with ld 4)."""
    return _ReflectPad.apply(x.contiguous(), pad, x.shape[3] if c is None else c)


class _Tanh(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous().float()
        y = torch.empty_like(x)
        L.call("crfr_tanh_fwd", ops.ptr(x), ops.ptr(y), x.numel(), ops.stream())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = dy.contiguous().float()
        dx = torch.empty_like(y)
        L.call("crfr_tanh_bwd", ops.ptr(y), ops.ptr(dy), ops.ptr(dx), y.numel(), ops.stream())
        return dx


def tanh(x):
    """nn.Tanh on an fp32 tensor (the 3-channel image heads)."""
    return _Tanh.apply(x)


class _Add(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        return ops.add_n(a.contiguous(), b.contiguous())

    @staticmethod
    def backward(ctx, g):
        return g, g


def add(a, b):
    """a + b on NHWC bf16 activations (fp32 sum, one rounding)."""
    return _Add.apply(a, b)


class _ConvTranspose2d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, stride, pad, out_pad, engine):
        cin, cout, k = weight.shape[0], weight.shape[1], weight.shape[2]
        n, h, w, _ = x.shape
        oh = (h - 1) * stride - 2 * pad + k + out_pad
        ow = (w - 1) * stride - 2 * pad + k + out_pad
        wp = ops.pack_conv_weight(weight.detach(), transposed=True)
        y, _, _ = ops.conv_fwd(x, wp, cin, cout, k, stride, pad, engine=engine, transposed=True, out_hw=(oh, ow))
        ctx.save_for_backward(x, weight)
        ctx.cfg = (stride, pad, cin, cout, k, engine)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        stride, pad, cin, cout, k, engine = ctx.cfg
        dy = dy.contiguous()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            wt = ops.pack_conv_weight(weight.detach(), for_dgrad=True, transposed=True)
            dx = ops.conv_dgrad(dy, wt, tuple(x.shape), cin, cout, k, stride, pad, engine=engine, transposed=True)
        if ctx.needs_input_grad[1]:
            dw, _ = ops.conv_wgrad(x, dy, cin, cout, k, stride, pad, engine=engine, transposed=True)
        return dx, dw, None, None, None, None


def conv_transpose2d(x, weight, stride, pad, out_pad, engine=L.ENGINE_AUTO):
    """nn.ConvTranspose2d (no bias) on an NHWC bf16 activation; ``weight`` fp32 [cin, cout, k, k]."""
    return _ConvTranspose2d.apply(x.contiguous(), weight, stride, pad, out_pad, engine)


class _MaxPool2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ops.maxpool2_fwd(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return ops.maxpool2_bwd(x, g.contiguous())


def max_pool2(x):
    """F.max_pool2d(x, 2, stride=2)."""
    return _MaxPool2.apply(x.contiguous())


class _UpAdd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, up, low):
        return ops.upnearest2_add_fwd(up.contiguous(), low.contiguous())

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        return g, ops.upnearest2_bwd(g)


def up2_add(up, low):
    """up + F.interpolate(low, scale_factor=2) (nearest)."""
    return _UpAdd.apply(up, low)
