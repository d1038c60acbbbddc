"""Drop-in loss modules (same class names / call signatures as the reference's loss/loss.py) backed by the fused
forward+backward loss kernels.  Each forward launches one kernel that produces the scalar and the input gradient
(bf16, NHWC) in the same pass; autograd only rescales that gradient by the incoming grad_output.
"""
import torch
import torch.nn as nn

from .. import ops


class _MSE97Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, t):
        loss, dx = ops.loss_mse97(x, t, 1.0, want_grad=ctx.needs_input_grad[0])
        ctx.dx, ctx.c = dx, x.shape[1]
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        return ops.nhwc_to_nchw(ctx.dx, ctx.c) * g, None


class _LandmarkFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, t):
        n, c, h, w = x.shape
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((n, h, w, (c + 7) // 8 * 8), dtype=torch.bfloat16, device=x.device)
        loss = ops.loss_landmark(x, t, 1.0, dx, 0)
        ctx.dx, ctx.c = dx, c
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        return ops.nhwc_to_nchw(ctx.dx, ctx.c) * g, None


class _CE2dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, t):
        n, c, h, w = x.shape
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((n, h, w, (c + 7) // 8 * 8), dtype=torch.bfloat16, device=x.device)
        loss = ops.loss_ce2d(x, t, 1.0, dx, 0)
        ctx.dx, ctx.c = dx, c
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        return ops.nhwc_to_nchw(ctx.dx, ctx.c) * g, None


class MSELossFunc(nn.Module):
    """ref: loss/loss.py:7-15 - mean((input - target)^2) * 97."""

    def forward(self, input, target):
        return _MSE97Fn.apply(input.float(), target)


class MSELoss_Landmark(nn.Module):
    """ref: loss/loss.py:17-32 - channel-sum of the 97 heat-maps, then mean squared error * 97."""

    def forward(self, input, target):
        return _LandmarkFn.apply(input.float(), target)


class CrossEntropyLoss2d(nn.Module):
    """ref: loss/loss.py:34-62 - NLLLoss(log_softmax(outputs, 1), squeeze(targets)), mean over B*H*W."""

    def __init__(self, weight=None):
        super().__init__()

    def forward(self, outputs, targets):
        return _CE2dFn.apply(outputs.float(), targets)
