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


class _KDFn(torch.autograd.Function):
    """mean(((t - s) - a)^2) with the three input gradients produced by the same fused kernel (crfr_loss_kd)."""

    @staticmethod
    def forward(ctx, t, s, a):
        t, a = t.contiguous().float(), a.contiguous().float()
        s = None if s is None else s.contiguous().float()
        want = (ctx.needs_input_grad[0], s is not None and ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        loss, grads = ops.loss_kd(t, s, a, 1.0, want=want)
        ctx.grads = grads
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        return tuple(None if d is None else d * g for d in ctx.grads)


class MSELoss(nn.Module):
    """ref: nn.MSELoss() as used at distill_main.py:63 - mean((input - target)^2)."""

    def forward(self, input, target):
        return _KDFn.apply(input, None, target)


class ResidualKDLoss(nn.Module):
    """ref: distill_main.py:68-70 - nn.MSELoss()(teacher - student, assistant): the assistant learns the residual
    between the HR teacher's and the LR student's features / embeddings."""

    def forward(self, teacher, student, assistant):
        return _KDFn.apply(teacher, student, assistant)
