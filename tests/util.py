"""Shared helpers for the parity tests."""
import torch


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def rel_err(a, b):
    """Norm-wise relative error ||a-b|| / ||b|| in float64."""
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def max_abs(a, b):
    return (a.detach().double().cpu() - b.detach().double().cpu()).abs().max().item()


def to_nchw(x_nhwc, c=None):
    """bf16 NHWC [N,H,W,ld] (device) -> fp32 NCHW (cpu)."""
    c = x_nhwc.shape[3] if c is None else c
    return x_nhwc[..., :c].float().permute(0, 3, 1, 2).contiguous().cpu()


def nhwc_from(x_nchw_cpu, ld=None, device="cuda"):
    """fp32 NCHW (cpu, already bf16-representable) -> bf16 NHWC [N,H,W,ld] on the device, zero channel padding."""
    n, c, h, w = x_nchw_cpu.shape
    ld = c if ld is None else ld
    out = torch.zeros((n, h, w, ld), dtype=torch.bfloat16)
    out[..., :c] = x_nchw_cpu.permute(0, 2, 3, 1).to(torch.bfloat16)
    return out.to(device)
