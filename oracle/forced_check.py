"""Teacher-forced checking of the native FSRNet program.  TEST INFRASTRUCTURE ONLY (used by tests/ and smoke()).

The stored bf16 forward tensors of a native call are read out of its workspace (crfr_fsrnet_tape gives their offsets)
and substituted into the CPU oracle at its storage points (fsrnet_oracle.ForcedPrecision): every layer is then checked
on identical inputs, and the backward pass - linear once the forward is fixed - must reproduce the parameter gradients
to rounding noise.  See tests/test_fsrnet_forced_gpu.py.
"""
import ctypes as C

import torch

LAYER_TOL = 2e-3        # stored layer output vs the oracle's bf16-rounded value on identical inputs
STORE_KINDS = (0, 1, 3)  # tape kinds whose bf16 output is a storage point of the oracle: conv, norm(+act), up+add


def make_net():
    from crfr_b200 import _lib as L
    from crfr_b200.model.FSRnet import OverallNetwork, weights_init
    torch.manual_seed(1234)
    net = OverallNetwork()
    net.apply(weights_init)
    net.engine = L.ENGINE_AUTO
    return net.cuda().train()


def tape_of(b, s):
    from crfr_b200 import _lib as L
    n = L.lib().crfr_fsrnet_tape(b, s, 1, None, 0)
    assert n > 0
    arr = (L.TapeEntry * n)()
    assert L.lib().crfr_fsrnet_tape(b, s, 1, arr, n) == n
    return list(arr)


def view(ws, off, n, h, w, c, ld):
    """bf16 NHWC view [n,h,w,c] (pixel stride ld) of the workspace at byte offset off."""
    return torch.as_strided(ws.view(torch.bfloat16), (n, h, w, c), (h * w * ld, w * ld, ld, 1), off // 2)


def stored(ws, e, lo=0, hi=None):
    """Stored activation of tape entry e for images [lo, hi) as fp32 NCHW on the CPU."""
    v = view(ws, e.out_off, e.n, e.h, e.w, e.c, e.ld)[lo:hi]
    return v.float().permute(0, 3, 1, 2).contiguous().cpu()


class Step:
    """One native call on a private workspace; keeps everything the checks need."""

    def __init__(self, net, x, targets=None, loss_div=None):
        from crfr_b200 import _lib as L, ops
        from crfr_b200.model import FSRnet as M
        self.L, self.M, self.ops = L, M, ops
        self.net, self.x = net, x
        self.b, self.s = x.shape[0], x.shape[2]
        self.params = net.ordered_parameters()
        self.pt = M._ParamTable([p.detach() for p in self.params])
        self.outs = M.alloc_outputs(x)
        self.ws = torch.empty(L.lib().crfr_fsrnet_workspace_bytes(self.b, self.s, 1), dtype=torch.uint8, device="cuda")
        self.targets = targets
        self.io = M._io(x, self.outs, targets, loss_div=loss_div or 2.0 * self.b, w_pix=5.0)

    def new_grads(self):
        grads = [torch.zeros_like(p) for p in self.params]
        return grads, self.M._ParamTable(grads)

    def train_step(self):
        grads, gt = self.new_grads()
        losses = torch.zeros(5, device="cuda")
        self.L.call("crfr_fsrnet_train_step", self.L.ENGINE_AUTO, self.pt.arr, gt.arr, C.byref(self.io),
                    losses.data_ptr(), self.ws.data_ptr(), self.ws.numel(), self.ops.stream())
        torch.cuda.synchronize()
        return losses.cpu(), grads

    def forward(self):
        self.L.call("crfr_fsrnet_forward", self.L.ENGINE_AUTO, self.pt.arr, C.byref(self.io), 1, self.ws.data_ptr(),
                    self.ws.numel(), self.ops.stream())
        torch.cuda.synchronize()

    def backward(self, d_outs):
        grads, gt = self.new_grads()
        dc, do, dl, dp = (t.contiguous() for t in d_outs)
        self.L.call("crfr_fsrnet_backward", self.L.ENGINE_AUTO, self.pt.arr, gt.arr, C.byref(self.io), dc.data_ptr(),
                    do.data_ptr(), dl.data_ptr(), dp.data_ptr(), self.ws.data_ptr(), self.ws.numel(),
                    self.ops.stream())
        torch.cuda.synchronize()
        return grads


def forced_feed(ws, tape, lo=0, hi=None):
    return [stored(ws, e, lo, hi) for e in tape if e.kind in STORE_KINDS]


def check_layers(pr, tape, what):
    """Per-layer forward errors of a ForcedPrecision run: every storage point within LAYER_TOL."""
    kinds = [e for e in tape if e.kind in STORE_KINDS]
    assert pr.pos == len(kinds) == len(pr.errors_q), (pr.pos, len(kinds))
    worst = max(range(len(kinds)), key=lambda i: pr.errors_q[i])
    e = kinds[worst]
    assert pr.errors_q[worst] < LAYER_TOL, "%s: storage point %d (kind %d, %dx%dx%d) deviates %.3e" % (
        what, worst, e.kind, e.c, e.h, e.w, pr.errors_q[worst])
    return pr.errors_q[worst]
