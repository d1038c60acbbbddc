"""Drop-in IR_50 teacher backed by the native sm_100a network program (forward only).

Mirror of the reference's DISTILLATION/model/model_irse.py for the configuration distill_main.py:14,201 uses
(``IR_50([112, 112])``): same class names (``Backbone``, ``bottleneck_IR``, ``Flatten``, ``l2_norm``, ``get_blocks``),
constructor arguments, sub-module / parameter / buffer names (identical ``state_dict`` keys, so the pretrained
teacher checkpoint loads) and construction + initialisation order.  The torch layers are parameter containers only;
``Backbone.forward`` runs ``crfr_ir50_forward``.

The teacher is frozen and evaluated in ``eval()`` mode by the reference (distill_main.py:43, 112-114); only that mode
is native: train-mode Dropout (model_irse.py:144) is stochastic and has no parity definition, so ``forward`` in
training mode raises.  The SE variants and the 100/152-layer variants are not on the hot path.
"""
import ctypes as C
from collections import namedtuple

import torch
import torch.nn as nn
from torch.nn import BatchNorm1d, BatchNorm2d, Conv2d, Dropout, Linear, MaxPool2d, Module, PReLU, Sequential

from .. import _lib as L
from .. import ops

__all__ = ["Backbone", "IR_50", "bottleneck_IR", "Flatten", "l2_norm", "get_blocks"]


class Flatten(Module):
    """ref: model_irse.py:10-12."""

    def forward(self, input):
        return input.view(input.size(0), -1)


def l2_norm(input, axis=1):
    """ref: model_irse.py:15-19."""
    norm = torch.norm(input, 2, axis, True)
    return torch.div(input, norm)


class bottleneck_IR(Module):
    """ref: model_irse.py:49-66 (shortcut: MaxPool2d(1, stride) or conv1x1 + BN; residual: BN, conv3x3, PReLU,
    conv3x3(stride), BN)."""

    def __init__(self, in_channel, depth, stride):
        super().__init__()
        if in_channel == depth:
            self.shortcut_layer = MaxPool2d(1, stride)
        else:
            self.shortcut_layer = Sequential(Conv2d(in_channel, depth, (1, 1), stride, bias=False), BatchNorm2d(depth))
        self.res_layer = Sequential(BatchNorm2d(in_channel),
                                    Conv2d(in_channel, depth, (3, 3), (1, 1), 1, bias=False), PReLU(depth),
                                    Conv2d(depth, depth, (3, 3), stride, 1, bias=False), BatchNorm2d(depth))


class Bottleneck(namedtuple("Block", ["in_channel", "depth", "stride"])):
    """A named tuple describing a ResNet block (ref: model_irse.py:92-93)."""


def get_block(in_channel, depth, num_units, stride=2):
    return [Bottleneck(in_channel, depth, stride)] + [Bottleneck(depth, depth, 1) for _ in range(num_units - 1)]


def get_blocks(num_layers):
    """ref: model_irse.py:101-126."""
    table = {50: (3, 4, 14, 3), 100: (3, 13, 30, 3), 152: (3, 8, 36, 3)}
    u = table[num_layers]
    return [get_block(64, 64, u[0]), get_block(64, 128, u[1]), get_block(128, 256, u[2]), get_block(256, 512, u[3])]


class _Table:
    def __init__(self, tensors, n):
        assert len(tensors) == n, (len(tensors), n)
        self.arr = (C.c_void_p * n)(*[t.data_ptr() for t in tensors])
        self.keep = tensors


class Backbone(Module):
    """ref: model_irse.py:129-188; forward(x) -> 512-d embedding."""

    def __init__(self, input_size, num_layers, mode="ir"):
        super().__init__()
        assert input_size[0] in [112, 224], "input_size should be [112, 112] or [224, 224]"
        assert num_layers in [50, 100, 152], "num_layers should be 50, 100 or 152"
        assert mode in ["ir", "ir_se"], "mode should be ir or ir_se"
        if mode != "ir":
            raise NotImplementedError("the squeeze-excitation variants are not part of the native hot path")
        blocks = get_blocks(num_layers)
        self.input_layer = Sequential(Conv2d(3, 64, (3, 3), 1, 1, bias=False), BatchNorm2d(64), PReLU(64))
        feat = 512 * 7 * 7 if input_size[0] == 112 else 512 * 14 * 14
        self.output_layer = Sequential(BatchNorm2d(512), Dropout(), Flatten(), Linear(feat, 512), BatchNorm1d(512))
        modules = []
        for block in blocks:
            for bottleneck in block:
                modules.append(bottleneck_IR(bottleneck.in_channel, bottleneck.depth, bottleneck.stride))
        self.body = Sequential(*modules)
        self._initialize_weights()
        self.engine = L.ENGINE_AUTO
        self._native = num_layers == 50 and input_size[0] == 112

    def _initialize_weights(self):
        """ref: model_irse.py:174-188."""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight.data)
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
            elif isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight.data)
                if m.bias is not None:
                    m.bias.data.zero_()

    def _run(self, x, want_features):
        if not self._native:
            raise NotImplementedError("only IR_50([112, 112]) has a native network program")
        if self.training:
            raise RuntimeError("the native IR_50 is the frozen teacher: call .eval() first (train-mode Dropout is "
                               "stochastic and has no parity definition)")
        if not x.is_cuda:
            raise RuntimeError("crfr_b200 IR_50 needs a CUDA tensor: the hot path has no CPU fallback")
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, 112, 112):
            raise ValueError("expected [B,3,112,112], got %s" % (tuple(x.shape),))
        x = x.contiguous().float()
        b = x.shape[0]
        emb = torch.empty((b, 512), dtype=torch.float32, device=x.device)
        feats = [torch.empty((b, c, s, s), dtype=torch.float32, device=x.device)
                 for c, s in ((64, 56), (128, 28), (256, 14), (512, 7))] if want_features else []
        params = [p.detach() for _, p in self.named_parameters()]
        buffers = [t for _, t in self.named_buffers()]
        ptab, btab = _Table(params, L.IR50_NPARAMS), _Table(buffers, 3 * L.IR50_NBN)
        io = L.ResnetIO()
        io.batch, io.size, io.x, io.emb = b, 112, x.data_ptr(), emb.data_ptr()
        for i, f in enumerate(feats):
            io.feat[i] = f.data_ptr()
        io.training, io.momentum, io.eps = 0, 0.1, 1e-5
        ws = ops.workspace(L.lib().crfr_ir50_workspace_bytes(b, 112))
        L.call("crfr_ir50_forward", self.engine, ptab.arr, btab.arr, C.byref(io), ws.data_ptr(), ws.numel(), ops.stream())
        return emb, feats

    def forward(self, x):
        """ref: model_irse.py:167-172: the embedding only."""
        return self._run(x, False)[0]

    def forward_features(self, x):
        """(embedding, x1, x2, x3, x4): the outputs of the four body stages next to the embedding - the tuple
        distill_main.py:59 unpacks from its teacher, and the 'extracted layers' use of DISTILLATION/model/utils.py:36-52
        (perceptual features, SUPER_RESOLUTION/train_FHN.py:255-265)."""
        emb, feats = self._run(x, True)
        return (emb,) + tuple(feats)


def IR_50(input_size):
    """Constructs a ir-50 model (ref: model_irse.py:191-197)."""
    return Backbone(input_size, 50, "ir")
