"""CPU oracle for the ResNet_34 embedding + residual knowledge-distillation path.  TEST INFRASTRUCTURE ONLY.

Plain-PyTorch (CPU, fp32) restatement of the reference algorithm, imported only by ``tests/``, ``oracle/make_golden.py``
and the ``cpu_baseline`` legs of the benches; the product package never imports it.

Parity pin: the reference ships no tests or golden vectors, so the pin is live execution of the
reference's own ``model/resnet.py`` in the build container: ``oracle/make_golden.py`` checks every function here
against it and writes ``tests/golden/resnet34.npz``; ``tests/test_oracle_golden.py`` re-checks this file against the
fixture on any box.

Reference sites restated (paths relative to /root/reference):
  BasicBlock                 model/resnet.py:18-47
  ResNet.__init__ / init     model/resnet.py:152-190   (kaiming-normal fan_out convs, zero-init last BN gamma)
  ResNet._make_layer         model/resnet.py:192-205
  ResNet.forward             model/resnet.py:207-225   (max-pool commented out, dropout commented out)
  ResNet_34                  model/resnet.py:231-236
  KD losses                  distill_main.py:63, 68-70 (one forward, no optimiser step in between: SURVEY.md 8c-iii)
"""
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle.fsrnet_oracle import FP32, Precision

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
LAYERS = (3, 4, 6, 3)
PLANES = (64, 128, 256, 512)
RESNET_NULL_GRAD = ("fc.bias", "bn_o1.bias")   # constant shifts removed by the train-mode BatchNorm1d: zero gradient


# ----------------------------------------------------------------------------------------------------------------
# seeded construction (draw-for-draw identical to ``torch.manual_seed(seed); ResNet_34()`` of the reference)
# ----------------------------------------------------------------------------------------------------------------
def build_resnet34_state_dict(seed):
    """state_dict (parameters + BatchNorm buffers, reference key order) of a freshly constructed ResNet_34."""
    torch.manual_seed(seed)
    mods = OrderedDict()                      # name -> torch layer, registered in the reference's module order
    mods["conv1"] = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
    mods["bn1"] = nn.BatchNorm2d(64)
    inplanes = 64
    for l, (nb, planes) in enumerate(zip(LAYERS, PLANES)):
        for b in range(nb):
            stride = 2 if (l > 0 and b == 0) else 1
            p = "layer%d.%d." % (l + 1, b)
            down = None
            if b == 0 and (stride != 1 or inplanes != planes):
                down = (nn.Conv2d(inplanes, planes, 1, stride, bias=False), nn.BatchNorm2d(planes))   # built first
            mods[p + "conv1"] = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
            mods[p + "bn1"] = nn.BatchNorm2d(planes)
            mods[p + "conv2"] = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
            mods[p + "bn2"] = nn.BatchNorm2d(planes)
            if down is not None:
                mods[p + "downsample.0"], mods[p + "downsample.1"] = down
            inplanes = planes
    mods["bn_o1"] = nn.BatchNorm2d(512)
    mods["fc"] = nn.Linear(25088, 512)
    mods["bn_o2"] = nn.BatchNorm1d(512)
    for m in mods.values():                   # model/resnet.py:175-180
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)
    for k, m in mods.items():                 # model/resnet.py:185-190 (zero_init_residual)
        if k.endswith(".bn2"):
            nn.init.constant_(m.weight, 0)
    sd = OrderedDict()
    for k, m in mods.items():
        for n, v in m.state_dict().items():
            sd[k + "." + n] = v.detach().clone()
    return sd


def randomize_norm_params(sd, seed):
    """Zero-initialised bn2 weights make every residual branch (and most gradients) vanish at init; the parity tests
    therefore draw all BatchNorm scales / shifts (and the running statistics) from a seeded distribution."""
    g = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for k, v in sd.items():
        if ".bn" in k or k.startswith("bn") or "downsample.1" in k:
            if k.endswith("weight"):
                v = 0.5 + torch.rand(v.shape, generator=g)
            elif k.endswith("bias"):
                v = 0.2 * torch.randn(v.shape, generator=g)
            elif k.endswith("running_mean"):
                v = 0.1 * torch.randn(v.shape, generator=g)
            elif k.endswith("running_var"):
                v = 0.5 + torch.rand(v.shape, generator=g)
        out[k] = v.clone()
    return out


def resnet34_param_names(sd):
    return [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var")
                                  or k.endswith("num_batches_tracked"))]


# ----------------------------------------------------------------------------------------------------------------
# forward restatement
# ----------------------------------------------------------------------------------------------------------------
def _bn(pr, sd, p, y, training, relu, res=None, new_buffers=None):
    """nn.BatchNorm2d / BatchNorm1d (+ residual, + ReLU) with the storage hooks of the CUDA program."""
    dims = (0, 2, 3) if y.dim() == 4 else (0,)
    shape = (1, -1, 1, 1) if y.dim() == 4 else (1, -1)
    if training:
        mean = y.mean(dim=dims)
        var = ((y - mean.view(shape)) ** 2).mean(dim=dims)
        if new_buffers is not None:
            n = y.numel() // y.shape[1]
            new_buffers[p + "running_mean"] = (1 - BN_MOMENTUM) * sd[p + "running_mean"] + BN_MOMENTUM * mean.detach()
            new_buffers[p + "running_var"] = ((1 - BN_MOMENTUM) * sd[p + "running_var"]
                                              + BN_MOMENTUM * var.detach() * n / max(n - 1, 1))
            new_buffers[p + "num_batches_tracked"] = sd[p + "num_batches_tracked"] + 1
    else:
        mean, var = sd[p + "running_mean"], sd[p + "running_var"]
    z = (y - mean.view(shape)) / torch.sqrt(var.view(shape) + BN_EPS)
    z = z * sd[p + "weight"].view(shape) + sd[p + "bias"].view(shape)
    if res is not None:
        z = z + res
    z = pr.qg(z)
    return pr.st(torch.relu(z) if relu else z)          # a stored activation (ForcedPrecision substitutes it)


def _conv(pr, x, w, stride, pad):
    return pr.sb(F.conv2d(pr.qg(x), pr.q(w), None, stride, pad))


def resnet34_forward(sd, x, training=True, pr=FP32, new_buffers=None):
    """-> (embedding [B,512], x1, x2, x3, x4); ``new_buffers`` (dict) receives the updated BatchNorm buffers."""
    a = _conv(pr, pr.q(x), sd["conv1.weight"], 2, 3)
    a = _bn(pr, sd, "bn1.", a, training, True, None, new_buffers)
    feats = []
    for l, nb in enumerate(LAYERS):
        for b in range(nb):
            p = "layer%d.%d." % (l + 1, b)
            stride = 2 if (l > 0 and b == 0) else 1
            y = _conv(pr, a, sd[p + "conv1.weight"], stride, 1)
            y = _bn(pr, sd, p + "bn1.", y, training, True, None, new_buffers)
            y = _conv(pr, y, sd[p + "conv2.weight"], 1, 1)
            res = a
            if (p + "downsample.0.weight") in sd:
                res = _conv(pr, a, sd[p + "downsample.0.weight"], stride, 0)
                res = _bn(pr, sd, p + "downsample.1.", res, training, False, None, new_buffers)
            a = _bn(pr, sd, p + "bn2.", y, training, True, res, new_buffers)
        feats.append(a)
    o = _bn(pr, sd, "bn_o1.", a, training, False, None, new_buffers)
    o = o.reshape(o.shape[0], -1)
    y = pr.sb(F.linear(pr.qg(o), pr.q(sd["fc.weight"]), sd["fc.bias"]))
    emb = _bn(pr, sd, "bn_o2.", y, training, False, None, new_buffers)
    return (emb,) + tuple(feats)


# ----------------------------------------------------------------------------------------------------------------
# residual knowledge distillation (distill_main.py:63, 68-70)
# ----------------------------------------------------------------------------------------------------------------
def kd_losses(t_outs, s_outs, a_outs):
    """(student_loss, assistant_loss): MSE(s_emb, t_emb.detach()) and sum_k MSE(t_k - s_k, a_k) over the four stage
    features and the embedding (teacher and student NOT detached in the assistant loss, as in the reference)."""
    mse = lambda a, b: ((a - b) ** 2).mean()
    student = mse(s_outs[0], t_outs[0].detach())
    assistant = sum(mse(t_outs[k] - s_outs[k], a_outs[k]) for k in (1, 2, 3, 4)) + mse(t_outs[0] - s_outs[0], a_outs[0])
    return student, assistant


def synthetic_faces(batch, seed=4321, size=112):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, 3, size, size, generator=g)


def kd_step(sd_t, sd_s, sd_a, x, precision="fp32"):
    """One KD evaluation: teacher in eval mode (frozen), student and assistant in train mode.
    -> (L_s, L_a, dL_s/dtheta_S, dL_a/dtheta_A, dL_a/dtheta_S, outputs)."""
    pr = Precision(precision)
    names = resnet34_param_names(sd_s)
    ls = {k: (sd_s[k].clone().requires_grad_(True) if k in names else sd_s[k]) for k in sd_s}
    la = {k: (sd_a[k].clone().requires_grad_(True) if k in names else sd_a[k]) for k in sd_a}
    with torch.no_grad():
        t_outs = resnet34_forward(sd_t, x, training=False, pr=pr)
    s_outs = resnet34_forward(ls, x, training=True, pr=pr)
    a_outs = resnet34_forward(la, x, training=True, pr=pr)
    l_s, l_a = kd_losses(t_outs, s_outs, a_outs)
    g_s = torch.autograd.grad(l_s, [ls[k] for k in names], retain_graph=True, allow_unused=True)
    g_a = torch.autograd.grad(l_a, [la[k] for k in names], retain_graph=True, allow_unused=True)
    g_as = torch.autograd.grad(l_a, [ls[k] for k in names], allow_unused=True)
    as_dict = lambda gs: {k: (torch.zeros_like(ls[k]) if g is None else g) for k, g in zip(names, gs)}
    return l_s.detach(), l_a.detach(), as_dict(g_s), as_dict(g_a), as_dict(g_as), (t_outs, s_outs, a_outs)


# ----------------------------------------------------------------------------------------------------------------
# IR_50 teacher (DISTILLATION/model/model_irse.py:49-66, 103-110, 129-197), eval mode
# ----------------------------------------------------------------------------------------------------------------
IR50_UNITS = (3, 4, 14, 3)


def ir50_block_specs():
    """(in_channel, depth, stride) of the 24 bottleneck_IR units, model_irse.py:96-110."""
    specs, in_ch = [], 64
    for units, depth in zip(IR50_UNITS, PLANES):
        for u in range(units):
            specs.append((in_ch, depth, 2 if u == 0 else 1))
            in_ch = depth
    return specs


def build_ir50_state_dict(seed):
    """state_dict of a freshly constructed ``IR_50([112, 112])`` (reference key order, draw-for-draw identical init)."""
    torch.manual_seed(seed)
    mods = OrderedDict()
    mods["input_layer.0"] = nn.Conv2d(3, 64, 3, 1, 1, bias=False)
    mods["input_layer.1"] = nn.BatchNorm2d(64)
    mods["input_layer.2"] = nn.PReLU(64)
    mods["output_layer.0"] = nn.BatchNorm2d(512)
    mods["output_layer.3"] = nn.Linear(512 * 7 * 7, 512)
    mods["output_layer.4"] = nn.BatchNorm1d(512)
    for i, (cin, d, stride) in enumerate(ir50_block_specs()):
        p = "body.%d." % i
        if cin != d:
            mods[p + "shortcut_layer.0"] = nn.Conv2d(cin, d, 1, stride, bias=False)
            mods[p + "shortcut_layer.1"] = nn.BatchNorm2d(d)
        mods[p + "res_layer.0"] = nn.BatchNorm2d(cin)
        mods[p + "res_layer.1"] = nn.Conv2d(cin, d, 3, 1, 1, bias=False)
        mods[p + "res_layer.2"] = nn.PReLU(d)
        mods[p + "res_layer.3"] = nn.Conv2d(d, d, 3, stride, 1, bias=False)
        mods[p + "res_layer.4"] = nn.BatchNorm2d(d)
    for m in mods.values():                   # model_irse.py:174-188
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            nn.init.xavier_uniform_(m.weight.data)
            if m.bias is not None:
                m.bias.data.zero_()
        elif isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
            m.weight.data.fill_(1)
            m.bias.data.zero_()
    sd = OrderedDict()
    for k, m in mods.items():
        for n, v in m.state_dict().items():
            sd[k + "." + n] = v.detach().clone()
    return sd


def randomize_bn_everywhere(sd, seed):
    """Seeded non-trivial BatchNorm scales / shifts / running statistics for every BatchNorm of a state_dict."""
    g = torch.Generator().manual_seed(seed)
    out = OrderedDict((k, v.clone()) for k, v in sd.items())
    for k in sd:
        if k.endswith("running_mean"):
            p = k[:-len("running_mean")]
            out[p + "weight"] = 0.5 + torch.rand(sd[p + "weight"].shape, generator=g)
            out[p + "bias"] = 0.2 * torch.randn(sd[p + "bias"].shape, generator=g)
            out[p + "running_mean"] = 0.1 * torch.randn(sd[k].shape, generator=g)
            out[p + "running_var"] = 0.5 + torch.rand(sd[k].shape, generator=g)
    return out


def ir50_forward(sd, x, pr=FP32, want_features=False):
    """Backbone.forward (model_irse.py:167-172) in eval mode -> embedding [B, 512]; with ``want_features`` also the
    outputs of the four body stages (units [3, 4, 14, 3], model_irse.py:103-110), i.e. what a FeatureExtractor
    (DISTILLATION/model/utils.py:36-52) walking ``body`` collects after blocks 2, 6, 20 and 23."""
    def prelu(y, a):
        a = a.view(1, -1, 1, 1)
        return pr.q(torch.clamp(y, min=0) + a * torch.clamp(y, max=0))

    def bn_prelu(y, p, a):   # BatchNorm (running statistics) + PReLU as one stored pass
        z = (y - sd[p + "running_mean"].view(1, -1, 1, 1)) / torch.sqrt(sd[p + "running_var"].view(1, -1, 1, 1) + BN_EPS)
        z = z * sd[p + "weight"].view(1, -1, 1, 1) + sd[p + "bias"].view(1, -1, 1, 1)
        a = a.view(1, -1, 1, 1)
        return pr.q(torch.clamp(z, min=0) + a * torch.clamp(z, max=0))

    feats = []
    a = _conv(pr, pr.q(x), sd["input_layer.0.weight"], 1, 1)
    a = bn_prelu(a, "input_layer.1.", sd["input_layer.2.weight"])
    for i, (cin, d, stride) in enumerate(ir50_block_specs()):
        p = "body.%d." % i
        if cin != d:
            sc = _conv(pr, a, sd[p + "shortcut_layer.0.weight"], stride, 0)
            sc = _bn(pr, sd, p + "shortcut_layer.1.", sc, False, False)
        else:
            sc = a[:, :, ::stride, ::stride]                       # MaxPool2d(1, stride)
        r = _bn(pr, sd, p + "res_layer.0.", a, False, False)
        r = prelu(_conv(pr, r, sd[p + "res_layer.1.weight"], 1, 1), sd[p + "res_layer.2.weight"])
        r = _conv(pr, r, sd[p + "res_layer.3.weight"], stride, 1)
        a = _bn(pr, sd, p + "res_layer.4.", r, False, False, res=sc)
        if i in (2, 6, 20, 23):
            feats.append(a)
    o = _bn(pr, sd, "output_layer.0.", a, False, False)
    o = o.reshape(o.shape[0], -1)
    y = pr.qb(F.linear(pr.qg(o), pr.q(sd["output_layer.3.weight"]), sd["output_layer.3.bias"]))
    emb = _bn(pr, sd, "output_layer.4.", y, False, False)
    return (emb, feats) if want_features else emb
