"""CPU oracle for the SUPER_RESOLUTION FSRNet variant.  TEST INFRASTRUCTURE ONLY.

Plain-PyTorch (CPU, fp32) restatement of SUPER_RESOLUTION/model/FSRnet.py:251-416 of the reference, written against a
flat ``state_dict`` (the reference's own key names) so the same weights drive the reference modules, this oracle and
the native drop-in modules.  Pinned in oracle/make_golden.py (``sr``) against the reference's modules loaded by file
path; the ``Precision`` hooks (oracle/fsrnet_oracle.py) give the bf16-storage evaluation used as the yardstick.

  _Residual_Block           :12-37      conv -> IN(affine) -> PReLU -> conv -> IN(affine) -> + x
  Bottleneck                :75-114     IN -> ReLU -> conv1x1 -> IN -> ReLU -> conv3x3 -> IN -> ReLU -> conv1x1, + x
  Hourglass                 :117-156
  Coarse_SR_Network         :251-301
  Fine_SR_Encoder           :304-346
  Prior_Estimation_Network  :349-376
  Fine_SR_Decoder           :379-416
"""
import torch
import torch.nn.functional as F

from .fsrnet_oracle import FP32, EPS


def _in(x, w=None, b=None):
    return F.instance_norm(x, None, None, w, b, True, 0.0, EPS)


def _pconv(pr, sd, key, x, pad, stride=1):
    """ReflectionPad2d(pad) + Conv2d(no bias): stored (bf16) padded input gradient, weights and output."""
    xp = pr.sb(F.pad(x, (pad, pad, pad, pad), mode="reflect")) if pr.on else F.pad(x, (pad, pad, pad, pad), mode="reflect")
    return pr.sb(F.conv2d(pr.qg(xp), pr.q(sd[key]), None, stride, 0))


def _in_relu(pr, x):
    return pr.st(F.relu(pr.qg(_in(x))))


def _res_block(pr, sd, p, x):
    y = pr.sb(F.conv2d(pr.qg(x), pr.q(sd[p + "conv1.weight"]), None, 1, 1))
    y = pr.st(F.prelu(pr.qg(_in(y, sd[p + "in1.weight"], sd[p + "in1.bias"])), sd[p + "relu.weight"]))
    y = pr.sb(F.conv2d(pr.qg(y), pr.q(sd[p + "conv2.weight"]), None, 1, 1))
    return pr.st(pr.qg(_in(y, sd[p + "in2.weight"], sd[p + "in2.bias"]) + x))


def _bottleneck(pr, sd, p, x):
    y = _in_relu(pr, x)
    y = pr.sb(F.conv2d(pr.qg(y), pr.q(sd[p + "conv1.weight"]), sd[p + "conv1.bias"]))
    y = _in_relu(pr, y)
    y = pr.sb(F.conv2d(pr.qg(y), pr.q(sd[p + "conv2.weight"]), sd[p + "conv2.bias"], 1, 1))
    y = _in_relu(pr, y)
    y = pr.sb(F.conv2d(pr.qg(y), pr.q(sd[p + "conv3.weight"]), sd[p + "conv3.bias"]))
    return pr.sb(y + x)


def _hourglass(pr, sd, p, n, x, num_blocks=3):
    def seq(s, t):
        for b in range(num_blocks):
            t = _bottleneck(pr, sd, p + "hg.%d.%d.%d." % (n - 1, s, b), t)
        return t
    x = pr.qg(x)
    up1 = seq(0, x)
    low1 = seq(1, F.max_pool2d(x, 2, stride=2))
    low2 = _hourglass(pr, sd, p, n - 1, low1, num_blocks) if n > 1 else seq(3, low1)
    low3 = seq(2, low2)
    return pr.sb(up1 + F.interpolate(pr.qg(low3), scale_factor=2))


def _trunk(pr, sd, p, i, x, n_blocks):
    """two stride-2 stages, residual blocks, two transposed-conv stages (Sequential indices from i)."""
    for j in (i, i + 6):
        x = _pconv(pr, sd, p + "%d.weight" % (j + 1), x, 1, 2)
        x = _pconv(pr, sd, p + "%d.weight" % (j + 3), x, 1, 1)
        x = _in_relu(pr, x)
    for b in range(n_blocks):
        x = _res_block(pr, sd, p + "%d." % (i + 12 + b), x)
    j = i + 12 + n_blocks
    for k in (j, j + 5):
        x = pr.sb(F.conv_transpose2d(pr.qg(x), pr.q(sd[p + "%d.weight" % k]), None, 2, 1, 1))
        x = _pconv(pr, sd, p + "%d.weight" % (k + 2), x, 1, 1)
        x = _in_relu(pr, x)
    return x


def _head(pr, sd, key, x):
    xp = F.pad(x, (1, 1, 1, 1), mode="reflect")
    return torch.tanh(F.conv2d(pr.qg(pr.sb(xp) if pr.on else xp), pr.q(sd[key]), None, 1, 0))


def coarse_features(sd, x, p="", pr=FP32, n_blocks=6):
    y = _pconv(pr, sd, p + "model.1.weight", pr.q(x), 3)
    return _trunk(pr, sd, p + "model.", 4, _in_relu(pr, y), n_blocks)


def coarse_forward(sd, x, p="", pr=FP32, n_blocks=6):
    return _head(pr, sd, p + "out.1.weight", coarse_features(sd, x, p, pr, n_blocks))


def encoder_forward(sd, x, p="", pr=FP32, n_blocks=6):
    y = _pconv(pr, sd, p + "model.1.weight", pr.q(x), 1)
    return _trunk(pr, sd, p + "model.", 2, y, n_blocks)


def prior_forward(sd, x, p="", pr=FP32, n_blocks=2, n_hourglass=4):
    y = _in_relu(pr, _pconv(pr, sd, p + "model.1.weight", pr.q(x), 3))
    for b in range(n_blocks):
        y = _res_block(pr, sd, p + "model.%d." % (4 + b), y)
    for h in range(n_hourglass):
        y = _hourglass(pr, sd, p + "model.%d." % (4 + n_blocks + h), 4, y)
    parsing = F.conv2d(pr.qg(y), pr.q(sd[p + "fc.weight"]), sd[p + "fc.bias"])
    landmark = F.conv2d(pr.qg(y), pr.q(sd[p + "fc_landmark.weight"]), None)
    return y, landmark, parsing


def decoder_forward(sd, x, p="", pr=FP32, n_blocks=6):
    return _head(pr, sd, p + "out.1.weight", _trunk(pr, sd, p + "model.", 0, x, n_blocks))


def sr_forward(sd, x, pr=FP32):
    """coarse -> (encoder, prior) -> cat(prior, encoder) -> decoder  (the wiring of :449-461 with the existing classes)."""
    coarse = coarse_forward(sd, x, "_coarse_sr_network.", pr)
    enc = encoder_forward(sd, coarse, "_fine_sr_encoder.", pr)
    pe, landmark, parsing = prior_forward(sd, coarse, "_prior_estimation_network.", pr)
    out = decoder_forward(sd, torch.cat((pe, enc), 1), "_fine_sr_decoder.", pr)
    return coarse, out, landmark, parsing
