"""Loads the reference's own modules by file path.  TEST INFRASTRUCTURE ONLY; build container only.

``/root/reference`` exists only in the build container, never on the GPU box, so this module is used by
``oracle/make_golden.py`` (fixture generation + pinning of the oracle restatement) and by the optional
``tests/test_oracle_vs_reference.py`` (skipped when the tree is absent).  Never import the reference as a package
(``utils/__init__.py`` and the training scripts pull missing dependencies / have side effects).
"""
import importlib.util
import os

REF_ROOT = os.environ.get("CRFR_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "model", "FSRnet.py"))


def load(relpath, name=None):
    path = os.path.join(REF_ROOT, relpath)
    name = name or ("crfr_ref_" + relpath.replace("/", "_").replace(".py", ""))
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_weights_init(m):
    """Behaviour of FSR_main.py:38-58 (the script itself cannot be imported: argparse + dataset side effects)."""
    import torch.nn as nn
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


def reference_fsrnet_forward(net, x):
    """The one wiring of OverallNetwork that runs (model/FSRnet.py:497-508 with :538-541's inputs)."""
    import torch
    _, coarse = net._coarse_sr_network(x)
    sr = net._fine_sr_encoder(coarse)
    pe, lm, ps = net._prior_estimation_network(coarse)
    out = net._fine_sr_decoder(torch.cat((pe, sr), 1))
    return coarse, out, lm, ps
