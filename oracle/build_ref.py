"""Builds oracle/_ref/: the reference's OWN modules for the hot path, byte-compiled.  TEST / BASELINE INFRASTRUCTURE ONLY.

The reference is pure Python; its sources stay where they lie under /root/reference (never copied into the repo).  This
recipe compiles the six modules of the path to CPython bytecode and writes ONLY those artefacts (*.refbin) into oracle/_ref/
(git-ignored, but shipped to the GPU box with the snapshot, like our own built .so).  `bench.py --impl reference` and the
`cpu_baseline` leg then time the real reference modules on the host cores (`kind: "reference"`); without oracle/_ref they
fall back to the oracle port (`kind: "port"`).  Run in the build container:  python -m oracle.build_ref

  model/FSRnet.py                        OverallNetwork and its sub-networks      (wiring: ref_loader.reference_fsrnet_forward)
  loss/loss.py                           MSELossFunc, MSELoss_Landmark, CrossEntropyLoss2d
  model/resnet.py                        ResNet_34
  DISTILLATION/model/model_irse.py       IR_50
  utils/eval.py, utils/utils.py          accuracy, calculate_accuracy, calculate_roc
"""
import importlib.machinery
import importlib.util
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_ROOT = os.environ.get("CRFR_REFERENCE_ROOT", "/root/reference")
EXT = ".refbin"    # CPython bytecode; not named .pyc so that snapshot tools that skip byte-code caches still ship it
MODULES = {"FSRnet": "model/FSRnet.py", "loss": "loss/loss.py", "resnet": "model/resnet.py",
           "model_irse": "DISTILLATION/model/model_irse.py", "eval": "utils/eval.py", "utils": "utils/utils.py"}


def build():
    """Returns True when oracle/_ref is usable (built now, or already present)."""
    if not os.path.isfile(os.path.join(REF_ROOT, MODULES["FSRnet"])):
        return available()
    os.makedirs(OUT, exist_ok=True)
    for name, rel in MODULES.items():
        py_compile.compile(os.path.join(REF_ROOT, rel), cfile=os.path.join(OUT, name + EXT), doraise=True,
                           dfile="reference/" + rel)
    with open(os.path.join(OUT, "VERSION"), "w") as f:
        f.write("python %d.%d\n" % sys.version_info[:2])
    return True


def available():
    if not os.path.isfile(os.path.join(OUT, "FSRnet" + EXT)):
        return False
    try:
        return open(os.path.join(OUT, "VERSION")).read().strip() == "python %d.%d" % sys.version_info[:2]
    except OSError:
        return False


def load(name):
    """Imports one byte-compiled reference module from oracle/_ref."""
    path = os.path.join(OUT, name + EXT)
    loader = importlib.machinery.SourcelessFileLoader("crfr_refbin_" + name, path)
    spec = importlib.util.spec_from_loader(loader.name, loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    ok = build()
    print("oracle/_ref:", "built" if ok else "reference tree not present, nothing built")
