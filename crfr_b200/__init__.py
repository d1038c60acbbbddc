"""Importable alias of the product package.

The package directory is ``cross-resolution-face-recognition_b200/`` (repo layout contract); hyphens are not legal
in a Python module name, so this shim makes it importable as ``crfr_b200`` by pointing ``__path__`` at it.
"""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "cross-resolution-face-recognition_b200")
__path__ = [_REAL]
with open(_os.path.join(_REAL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, "__init__.py"), "exec"))
