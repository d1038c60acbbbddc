"""crfr_b200: B200-native (sm_100a) hot path of HyoKong/Cross-Resolution-Face-Recognition.

Python is the host mirror of the reference's operator interface (nn.Module.forward / loss / eval signatures); all
arithmetic runs in hand-written CUDA behind the C-ABI of include/crfr.h (libcrfr.so, loaded with ctypes).
"""
from . import _lib  # noqa: F401
from ._lib import ENGINE_AUTO, ENGINE_DIRECT, ENGINE_TCGEN05, lib  # noqa: F401


def build(force=False):
    """Compiles csrc/*.cu for sm_100a and links libcrfr.so next to this package (no-op when up to date)."""
    from .buildlib import build as _b
    return _b(force=force)
