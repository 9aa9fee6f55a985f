"""rustyread_b200 — B200-native implementation of rustyread's `simulate` hot path.

The product is the C-ABI shared library (include/rustyread_b200.h, csrc/); this package is the
thin Python host mirror used by the tests and bench.py.  There is no CPU fallback: every compute
entry point needs the CUDA library and a GPU.
"""
from .capi import Library, Context, load_library, RRError  # noqa: F401
from ._build import build_library  # noqa: F401

__all__ = ["Library", "Context", "load_library", "build_library", "RRError"]
