"""luminary_b200 - B200-native wavefront path-tracing hot path behind Luminary's device interface.

The product is the C-ABI shared library `liblumb200.so` (CUDA, sm_100a; sources in csrc/, interface in
include/lumb200.h). This package is only the Python mirror of that interface (ctypes) plus the procedural
scene generators used by tests and bench.py. There is no CPU fallback: importing `luminary_b200.api` raises
if the library has not been built.
"""
from . import scenes  # noqa: F401

__all__ = ["scenes"]
