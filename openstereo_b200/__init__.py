"""B200-native cost-volume hot path for OpenStereo (sm_100a CUDA kernels behind a C ABI).

Importing the package loads ``lib/libopenstereo_b200.so``; if it has not been built the import
raises -- there is no PyTorch/CPU fallback for the product path.
"""
from . import _lib  # noqa: F401  (fails loudly when the native library is missing)
from . import ops  # noqa: F401

__all__ = ["ops"]
