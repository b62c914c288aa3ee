"""abnn_b200 — B200-native (sm_100a) implementation of ABNN's event-driven Monte-Carlo traversal.

The product is the CUDA library libabnn_b200.so behind the C-ABI in include/abnn.h; this package
is its ctypes binding (capi) and a Python mirror of the reference's Brain / BrainEngine (brain).
There is no CPU fallback: importing `capi.load()` fails loudly if the library has not been built.
"""
from . import capi  # noqa: F401
from .brain import Brain, BrainEngine, FunctionalDataset, SYN_DTYPE  # noqa: F401

__all__ = ["capi", "Brain", "BrainEngine", "FunctionalDataset", "SYN_DTYPE"]
