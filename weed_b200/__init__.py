"""weed_b200 — B200-native (sm_100a) compute backend for the Weed tensor/autograd library.

The product is native: ``libweedcu.so`` (hand-written CUDA kernels behind the C-ABI of
``include/weedcu.h``) and ``libweed_b200.so`` (the C++ host library that mirrors Weed's
Tensor / Storage / Module / optimiser API on top of it). This Python package only loads those
libraries with ctypes for tests and benchmarks; it contains no compute and no CPU fallback —
importing a binding without the built extension raises.
"""
from ._lib import weedcu, WeedcuError, View, Mat, make_view, contiguous_view, check  # noqa: F401

__all__ = ["weedcu", "WeedcuError", "View", "Mat", "make_view", "contiguous_view", "check"]
