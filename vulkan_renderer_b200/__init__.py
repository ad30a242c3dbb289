"""B200-native shading pass of MomentsInGraphics/vulkan_renderer behind the reference's C host surface.

The product is libvkr_b200.so (hand-written sm_100a CUDA + C++ host code, C-ABI in include/vkr_b200.h).
This Python package is a thin ctypes mirror for tests and benchmarks. Importing it loads the library
and verifies that every symbol of the header is exported; it raises if the library was not built.
"""
from . import api
from .api import load_library

_lib = load_library()

from .frame import Frame  # noqa: E402

__all__ = ["api", "load_library", "Frame"]
