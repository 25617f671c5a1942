"""mray_b200 — B200-native (sm_100a) implementation of MRay's wavefront path-tracing hot path.

The product is the C-ABI shared library ``mray_b200/lib/libmray_b200.so`` (include/mray_b200.h)
built from the CUDA sources in ``mray_b200/csrc``; this Python package is only the loader /
ctypes mirror used by tests and bench.py. There is no CPU fallback: importing works anywhere, but
every compute entry point raises when the CUDA extension or a CUDA device is missing.
"""
from . import scenes  # noqa: F401
from .capi import Accelerator, Context, MrbError, Renderer, Scene, Spectrum, load_library  # noqa: F401

__all__ = ["scenes", "Accelerator", "Context", "MrbError", "Renderer", "Scene", "Spectrum", "load_library"]
