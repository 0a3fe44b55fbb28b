"""azdopt_b200 — B200-native batched search step of azdopt's c21 example.

The product is ``lib/libazb.so`` (CUDA, sm_100a) behind the C ABI in
``include/azb.h``.  This package is the thin host-side mirror of the
reference's ``NablaOptimizer`` / ``NablaStateActionSpace`` / ``NablaModel``
surface on top of that ABI (ctypes, numpy host buffers).  There is no CPU
fallback: importing :mod:`azdopt_b200.capi` without the built library raises.
"""
from .build import build, lib_path  # noqa: F401

__all__ = ["build", "lib_path"]
