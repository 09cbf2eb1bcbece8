"""mepinterp_b200 — B200-native per-k-point Wannier interpolation (drop-in for MEPinterp's core_worker hot path).

The product is libmepb.so (CUDA, C ABI in include/mepb.h) plus the C++ `mepInterp` driver; this package is the thin
Python binding used by the tests and bench.py.
"""
from .api import Context, MultiContext, MepbError, make_model, make_params, lib, LIB_PATH, EXPORTS  # noqa: F401

__all__ = ["Context", "MultiContext", "MepbError", "make_model", "make_params", "lib", "LIB_PATH", "EXPORTS"]
