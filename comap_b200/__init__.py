"""comap_b200 -- B200-native implementation of CoMap's data-parallel hot path.

The product is libcomap_b200.so (CUDA sm_100a kernels behind the C ABI declared in
include/comap_b200.h) plus the C++ `comap_b200` front-end that keeps CoMap's option-file
interface.  This Python package is the thin ctypes binding used by tests and bench.py.
"""
__all__ = ["synthetic"]
