"""flechasdb_b200 -- B200 (sm_100a) engine behind flechasdb's IVF-PQ build and query path.

The product is libflechasdb_b200.so (CUDA kernels + the C ABI of include/flechasdb_b200.h).
This package only loads it (ctypes) and mirrors the reference's host-side API on top of it.
"""
from . import _capi as capi  # noqa: F401
from ._capi import FdbError  # noqa: F401
