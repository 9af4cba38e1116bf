"""hexed_b200 -- B200-native implementation of Hexed's per-stage DG residual update.

The product is `libhexed_b200.so` (hand-written sm_100a CUDA behind the C ABI of include/hexed_b200.h).
This package is the thin host side used by the tests and the benchmark: basis tables, mesh flattening and a
ctypes binding that mirrors the reference's `kernels.hpp` entry points. There is no CPU fallback: if the
library is missing or no GPU is present, compute calls raise.
"""
from . import basis, tables, mesh, cases  # noqa: F401
from .basis import gauss_legendre, gauss_lobatto  # noqa: F401
