"""exonerate_b200 -- B200-native C4 Viterbi engine behind exonerate's C4 API.

The product is the CUDA library `libc4b200.so` (C ABI in include/c4b200.h,
sources in exonerate_b200/csrc/).  This package only holds the ctypes binding.
"""
from . import abi  # noqa: F401
from .engine import Batch, C4BError, Engine, Group, HSPset, Optimal, PairSet, load_library  # noqa: F401
