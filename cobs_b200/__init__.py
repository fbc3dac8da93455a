"""cobs_b200 -- B200-native (sm_100a) implementation of the COBS query hot path.

The compute lives in cobs_b200/lib/libcobsgpu.so (hand-written CUDA behind the C ABI of
include/cobsgpu.h); this package is the ctypes binding plus a mirror of the reference's
Python `Search` class.  No CPU fallback exists.
"""
from ._lib import CobsGpuError, KIND_CLASSIC, KIND_COMPACT, LIB_PATH, lib  # noqa: F401
from .api import GpuGroup, GpuIndex, Search, SearchResult, decode_keys, merge_device  # noqa: F401
