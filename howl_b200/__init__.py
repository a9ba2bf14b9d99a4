"""howl_b200 -- B200-native (sm_100a) implementation of castorini/howl's data-parallel hot path.

The product is libhowl_b200.so (hand-written CUDA behind the C ABI of include/howl_b200.h); this package is the
thin Python host layer that mirrors howl's operator interface on top of it.
"""
from ._lib import HowlB200Error, load  # noqa: F401
from .runtime import Context  # noqa: F401

__version__ = "0.1.0"
