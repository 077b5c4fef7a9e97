"""prealpha_b200 -- B200-native pair-distribution path of prealpha.

The product is the C-ABI library built from ``prealpha_b200/csrc`` (declared in
``include/prealpha_b200.h``); this package only holds its ctypes binding and the
Python mirror of the reference's `distribution` input interface.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
