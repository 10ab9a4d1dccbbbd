"""bart_b200 -- B200-native forward-model hot path of exosports/BART's `transit`.

The product is the C-ABI library `libbart_b200.so` (include/bart_b200.h, sources in
bart_b200/csrc: C++ host + hand-written CUDA for sm_100a) and the CPython module
`bart_b200/python/transit_module` with the reference's SWIG surface.  This package only holds
the ctypes mirror (`api`), the BARTfunc-style driver (`driver`) and synthetic-input generators
(`synth`).  Nothing here computes on the CPU.
"""
from . import api  # noqa: F401

__all__ = ["api"]
