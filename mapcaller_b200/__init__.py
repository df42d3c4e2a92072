"""mapcaller_b200 -- B200-native (sm_100a) replacement for MapCaller's read-mapping hot path.

The product is the C-ABI shared library `libmapcaller_b200.so` (include/mapcaller_b200.h) built from
mapcaller_b200/csrc by `__graft_entry__.build()`; `mapcaller_b200.api` is a thin ctypes binding used by
the tests and bench.py, `mapcaller_b200.simulate` fabricates seeded synthetic inputs.
"""
from . import simulate  # noqa: F401

__all__ = ["api", "simulate"]
