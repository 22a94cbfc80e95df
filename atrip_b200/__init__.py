"""atrip_b200 -- B200-native (T) triples-energy hot path of atrip.

The product is the C-ABI library ``atrip_b200/csrc/libatrip_b200.so`` (include/atrip_b200.h) and
the C++ host API over it (include/atrip/Atrip.hpp).  This Python package is plumbing only: a
ctypes binding of the C-ABI for the tests and bench.py, and the process-per-GPU launcher glue.
There is no CPU fallback: importing works anywhere, computing needs the CUDA library and a B200.
"""
from .capi import Engine, EngineError, lib_path, load_library, build_library  # noqa: F401
