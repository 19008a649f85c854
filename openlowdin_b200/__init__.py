"""openlowdin_b200 -- B200 (sm_100a) four-index AO->MO integral transformation for openLOWDIN.

The product is the C-ABI shared library ``liblowdin_itgpu.so`` (include/lowdin_it.h) built from
``csrc/`` by ``build.py``.  This package is only the thin ctypes binding that tests and bench.py
use to call that ABI; there is no Python or CPU implementation of the transform here, and
importing :mod:`openlowdin_b200.capi` raises if the CUDA library is missing.
"""
from .capi import (CONV_C, CONV_E, GEN_HASH, GEN_FOLD, Transformer, LowdinITError, lib_path, load,  # noqa: F401
                   transform_all, transform_inter_all)
