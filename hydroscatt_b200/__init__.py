"""hydroscatt_b200 -- B200-native T-matrix scattering engine behind hydroscatt's pytmatrix / tm2l surfaces.

Package layout (only what the hot path needs):
  csrc/      hand-written sm_100a CUDA kernels + the C-ABI (include/hydrotm.h) -> libhydrotm.so
  _lib.py    ctypes loader (fails loudly if the library is missing; there is no CPU fallback)
  engine.py  batched host API on torch tensors (device memory + streams only)
  shim/      drop-in `pytmatrix` and `f_2l_tmatrix.tm2l` packages for the unmodified hydroscatt scripts
"""
__version__ = "0.1.0"

import os as _os
import sys as _sys

SHIM_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "shim")


def install_shims():
    """Put the drop-in ``pytmatrix`` and ``f_2l_tmatrix`` packages first on sys.path."""
    if SHIM_DIR not in _sys.path:
        _sys.path.insert(0, SHIM_DIR)
    return SHIM_DIR
