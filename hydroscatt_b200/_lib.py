"""ctypes loader for libhydrotm.so (the C-ABI of include/hydrotm.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``hydroscatt_b200.build.build()``.  There is
no CPU fallback: if the shared object is missing, import of the engine fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhydrotm.so")

_lib = None

_dp = C.c_void_p  # device pointers travel as integers

# name -> (restype, argtypes); every symbol include/hydrotm.h declares
SIGNATURES = {
    "hs_version": (C.c_char_p, []),
    "hs_tmat_size": (C.c_longlong, [C.c_int]),
    "hs_tmat_mode_offset": (C.c_longlong, [C.c_int, C.c_int]),
    "hs_nmax_start": (C.c_int, [C.c_double, C.c_double]),
    "hs_ctx_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "hs_ctx_destroy": (None, [C.c_void_p]),
    "hs_last_error": (C.c_char_p, [C.c_void_p]),
    "hs_launch_count": (C.c_longlong, [C.c_void_p]),
    "hs_calctmat_batch": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_double, C.c_int,
                                    C.c_double, C.c_int, C.c_int, _dp, C.c_longlong, _dp, _dp, _dp, _dp,
                                    _dp, _dp, C.c_void_p]),
    "hs_calcampl_batch": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_int, _dp, C.c_int,
                                    _dp, _dp, _dp, C.c_double, C.c_int, _dp, _dp, C.c_void_p]),
    "hs_xsect_batch": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_int, _dp, C.c_int, _dp, _dp,
                                 _dp, C.c_double, _dp, C.c_void_p]),
    "hs_scatter_batch": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_int,
                                   C.POINTER(C.c_double), C.c_int, _dp, _dp, _dp, C.c_double, _dp, _dp, _dp,
                                   C.c_void_p]),
    "hs_calctmat_scatter_batch": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_double, C.c_int, C.c_double, C.c_int,
                                            C.c_int, _dp, C.c_longlong, _dp, _dp, _dp, _dp, _dp, _dp, C.c_int,
                                            C.POINTER(C.c_double), C.c_int, _dp, _dp, _dp, C.c_double, _dp, _dp, _dp, C.c_void_p]),
    "hs_psd_weights": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_int, _dp, _dp, _dp,
                                 C.c_void_p]),
    "hs_psd_integrate": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, C.c_void_p]),
    "hs_observables": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp,
                                 C.c_int, C.c_double, _dp, C.c_void_p]),
    "hs_psd_observables": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, C.c_int, C.c_int,
                                     _dp, C.c_void_p]),
    "hs_psd_observables_sel": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, C.c_int, C.c_int,
                                         C.POINTER(C.c_int), _dp, C.c_void_p]),
    "hs_pack_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp,
                               C.c_void_p]),
    "hs_gather_rows": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, _dp, _dp, _dp, C.c_void_p]),
    "hs_canting_observables": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _dp, _dp, C.c_longlong, C.c_longlong, _dp, C.c_double,
                                         C.c_int, _dp, C.c_int, _dp, _dp, _dp, C.c_void_p]),
    "hs_canting_mc_scratch_bytes": (C.c_longlong, [C.c_int, C.c_int, C.c_double]),
    "hs_canting_moments_mc": (C.c_int, [C.c_void_p, C.c_int, _dp, C.c_double, C.c_int, C.POINTER(C.c_ulonglong), C.c_double, _dp,
                                        C.c_longlong, _dp, C.c_void_p]),
    "hs_canting_mixture": (C.c_int, [C.c_void_p, C.c_longlong, _dp, _dp, _dp, C.c_void_p]),
    "hs_tm2l_batch": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_double, _dp, _dp,
                                C.c_void_p]),
    "hs_fp64_peak": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_void_p]),
    "hs_fp64_tensor_peak": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_void_p]),
}


class EngineMissingError(ImportError):
    pass


def load():
    """Load libhydrotm.so and attach prototypes.  Raises EngineMissingError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineMissingError(
            f"{LIB_PATH} not found: build the CUDA engine first (python -c 'import __graft_entry__ as g; "
            "g.build()').  hydroscatt_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
