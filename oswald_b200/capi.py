"""ctypes binding of the C ABI in include/oswald_cuda.h (liboswald_cuda.so, built in-tree).

There is no fallback: if the shared library is missing this module raises, and every
compute entry point needs a B200 (the library refuses other devices).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OSWALD_CUDA_LIB") or os.path.join(_HERE, "liboswald_cuda.so")   # override: experiments only

OSW_OK = 0
OSW_K_U16, OSW_K_I32, OSW_K_DEFAULT = 1, 2, 3
OSW_K_TWO_TRACK, OSW_K_PAIR_DB, OSW_K_TRANSPOSED = 4, 8, 16
OSW_SCORE_FLAGGED = 0x7FFFFFFF


class OswHit(C.Structure):
    _fields_ = [("score", C.c_int32), ("index", C.c_uint32)]


class OswTiming(C.Structure):
    _fields_ = [("device_ms", C.c_double), ("score_ms", C.c_double), ("rescore_ms", C.c_double),
                ("topr_ms", C.c_double), ("h2d_ms", C.c_double), ("wall_ms", C.c_double),
                ("cells", C.c_uint64), ("padded_cells", C.c_uint64), ("rescored_pairs", C.c_uint64),
                ("launches", C.c_uint64), ("sm_cycles", C.c_uint64), ("db_stream_bytes", C.c_uint64),
                ("bound_bytes", C.c_uint64), ("score_launches", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/oswald_cuda.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "osw_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "osw_device_info": (C.c_int, [C.c_int, C.c_char_p, C.c_size_t]),
    "osw_init": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p)]),
    "osw_free": (None, [C.c_void_p]),
    "osw_db_load": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_uint64]),
    "osw_db_write_file": (C.c_int, [C.c_char_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]),
    "osw_db_load_file": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_int]),
    "osw_db_file_info": (C.c_int, [C.c_char_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                   C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "osw_set_device_window": (C.c_int, [C.c_void_p, C.c_uint64]),
    "osw_db_upload": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "osw_db_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "osw_search": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                             C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(OswTiming)]),
    "osw_set_kernels": (C.c_int, [C.c_void_p, C.c_int]),
    "osw_merge_hits": (C.c_size_t, [C.POINTER(C.c_void_p), C.POINTER(C.c_uint32), C.c_int, C.c_uint32, C.c_void_p]),
    "osw_calibrate": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "osw_calibrate_mix": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.c_int]),
    "osw_matrix_count": (C.c_int, []),
    "osw_matrix_name": (C.c_char_p, [C.c_int]),
    "osw_matrix_by_name": (C.c_int, [C.c_char_p, C.c_void_p]),
    "osw_strerror": (C.c_char_p, [C.c_int]),
    "osw_last_error": (C.c_char_p, []),
}

_LIB = None


def lib():
    """The loaded library.  Raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("oswald_b200: %s is missing - build it with `make lib` "
                               "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


class OswError(RuntimeError):
    pass


def check(rc, what):
    if rc != OSW_OK:
        L = lib()
        raise OswError("%s: %s (%s)" % (what, L.osw_strerror(rc).decode(), L.osw_last_error().decode()))
