"""Loader of the in-tree CUDA library.  There is deliberately no fallback: if
lib/libpfdr_b200.so is missing or cannot be loaded the import of the operators fails loudly."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PFDR_B200_LIBPATH") or os.path.join(_HERE, "lib", "libpfdr_b200.so")   # env: A/B of builds
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "pfdr_b200.h")


class Stats(C.Structure):
    _fields_ = [("setup_ms", C.c_double), ("loop_ms", C.c_double), ("total_ms", C.c_double),
                ("launches", C.c_longlong), ("loop_launches", C.c_longlong), ("iterations", C.c_int),
                ("h2d_bytes", C.c_longlong), ("d2h_bytes", C.c_longlong), ("slots", C.c_longlong),
                ("layout", C.c_int), ("gpus", C.c_int)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                              f"g.build()'` (cp_pfdr_graph_d1_b200 has no CPU fallback)")
        _lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        _lib.pfdr_b200_last_error.restype = C.c_char_p
        _lib.pfdr_b200_get_last_stats.argtypes = [C.POINTER(Stats)]
        _lib.pfdr_b200_set_stream.argtypes = [C.c_void_p]
    return _lib
