"""Stand-in for the reference's Boost.Python module `libCP` (python/CP_quadratic_l1_py.cpp:107-160): same function, argument
order, keyword names and defaults (:68-76), same return value `(Cv, rX)`, so that python/example_EEG_CP.py runs with
`from cp_pfdr_graph_d1_b200 import libCP`.

The cut-pursuit outer loop itself is the reference's own code (graph cuts stay on the CPU, BASELINE.json north_star); this
module drives it through the gateway csrc/cp_gateway.cpp, built by the reference's maintainer against libpfdr_b200_shim.so
(INTEGRATION.md "Python module") so that every PFDR solve and operator norm inside it runs on the GPU.  The shared library is
looked up in $PFDR_B200_CP_LIB, then lib/libCP_b200.so next to this file.  Nothing is computed in Python.

Differences from the Boost module, all deliberate:
 * a scalar `edge_weight` is broadcast to all E edges (the reference fills V of the E entries, :133-134, and truncates the
   scalar to float for float64 inputs, :132/:138);
 * `source` / `target` of any integer type are accepted (the example passes uint32; the reference reinterprets the buffer).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _load():
    global _lib
    if _lib is None:
        path = os.environ.get("PFDR_B200_CP_LIB") or os.path.join(_HERE, "lib", "libCP_b200.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} not found: build the cut-pursuit gateway against libpfdr_b200_shim.so (INTEGRATION.md, "
                              "'Python module') or point PFDR_B200_CP_LIB at it")
        from . import _lib as product
        product.load()                               # libpfdr_b200.so first (RTLD_GLOBAL): fails loudly without the CUDA library
        _lib = C.CDLL(path)
    return _lib


def CP_quadratic_l1(obs, source, target, edge_weight, A, l1_weight, positivity=0, PFDR_rho=1.0, PFDR_condMin=1e-3,
                    CP_difTol=1e-3, PFDR_difRcd=0.0, PFDR_difTol=1e-4, CP_itMax=10, PFDR_itMax=10000, verbose=0):
    """minimise 1/2 ||obs - A x||^2 + sum_e edge_weight_e |x_u - x_v| + sum_v l1_weight_v |x_v| (+ x >= 0 if positivity)
    by cut-pursuit.  A: N-by-V matrix, a vector of length N (diagonal: V = N) or a scalar / length-1 array (identity);
    edge_weight, l1_weight: arrays or scalars.  Returns (Cv, rX): component of every vertex, value of every component;
    the solution is rX[Cv].  The element type of `obs` (float32 / float64) selects the arithmetic."""
    obs = np.asarray(obs)
    if obs.dtype == np.float32:
        R, suf = C.c_float, "f32"
    elif obs.dtype == np.float64:
        R, suf = C.c_double, "f64"
    else:
        raise ValueError("Type unknown, must be float32, float64, f4, or f8.")          # the reference's message (:253)
    dt = obs.dtype
    obs = np.ascontiguousarray(obs.ravel())
    N = obs.shape[0]
    Eu = np.ascontiguousarray(np.asarray(source).ravel(), dtype=np.int32)
    Ev = np.ascontiguousarray(np.asarray(target).ravel(), dtype=np.int32)
    E = Eu.shape[0]
    A = np.asarray(A, dtype=dt)
    bad = "A should be either a scalar, a vector of size N, or a N-by-V matrix "
    n_col = A.shape[1] if A.ndim == 2 else 1
    if A.ndim > 2:
        raise ValueError(bad)
    if n_col == 1:                                        # reference :203-214
        flat = A.ravel()
        if flat.shape[0] == 1:
            Y, Aarg = obs, None                           # identity
        elif flat.shape[0] == N:
            Y, Aarg = (obs * flat).astype(dt), (flat * flat).astype(dt)     # diagonal: observation and operator pre-multiplied (:119-128)
        else:
            raise ValueError(bad)
        V, N = N, 0
    else:
        if A.shape[0] != N:
            raise ValueError(bad)
        V = A.shape[1]
        Y, Aarg = obs, np.asfortranarray(A)               # column-major N x V as the C++ side expects
    w = np.asarray(edge_weight, dtype=dt).ravel()
    w = np.full(E, w[0], dt) if w.shape[0] == 1 else np.ascontiguousarray(w)
    l1 = np.asarray(l1_weight, dtype=dt).ravel()
    l1 = np.full(V, l1[0], dt) if l1.shape[0] == 1 else np.ascontiguousarray(l1)
    if w.shape[0] != E or l1.shape[0] != V:
        raise ValueError("edge_weight must have one entry per edge (or be a scalar), l1_weight one per vertex (or be a scalar)")
    if E and (min(Eu.min(), Ev.min()) < 0 or max(Eu.max(), Ev.max()) >= V):
        raise ValueError("source / target hold a vertex number outside [0, V)")
    f = getattr(_load(), f"cp_b200_quadratic_d1_l1_{suf}")
    PR, PI = C.POINTER(R), C.POINTER(C.c_int)
    f.restype = C.c_int
    f.argtypes = [C.c_int, C.c_int, C.c_int, PR, PR, PI, PI, PR, PR, C.c_int, R, C.c_int, R, R, R, R, C.c_int, C.c_int, PI, PR, PI]
    Cv, rX, it = np.zeros(V, np.int32), np.zeros(V, dt), C.c_int(0)
    p = lambda a, t: None if a is None else a.ctypes.data_as(C.POINTER(t))
    rV = f(V, E, N, p(Y, R), p(Aarg, R), p(Eu, C.c_int), p(Ev, C.c_int), p(w, R), p(l1, R), int(positivity), CP_difTol,
           int(CP_itMax), PFDR_rho, PFDR_condMin, PFDR_difRcd, PFDR_difTol, int(PFDR_itMax), int(verbose), p(Cv, C.c_int),
           p(rX, R), C.byref(it))
    return Cv, rX[:rV].copy()
