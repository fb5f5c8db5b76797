"""Host-side mirror of the reference's operator interface for the PFDR hot path.

Same routine names and argument meaning as the reference's gateways
(/root/reference/octave/mex/PFDR_graph_quadratic_d1_l1_mex.cpp:11-88,
 .../PFDR_graph_quadratic_d1_bounds_mex.cpp, .../PFDR_graph_loss_d1_simplex_mex.cpp,
 .../PFDR_graph_l22_d1_l1_mex.cpp:54-83), each a thin call into the C ABI of
lib/libpfdr_b200.so (include/pfdr_b200.h).  Arrays may be numpy (host) or torch CUDA tensors
(used in place on the device).  Nothing here computes: without the CUDA library the import fails.
"""
import numpy as np

from . import _lib
from ._abi import DIAG, SCAL, Binder, PfdrError, Result  # noqa: F401

_binder = None


def _b():
    global _binder
    if _binder is None:
        _binder = Binder(_lib.load(), "pfdr_b200_", returns_status=True)
    return _binder


def PFDR_graph_quadratic_d1_l1(*a, **k):
    return _b().PFDR_graph_quadratic_d1_l1(*a, **k)


def PFDR_graph_quadratic_d1_bounds(*a, **k):
    return _b().PFDR_graph_quadratic_d1_bounds(*a, **k)


def PFDR_graph_loss_d1_simplex(*a, **k):
    return _b().PFDR_graph_loss_d1_simplex(*a, **k)


def proj_simplex_metric(*a, **k):
    return _b().proj_simplex_metric(*a, **k)


def operator_norm_matrix(*a, **k):
    return _b().operator_norm_matrix(*a, **k)


for _f in (PFDR_graph_quadratic_d1_l1, PFDR_graph_quadratic_d1_bounds, PFDR_graph_loss_d1_simplex,
           proj_simplex_metric, operator_norm_matrix):
    _f.__doc__ = getattr(Binder, _f.__name__).__doc__


def PFDR_graph_l22_d1_l1(Y, La_l2, Eu, Ev, La_d1, La_l1=None, positivity=0, rho=1.5, condMin=1e-3, difRcd=0.0,
                         difTol=1e-5, itMax=1000, want_obj=False, want_dif=False, verbose=0, backend=None):
    """1/2 ||y - x||^2_{La_l2} + d1 + l1 : the reference's l22 gateway
    (octave/mex/PFDR_graph_l22_d1_l1_mex.cpp:41-83): X starts at zero, Y is pre-scaled by La_l2, the
    solver runs with N = 0, A = L = La_l2 (DIAG) and the constant 1/2||y||^2_{La_l2} is added to Obj.
    `backend` (tests only) substitutes another Binder for the CUDA library."""
    b = backend or _b()
    Y = np.asarray(Y)
    dt = Y.dtype
    X = np.zeros_like(Y)
    if La_l2 is None:
        Ys, A, L, Ltype = Y, None, None, SCAL
        shift = 0.5 * float(np.sum(Y.astype(np.float64) ** 2))
    else:
        La_l2 = np.ascontiguousarray(La_l2, dtype=dt)
        Ys, A, L, Ltype = (La_l2 * Y).astype(dt), La_l2, La_l2, DIAG
        shift = 0.5 * float(np.sum(La_l2.astype(np.float64) * Y.astype(np.float64) ** 2))
    r = b.PFDR_graph_quadratic_d1_l1(X, Ys, A, Eu, Ev, La_d1, La_l1, positivity, Ltype, L, rho, condMin, difRcd,
                                     difTol, itMax, 0, want_obj, want_dif, verbose)
    return Result(r.X, r.it, None if r.Obj is None else r.Obj + dt.type(shift), r.Dif)


def set_stream(cuda_stream):
    """Launch stream for subsequent calls of this thread (e.g. torch.cuda.current_stream().cuda_stream)."""
    _lib.load().pfdr_b200_set_stream(cuda_stream)


def last_stats():
    s = _lib.Stats()
    _lib.load().pfdr_b200_get_last_stats(s)
    return s.asdict()


# ---------------------------------------------------------------------------------------------------
# resident solver handles (include/pfdr_b200.h "resident solver handles")
# ---------------------------------------------------------------------------------------------------
import ctypes as _C

from ._abi import _CT, _as_int, _as_real, _as_rows, _ptr, _real_dtype


class _Solver:
    """Inputs, half-edge layout and work arrays stay in HBM; reset()/run() are device-only."""

    def __init__(self):
        self._h = _C.c_void_p()
        self._lib = _lib.load()
        self._keep = []

    def _chk(self, rc):
        if rc != 0:
            raise PfdrError(f"libpfdr_b200 status {rc}: {self._lib.pfdr_b200_last_error().decode()}")

    def reset(self):
        self._chk(self._lib.pfdr_b200_solver_reset(self._h))

    def run(self):
        self._chk(self._lib.pfdr_b200_solver_run(self._h))
        return last_stats()

    def fetch(self, want_obj=False, want_dif=False):
        R = _CT[self._dt][0]
        X = np.zeros(self._shape, self._dt)
        it = _C.c_int(0)
        Obj = np.zeros(self._itMax + 1, self._dt) if want_obj else None
        Dif = np.zeros(max(self._itMax, 1), self._dt) if want_dif else None
        self._lib.pfdr_b200_solver_fetch.argtypes = [_C.c_void_p, _C.c_void_p, _C.POINTER(_C.c_int), _C.c_void_p, _C.c_void_p]
        self._chk(self._lib.pfdr_b200_solver_fetch(self._h, X.ctypes.data, _C.byref(it),
                                                   None if Obj is None else Obj.ctypes.data,
                                                   None if Dif is None else Dif.ctypes.data))
        n = it.value
        return Result(X, n, None if Obj is None else Obj[:n + 1], None if Dif is None else Dif[:n])

    def close(self):
        if self._h:
            self._lib.pfdr_b200_solver_destroy.argtypes = [_C.c_void_p]
            self._lib.pfdr_b200_solver_destroy(self._h)
            self._h = _C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class QuadraticSolver(_Solver):
    """Handle on PFDR_graph_quadratic_d1_l1 (prox='l1') / _bounds (prox='bounds'); arrays may be numpy or
    torch CUDA tensors (kept referenced by the handle, used in place)."""

    def __init__(self, X0, Y, A, Eu, Ev, La_d1, La_l1=None, positivity=0, min=-np.inf, max=np.inf, prox="l1",
                 Ltype=SCAL, L=None, rho=1.5, condMin=1e-3, difRcd=0.0, difTol=1e-5, itMax=1000, N=0,
                 want_obj=False, want_dif=False):
        super().__init__()
        dt = _real_dtype(X0)
        R, suf = _CT[dt]
        P, PI = _C.POINTER(R), _C.POINTER(_C.c_int)
        self._dt, self._shape, self._itMax = dt, (int(X0.shape[0]),), itMax
        Y, A, La_d1, La_l1, L, X0 = (_as_real(a, dt) for a in (Y, A, La_d1, La_l1, L, X0))
        Eu, Ev = _as_int(Eu), _as_int(Ev)
        self._keep = [X0, Y, A, Eu, Ev, La_d1, La_l1, L]
        f = getattr(self._lib, f"pfdr_b200_solver_create_quadratic_{suf}")
        f.argtypes = [_C.POINTER(_C.c_void_p), _C.c_int, _C.c_int, _C.c_int, _C.c_int, P, P, P, PI, PI, P, P, _C.c_int,
                      R, R, _C.c_int, P, R, R, R, R, _C.c_int, _C.c_int, _C.c_int]
        self._chk(f(_C.byref(self._h), 0 if prox == "l1" else 1, int(X0.shape[0]), int(Eu.shape[0]), int(N),
                    _ptr(X0, R), _ptr(Y, R), _ptr(A, R), _ptr(Eu, _C.c_int), _ptr(Ev, _C.c_int), _ptr(La_d1, R),
                    _ptr(La_l1, R), int(positivity), float(min), float(max), int(Ltype), _ptr(L, R), rho, condMin,
                    difRcd, difTol, itMax, int(want_obj), int(want_dif)))
        self.setup_stats = last_stats()


class SimplexSolver(_Solver):
    """Handle on PFDR_graph_loss_d1_simplex; P0, Q are (V, K)."""

    def __init__(self, P0, Q, al, Eu, Ev, La_d1, La_f=None, rho=1.5, condMin=1e-2, difRcd=0.0, difTol=1e-4,
                 itMax=500, want_obj=False, want_dif=False):
        super().__init__()
        dt = _real_dtype(P0)
        R, suf = _CT[dt]
        P, PI = _C.POINTER(R), _C.POINTER(_C.c_int)
        V, K = int(P0.shape[0]), int(P0.shape[1])
        self._dt, self._shape, self._itMax = dt, (V, K), itMax
        P0, Q, La_d1, La_f = (_as_rows(a, dt) for a in (P0, Q, La_d1, La_f))
        Eu, Ev = _as_int(Eu), _as_int(Ev)
        self._keep = [P0, Q, Eu, Ev, La_d1, La_f]
        f = getattr(self._lib, f"pfdr_b200_solver_create_simplex_{suf}")
        f.argtypes = [_C.POINTER(_C.c_void_p), _C.c_int, _C.c_int, _C.c_int, R, P, P, P, PI, PI, P, R, R, R, R,
                      _C.c_int, _C.c_int, _C.c_int]
        self._chk(f(_C.byref(self._h), K, V, int(Eu.shape[0]), al, _ptr(La_f, R), _ptr(P0, R), _ptr(Q, R),
                    _ptr(Eu, _C.c_int), _ptr(Ev, _C.c_int), _ptr(La_d1, R), rho, condMin, difRcd, difTol, itMax,
                    int(want_obj), int(want_dif)))
        self.setup_stats = last_stats()
