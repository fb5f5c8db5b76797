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


class CpOperator:
    """The dense passes of the cut-pursuit's outer iteration with the operator resident in HBM (include/pfdr_b200.h,
    reference src/CP_PFDR_graph_quadratic_d1_l1.cpp:340-374 gradient, :664-772 reduced operator, :889-903 residual).
    A: (N, V) Fortran-ordered with N > 0; (V, V) A^tA with N = -V; 1-D diagonal of A^tA or None (identity) with N = 0.
    A torch CUDA tensor is used in place (and must outlive the handle).  dtype: only read when A is None."""

    def __init__(self, A, V, N, dtype=np.float64):
        self._lib = _lib.load()
        self._h = _C.c_void_p()
        self.V, self.N = int(V), int(N)
        if A is None:                                   # identity: `dtype` says which arithmetic the calls use
            if N != 0:
                raise PfdrError("A is required unless N == 0")
            dt = np.dtype(dtype)
        else:
            dt = _real_dtype(A)
            A = _as_real(A, dt)
        self._dt, self._keep, self.rV = dt, A, 0
        R, suf = _CT[dt]
        f = getattr(self._lib, f"pfdr_b200_cp_operator_create_{suf}")
        f.argtypes = [_C.POINTER(_C.c_void_p), _C.c_int, _C.c_int, _C.POINTER(R)]
        self._chk(f(_C.byref(self._h), self.V, self.N, _ptr(A, R)))
        for name, n in (("set_components", 3), ("reduce", 6), ("gradient", 4), ("residual", 4)):
            getattr(self._lib, f"pfdr_b200_cp_{name}").argtypes = [_C.c_void_p] * n
        self._lib.pfdr_b200_cp_set_components.argtypes = [_C.c_void_p, _C.c_int, _C.POINTER(_C.c_int)]
        self._lib.pfdr_b200_cp_reduce.argtypes = [_C.c_void_p, _C.c_void_p, _C.c_int, _C.c_void_p, _C.c_void_p, _C.c_void_p]

    def _chk(self, rc):
        if rc != 0:
            raise PfdrError(f"libpfdr_b200 status {rc}: {self._lib.pfdr_b200_last_error().decode()}")

    def _vp(self, a):
        return None if a is None else _C.cast(_ptr(_as_real(a, self._dt), _CT[self._dt][0]), _C.c_void_p)

    def set_components(self, Cv, rV):
        Cv = _as_int(Cv)
        self._chk(self._lib.pfdr_b200_cp_set_components(self._h, int(rV), _ptr(Cv, _C.c_int)))
        self.rV = int(rV)

    def reduce(self, Y, preAt=True, want_rA=True, on_device=False):
        """-> (rA or None, rAA or None, rY or None): numpy arrays (rA, rAA Fortran-ordered), or with on_device torch CUDA
        tensors holding the same column-major data (rA as (rV, N), rAA as (rV, rV): transposes of the matrices)"""
        N, rV, dt = self.N, self.rV, self._dt
        pre = bool(preAt) or N <= 0
        if on_device:
            import torch
            tdt = torch.float32 if dt == np.float32 else torch.float64
            new = lambda *shape: torch.empty(*shape, dtype=tdt, device="cuda")
            rA = new(rV, N) if (N > 0 and want_rA) else None
            rAA = (new(rV) if N == 0 else new(rV, rV)) if pre else None
            rY = new(rV) if pre else None
        else:
            rA = np.zeros((N, rV), dt, order="F") if (N > 0 and want_rA) else None
            rAA = (np.zeros(rV, dt) if N == 0 else np.zeros((rV, rV), dt, order="F")) if pre else None
            rY = np.zeros(rV, dt) if pre else None
        Y = None if Y is None else _as_real(Y, dt)
        self._chk(self._lib.pfdr_b200_cp_reduce(self._h, self._vp(Y), int(pre), self._vp(rA), self._vp(rAA), self._vp(rY)))
        return rA, rAA, rY

    def gradient(self, rX, YorR=None, on_device=False):
        if on_device:
            import torch
            out = torch.empty(self.V, dtype=torch.float32 if self._dt == np.float32 else torch.float64, device="cuda")
        else:
            out = np.zeros(self.V, self._dt)
        self._chk(self._lib.pfdr_b200_cp_gradient(self._h, self._vp(rX),
                                                  self._vp(YorR), self._vp(out)))
        return out

    def residual(self, rX, Y):
        out = np.zeros(self.N, self._dt)
        self._chk(self._lib.pfdr_b200_cp_residual(self._h, self._vp(rX), self._vp(Y), self._vp(out)))
        return out

    def close(self):
        if self._h:
            self._lib.pfdr_b200_cp_operator_destroy.argtypes = [_C.c_void_p]
            self._lib.pfdr_b200_cp_operator_destroy(self._h)
            self._h = _C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- graph passes of the cut-pursuit's outer iteration (csrc/cpgraph.cu; reference src/CP_PFDR_graph_quadratic_d1_l1.cpp) ----
def _u8(a):
    return a if hasattr(a, "data_ptr") else np.ascontiguousarray(a, dtype=np.uint8)


def _u8_ptr(a):
    PB = _C.POINTER(_C.c_ubyte)
    if hasattr(a, "data_ptr"):
        import torch
        if a.dtype != torch.uint8 or not a.is_cuda or not a.is_contiguous():
            raise TypeError("edge flags on the device must be a contiguous torch.uint8 CUDA tensor")
        return _C.cast(_C.c_void_p(a.data_ptr()), PB)
    return a.ctypes.data_as(PB)


def cp_components(V, Eu, Ev, active, on_device=False):
    """connected components of the non-active edges, numbered by smallest vertex (:570-600) -> (Cv, rV); inputs numpy or
    torch CUDA tensors; Cv a numpy array, or a torch CUDA tensor with on_device"""
    lib = _lib.load()
    Eu, Ev, active = _as_int(Eu), _as_int(Ev), _u8(active)
    rV = _C.c_int(0)
    if on_device:
        import torch
        Cv = torch.empty(V, dtype=torch.int32, device="cuda")
    else:
        Cv = np.zeros(V, np.int32)
    PI, PB = _C.POINTER(_C.c_int), _C.POINTER(_C.c_ubyte)
    f = lib.pfdr_b200_cp_components
    f.argtypes = [_C.c_int, _C.c_int, PI, PI, PB, PI, PI]
    rc = f(int(V), int(Eu.shape[0]), _ptr(Eu, _C.c_int), _ptr(Ev, _C.c_int), _u8_ptr(active), _ptr(Cv, _C.c_int), _C.byref(rV))
    if rc != 0:
        raise PfdrError(f"libpfdr_b200 status {rc}: {lib.pfdr_b200_last_error().decode()}")
    return Cv, rV.value


def cp_reduced_edges(V, Eu, Ev, active, La_d1, Cv, rV, eps=None, capacity=None):
    """reduced connectivity (:603-661) -> (rEu, rEv, rLa_d1), sorted by (rEu, rEv)"""
    lib = _lib.load()
    dt = _real_dtype(La_d1)
    R, suf = _CT[dt]
    Eu, Ev, active, Cv, La_d1 = _as_int(Eu), _as_int(Ev), _u8(active), _as_int(Cv), _as_real(La_d1, dt)
    E = int(Eu.shape[0])
    cap = int(capacity) if capacity is not None else E + int(rV)
    eps = float(np.finfo(dt).eps) if eps is None else float(eps)
    rEu, rEv, rLa, rE = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, dt), _C.c_int(0)
    PI, PB, PR = _C.POINTER(_C.c_int), _C.POINTER(_C.c_ubyte), _C.POINTER(R)
    f = getattr(lib, f"pfdr_b200_cp_reduced_edges_{suf}")
    f.argtypes = [_C.c_int, _C.c_int, PI, PI, PB, PR, PI, _C.c_int, R, _C.c_int, PI, PI, PI, PR]
    rc = f(int(V), E, _ptr(Eu, _C.c_int), _ptr(Ev, _C.c_int), _u8_ptr(active), _ptr(La_d1, R), _ptr(Cv, _C.c_int), int(rV), eps,
           cap, _C.byref(rE), _ptr(rEu, _C.c_int), _ptr(rEv, _C.c_int), _ptr(rLa, R))
    if rc != 0:
        raise PfdrError(f"libpfdr_b200 status {rc}: {lib.pfdr_b200_last_error().decode()}")
    n = rE.value
    return rEu[:n].copy(), rEv[:n].copy(), rLa[:n].copy()


def cp_merge(Eu, Ev, Cv, rX, CP_difTol, active, eps=None):
    """merge step (:863-886): clears `active` (uint8 array, in place) between components of relatively equal value;
    returns the number of edges deactivated"""
    lib = _lib.load()
    dt = _real_dtype(rX)
    R, suf = _CT[dt]
    if not hasattr(active, "data_ptr") and not (isinstance(active, np.ndarray) and active.dtype == np.uint8 and active.flags.c_contiguous):
        raise TypeError("active is updated in place: it must be a contiguous uint8 array (or torch.uint8 CUDA tensor)")
    Eu, Ev, Cv, rX = _as_int(Eu), _as_int(Ev), _as_int(Cv), _as_real(rX, dt)
    eps = float(np.finfo(dt).eps) if eps is None else float(eps)
    n = _C.c_int(0)
    PI, PB, PR = _C.POINTER(_C.c_int), _C.POINTER(_C.c_ubyte), _C.POINTER(R)
    f = getattr(lib, f"pfdr_b200_cp_merge_{suf}")
    f.argtypes = [_C.c_int, _C.c_int, PI, PI, PI, _C.c_int, PR, R, R, PB, PI]
    rc = f(int(Cv.shape[0]), int(Eu.shape[0]), _ptr(Eu, _C.c_int), _ptr(Ev, _C.c_int), _ptr(Cv, _C.c_int), int(rX.shape[0]), _ptr(rX, R),
           float(CP_difTol), eps, _u8_ptr(active), _C.byref(n))
    if rc != 0:
        raise PfdrError(f"libpfdr_b200 status {rc}: {lib.pfdr_b200_last_error().decode()}")
    return n.value
