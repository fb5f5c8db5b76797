"""ctypes plumbing shared by every implementation of the PFDR C signatures.

The product library (lib/libpfdr_b200.so, prefix ``pfdr_b200_``) and the two CPU checkers that the
test infrastructure under oracle/ builds (prefixes ``oracle_`` and ``ref_``; this package never loads
them) all expose the reference's argument lists
(/root/reference/include/PFDR_graph_quadratic_d1_l1.hpp:36-42, ..._bounds.hpp:34-40,
PFDR_graph_loss_d1_simplex.hpp:24-30, proj_simplex.hpp:33-35, operator_norm_matrix.hpp:12-14),
so one binder serves all three and the parity tests call them through identical code.
"""
import ctypes as C
from collections import namedtuple

import numpy as np

SCAL, DIAG = 0, 1

Result = namedtuple("Result", "X it Obj Dif")

_CT = {np.dtype(np.float32): (C.c_float, "f32"), np.dtype(np.float64): (C.c_double, "f64")}


class PfdrError(RuntimeError):
    pass


_NP_OF = {C.c_float: np.float32, C.c_double: np.float64, C.c_int: np.int32}


def _ptr(a, ctype):
    """host numpy array / torch cuda tensor / None -> pointer argument.  The array must already have the element type the
    C signature expects and be contiguous: a pointer to anything else would be reinterpreted silently."""
    if a is None:
        return None
    want = np.dtype(_NP_OF[ctype])
    if isinstance(a, np.ndarray):
        if a.dtype != want:
            raise TypeError(f"array of {a.dtype} where the C signature takes {want}")
        if not (a.flags.c_contiguous or a.flags.f_contiguous):
            raise TypeError("array must be contiguous (C- or Fortran-ordered)")
        return a.ctypes.data_as(C.POINTER(ctype))
    if hasattr(a, "data_ptr"):          # torch tensor (device pointers are accepted by libpfdr_b200 only)
        import torch
        tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64, np.dtype(np.int32): torch.int32}[want]
        if a.dtype != tdt:
            raise TypeError(f"tensor of {a.dtype} where the C signature takes {want}")
        if not a.is_cuda:
            raise TypeError("torch tensors must live on a CUDA device (pass numpy arrays for host data)")
        if not (a.is_contiguous() or a.t().is_contiguous()):
            raise TypeError("tensor must be contiguous")
        return C.cast(C.c_void_p(a.data_ptr()), C.POINTER(ctype))
    raise TypeError(f"unsupported array type {type(a)}")


def _real_dtype(a):
    if isinstance(a, np.ndarray):
        if a.dtype not in _CT:
            raise TypeError(f"the in/out array must be float32 or float64, not {a.dtype}")
        return a.dtype
    import torch
    if a.dtype == torch.float32:
        return np.dtype(np.float32)
    if a.dtype == torch.float64:
        return np.dtype(np.float64)
    raise TypeError(f"the in/out tensor must be float32 or float64, not {a.dtype}")


def _as_real(a, dt):
    """1-D arrays: contiguous; 2-D matrices: COLUMN-major (the reference's layout for A, A^tA; a (V, K) label
    array is passed C-ordered on purpose and handled by the simplex binder itself)"""
    if a is None or not isinstance(a, np.ndarray):
        return a
    if a.ndim == 2:
        return np.asfortranarray(a, dtype=dt)
    return np.ascontiguousarray(a, dtype=dt)


def _as_rows(a, dt):
    if a is None or not isinstance(a, np.ndarray):
        return a
    return np.ascontiguousarray(a, dtype=dt)


def _as_int(a):
    if not isinstance(a, np.ndarray):
        return a
    return np.ascontiguousarray(a, dtype=np.int32)


class Binder:
    """Binds one shared library exposing ``<prefix><routine>_{f32,f64}``."""

    def __init__(self, lib, prefix, returns_status, opnorm_seeded=False):
        self.lib, self.prefix, self.returns_status = lib, prefix, returns_status
        self.opnorm_seeded = opnorm_seeded

    def _fn(self, name, suf, argtypes, restype=None):
        f = getattr(self.lib, f"{self.prefix}{name}_{suf}")
        f.argtypes = argtypes
        f.restype = C.c_int if self.returns_status else restype
        return f

    def _check(self, rc):
        if self.returns_status and rc != 0:
            msg = self.lib.pfdr_b200_last_error()
            raise PfdrError(f"libpfdr_b200 status {rc}: {msg.decode() if msg else ''}")

    # -- quadratic family -------------------------------------------------------------------
    def _quad(self, name, X, Y, A, Eu, Ev, La_d1, mid_args, mid_types, Ltype, L, rho, condMin, difRcd, difTol,
              itMax, N, want_obj, want_dif, verbose):
        dt = _real_dtype(X)
        R, suf = _CT[dt]
        P, PI = C.POINTER(R), C.POINTER(C.c_int)
        if isinstance(X, np.ndarray) and not X.flags.c_contiguous:
            raise TypeError("X is updated in place: it must be a C-contiguous array")
        V = int(X.shape[0])
        E = int(Eu.shape[0])
        Y, A, La_d1, L = (_as_real(a, dt) for a in (Y, A, La_d1, L))
        Eu, Ev = _as_int(Eu), _as_int(Ev)
        if N is None:
            N = 0 if (A is None or A.ndim == 1) else int(A.shape[0] if isinstance(A, np.ndarray) else A.shape[0])
        it = C.c_int(0)
        Obj = np.zeros(itMax + 1, dt) if want_obj else None
        Dif = np.zeros(max(itMax, 1), dt) if want_dif else None
        f = self._fn(name, suf, [C.c_int, C.c_int, C.c_int, P, P, P, PI, PI, P] + mid_types +
                     [C.c_int, P, R, R, R, R, C.c_int, PI, P, P, C.c_int])
        rc = f(V, E, N, _ptr(X, R), _ptr(Y, R), _ptr(A, R), _ptr(Eu, C.c_int), _ptr(Ev, C.c_int), _ptr(La_d1, R),
               *mid_args(R, dt), Ltype, _ptr(L, R), rho, condMin, difRcd, difTol, itMax, C.byref(it),
               _ptr(Obj, R), _ptr(Dif, R), verbose)
        self._check(rc)
        n = it.value
        return Result(X, n, None if Obj is None else Obj[:n + 1], None if Dif is None else Dif[:n])

    def PFDR_graph_quadratic_d1_l1(self, X, Y, A, Eu, Ev, La_d1, La_l1=None, positivity=0, Ltype=SCAL, L=None,
                                   rho=1.5, condMin=1e-3, difRcd=0.0, difTol=1e-5, itMax=1000, N=None,
                                   want_obj=False, want_dif=False, verbose=0):
        """X is updated in place (and returned).  A: None (identity), 1-D (diagonal, N=0), 2-D Fortran-ordered
        N-by-V (pass N>0) or V-by-V A^tA (pass N=-V)."""
        return self._quad("PFDR_graph_quadratic_d1_l1" if self.prefix != "pfdr_b200_" else "quadratic_d1_l1",
                          X, Y, A, Eu, Ev, La_d1,
                          lambda R, dt: (_ptr(_as_real(La_l1, dt), R), int(positivity)),
                          [C.POINTER(_CT[_real_dtype(X)][0]), C.c_int],
                          Ltype, L, rho, condMin, difRcd, difTol, itMax, N, want_obj, want_dif, verbose)

    def PFDR_graph_quadratic_d1_bounds(self, X, Y, A, Eu, Ev, La_d1, min=-np.inf, max=np.inf, Ltype=SCAL, L=None,
                                       rho=1.5, condMin=1e-3, difRcd=0.0, difTol=1e-5, itMax=1000, N=None,
                                       want_obj=False, want_dif=False, verbose=0):
        R0 = _CT[_real_dtype(X)][0]
        return self._quad("PFDR_graph_quadratic_d1_bounds" if self.prefix != "pfdr_b200_" else "quadratic_d1_bounds",
                          X, Y, A, Eu, Ev, La_d1, lambda R, dt: (float(min), float(max)), [R0, R0],
                          Ltype, L, rho, condMin, difRcd, difTol, itMax, N, want_obj, want_dif, verbose)

    # -- simplex family ---------------------------------------------------------------------
    def PFDR_graph_loss_d1_simplex(self, P, Q, al, Eu, Ev, La_d1, La_f=None, rho=1.5, condMin=1e-2, difRcd=0.0,
                                   difTol=1e-4, itMax=500, want_obj=False, want_dif=False, verbose=0):
        """P, Q: (V, K) C-ordered arrays (== the reference's K-by-V column-major).  P in/out."""
        dt = _real_dtype(P)
        R, suf = _CT[dt]
        PR, PI = C.POINTER(R), C.POINTER(C.c_int)
        if isinstance(P, np.ndarray) and not P.flags.c_contiguous:
            raise TypeError("P is updated in place: it must be a C-contiguous (V, K) array")
        V, K = int(P.shape[0]), int(P.shape[1])
        E = int(Eu.shape[0])
        Q, La_d1, La_f = (_as_rows(a, dt) for a in (Q, La_d1, La_f))
        Eu, Ev = _as_int(Eu), _as_int(Ev)
        it = C.c_int(0)
        Obj = np.zeros(itMax + 1, dt) if want_obj else None
        Dif = np.zeros(max(itMax, 1), dt) if want_dif else None
        name = "PFDR_graph_loss_d1_simplex" if self.prefix != "pfdr_b200_" else "loss_d1_simplex"
        f = self._fn(name, suf, [C.c_int, C.c_int, C.c_int, R, PR, PR, PR, PI, PI, PR, R, R, R, R, C.c_int, PI,
                                 PR, PR, C.c_int])
        rc = f(K, V, E, al, _ptr(La_f, R), _ptr(P, R), _ptr(Q, R), _ptr(Eu, C.c_int), _ptr(Ev, C.c_int),
               _ptr(La_d1, R), rho, condMin, difRcd, difTol, itMax, C.byref(it), _ptr(Obj, R), _ptr(Dif, R), verbose)
        self._check(rc)
        n = it.value
        return Result(P, n, None if Obj is None else Obj[:n + 1], None if Dif is None else Dif[:n])

    # -- helpers ----------------------------------------------------------------------------
    def proj_simplex_metric(self, X, M, A=None):
        """X: (N, D) C-ordered, projected in place; M: (nm, D); A: (na,) target sums (default 1)."""
        dt = _real_dtype(X)
        R, suf = _CT[dt]
        PR = C.POINTER(R)
        N, D = int(X.shape[0]), int(X.shape[1])
        M = _as_rows(M, dt)
        A = np.ones(1, dt) if A is None else _as_rows(np.atleast_1d(A), dt)
        f = self._fn("proj_simplex_metric", suf, [PR, PR, C.c_int, C.c_int, C.c_int, PR, C.c_int])
        self._check(f(_ptr(X, R), _ptr(M, R), D, N, int(M.shape[0]), _ptr(A, R), int(A.shape[0])))
        return X

    def operator_norm_matrix(self, A, M=None, N=None, nTol=1e-3, itMax=100, nbInit=10, verbose=0, seed=1):
        """A: Fortran-ordered M-by-N (or the symmetrised matrix with M or N = 0)."""
        dt = _real_dtype(A)
        R, suf = _CT[dt]
        PR = C.POINTER(R)
        if M is None:
            M, N = int(A.shape[0]), int(A.shape[1])
        if self.returns_status:
            out = R(0)
            f = self._fn("operator_norm_matrix", suf, [C.c_int, C.c_int, PR, R, C.c_int, C.c_int, C.c_int, PR])
            self._check(f(M, N, _ptr(A, R), nTol, itMax, nbInit, verbose, C.byref(out)))
            return float(out.value)
        last = C.c_ulonglong if self.opnorm_seeded else C.c_int
        f = self._fn("operator_norm_matrix", suf, [C.c_int, C.c_int, PR, R, C.c_int, C.c_int, last], restype=R)
        return float(f(M, N, _ptr(A, R), nTol, itMax, nbInit, seed if self.opnorm_seeded else verbose))
