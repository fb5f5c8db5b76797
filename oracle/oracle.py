"""TEST INFRASTRUCTURE ONLY: ctypes access to the two CPU checkers.

  oracle()    -> Binder over oracle/liboracle.so  (serial C restatement, oracle/pfdr_oracle_impl.h)
  reference() -> Binder over oracle/_ref/libref.so (the unmodified reference sources compiled by
                 oracle/build_ref.sh), or None when it has not been built.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from cp_pfdr_graph_d1_b200._abi import Binder  # noqa: E402

ORACLE_SO = os.path.join(_HERE, "liboracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libref.so")


def build(ref=True):
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    if ref:
        subprocess.check_call([os.path.join(_HERE, "build_ref.sh")], stderr=subprocess.DEVNULL)


_o = _r = None


def oracle():
    global _o
    if _o is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        _o = Binder(C.CDLL(ORACLE_SO), "oracle_", returns_status=False, opnorm_seeded=True)
    return _o


def reference():
    global _r
    if _r is None and os.path.exists(REF_SO):
        lib = C.CDLL(REF_SO)
        lib.ref_num_procs.restype = C.c_int
        _r = Binder(lib, "ref_", returns_status=False)
    return _r


# ---------------------------------------------------------------------------------------------------
# cut-pursuit gateways (the reference's UNMODIFIED CP_*.cpp): libref.so runs them on the reference's own
# PFDR, libcp_b200*.so on the B200 PFDR through the drop-in shim (oracle/build_ref.sh).
# ---------------------------------------------------------------------------------------------------
import numpy as np  # noqa: E402

CP_B200_SO = os.path.join(_HERE, "_ref", "libcp_b200.so")
CP_B200_REFNORM_SO = os.path.join(_HERE, "_ref", "libcp_b200_refnorm.so")


class CutPursuit:
    def __init__(self, path):
        self.lib = C.CDLL(path)

    def _common(self, dt):
        R = C.c_float if dt == np.float32 else C.c_double
        return R, "f32" if dt == np.float32 else "f64", C.POINTER(R), C.POINTER(C.c_int)

    def _finish(self, V, rV, Cv, rXp, dt, CP_it, Time, Obj, Dif):
        n = rV.value
        rX = np.ctypeslib.as_array(rXp, shape=(n,)).copy() if n else np.zeros(0, dt)
        self.lib.ref_free.argtypes = [C.c_void_p]
        self.lib.ref_free(C.cast(rXp, C.c_void_p))
        k = CP_it.value
        return dict(rV=n, Cv=Cv, rX=rX, CP_it=k, Time=Time[:k + 1], Obj=Obj[:k + 1], Dif=Dif[:k])

    def quadratic_d1_l1(self, Y, A, Eu, Ev, La_d1, La_l1, positivity, N, CP_difTol=1e-4, CP_itMax=15, rho=1.5,
                        condMin=1e-3, difRcd=0.0, difTol=1e-7, itMax=10000, verbose=0, duplex=0):
        dt = Y.dtype.type
        R, suf, PR, PI = self._common(dt)
        V, E = (A.shape[1] if N > 0 else Y.shape[0]), Eu.shape[0]
        f = getattr(self.lib, f"ref_CP_PFDR_graph_quadratic_d1_l1_{suf}")
        f.restype = None
        f.argtypes = [C.c_int, C.c_int, C.c_int, PI, PI, C.POINTER(PR), PR, PR, PI, PI, PR, PR, C.c_int, R, C.c_int,
                      PI, R, R, R, R, C.c_int, C.POINTER(C.c_double), PR, PR, C.c_int, C.c_int]
        rV, CP_it, rXp = C.c_int(0), C.c_int(0), PR()
        Cv = np.zeros(V, np.int32)
        Time, Obj, Dif = np.zeros(CP_itMax + 1), np.zeros(CP_itMax + 1, dt), np.zeros(max(CP_itMax, 1), dt)
        p = lambda a, t=R: None if a is None else a.ctypes.data_as(C.POINTER(t))
        f(V, E, N, C.byref(rV), p(Cv, C.c_int), C.byref(rXp), p(Y), p(A), p(Eu, C.c_int), p(Ev, C.c_int), p(La_d1),
          p(La_l1), positivity, CP_difTol, CP_itMax, C.byref(CP_it), rho, condMin, difRcd, difTol, itMax,
          Time.ctypes.data_as(C.POINTER(C.c_double)), p(Obj), p(Dif), verbose, duplex)
        return self._finish(V, rV, Cv, rXp, dt, CP_it, Time, Obj, Dif)

    def quadratic_d1_bounds(self, Y, A, Eu, Ev, La_d1, lo, hi, N, CP_difTol=1e-4, CP_itMax=10, rho=1.5, condMin=1e-3,
                            difRcd=0.0, difTol=1e-7, itMax=10000, verbose=0):
        dt = Y.dtype.type
        R, suf, PR, PI = self._common(dt)
        V, E = (A.shape[1] if N > 0 else Y.shape[0]), Eu.shape[0]
        f = getattr(self.lib, f"ref_CP_PFDR_graph_quadratic_d1_bounds_{suf}")
        f.restype = None
        f.argtypes = [C.c_int, C.c_int, C.c_int, PI, PI, C.POINTER(PR), PR, PR, PI, PI, PR, R, R, R, C.c_int, PI, R, R,
                      R, R, C.c_int, C.POINTER(C.c_double), PR, PR, C.c_int]
        rV, CP_it, rXp = C.c_int(0), C.c_int(0), PR()
        Cv = np.zeros(V, np.int32)
        Time, Obj, Dif = np.zeros(CP_itMax + 1), np.zeros(CP_itMax + 1, dt), np.zeros(max(CP_itMax, 1), dt)
        p = lambda a, t=R: None if a is None else a.ctypes.data_as(C.POINTER(t))
        f(V, E, N, C.byref(rV), p(Cv, C.c_int), C.byref(rXp), p(Y), p(A), p(Eu, C.c_int), p(Ev, C.c_int), p(La_d1),
          lo, hi, CP_difTol, CP_itMax, C.byref(CP_it), rho, condMin, difRcd, difTol, itMax,
          Time.ctypes.data_as(C.POINTER(C.c_double)), p(Obj), p(Dif), verbose)
        return self._finish(V, rV, Cv, rXp, dt, CP_it, Time, Obj, Dif)

    def loss_d1_simplex(self, Q, al, Eu, Ev, La_d1, CP_difTol=1e-3, CP_itMax=10, rho=1.5, condMin=1e-2, difRcd=0.0,
                        difTol=1e-4, itMax=1000, verbose=0):
        dt = Q.dtype.type
        R, suf, PR, PI = self._common(dt)
        V, K, E = Q.shape[0], Q.shape[1], Eu.shape[0]
        f = getattr(self.lib, f"ref_CP_PFDR_graph_loss_d1_simplex_{suf}")
        f.restype = None
        f.argtypes = [C.c_int, C.c_int, C.c_int, R, PI, PI, C.POINTER(PR), PR, PI, PI, PR, R, C.c_int, PI, R, R, R, R,
                      C.c_int, C.POINTER(C.c_double), PR, PR, C.c_int]
        rV, CP_it, rPp = C.c_int(0), C.c_int(0), PR()
        Cv = np.zeros(V, np.int32)
        Time, Obj, Dif = np.zeros(CP_itMax + 1), np.zeros(CP_itMax + 1, dt), np.zeros(max(CP_itMax, 1), dt)
        p = lambda a, t=R: None if a is None else a.ctypes.data_as(C.POINTER(t))
        f(K, V, E, al, C.byref(rV), p(Cv, C.c_int), C.byref(rPp), p(Q), p(Eu, C.c_int), p(Ev, C.c_int), p(La_d1),
          CP_difTol, CP_itMax, C.byref(CP_it), rho, condMin, difRcd, difTol, itMax,
          Time.ctypes.data_as(C.POINTER(C.c_double)), p(Obj), p(Dif), verbose)
        n = rV.value
        rP = np.ctypeslib.as_array(rPp, shape=(n, K)).copy()
        self.lib.ref_free.argtypes = [C.c_void_p]
        self.lib.ref_free(C.cast(rPp, C.c_void_p))
        k = CP_it.value
        return dict(rV=n, Cv=Cv, rX=rP, CP_it=k, Time=Time[:k + 1], Obj=Obj[:k + 1], Dif=Dif[:k])


def cp_reference():
    return CutPursuit(REF_SO) if os.path.exists(REF_SO) else None


def cp_b200(refnorm=False):
    path = CP_B200_REFNORM_SO if refnorm else CP_B200_SO
    return CutPursuit(path) if os.path.exists(path) else None


# ---------------------------------------------------------------------------------------------------
# reduced-problem construction of the cut-pursuit (oracle/cp_reduce_oracle_impl.h)
# ---------------------------------------------------------------------------------------------------
def component_lists(Cv, rV):
    """(Vc, rVc): vertices grouped by component, ascending vertex number inside a component"""
    Cv = np.asarray(Cv, np.int32)
    Vc = np.argsort(Cv, kind="stable").astype(np.int32)
    rVc = np.concatenate([[0], np.cumsum(np.bincount(Cv, minlength=rV))]).astype(np.int32)
    return Vc, rVc


class CpReduceOracle:
    """ctypes access to oracle_cp_{gradient,reduce,residual,equilibrate}_{f32,f64}"""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        self.lib = C.CDLL(ORACLE_SO)

    @staticmethod
    def _t(dt):
        return (C.c_float, "f32") if np.dtype(dt) == np.float32 else (C.c_double, "f64")

    @staticmethod
    def _p(a, t):
        return None if a is None else np.ascontiguousarray(a).ctypes.data_as(C.POINTER(t)) if a.ndim == 1 else a.ctypes.data_as(C.POINTER(t))

    def gradient(self, V, N, A, YorR, rX, Cv, Vc, rVc):
        dt = rX.dtype
        R, suf = self._t(dt)
        f = getattr(self.lib, f"oracle_cp_gradient_{suf}")
        f.restype = None
        out = np.zeros(V, dt)
        p, I = self._p, C.c_int
        f.argtypes = [I, I, I] + [C.POINTER(R)] * 3 + [C.POINTER(I)] * 3 + [C.POINTER(R)]
        f(V, N, int(rX.shape[0]), p(A, R), p(YorR, R), p(rX, R), p(Cv, I), p(Vc, I), p(rVc, I), p(out, R))
        return out

    def reduce(self, V, N, A, Y, Vc, rVc, preAt):
        dt = Y.dtype
        R, suf = self._t(dt)
        rV = int(rVc.shape[0]) - 1
        f = getattr(self.lib, f"oracle_cp_reduce_{suf}")
        f.restype = None
        rA = np.zeros((N, rV), dt, order="F") if N > 0 else None
        rAA = (np.zeros(rV, dt) if N == 0 else np.zeros((rV, rV), dt, order="F")) if (preAt or N <= 0) else None
        rY = np.zeros(rV, dt) if (preAt or N <= 0) else None
        p, I = self._p, C.c_int
        f.argtypes = [I, I, I, C.POINTER(R), C.POINTER(R), C.POINTER(I), C.POINTER(I), I] + [C.POINTER(R)] * 3
        f(V, N, rV, p(A, R), p(Y, R), p(Vc, I), p(rVc, I), int(preAt), p(rA, R), p(rAA, R), p(rY, R))
        return rA, rAA, rY

    def residual(self, rA, rX, Y):
        dt = Y.dtype
        R, suf = self._t(dt)
        f = getattr(self.lib, f"oracle_cp_residual_{suf}")
        f.restype = None
        N, rV = rA.shape
        out = np.zeros(N, dt)
        p = self._p
        f.argtypes = [C.c_int, C.c_int] + [C.POINTER(R)] * 4
        f(N, rV, p(rA, R), p(rX, R), p(Y, R), p(out, R))
        return out

    def equilibrate(self, N, M, preAt):
        dt = M.dtype
        R, suf = self._t(dt)
        rV = M.shape[1]
        f = getattr(self.lib, f"oracle_cp_equilibrate_{suf}")
        f.restype = None
        S, d = np.zeros_like(M, order="F"), np.zeros(rV, dt)
        p = self._p
        f.argtypes = [C.c_int, C.c_int, C.c_int] + [C.POINTER(R)] * 3
        f(N, rV, int(preAt), p(M, R), p(S, R), p(d, R))
        return S, d


def reduced_graph(Cv, rV, Eu, Ev, La_d1, La_l1, eps):
    """reduced connectivity of src/CP_PFDR_graph_quadratic_d1_l1.cpp:603-661 as far as the minimiser of the reduced problem
    depends on it: one edge per pair of neighbouring components carrying the sum of the weights between them, a self-loop
    of weight eps on isolated components, per-component sums of the l1 weights.  (Edge order and self-loops on cut edges
    inside one component - which the reference also emits - only change the preconditioner, not the minimiser.)"""
    cu, cv = Cv[Eu].astype(np.int64), Cv[Ev].astype(np.int64)
    keep = (cu != cv) & (La_d1 != 0)
    lo, hi = np.minimum(cu, cv)[keep], np.maximum(cu, cv)[keep]
    key, inv = np.unique(lo * rV + hi, return_inverse=True)
    w = np.bincount(inv, weights=La_d1[keep].astype(np.float64), minlength=key.shape[0])
    rEu, rEv = (key // rV).astype(np.int32), (key % rV).astype(np.int32)
    touched = np.zeros(rV, bool)
    touched[rEu] = True
    touched[rEv] = True
    iso = np.nonzero(~touched)[0].astype(np.int32)
    rEu, rEv = np.concatenate([rEu, iso]), np.concatenate([rEv, iso])
    w = np.concatenate([w, np.full(iso.shape[0], eps)]).astype(La_d1.dtype)
    rl1 = None if La_l1 is None else np.bincount(Cv, weights=La_l1.astype(np.float64), minlength=rV).astype(La_l1.dtype)
    return rEu, rEv, w, rl1


class CpGraphOracle:
    """ctypes access to oracle_cp_components / oracle_cp_reduced_edges_* / oracle_cp_merge_* """

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        self.lib = C.CDLL(ORACLE_SO)

    def components(self, V, Eu, Ev, active):
        I, B = C.POINTER(C.c_int), C.POINTER(C.c_ubyte)
        Eu, Ev, active = (np.ascontiguousarray(Eu, np.int32), np.ascontiguousarray(Ev, np.int32), np.ascontiguousarray(active, np.uint8))
        Cv, Vc, rVc = np.zeros(V, np.int32), np.zeros(V, np.int32), np.zeros(V + 1, np.int32)
        f = self.lib.oracle_cp_components
        f.restype, f.argtypes = C.c_int, [C.c_int, C.c_int, I, I, B, I, I, I]
        rV = f(V, Eu.shape[0], Eu.ctypes.data_as(I), Ev.ctypes.data_as(I), active.ctypes.data_as(B), Cv.ctypes.data_as(I), Vc.ctypes.data_as(I),
               rVc.ctypes.data_as(I))
        return Cv, rV, Vc, rVc[:rV + 1].copy()

    def reduced_edges(self, V, Eu, Ev, active, La_d1, Cv, rV, eps):
        dt = La_d1.dtype
        R, suf = (C.c_float, "f32") if dt == np.float32 else (C.c_double, "f64")
        I, B, P = C.POINTER(C.c_int), C.POINTER(C.c_ubyte), C.POINTER(R)
        Eu, Ev, active, Cv = (np.ascontiguousarray(Eu, np.int32), np.ascontiguousarray(Ev, np.int32), np.ascontiguousarray(active, np.uint8),
                              np.ascontiguousarray(Cv, np.int32))
        E = Eu.shape[0]
        rEu, rEv, rLa = np.zeros(E + rV, np.int32), np.zeros(E + rV, np.int32), np.zeros(E + rV, dt)
        f = getattr(self.lib, f"oracle_cp_reduced_edges_{suf}")
        f.restype, f.argtypes = C.c_int, [C.c_int, C.c_int, I, I, B, P, I, C.c_int, R, I, I, P]
        n = f(V, E, Eu.ctypes.data_as(I), Ev.ctypes.data_as(I), active.ctypes.data_as(B), np.ascontiguousarray(La_d1).ctypes.data_as(P),
              Cv.ctypes.data_as(I), rV, eps, rEu.ctypes.data_as(I), rEv.ctypes.data_as(I), rLa.ctypes.data_as(P))
        return rEu[:n].copy(), rEv[:n].copy(), rLa[:n].copy()

    def merge(self, Eu, Ev, Cv, rX, tol, eps, active):
        dt = rX.dtype
        R, suf = (C.c_float, "f32") if dt == np.float32 else (C.c_double, "f64")
        I, B, P = C.POINTER(C.c_int), C.POINTER(C.c_ubyte), C.POINTER(R)
        Eu, Ev, Cv = (np.ascontiguousarray(Eu, np.int32), np.ascontiguousarray(Ev, np.int32), np.ascontiguousarray(Cv, np.int32))
        f = getattr(self.lib, f"oracle_cp_merge_{suf}")
        f.restype, f.argtypes = C.c_int, [C.c_int, I, I, I, P, R, R, B]
        return f(Eu.shape[0], Eu.ctypes.data_as(I), Ev.ctypes.data_as(I), Cv.ctypes.data_as(I), np.ascontiguousarray(rX).ctypes.data_as(P), tol, eps,
                 active.ctypes.data_as(B))
