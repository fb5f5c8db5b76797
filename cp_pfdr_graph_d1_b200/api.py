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
