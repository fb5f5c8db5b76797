"""CPU-only checks of the drop-in boundary: the CUDA library loads, exports every symbol that
include/pfdr_b200.h declares, and fails loudly (no CPU fallback) when no device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from cp_pfdr_graph_d1_b200 import _lib


def declared_symbols():
    src = open(_lib.HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pfdr_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/pfdr_b200.h but not exported"
    assert lib.pfdr_b200_version() >= 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from cp_pfdr_graph_d1_b200 import api
    X = np.zeros(4, np.float32)
    with pytest.raises(api.PfdrError, match="no usable CUDA device"):
        api.PFDR_graph_quadratic_d1_l1(X, np.ones(4, np.float32), None, np.array([0, 1, 2], np.int32),
                                       np.array([1, 2, 3], np.int32), np.ones(3, np.float32), N=0, itMax=3)
    with pytest.raises(api.PfdrError):
        api.proj_simplex_metric(np.ones((2, 3), np.float32), np.ones((1, 3), np.float32))


def test_stats_struct_layout():
    s = _lib.Stats()
    assert _lib.load().pfdr_b200_get_last_stats(C.byref(s)) == 0
    assert set(s.asdict()) >= {"setup_ms", "loop_ms", "launches", "iterations"}


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: nothing under cp_pfdr_graph_d1_b200/ may reference it"""
    pkg = os.path.dirname(_lib.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "liboracle" not in txt and "libref" not in txt and "from oracle" not in txt, f


def test_binder_rejects_wrong_element_types():
    """pointers are passed raw: an array of another element type, or a non-contiguous one, must be refused, not reinterpreted"""
    import ctypes as C
    import numpy as np
    import pytest
    from cp_pfdr_graph_d1_b200 import _abi
    with pytest.raises(TypeError):
        _abi._ptr(np.zeros(4, np.float64), C.c_float)
    with pytest.raises(TypeError):
        _abi._ptr(np.zeros(4, np.int64), C.c_int)
    with pytest.raises(TypeError):
        _abi._ptr(np.zeros((4, 6), np.float32)[::2, ::2], C.c_float)
    with pytest.raises(TypeError):
        _abi._real_dtype(np.zeros(4, np.float16))
    import torch
    with pytest.raises(TypeError):
        _abi._ptr(torch.zeros(4, dtype=torch.float32), C.c_float)          # a CPU tensor is not a device pointer
    with pytest.raises(TypeError):
        _abi._real_dtype(torch.zeros(4, dtype=torch.bfloat16))
    assert _abi._ptr(np.zeros(4, np.float32), C.c_float) is not None and _abi._ptr(None, C.c_float) is None


def test_libcp_signature_matches_the_reference_binding():
    """keyword names and defaults of python/CP_quadratic_l1_py.cpp:68-76 / :111-116"""
    import inspect
    from cp_pfdr_graph_d1_b200 import libCP
    sig = inspect.signature(libCP.CP_quadratic_l1)
    assert list(sig.parameters) == ["obs", "source", "target", "edge_weight", "A", "l1_weight", "positivity", "PFDR_rho",
                                    "PFDR_condMin", "CP_difTol", "PFDR_difRcd", "PFDR_difTol", "CP_itMax", "PFDR_itMax", "verbose"]
    d = {k: v.default for k, v in sig.parameters.items() if v.default is not inspect.Parameter.empty}
    assert d == dict(positivity=0, PFDR_rho=1.0, PFDR_condMin=1e-3, CP_difTol=1e-3, PFDR_difRcd=0.0, PFDR_difTol=1e-4,
                     CP_itMax=10, PFDR_itMax=10000, verbose=0)


def test_libcp_fails_loudly_without_its_library(monkeypatch, tmp_path):
    from cp_pfdr_graph_d1_b200 import libCP
    monkeypatch.setenv("PFDR_B200_CP_LIB", str(tmp_path / "absent.so"))
    monkeypatch.setattr(libCP, "_lib", None)
    with pytest.raises(ImportError):
        libCP.CP_quadratic_l1(np.zeros(3), np.array([0, 1]), np.array([1, 2]), 0.1, 1.0, 0.0)


_STUB_C = r"""
#include <string.h>
double seen[16];
#define STUB(SUF, real)                                                                                         \
int cp_b200_quadratic_d1_l1_##SUF(int V, int E, int N, const real *Y, const real *A, const int *Eu,             \
    const int *Ev, const real *La_d1, const real *La_l1, int positivity, real CP_difTol, int CP_itMax,          \
    real PFDR_rho, real PFDR_condMin, real PFDR_difRcd, real PFDR_difTol, int PFDR_itMax, int verbose,          \
    int *Cv, real *rX, int *CP_it)                                                                              \
{                                                                                                               \
    double s[16] = {V, E, N, Y[V > 1 ? 1 : 0], A ? A[1] : -1.0, Eu[E - 1], Ev[E - 1], La_d1[E - 1], La_l1[V - 1], \
                    positivity, CP_difTol, CP_itMax, PFDR_rho, PFDR_difTol, PFDR_itMax, sizeof(real)};          \
    memcpy(seen, s, sizeof(s));                                                                                 \
    for (int v = 0; v < V; v++) Cv[v] = v & 1;                                                                  \
    rX[0] = (real) 1.5; rX[1] = (real) 2.5; *CP_it = 3;                                                         \
    return 2;                                                                                                   \
}
STUB(f32, float)
STUB(f64, double)
"""


def test_libcp_argument_handling_against_a_stub_gateway(monkeypatch, tmp_path):
    """the host logic of libCP.CP_quadratic_l1 (A modes, scalar overloads, integer edge types, element type) checked on the
    arguments a stub of csrc/cp_gateway.cpp receives - what the Boost module derives at python/CP_quadratic_l1_py.cpp:115-170"""
    import subprocess
    from cp_pfdr_graph_d1_b200 import libCP
    src, so = tmp_path / "stub.c", tmp_path / "libstub.so"
    src.write_text(_STUB_C)
    subprocess.check_call(["gcc", "-shared", "-fPIC", "-O1", "-o", str(so), str(src)])
    monkeypatch.setenv("PFDR_B200_CP_LIB", str(so))
    monkeypatch.setattr(libCP, "_lib", None)
    seen = lambda: np.array((C.c_double * 16).in_dll(libCP._load(), "seen"))
    Eu, Ev = np.array([0, 1, 2], np.uint32), np.array([1, 2, 3], np.int64)
    # matrix mode, float64, array weights
    obs = np.arange(3.0); A = np.arange(12.0).reshape(3, 4)
    Cv, rX = libCP.CP_quadratic_l1(obs, Eu, Ev, np.array([.1, .2, .3]), A, np.array([1., 2., 3., 4.]), positivity=1, PFDR_rho=1.5)
    assert Cv.tolist() == [0, 1, 0, 1] and rX.tolist() == [1.5, 2.5] and rX.dtype == np.float64
    s = seen()
    assert s[:3].tolist() == [4, 3, 3]                 # V = columns of A, E, N = len(obs)
    assert s[3] == 1.0 and s[4] == A[1, 0]              # A handed over column major: element 1 is A[1, 0]
    assert s[5:10].tolist() == [2, 3, .3, 4., 1] and s[12] == 1.5 and s[15] == 8
    assert s[10:12].tolist() == [1e-3, 10] and s[13:15].tolist() == [1e-4, 10000]        # the Boost module's defaults
    # diagonal mode, float32, scalar weights: V = N, N -> 0, observation and operator pre-multiplied
    obs = np.array([1., 2., 3., 4.], np.float32); a = np.array([2., 3., 4., 5.], np.float32)
    Cv, rX = libCP.CP_quadratic_l1(obs, Eu, Ev, 0.25, a, 0.5)
    s = seen()
    assert rX.dtype == np.float32 and s[:3].tolist() == [4, 3, 0] and s[15] == 4
    assert s[3] == 2. * 3. and s[4] == 3. * 3. and s[7] == 0.25 and s[8] == 0.5 and s[9] == 0
    # identity mode: A a scalar
    libCP.CP_quadratic_l1(obs, Eu, Ev, 0.25, 1.0, np.float32(0.5))
    s = seen()
    assert s[:3].tolist() == [4, 3, 0] and s[3] == 2. and s[4] == -1.0
    monkeypatch.setattr(libCP, "_lib", None)
