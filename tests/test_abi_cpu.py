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
