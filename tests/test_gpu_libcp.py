"""SURVEY 8f N2 - the ctypes stand-in for the reference's Boost.Python module (cp_pfdr_graph_d1_b200/libCP.py): same call as
python/example_EEG_CP.py:46, compared with the reference's own cut-pursuit (oracle/_ref/libref.so) called with the arguments
the Boost module would derive (python/CP_quadratic_l1_py.cpp:115-170)."""
import os

import numpy as np
import pytest

from cp_pfdr_graph_d1_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture()
def libcp(gpu_api, monkeypatch):
    from oracle import oracle as O
    if O.cp_reference() is None or not os.path.exists(O.CP_B200_SO):
        pytest.skip("oracle/_ref cut-pursuit libraries not built (need /root/reference at build time)")
    monkeypatch.setenv("PFDR_B200_CP_LIB", O.CP_B200_SO)
    monkeypatch.setenv("PFDR_B200_OPNORM_TIME", "42")        # same power-method starts as oracle/_ref (tests/test_gpu_cp_dropin.py)
    from cp_pfdr_graph_d1_b200 import libCP
    return libCP, O.cp_reference()


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(float(np.linalg.norm(b)), 1e-300))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_matrix_mode_as_the_eeg_example(libcp, dtype):
    libCP, ref = libcp
    w = synth.eeg_like(4, 60, dtype)
    Cv, rX = libCP.CP_quadratic_l1(w["Y"], w["Eu"].astype(np.uint32), w["Ev"].astype(np.uint32), w["La_d1"], np.ascontiguousarray(w["A"]),
                                   w["La_l1"], positivity=1, PFDR_rho=1.5)
    r0 = ref.quadratic_d1_l1(w["Y"], w["A"], w["Eu"], w["Ev"], w["La_d1"], w["La_l1"], 1, w["N"], CP_difTol=1e-3, CP_itMax=10,
                             rho=1.5, condMin=1e-3, difRcd=0.0, difTol=1e-4, itMax=10000)
    assert Cv.dtype == np.int32 and rX.dtype == dtype and rX.shape[0] == Cv.max() + 1
    if dtype == np.float64:
        assert rX.shape[0] == r0["rV"] and np.array_equal(Cv, r0["Cv"])
        assert rel(rX, r0["rX"]) < 1e-5
    else:
        # float32 cut decisions need not coincide (the reference's own float32 run depends on its summation order as much):
        # the two piecewise-constant solutions must be equally good minimisers of the float64 objective
        def objective(x):
            x = x.astype(np.float64)
            r = w["Y"].astype(np.float64) - w["A"].astype(np.float64) @ x
            return 0.5 * r @ r + np.sum(w["La_d1"] * np.abs(x[w["Eu"]] - x[w["Ev"]])) + np.sum(w["La_l1"] * np.abs(x))
        f1, f0 = objective(rX[Cv]), objective(r0["rX"][r0["Cv"]])
        assert abs(f1 - f0) <= 2e-2 * abs(f0), (f1, f0)
        assert rX.min() >= 0


def test_diagonal_mode_and_scalars(libcp):
    libCP, ref = libcp
    w = synth.c2_denoise(12, np.float64, diag=True)
    a = np.sqrt(w["La_l2"])
    Cv, rX = libCP.CP_quadratic_l1(w["Y"], w["Eu"], w["Ev"], 0.4, a, 0.01, positivity=0, PFDR_rho=1.5, PFDR_difTol=1e-6, CP_itMax=4)
    E, V = w["E"], w["V"]
    r0 = ref.quadratic_d1_l1((w["Y"] * a), (a * a), w["Eu"], w["Ev"], np.full(E, 0.4), np.full(V, 0.01), 0, 0, CP_difTol=1e-3,
                             CP_itMax=4, rho=1.5, condMin=1e-3, difRcd=0.0, difTol=1e-6, itMax=10000)
    assert np.array_equal(Cv, r0["Cv"]) and rel(rX, r0["rX"]) < 1e-5


def test_identity_mode(libcp):
    libCP, ref = libcp
    w = synth.c2_denoise(10, np.float64, diag=False)
    Cv, rX = libCP.CP_quadratic_l1(w["Y"], w["Eu"], w["Ev"], w["La_d1"], np.array([1.0]), np.array([0.0]), CP_itMax=3, PFDR_difTol=1e-6)
    r0 = ref.quadratic_d1_l1(w["Y"], None, w["Eu"], w["Ev"], w["La_d1"], np.zeros(w["V"]), 0, 0, CP_difTol=1e-3, CP_itMax=3, rho=1.0,
                             condMin=1e-3, difRcd=0.0, difTol=1e-6, itMax=10000)
    assert np.array_equal(Cv, r0["Cv"]) and rel(rX, r0["rX"]) < 1e-5


def test_argument_errors(libcp):
    libCP, _ = libcp
    w = synth.c2_denoise(5, np.float64, diag=False)
    with pytest.raises(ValueError):
        libCP.CP_quadratic_l1(w["Y"].astype(np.int32), w["Eu"], w["Ev"], 0.1, 1.0, 0.0)
    with pytest.raises(ValueError):
        libCP.CP_quadratic_l1(w["Y"], w["Eu"], w["Ev"], 0.1, np.ones(7), 0.0)              # neither scalar, nor N, nor N x V
    with pytest.raises(ValueError):
        libCP.CP_quadratic_l1(w["Y"], w["Eu"], w["Ev"] + w["V"], 0.1, 1.0, 0.0)            # vertex number out of range
