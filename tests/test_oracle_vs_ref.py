"""Pins the CPU restatement (oracle/liboracle.so) against the unmodified reference compiled from
/root/reference (oracle/_ref/libref.so) on seeded inputs - the reference ships no golden vectors
(SURVEY section 4), so outputs of the reference itself are the pin.  Also checks the numpy model of the
GPU schedule (tests/sell_model.py) against the oracle.  CPU only."""
import numpy as np
import pytest

import problems

from cp_pfdr_graph_d1_b200 import synth
from cp_pfdr_graph_d1_b200.api import PFDR_graph_l22_d1_l1

import sell_model

TOL = {np.float32: 2e-4, np.float64: 1e-10}


def rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(float(np.linalg.norm(b)), 1e-300))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("variant", ["identity", "diag", "diag_l1_pos", "l1", "recond"])
def test_quadratic_n0(oracle, reference, dtype, variant):
    w = synth.c2_denoise(12, dtype, diag=(variant != "identity"))
    kw = dict(difTol=1e-5, itMax=200, want_obj=True, want_dif=True)
    if variant == "diag_l1_pos":
        kw.update(La_l1=np.full(w["V"], 0.05, dtype), positivity=1)
    if variant == "l1":
        kw.update(La_l1=np.full(w["V"], 0.05, dtype), positivity=0, want_obj=False)
    if variant == "recond":
        kw.update(difRcd=1e-3, difTol=1e-6, itMax=400)
    a = PFDR_graph_l22_d1_l1(w["Y"], w["La_l2"], w["Eu"], w["Ev"], w["La_d1"], backend=oracle, **kw)
    b = PFDR_graph_l22_d1_l1(w["Y"], w["La_l2"], w["Eu"], w["Ev"], w["La_d1"], backend=reference, **kw)
    assert abs(a.it - b.it) <= (0 if dtype is np.float64 else 2)
    assert rel(a.X, b.X) < TOL[dtype]
    n = min(a.it, b.it)
    if kw.get("want_obj", True):
        assert np.allclose(a.Obj[:n], b.Obj[:n], rtol=50 * TOL[dtype])
    assert np.allclose(a.Dif[:n], b.Dif[:n], rtol=1e-2 if dtype is np.float32 else 1e-6, atol=1e-12)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mode", ["direct", "ata"])
@pytest.mark.parametrize("family", ["l1", "bounds"])
def test_quadratic_dense(oracle, reference, dtype, mode, family):
    w = synth.dense_problem(7, 24, dtype)
    V = w["V"]
    A64 = w["A"].astype(np.float64)
    Lam = float(np.linalg.eigvalsh(A64.T @ A64).max())
    if mode == "direct":
        A, Y, N = w["A"], w["Y"], w["N"]
    else:
        A, Y, N = np.asfortranarray((A64.T @ A64).astype(dtype)), (A64.T @ w["Y"]).astype(dtype), -V
    out = []
    for be in (oracle, reference):
        X = np.zeros(V, dtype)
        kw = dict(Ltype=0, L=np.array([Lam], dtype), rho=1.5, condMin=1e-3, difRcd=0.0, difTol=1e-5, itMax=300,
                  N=N, want_obj=True, want_dif=True)
        if family == "l1":
            r = be.PFDR_graph_quadratic_d1_l1(X, Y, A, w["Eu"], w["Ev"], w["La_d1"], np.full(V, 0.01, dtype), 1, **kw)
        else:
            r = be.PFDR_graph_quadratic_d1_bounds(X, Y, A, w["Eu"], w["Ev"], w["La_d1"], 0.0, 1.5, **kw)
        out.append(r)
    a, b = out
    assert abs(a.it - b.it) <= (0 if dtype is np.float64 else 3)
    assert rel(a.X, b.X) < 10 * TOL[dtype]
    n = min(a.it, b.it)
    assert np.allclose(a.Obj[:n], b.Obj[:n], rtol=1e-3 if dtype is np.float32 else 1e-8, atol=1e-6)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("al", [0.0, 0.05, 1.0])
@pytest.mark.parametrize("variant", ["plain", "weighted", "recond", "labels"])
def test_simplex(oracle, reference, dtype, al, variant):
    w = synth.c4_simplex(8, 5, dtype)
    kw = dict(rho=1.5, condMin=1e-2, difRcd=0.0, difTol=1e-4, itMax=150, want_obj=True, want_dif=True)
    if variant == "weighted":
        if al == 0.0:
            pytest.skip("La_f is ignored by the linear loss")
        kw["La_f"] = np.random.default_rng(9).uniform(0.5, 2.0, w["V"]).astype(dtype)
    if variant == "recond":
        kw.update(difRcd=1e-2, difTol=1e-5, itMax=300)
    if variant == "labels":
        kw.update(difTol=1.0, itMax=60)
    out = []
    for be in (oracle, reference):
        P = w["Q"].copy()
        out.append(be.PFDR_graph_loss_d1_simplex(P, w["Q"], al, w["Eu"], w["Ev"], w["La_d1"], **kw))
    a, b = out
    assert abs(a.it - b.it) <= (0 if dtype is np.float64 else 3)
    assert rel(a.X, b.X) < 20 * TOL[dtype]
    n = min(a.it, b.it)
    assert np.allclose(a.Obj[:n], b.Obj[:n], rtol=1e-3 if dtype is np.float32 else 1e-8)
    assert np.allclose(a.X.sum(axis=1), 1, atol=1e-4 if dtype is np.float32 else 1e-10)


@pytest.mark.parametrize("V", [2, 5, 33, 100, 257])
@pytest.mark.parametrize("kind", ["diag", "dense", "ata"])
def test_quadratic_irregular_multigraphs(oracle, reference, V, kind):
    """pins the oracle on what cut-pursuit's reduced graphs contain (self-loops, duplicate edges, a hub) from two
    vertices up - the inputs of tests/test_gpu_quadratic.py::test_paths_agree_on_irregular_graphs"""
    fam, args, kw = problems.irregular_case(V, kind)
    a, b = (be.PFDR_graph_quadratic_d1_l1(*(x.copy() if isinstance(x, np.ndarray) else x for x in args), **kw)
            for be in (oracle, reference))
    assert np.isfinite(a.X).all() and a.it == b.it and rel(a.X, b.X) < 1e-9


@pytest.mark.parametrize("K", [1, 3, 16, 33, 50])
def test_simplex_label_counts(oracle, reference, K):
    """pins the oracle for the label counts the GPU parity cases use (tests/problems.py: K = 1 ... 127), among
    them more than 32 labels and counts that are not a multiple of 4"""
    w = synth.c4_simplex(5, K, np.float64)
    kw = dict(rho=1.5, condMin=1e-2, difRcd=1e-2, difTol=1e-5, itMax=120, want_obj=True, want_dif=True)
    a, b = (be.PFDR_graph_loss_d1_simplex(w["Q"].copy(), w["Q"], 0.05, w["Eu"], w["Ev"], w["La_d1"], **kw)
            for be in (oracle, reference))
    assert a.it == b.it and rel(a.X, b.X) < 1e-9
    assert np.allclose(a.Obj[:a.it], b.Obj[:a.it], rtol=1e-8)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_proj_simplex_metric(oracle, reference, dtype):
    rng = np.random.default_rng(11)
    X = rng.normal(0.3, 1.0, (257, 9)).astype(dtype)
    M = rng.uniform(0.2, 1.0, (257, 9)).astype(dtype)
    A = rng.uniform(0.5, 2.0, 3).astype(dtype)
    a = oracle.proj_simplex_metric(X.copy(), M, A)
    b = reference.proj_simplex_metric(X.copy(), M, A)
    assert np.array_equal(a, b)                  # same operation order: bit-identical
    tgt = np.concatenate([A, np.full(257 - 3, A[-1])])
    assert np.allclose(a.sum(axis=1), tgt, rtol=1e-4 if dtype is np.float32 else 1e-12)
    assert (a >= 0).all()
    # single shared metric / single target (nm = na = 1)
    a = oracle.proj_simplex_metric(X.copy(), M[:1])
    b = reference.proj_simplex_metric(X.copy(), M[:1])
    assert np.array_equal(a, b)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(20, 50), (50, 20), (0, 30)])
def test_operator_norm(oracle, reference, dtype, shape):
    rng = np.random.default_rng(12)
    M, N = shape
    if M == 0:
        B = rng.standard_normal((40, N))
        A = np.asfortranarray((B.T @ B).astype(dtype))
        lam = np.linalg.eigvalsh(A.astype(np.float64)).max()
    else:
        A = np.asfortranarray(rng.standard_normal((M, N)).astype(dtype))
        lam = np.linalg.norm(A.astype(np.float64), 2) ** 2
    a = oracle.operator_norm_matrix(A, M, N, 1e-4, 200, 8, seed=3)
    b = reference.operator_norm_matrix(A, M, N, 1e-4, 200, 8)
    # power method from random starts: both approach lambda_max from below
    assert a <= lam * (1 + 1e-4) and b <= lam * (1 + 1e-4)
    assert a > lam * 0.97 and b > lam * 0.97


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("variant", ["identity", "diag_l1_pos", "box"])
def test_gpu_schedule_model(oracle, dtype, variant):
    """the half-edge / recomputed-constants schedule of quad.cu reproduces the reference iteration"""
    w = synth.c2_denoise(10, dtype, diag=(variant != "identity"))
    V = w["V"]
    A = w["La_l2"]
    Y = w["Y"] if A is None else (A * w["Y"]).astype(dtype)
    La_l1 = np.full(V, 0.05, dtype) if variant == "diag_l1_pos" else None
    pos = int(variant == "diag_l1_pos")
    kw = dict(rho=1.5, condMin=1e-3, difRcd=0.0, difTol=1e-5, itMax=150)
    Ltype, L = (1, A) if A is not None else (0, None)
    if variant == "box":
        r = oracle.PFDR_graph_quadratic_d1_bounds(np.zeros(V, dtype), Y, A, w["Eu"], w["Ev"], w["La_d1"], 0.5, 2.5,
                                                  Ltype, L, N=0, **kw)
        box = (dtype(0.5), dtype(2.5))
    else:
        r = oracle.PFDR_graph_quadratic_d1_l1(np.zeros(V, dtype), Y, A, w["Eu"], w["Ev"], w["La_d1"], La_l1, pos,
                                              Ltype, L, N=0, **kw)
        box = None
    x, it = sell_model.quad_n0(V, np.zeros(V, dtype), Y, A, w["Eu"], w["Ev"], w["La_d1"], La_l1, pos, box, Ltype, L,
                               1.5, 1e-3, 1e-5, 150)
    assert abs(it - r.it) <= (0 if dtype is np.float64 else 2)
    assert rel(x, r.X) < (1e-4 if dtype is np.float32 else 1e-11)


# ---------------------------------------------------------------------------------------------------------------------
# the reduced-problem construction has no entry point in the reference (inline blocks of the cut-pursuit).
# Pin: the reference's own cut-pursuit, stopped after a few outer iterations, returns the components Cv and the reduced
# values rX = argmin of the last reduced problem; the restated blocks (oracle/cp_reduce_oracle_impl.h) must rebuild a reduced
# problem with the same minimiser from the same Cv.
# ---------------------------------------------------------------------------------------------------------------------
CP_CHAIN_KW = dict(CP_difTol=0.0, CP_itMax=6, rho=1.5, condMin=1e-3, difRcd=0.0, difTol=1e-10, itMax=200000)


def cp_chain_problem(mode):
    """seeded inputs of the three operator modes: (V, N, A, Y, Eu, Ev, La_d1, La_l1, positivity)"""
    dt = np.float64
    if mode == "diag":
        w = synth.c2_denoise(9, dt, diag=True)
        return w["V"], 0, w["La_l2"], (w["La_l2"] * w["Y"]).astype(dt), w["Eu"], w["Ev"], w["La_d1"], None, 0
    w = synth.eeg_like(2, 24, dt)
    V = w["V"]
    if mode == "direct":
        return V, w["N"], w["A"], w["Y"], w["Eu"], w["Ev"], w["La_d1"], w["La_l1"], 1
    # the l1 cut-pursuit of the reference cannot run with N < 0 (src/CP_PFDR_graph_quadratic_d1_l1.cpp:81 derives a thread
    # count from N and hands libgomp a negative one); its bounds twin (same blocks, src/CP_PFDR_graph_quadratic_d1_bounds.cpp:
    # 640-750) can: x >= 0 as the box [0, inf), no l1 term
    return V, -V, np.asfortranarray(w["A"].T @ w["A"]), w["A"].T @ w["Y"], w["Eu"], w["Ev"], w["La_d1"], None, 1


def cp_chain_reference(reference_cp, mode):
    """the reference's own cut-pursuit stopped after CP_itMax outer iterations -> dict(rV, Cv, rX) of the last reduced problem"""
    V, N, A, Y, Eu, Ev, La_d1, La_l1, pos = cp_chain_problem(mode)
    if mode == "ata":
        return reference_cp.quadratic_d1_bounds(Y, A, Eu, Ev, La_d1, 0.0, np.inf, N, **CP_CHAIN_KW)
    return reference_cp.quadratic_d1_l1(Y, A, Eu, Ev, La_d1, La_l1, pos, N, **CP_CHAIN_KW)


def cp_chain_oracle(mode, Cv, rV, preAt_wanted=None):
    """reduced problem rebuilt from the components by the restated blocks, solved by the oracle PFDR -> reduced values"""
    from oracle import oracle as O
    dt = np.float64
    V, N, A, Y, Eu, Ev, La_d1, La_l1, pos = cp_chain_problem(mode)
    assert 1 < rV < V
    red = O.CpReduceOracle()
    Vc, rVc = O.component_lists(Cv, rV)
    # the rule of :671 (PFDR_it = iterations of the previous reduced solve; PFDR_itMax on the first outer iteration)
    preAt = (N <= 0) or (rV < (2 * N * CP_CHAIN_KW["itMax"]) // (N + CP_CHAIN_KW["itMax"]))
    if preAt_wanted is not None:
        preAt = preAt_wanted                     # either form of the reduced problem has the same minimiser
    rA, rAA, rY = red.reduce(V, N, A, Y, Vc, rVc, preAt)
    rEu, rEv, rLa, rl1 = O.reduced_graph(Cv, rV, Eu, Ev, La_d1, La_l1, np.finfo(dt).eps)
    orc = O.oracle()
    X = np.zeros(rV, dt)
    common = dict(rho=1.5, condMin=1e-3, difRcd=0.0, difTol=1e-10, itMax=200000)
    if N == 0:
        out = orc.PFDR_graph_quadratic_d1_l1(X, rY, rAA, rEu, rEv, rLa, rl1, pos, 1, rAA, N=0, **common)
    elif preAt:
        S, d = red.equilibrate(N, rAA, True)
        c = np.linalg.eigvalsh(S).max()
        if mode == "ata":
            out = orc.PFDR_graph_quadratic_d1_bounds(X, rY, rAA, rEu, rEv, rLa, 0.0, np.inf, 1, (d * d * c).astype(dt), N=-rV, **common)
        else:
            out = orc.PFDR_graph_quadratic_d1_l1(X, rY, rAA, rEu, rEv, rLa, rl1, pos, 1, (d * d * c).astype(dt), N=-rV, **common)
    else:
        S, d = red.equilibrate(N, rA, False)
        c = np.linalg.norm(S, 2) ** 2
        out = orc.PFDR_graph_quadratic_d1_l1(X, Y, rA, rEu, rEv, rLa, rl1, pos, 1, (d * d * c).astype(dt), N=N, **common)
    return out.X


CP_CHAIN_CASES = [("direct", True), ("direct", False), ("ata", None), ("diag", None)]


@pytest.mark.parametrize("mode,preAt", CP_CHAIN_CASES)
def test_cp_reduce_chain(reference, mode, preAt):
    from oracle import oracle as O
    cp = O.cp_reference()
    if cp is None:
        pytest.skip("oracle/_ref/libref.so not built")
    r = cp_chain_reference(cp, mode)
    X = cp_chain_oracle(mode, r["Cv"], r["rV"], preAt)
    assert rel(X, r["rX"]) < 1e-6, (mode, rel(X, r["rX"]))


def test_cp_reduce_identities():
    """algebra the blocks must satisfy whatever the reference does: with S the V x rV indicator matrix of the components,
    rA = A S, rAA = S^t (A^t A) S, rY = S^t A^t Y; the gradient blocks agree with a dense evaluation; R = Y - A x"""
    from oracle import oracle as O
    red = O.CpReduceOracle()
    rng = np.random.default_rng(3)
    V, N, rV = 57, 13, 7
    Cv = rng.integers(0, rV, V).astype(np.int32)
    Cv[:rV] = np.arange(rV)
    Vc, rVc = O.component_lists(Cv, rV)
    S = np.zeros((V, rV)); S[np.arange(V), Cv] = 1
    A = np.asfortranarray(rng.normal(0, 1, (N, V))); Y = rng.normal(0, 1, N)
    rA, rAA, rY = red.reduce(V, N, A, Y, Vc, rVc, True)
    assert np.allclose(rA, A @ S, rtol=1e-12) and np.allclose(rAA, S.T @ A.T @ A @ S, rtol=1e-11) and np.allclose(rY, S.T @ A.T @ Y, rtol=1e-11)
    assert np.array_equal(rAA, rAA.T)
    M = np.asfortranarray(A.T @ A); AtY = A.T @ Y
    _, rAA2, rY2 = red.reduce(V, -V, M, AtY, Vc, rVc, True)
    assert np.allclose(rAA2, S.T @ M @ S, rtol=1e-11) and np.allclose(rY2, S.T @ AtY, rtol=1e-11) and np.array_equal(rAA2, rAA2.T)
    a = rng.uniform(0.5, 2, V)
    _, rAA3, rY3 = red.reduce(V, 0, a, AtY, Vc, rVc, True)
    assert np.allclose(rAA3, S.T @ a) and np.allclose(rY3, S.T @ AtY)
    _, rAA4, _ = red.reduce(V, 0, None, AtY, Vc, rVc, True)
    assert np.array_equal(rAA4, np.bincount(Cv, minlength=rV).astype(float))
    rX = rng.normal(0, 1, rV); rX[2] = 0.0
    x = rX[Cv]
    R = red.residual(rA, rX, Y)
    assert np.allclose(R, Y - A @ x, rtol=1e-11, atol=1e-12)
    assert np.allclose(red.gradient(V, N, A, R, rX, Cv, Vc, rVc), -A.T @ R, rtol=1e-11, atol=1e-12)
    assert np.allclose(red.gradient(V, -V, M, AtY, rX, Cv, Vc, rVc), M @ x - AtY, rtol=1e-10, atol=1e-11)
    assert np.allclose(red.gradient(V, 0, a, AtY, rX, Cv, Vc, rVc), a * x - AtY, rtol=1e-12)
    assert np.allclose(red.gradient(V, 0, None, AtY, rX, Cv, Vc, rVc), x - AtY, rtol=1e-12)


# the component numbering of the restated sweep against the reference's own (:570-600) - every cut-pursuit
# run returns Cv; cutting exactly the edges whose ends lie in different components of that Cv and sweeping again must give
# the same numbering back (components of the reference are connected through their non-cut edges by construction).
CP_NUMBERING_KW = dict(CP_difTol=1e-4, CP_itMax=4, rho=1.5, condMin=1e-3, difRcd=0.0, difTol=1e-6, itMax=2000)


def cp_numbering_graph(kind):
    dt = np.float64
    return synth.eeg_like(3, 30, dt) if kind == "eeg" else (synth.c2_denoise(12, dt, diag=True) if kind == "grid" else synth.c3_bounds_ata(400, 8, 128, dt))


def cp_numbering_reference(cp, kind):
    w, kw, dt = cp_numbering_graph(kind), CP_NUMBERING_KW, np.float64
    if kind == "eeg":
        return cp.quadratic_d1_l1(w["Y"], w["A"], w["Eu"], w["Ev"], w["La_d1"], w["La_l1"], 1, w["N"], **kw)
    if kind == "grid":
        return cp.quadratic_d1_l1((w["La_l2"] * w["Y"]).astype(dt), w["La_l2"], w["Eu"], w["Ev"], w["La_d1"], None, 0, 0, **kw)
    return cp.quadratic_d1_bounds(w["AtY"], w["AtA"], w["Eu"], w["Ev"], w["La_d1"], 0.0, 1.0, -w["V"], **kw)


def check_cp_numbering(kind, Cv, rV):
    from oracle import oracle as O
    w = cp_numbering_graph(kind)
    assert rV > 1
    active = (Cv[w["Eu"]] != Cv[w["Ev"]]).astype(np.uint8)
    oCv, orV, Vc, rVc = O.CpGraphOracle().components(w["V"], w["Eu"], w["Ev"], active)
    assert orV == rV and np.array_equal(oCv, Cv)
    assert np.array_equal(np.sort(Vc), np.arange(w["V"])) and np.array_equal(np.diff(rVc), np.bincount(Cv, minlength=rV))
    assert all(np.all(Cv[Vc[rVc[c]:rVc[c + 1]]] == c) for c in range(rV))


@pytest.mark.parametrize("kind", ["eeg", "grid", "knn"])
def test_cp_components_numbering(reference, kind):
    from oracle import oracle as O
    cp = O.cp_reference()
    if cp is None:
        pytest.skip("oracle/_ref/libref.so not built")
    r = cp_numbering_reference(cp, kind)
    check_cp_numbering(kind, r["Cv"], r["rV"])


def test_cp_reduced_edges_and_merge_rules():
    """hand-checked instance of :603-661 and :863-886: path 0-1-2-3-4 plus the chord 0-2 and an isolated vertex 5"""
    from oracle import oracle as O
    g = O.CpGraphOracle()
    Eu = np.array([0, 1, 2, 3, 0], np.int32); Ev = np.array([1, 2, 3, 4, 2], np.int32)
    La = np.array([1.0, 2.0, 4.0, 0.0, 8.0])
    active = np.array([0, 1, 1, 1, 1], np.uint8)             # 0-1 joined; 1-2, 2-3 cut; 3-4 cut but weightless; 0-2 cut
    Cv, rV, _, _ = g.components(6, Eu, Ev, active)
    assert rV == 5 and np.array_equal(Cv, [0, 0, 1, 2, 3, 4])
    rEu, rEv, rLa = g.reduced_edges(6, Eu, Ev, active, La, Cv, rV, 1e-16)
    # {0,1}-{2}: 2 + 8; {2}-{3}: 4; {4} and {5} untouched by a nonzero cut edge -> loops of weight eps
    assert list(zip(rEu, rEv)) == [(0, 1), (1, 2), (3, 3), (4, 4)] and np.allclose(rLa, [10.0, 4.0, 1e-16, 1e-16])
    active2 = np.array([1, 0, 0, 0, 0], np.uint8)            # cut edge 0-1 inside a component (0 and 1 stay joined through 2): a loop, met twice
    Cv2, rV2, _, _ = g.components(6, Eu, Ev, active2)
    assert rV2 == 2 and np.array_equal(Cv2, [0, 0, 0, 0, 0, 1])
    rEu, rEv, rLa = g.reduced_edges(6, Eu, Ev, active2, La, Cv2, rV2, 1e-16)
    assert list(zip(rEu, rEv)) == [(0, 0), (1, 1)] and np.allclose(rLa, [2 * 1.0, 1e-16])
    rX = np.array([1.0, 1.0005, 2.0, 0.0, 1e-20])
    act = active.copy()
    n = g.merge(Eu, Ev, Cv, rX, 1e-3, 1e-16, act)            # 1-2 and 0-2 join components 0 and 1 (values 1 / 1.0005: within 1e-3 relative)
    assert n == 2 and np.array_equal(act, [0, 0, 1, 1, 0])
