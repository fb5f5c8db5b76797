"""Golden vectors: outputs of the reference itself (tests/golden/make_golden.py) for seeded inputs.
CPU: the oracle restatement reproduces them.  GPU (-m gpu): the CUDA path reproduces them through the C ABI."""
import os
import sys

import numpy as np
import pytest

import problems

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden as G  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_outputs.npz"))
TOL_ORACLE = {np.float32: 2e-4, np.float64: 1e-9}
TOL_GPU = {np.float32: 1e-3, np.float64: 1e-5}


def rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(float(np.linalg.norm(b)), 1e-300))


def check(kind, name, dtype, binder, tol):
    tag = np.dtype(dtype).name
    if kind == "quad":
        fam, args, kw = problems.quad_case(name, dtype)
        r = problems.run_quad(binder, name, dtype)
    else:
        args, kw = problems.simplex_case(name, dtype)
        r = problems.run_simplex(binder, name, dtype)
    assert G.digest(args) == str(GOLD[f"{kind}/{name}/{tag}/sha"]), "seeded inputs drifted: regenerate the fixtures"
    git = int(GOLD[f"{kind}/{name}/{tag}/it"])
    assert abs(r.it - git) <= (0 if dtype is np.float64 else max(2, git // 10))
    assert rel(r.X, GOLD[f"{kind}/{name}/{tag}/X"]) < tol
    key = f"{kind}/{name}/{tag}/Obj"
    if key in GOLD.files and r.Obj is not None:
        n = min(r.it, git) + 1
        g = GOLD[key]
        assert np.allclose(r.Obj[:n], g[:n], rtol=10 * tol, atol=tol * max(float(np.abs(g).max()), 1.0) * 0.1)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name", G.QUAD)
def test_oracle_quadratic_golden(oracle, name, dtype):
    check("quad", name, dtype, oracle, TOL_ORACLE[dtype])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name", G.SIMPLEX)
def test_oracle_simplex_golden(oracle, name, dtype):
    check("simplex", name, dtype, oracle, TOL_ORACLE[dtype])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_oracle_proj_golden(oracle, dtype):
    rng = np.random.default_rng(11)
    X = rng.normal(0.3, 1.0, (300, 9)).astype(dtype)
    M = rng.uniform(0.2, 1.0, (300, 9)).astype(dtype)
    A = rng.uniform(0.5, 2.0, 3).astype(dtype)
    assert np.array_equal(oracle.proj_simplex_metric(X.copy(), M, A), GOLD[f"proj/{np.dtype(dtype).name}/X"])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name", G.QUAD)
def test_gpu_quadratic_golden(gpu_api, name, dtype):
    check("quad", name, dtype, gpu_api._b(), TOL_GPU[dtype])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name", G.SIMPLEX)
def test_gpu_simplex_golden(gpu_api, name, dtype):
    check("simplex", name, dtype, gpu_api._b(), TOL_GPU[dtype])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_gpu_proj_golden(gpu_api, dtype):
    rng = np.random.default_rng(11)
    X = rng.normal(0.3, 1.0, (300, 9)).astype(dtype)
    M = rng.uniform(0.2, 1.0, (300, 9)).astype(dtype)
    A = rng.uniform(0.5, 2.0, 3).astype(dtype)
    got = gpu_api.proj_simplex_metric(X.copy(), M, A)
    assert np.allclose(got, GOLD[f"proj/{np.dtype(dtype).name}/X"], rtol=1e-4 if dtype is np.float32 else 1e-11,
                       atol=1e-6 if dtype is np.float32 else 1e-14)


# ---- cut-pursuit blocks: the pins of tests/test_oracle_vs_ref.py on stored outputs of the reference's
# cut-pursuit, so that they also hold where oracle/_ref cannot be built ------------------------------------------------
GOLD_CP = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_cp_outputs.npz"))


@pytest.mark.parametrize("mode,preAt", [("direct", True), ("direct", False), ("ata", None), ("diag", None)])
def test_oracle_cp_reduce_chain_golden(oracle, mode, preAt):
    import test_oracle_vs_ref as T
    assert G.digest([a for a in T.cp_chain_problem(mode) if isinstance(a, np.ndarray)]) == str(GOLD_CP[f"chain/{mode}/sha"]), \
        "seeded inputs drifted: regenerate the fixtures"
    Cv, rX = GOLD_CP[f"chain/{mode}/Cv"], GOLD_CP[f"chain/{mode}/rX"]
    X = T.cp_chain_oracle(mode, Cv, rX.shape[0], preAt)
    assert rel(X, rX) < 1e-6


@pytest.mark.parametrize("kind", ["eeg", "grid", "knn"])
def test_oracle_cp_components_numbering_golden(kind):
    import test_oracle_vs_ref as T
    Cv = GOLD_CP[f"numbering/{kind}/Cv"]
    T.check_cp_numbering(kind, Cv, int(Cv.max()) + 1)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["eeg", "grid", "knn"])
def test_gpu_cp_components_numbering_golden(gpu_api, kind):
    """the device sweep reproduces the reference's own component numbering from its cut"""
    import test_oracle_vs_ref as T
    w = T.cp_numbering_graph(kind)
    Cv = GOLD_CP[f"numbering/{kind}/Cv"]
    active = (Cv[w["Eu"]] != Cv[w["Ev"]]).astype(np.uint8)
    got, rV = gpu_api.cp_components(w["V"], w["Eu"], w["Ev"], active)
    assert rV == int(Cv.max()) + 1 and np.array_equal(got, Cv)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["direct", "ata", "diag"])
def test_gpu_cp_reduce_on_the_reference_components_golden(gpu_api, mode):
    """the device passes on the components the reference's cut-pursuit actually produced (stored fixtures): reduced operator
    against the restated blocks, and - closing the chain on the device - the library's own PFDR on the device-built reduced
    problem reproduces the reference's reduced values"""
    import test_oracle_vs_ref as T
    from oracle import oracle as O
    V, N, A, Y, Eu, Ev, La_d1, La_l1, pos = T.cp_chain_problem(mode)
    Cv, rX = GOLD_CP[f"chain/{mode}/Cv"], GOLD_CP[f"chain/{mode}/rX"]
    rV = rX.shape[0]
    Vc, rVc = O.component_lists(Cv, rV)
    op = gpu_api.CpOperator(A, V, N)
    op.set_components(Cv, rV)
    rA, rAA, rY = op.reduce(Y, True)
    oA, oAA, oY = O.CpReduceOracle().reduce(V, N, A, Y, Vc, rVc, True)
    assert rel(rAA, oAA) < 1e-10 and rel(rY, oY) < 1e-10 and (N <= 0 or rel(rA, oA) < 1e-10)
    active = (Cv[Eu] != Cv[Ev]).astype(np.uint8)
    rEu, rEv, rLa = gpu_api.cp_reduced_edges(V, Eu, Ev, active, La_d1, Cv, rV)
    rl1 = None if La_l1 is None else np.bincount(Cv, weights=La_l1, minlength=rV)
    kw = dict(rho=1.5, condMin=1e-3, difRcd=0.0, difTol=1e-10, itMax=200000)
    if N == 0:
        x = gpu_api.PFDR_graph_quadratic_d1_l1(np.zeros(rV), rY, rAA, rEu, rEv, rLa, rl1, pos, 1, rAA, N=0, **kw).X
    else:
        d = np.sqrt(np.diag(rAA))
        L = d * d * np.linalg.eigvalsh(rAA / np.outer(d, d)).max()
        if mode == "ata":
            x = gpu_api.PFDR_graph_quadratic_d1_bounds(np.zeros(rV), rY, rAA, rEu, rEv, rLa, 0.0, np.inf, 1, L, N=-rV, **kw).X
        else:
            x = gpu_api.PFDR_graph_quadratic_d1_l1(np.zeros(rV), rY, rAA, rEu, rEv, rLa, rl1, pos, 1, L, N=-rV, **kw).X
    assert rel(x, rX) < 1e-6
    op.close()
