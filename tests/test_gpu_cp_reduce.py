"""SURVEY 8f N1 - the dense passes of the cut-pursuit's outer iteration on the device (csrc/cpreduce.cu) against the oracle
restatement of the reference's blocks (oracle/cp_reduce_oracle_impl.h; pinned to the reference in
tests/test_oracle_vs_ref.py::test_cp_reduce_chain).  Tolerances: 1e-10 relative in float64, 1e-3 in float32 (the reference
accumulates running sums in `real`, the device in double)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = {np.float32: 1e-3, np.float64: 1e-10}


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / max(float(np.linalg.norm(np.asarray(b, np.float64))), 1e-300))


def components(V, rV, kind, rng):
    if kind == "random":
        Cv = rng.integers(0, rV, V)
        Cv[rng.permutation(V)[:rV]] = np.arange(rV)                  # no empty component
    elif kind == "blocks":                                           # contiguous ranges of very different sizes
        cuts = np.sort(rng.choice(np.arange(1, V), rV - 1, replace=False))
        Cv = np.searchsorted(cuts, np.arange(V), side="right")
    elif kind == "one":
        Cv = np.zeros(V, int)
    elif kind == "all":
        Cv = rng.permutation(V)
    else:
        raise KeyError(kind)
    return Cv.astype(np.int32), int(Cv.max()) + 1


CASES = [(57, 13, 7, "random"), (300, 40, 23, "blocks"), (1000, 91, 1, "one"), (200, 16, 200, "all"), (4096, 64, 300, "random"),
         (2500, 33, 40, "blocks"), (777, 128, 5, "blocks")]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("V,N,rV,kind", CASES)
def test_direct_operator(gpu_api, V, N, rV, kind, dtype):
    from oracle import oracle as O
    rng = np.random.default_rng(V + N)
    red = O.CpReduceOracle()
    Cv, rV = components(V, rV, kind, rng)
    Vc, rVc = O.component_lists(Cv, rV)
    A = np.asfortranarray(rng.normal(0, 1, (N, V)).astype(dtype))
    Y = rng.normal(0, 1, N).astype(dtype)
    op = gpu_api.CpOperator(A, V, N)
    op.set_components(Cv, rV)
    for preAt in (True, False):
        rA, rAA, rY = op.reduce(Y, preAt)
        oA, oAA, oY = red.reduce(V, N, A, Y, Vc, rVc, preAt)
        assert rel(rA, oA) < TOL[dtype]
        if preAt:
            assert rel(rAA, oAA) < TOL[dtype] and rel(rY, oY) < TOL[dtype]
            assert np.array_equal(rAA, rAA.T)
        else:
            assert rAA is None and rY is None
    rX = rng.normal(0, 1, rV).astype(dtype)
    R = op.residual(rX, Y)
    assert rel(R, red.residual(oA, rX, Y)) < 10 * TOL[dtype]
    g_kept = op.gradient(rX)                                         # residual kept on the device by residual()
    g_given = op.gradient(rX, R)
    assert np.array_equal(g_kept, g_given)
    assert rel(g_given, red.gradient(V, N, A, R, rX, Cv, Vc, rVc)) < 10 * TOL[dtype]
    op.close()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("V,N,rV,kind", [c for c in CASES if c[0] <= 2500])
def test_premultiplied_operator(gpu_api, V, N, rV, kind, dtype):
    from oracle import oracle as O
    rng = np.random.default_rng(V + 3 * N)
    red = O.CpReduceOracle()
    Cv, rV = components(V, rV, kind, rng)
    Vc, rVc = O.component_lists(Cv, rV)
    B = rng.normal(0, 1, (N, V))
    M = np.asfortranarray((B.T @ B / N + 0.1 * np.eye(V)).astype(dtype))
    AtY = (B.T @ rng.normal(0, 1, N)).astype(dtype)
    op = gpu_api.CpOperator(M, V, -V)
    op.set_components(Cv, rV)
    _, rAA, rY = op.reduce(AtY)
    _, oAA, oY = red.reduce(V, -V, M, AtY, Vc, rVc, True)
    assert rel(rAA, oAA) < TOL[dtype] and rel(rY, oY) < TOL[dtype]
    assert np.array_equal(rAA, rAA.T)
    rX = rng.normal(0, 1, rV).astype(dtype)
    rX[rV // 2] = 0
    assert rel(op.gradient(rX, AtY), red.gradient(V, -V, M, AtY, rX, Cv, Vc, rVc)) < 10 * TOL[dtype]
    op.close()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("identity", [False, True])
@pytest.mark.parametrize("V,rV,kind", [(1000, 1, "one"), (5000, 77, "random"), (100000, 900, "blocks"), (64, 64, "all")])
def test_diagonal_operator(gpu_api, V, rV, kind, identity, dtype):
    from oracle import oracle as O
    rng = np.random.default_rng(V)
    red = O.CpReduceOracle()
    Cv, rV = components(V, rV, kind, rng)
    Vc, rVc = O.component_lists(Cv, rV)
    a = None if identity else rng.uniform(0.5, 1.5, V).astype(dtype)
    Y = rng.normal(0, 1, V).astype(dtype)
    op = gpu_api.CpOperator(a, V, 0, dtype=dtype)
    op.set_components(Cv, rV)
    _, rAA, rY = op.reduce(Y)
    _, oAA, oY = red.reduce(V, 0, a, Y, Vc, rVc, True)
    assert rel(rAA, oAA) < TOL[dtype] and rel(rY, oY) < 30 * TOL[dtype]
    if identity:
        assert np.array_equal(rAA, np.bincount(Cv, minlength=rV).astype(dtype))
    rX = rng.normal(0, 1, rV).astype(dtype)
    assert np.array_equal(op.gradient(rX, Y), red.gradient(V, 0, a, Y, rX, Cv, Vc, rVc))
    op.close()


def test_argument_errors(gpu_api):
    from cp_pfdr_graph_d1_b200.api import PfdrError
    A = np.asfortranarray(np.ones((4, 10)))
    op = gpu_api.CpOperator(A, 10, 4)
    with pytest.raises(PfdrError):
        op.reduce(np.ones(4))                       # no components yet
    with pytest.raises(PfdrError):
        op.set_components(np.full(10, 3, np.int32), 2)   # label outside [0, rV)
    op.set_components(np.zeros(10, np.int32), 1)
    with pytest.raises(PfdrError):
        op.residual(np.ones(1), np.ones(4))         # no reduce for these components yet
    with pytest.raises(PfdrError):
        gpu_api.CpOperator(np.asfortranarray(np.ones((5, 5))), 5, -4)


def test_device_resident_operator_and_chain(gpu_api):
    """one outer iteration the way a patched cut-pursuit would run it: operator uploaded once (a torch tensor), gradient ->
    (cut on the host) -> components -> reduce -> PFDR on the reduced problem (the library's own) -> residual -> next gradient;
    compared with the same chain on the oracle"""
    import torch
    from oracle import oracle as O
    from cp_pfdr_graph_d1_b200 import synth
    w = synth.eeg_like(3, 40, np.float64)
    V, N, A, Y = w["V"], w["N"], w["A"], w["Y"]
    red, orc = O.CpReduceOracle(), O.oracle()
    At = torch.from_numpy(np.ascontiguousarray(A.T)).cuda().t()     # (N, V) view, column major in memory
    op = gpu_api.CpOperator(At, V, N)
    Cv = (w["points"][:, 2] > 0).astype(np.int32) + 2 * (w["points"][:, 0] > 0.3).astype(np.int32)
    rV = int(Cv.max()) + 1
    Vc, rVc = O.component_lists(Cv, rV)
    op.set_components(Cv, rV)
    rA, rAA, rY = op.reduce(Y, True)
    oA, oAA, oY = red.reduce(V, N, A, Y, Vc, rVc, True)
    assert rel(rAA, oAA) < 1e-10
    rEu, rEv, rLa, rl1 = O.reduced_graph(Cv, rV, w["Eu"], w["Ev"], w["La_d1"], w["La_l1"], np.finfo(np.float64).eps)
    L = np.array([np.linalg.eigvalsh(oAA).max()])
    kw = dict(rho=1.5, condMin=1e-3, difRcd=0.0, difTol=1e-9, itMax=20000, N=-rV)
    x_gpu = gpu_api.PFDR_graph_quadratic_d1_l1(np.zeros(rV), rY, rAA, rEu, rEv, rLa, rl1, 1, 0, L, **kw).X
    x_or = orc.PFDR_graph_quadratic_d1_l1(np.zeros(rV), oY, oAA, rEu, rEv, rLa, rl1, 1, 0, L, **kw).X
    assert rel(x_gpu, x_or) < 1e-8
    R = op.residual(x_gpu, Y)
    assert rel(R, red.residual(oA, x_or, Y)) < 1e-8
    assert rel(op.gradient(x_gpu), red.gradient(V, N, A, red.residual(oA, x_or, Y), x_or, Cv, Vc, rVc)) < 1e-7
    op.close()
