"""GPU parity of the simplex family, proj_simplex_metric and operator_norm_matrix (through the C ABI)
against the CPU oracle.  Tolerances: 1e-5 relative (float64), 1e-3 (float32) - BASELINE.json."""
import numpy as np
import pytest

import problems

pytestmark = pytest.mark.gpu
TOL = {np.float32: 1e-3, np.float64: 1e-5}


def rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(float(np.linalg.norm(b)), 1e-300))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name", problems.SIMPLEX_CASES)
def test_simplex_parity(gpu_api, oracle, name, dtype):
    from cp_pfdr_graph_d1_b200.api import _b
    got = problems.run_simplex(_b(), name, dtype)
    ref = problems.run_simplex(oracle, name, dtype)
    tol = TOL[dtype]
    slack = 0 if dtype is np.float64 else max(2, ref.it // 20)
    assert abs(got.it - ref.it) <= slack, (got.it, ref.it)
    assert rel(got.X, ref.X) < tol, rel(got.X, ref.X)
    assert np.allclose(got.X.sum(axis=1), 1.0, atol=1e-5 if dtype is np.float32 else 1e-12)
    assert (got.X >= 0).all()
    n = min(got.it, ref.it)
    assert got.Obj.shape[0] == got.it + 1
    assert np.allclose(got.Obj[:n + 1], ref.Obj[:n + 1], rtol=tol, atol=tol * max(float(np.abs(ref.Obj).max()), 1.0) * 1e-1)
    if ref.Dif is not None and n and "labels" not in name:
        m = max(1, n // 2) if dtype is np.float32 else n
        assert np.allclose(got.Dif[:m], ref.Dif[:m], rtol=5e-2 if dtype is np.float32 else 1e-4, atol=1e-6 if dtype is np.float32 else 1e-12)   # f32: rounding floor of a mean absolute change
    if "labels" in name and dtype is np.float64:
        assert np.array_equal(got.Dif[:n], ref.Dif[:n])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name", ["kla", "kla_k3", "kla_k1", "kla_recond", "kla_labels", "lin", "quad_weighted", "kla_rand"])
def test_simplex_parity_scalar_lanes(gpu_api, oracle, monkeypatch, name, dtype):
    """K not a multiple of the labels per lane runs on padded rows by default (test_simplex_parity: K = 5, 3, 33, 50,
    127 ...); this forces the one-label-per-lane kernels (the default only for K <= 2) on the same cases"""
    from cp_pfdr_graph_d1_b200.api import _b
    monkeypatch.setenv("PFDR_B200_LPT", "1")
    got = problems.run_simplex(_b(), name, dtype)
    ref = problems.run_simplex(oracle, name, dtype)
    tol = TOL[dtype]
    assert abs(got.it - ref.it) <= (0 if dtype is np.float64 else max(2, ref.it // 20)), (got.it, ref.it)
    assert rel(got.X, ref.X) < tol, rel(got.X, ref.X)
    assert got.X.shape == ref.X.shape and np.isfinite(got.X).all()
    assert np.allclose(got.X.sum(axis=1), 1.0, atol=1e-5 if dtype is np.float32 else 1e-12) and (got.X >= 0).all()
    n = min(got.it, ref.it)
    assert np.allclose(got.Obj[:n + 1], ref.Obj[:n + 1], rtol=tol, atol=tol * max(float(np.abs(ref.Obj).max()), 1.0) * 1e-1)


def chain_or_image_case(kind, K, dtype):
    """1-D chain (one offset) / 2-D image (two offsets) label problems: the other shapes of the stencil layout"""
    rng = np.random.default_rng(K + len(kind))
    if kind == "chain":
        V = 501
        Eu = np.arange(V - 1, dtype=np.int32); Ev = Eu + 1
    else:
        nx, ny = 23, 17
        V = nx * ny
        v = np.arange(V).reshape(ny, nx)
        eu = np.stack([v, v], -1); ev = np.stack([v + 1, v + nx], -1)
        ok = np.zeros(eu.shape, bool); ok[:, :-1, 0] = True; ok[:-1, :, 1] = True
        Eu, Ev = eu[ok].astype(np.int32), ev[ok].astype(np.int32)
    Q = rng.uniform(0, 1, (V, K))
    Q[np.arange(V), (np.arange(V) * 3 // V) % K] += 1.5
    Q = (Q / Q.sum(axis=1, keepdims=True)).astype(dtype)
    La_d1 = rng.uniform(0.05, 0.3, Eu.shape[0]).astype(dtype)
    return Q, Eu, Ev, La_d1


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name", ["kla", "kla_weighted", "kla_recond", "kla_labels", "kla_k16", "kla_k3", "kla_k12", "kla_k64",
                                  "kla_recond_k16", "lin_k16", "lin_recond_k16", "quad_k16", "quad_recond_k16",
                                  "quad_weighted_k16", "kla_k33", "kla_recond_k33", "quad_weighted_k50", "kla_notrack"])
def test_stencil_layout_matches_table(gpu_api, monkeypatch, name, dtype):
    """grid problems run on the constant-offset layout (no index stream); the half-edge table (PFDR_B200_DIA=0) must
    give the same bits: same arithmetic, same summation order"""
    from cp_pfdr_graph_d1_b200.api import _b
    got = problems.run_simplex(_b(), name, dtype)
    assert gpu_api.last_stats()["layout"] == 1
    monkeypatch.setenv("PFDR_B200_DIA", "0")
    if dtype is np.float64:
        monkeypatch.setenv("PFDR_B200_LPT", "4")           # same lane layout as the stencil kernels -> same reduction trees
    tab = problems.run_simplex(_b(), name, dtype)
    assert gpu_api.last_stats()["layout"] == 0
    assert got.it == tab.it
    assert np.array_equal(got.X, tab.X)
    assert np.array_equal(got.Obj, tab.Obj)
    if got.Dif is not None and "labels" in name:
        assert np.array_equal(got.Dif, tab.Dif)
    elif got.Dif is not None:        # the evolution sums are associated differently (per-lane double accumulators)
        assert np.allclose(got.Dif, tab.Dif, rtol=1e-5 if dtype is np.float32 else 1e-12, atol=0)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("K", [4, 7, 16])
@pytest.mark.parametrize("kind", ["chain", "image"])
@pytest.mark.parametrize("recond", [0.0, 1e-2])
def test_stencil_layout_chain_and_image(gpu_api, oracle, kind, K, dtype, recond):
    from cp_pfdr_graph_d1_b200.api import _b
    Q, Eu, Ev, La_d1 = chain_or_image_case(kind, K, dtype)
    kw = dict(rho=1.5, condMin=1e-2, difRcd=recond, difTol=1e-5, itMax=200, want_obj=True, want_dif=True)
    got = _b().PFDR_graph_loss_d1_simplex(Q.copy(), Q, 0.05, Eu, Ev, La_d1, **kw)
    assert gpu_api.last_stats()["layout"] == 1
    ref = oracle.PFDR_graph_loss_d1_simplex(Q.copy(), Q, 0.05, Eu, Ev, La_d1, **kw)
    tol = TOL[dtype]
    assert abs(got.it - ref.it) <= (0 if dtype is np.float64 else max(2, ref.it // 20)), (got.it, ref.it)
    assert rel(got.X, ref.X) < tol, rel(got.X, ref.X)
    n = min(got.it, ref.it)
    assert np.allclose(got.Obj[:n + 1], ref.Obj[:n + 1], rtol=tol, atol=tol * max(float(np.abs(ref.Obj).max()), 1.0) * 1e-1)


def test_stencil_layout_not_used_for_other_numberings(gpu_api):
    """a grid whose edge list is shuffled (or holds u > v) keeps the half-edge table: the per-vertex summation order
    of the reference depends on the edge numbering"""
    from cp_pfdr_graph_d1_b200.api import _b
    args, kw = problems.simplex_case("kla_k16", np.float32)
    P0, Q, al, Eu, Ev, La_d1 = args
    ref = _b().PFDR_graph_loss_d1_simplex(P0.copy(), Q, al, Eu, Ev, La_d1, **kw)
    assert gpu_api.last_stats()["layout"] == 1
    perm = np.random.default_rng(0).permutation(Eu.shape[0])
    got = _b().PFDR_graph_loss_d1_simplex(P0.copy(), Q, al, Eu[perm], Ev[perm], La_d1[perm], **kw)
    assert gpu_api.last_stats()["layout"] == 0
    assert abs(got.it - ref.it) <= 2 and rel(got.X, ref.X) < 1e-4
    got = _b().PFDR_graph_loss_d1_simplex(P0.copy(), Q, al, Ev, Eu, La_d1, **kw)
    assert gpu_api.last_stats()["layout"] == 0
    assert abs(got.it - ref.it) <= 2 and rel(got.X, ref.X) < 1e-4


@pytest.mark.parametrize("knobs", [{"PFDR_B200_TILE": "32"}, {"PFDR_B200_TILE": "32", "PFDR_B200_DIA": "0"},
                                   {"PFDR_B200_TILE": "8", "PFDR_B200_IDX32": "0", "PFDR_B200_DIA": "0"}, {"PFDR_B200_TILE": "8"},
                                   {"PFDR_B200_LPT": "4"}, {"PFDR_B200_LPT": "4", "PFDR_B200_TILE": "16"},
                                   {"PFDR_B200_IDX32": "0", "PFDR_B200_DIA": "0"}, {"PFDR_B200_LPT": "2", "PFDR_B200_TILE": "24"}],
                         ids=lambda k: "+".join(f"{a[10:]}={b}" for a, b in k.items()))
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name", ["kla_k16", "kla_recond_k16", "kla_labels_k16", "quad_weighted_k16", "lin_k16", "kla_k33", "kla"])
def test_simplex_parity_size_gated_paths(gpu_api, oracle, monkeypatch, name, dtype, knobs):
    """the code paths that full-size problems switch on by size - the tiled slice order (252^3), four labels per lane
    in float64 (V.K > 2^23), 64-bit element offsets (> 2^32 elements) - forced on the small cases the oracle can run"""
    from cp_pfdr_graph_d1_b200.api import _b
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    got = problems.run_simplex(_b(), name, dtype)
    ref = problems.run_simplex(oracle, name, dtype)
    tol = TOL[dtype]
    assert abs(got.it - ref.it) <= (0 if dtype is np.float64 else max(2, ref.it // 20)), (got.it, ref.it)
    assert rel(got.X, ref.X) < tol, rel(got.X, ref.X)
    assert np.allclose(got.X.sum(axis=1), 1.0, atol=1e-5 if dtype is np.float32 else 1e-12) and (got.X >= 0).all()
    n = min(got.it, ref.it)
    assert np.allclose(got.Obj[:n + 1], ref.Obj[:n + 1], rtol=tol, atol=tol * max(float(np.abs(ref.Obj).max()), 1.0) * 1e-1)
    if "labels" in name and dtype is np.float64:
        assert np.array_equal(got.Dif[:n], ref.Dif[:n])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("D", [1, 2, 5, 16, 32, 33, 100])
def test_proj_simplex_metric(gpu_api, oracle, dtype, D):
    rng = np.random.default_rng(D)
    N = 1000
    X = rng.normal(0.3, 1.0, (N, D)).astype(dtype)
    M = rng.uniform(0.2, 1.0, (N, D)).astype(dtype)
    A = rng.uniform(0.5, 2.0, 3).astype(dtype)
    for Mi, Ai in ((M, A), (M[:1], None), (M[:7], A[:1])):
        got = gpu_api.proj_simplex_metric(X.copy(), Mi, Ai)
        ref = oracle.proj_simplex_metric(X.copy(), Mi, Ai)
        assert np.allclose(got, ref, rtol=1e-4 if dtype is np.float32 else 1e-11, atol=1e-5 if dtype is np.float32 else 1e-13)
        assert (got >= 0).all()
        assert np.array_equal(got == 0, ref == 0) or dtype is np.float32


@pytest.mark.parametrize("D", [4096, 5000])
def test_proj_simplex_metric_many_coordinates(gpu_api, oracle, D):
    """no limit on D (the reference has none, src/proj_simplex_metric.cpp:18-83): flags in local memory up to 4096
    coordinates, in a scratch array beyond"""
    rng = np.random.default_rng(D)
    X = rng.normal(0.0, 1.0, (40, D))
    M = rng.uniform(0.2, 1.0, (40, D))
    got = gpu_api.proj_simplex_metric(X.copy(), M, None)
    ref = oracle.proj_simplex_metric(X.copy(), M, None)
    assert np.allclose(got, ref, rtol=1e-10, atol=1e-13) and np.array_equal(got == 0, ref == 0)


def test_proj_simplex_empty(gpu_api):
    X = np.zeros((0, 4), np.float32)
    gpu_api.proj_simplex_metric(X, np.ones((1, 4), np.float32))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(20, 50), (50, 20), (0, 30), (37, 0), (91, 700), (300, 300), (0, 1500), (1200, 1300), (64, 5000)])
def test_operator_norm(gpu_api, monkeypatch, dtype, shape):
    """same estimator as the reference, start for start (glibc rand_r streams seeded time + thread, nbInit rounded up to
    the core count, the reference's stopping rule): with the clock pinned like oracle/_ref's, the two values agree to
    rounding; and with a tight tolerance the value is lambda_max.  n > 1024 takes the batched all-SM path."""
    from oracle import oracle as O
    monkeypatch.setenv("PFDR_B200_OPNORM_TIME", "42")
    rng = np.random.default_rng(12)
    M, N = shape
    if M == 0 or N == 0:
        n = max(M, N)
        B = rng.standard_normal((n + 10, n))
        A = np.asfortranarray((B.T @ B).astype(dtype))
        lam = np.linalg.eigvalsh(A.astype(np.float64)).max()
    else:
        A = np.asfortranarray(rng.standard_normal((M, N)).astype(dtype))
        lam = np.linalg.norm(A.astype(np.float64), 2) ** 2
    got = gpu_api.operator_norm_matrix(A, M, N, 1e-3, 100, 10)
    again = gpu_api.operator_norm_matrix(A, M, N, 1e-3, 100, 10)
    assert got == again
    ref = O.reference()
    if ref is not None:                        # the reference's own value at its cut-pursuit settings (nTol 1e-3, 100, 10)
        want = ref.operator_norm_matrix(A, M, N, 1e-3, 100, 10)
        assert abs(got - want) <= (2e-3 if dtype is np.float32 else 1e-9) * want, (got, want, lam)
    tight = gpu_api.operator_norm_matrix(A, M, N, 1e-7 if dtype is np.float64 else 1e-6, 3000, 10)
    assert lam * (1 - (2e-3 if dtype is np.float32 else 2e-4)) < tight <= lam * (1 + (1e-4 if dtype is np.float32 else 1e-9)), (tight, lam)


def test_simplex_rejects_bad_args(gpu_api):
    from cp_pfdr_graph_d1_b200.api import _b, PfdrError
    args, kw = problems.simplex_case("kla", np.float32)
    args = list(args)
    args[2] = 1.5
    with pytest.raises(PfdrError):
        _b().PFDR_graph_loss_d1_simplex(*args, **kw)


def test_resident_simplex_handle(gpu_api):
    args, kw = problems.simplex_case("kla_k16", np.float32)
    ref = problems.run_simplex(gpu_api._b(), "kla_k16", np.float32)
    P0, Q, al, Eu, Ev, La_d1 = args
    s = gpu_api.SimplexSolver(P0, Q, al, Eu, Ev, La_d1, rho=kw["rho"], condMin=kw["condMin"], difRcd=kw["difRcd"],
                              difTol=kw["difTol"], itMax=kw["itMax"], want_obj=True, want_dif=True)
    for _ in range(2):
        s.reset()
        s.run()
        r = s.fetch(want_obj=True, want_dif=True)
        assert r.it == ref.it and np.array_equal(r.X, ref.X) and np.array_equal(r.Obj, ref.Obj)
    s.close()
