"""GPU parity of the quadratic family: CUDA path (through the C ABI) vs the CPU oracle on the same
seeded inputs.  Tolerances are BASELINE.json's: 1e-5 relative in float64, 1e-3 in float32."""
import numpy as np
import pytest

import problems

pytestmark = pytest.mark.gpu
TOL = {np.float32: 1e-3, np.float64: 1e-5}


def rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(float(np.linalg.norm(b)), 1e-300))


@pytest.mark.parametrize("mode", ["large", "table", "small", "small_c1", "small_c16"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name", problems.QUAD_CASES)
def test_quadratic_parity(gpu_api, oracle, monkeypatch, name, dtype, mode):
    """mode "small": problems that fit run as ONE launch of the single-CTA persistent kernel (the path the
    cut-pursuit reduced problems take); mode "large": always the one-kernel-per-iteration path (grid cases on the
    stencil layout); mode "table": the same on the half-edge table whatever the graph"""
    from cp_pfdr_graph_d1_b200.api import _b
    monkeypatch.setenv("PFDR_B200_SMALL", "0" if mode in ("large", "table") else "1")
    if mode == "table":
        monkeypatch.setenv("PFDR_B200_DIA", "0")
    if "_c" in mode:
        monkeypatch.setenv("PFDR_B200_SMALL_C", mode.split("_c")[1])      # CTAs of the cluster (default: by size)
    got = problems.run_quad(_b(), name, dtype)
    ref = problems.run_quad(oracle, name, dtype)
    tol = TOL[dtype]
    slack = 0 if dtype is np.float64 else max(2, ref.it // 50)
    if "recond" in name and dtype is np.float32:
        slack = max(slack, ref.it // 10)       # a reconditioning firing one iteration apart shifts the tail
    assert abs(got.it - ref.it) <= slack, (got.it, ref.it)
    assert rel(got.X, ref.X) < tol, rel(got.X, ref.X)
    n = min(got.it, ref.it)
    if ref.Obj is not None:
        assert got.Obj.shape[0] == got.it + 1
        assert np.allclose(got.Obj[:n + 1], ref.Obj[:n + 1], rtol=tol, atol=tol * abs(ref.Obj[0]) * 1e-2), \
            np.abs(got.Obj[:n + 1] - ref.Obj[:n + 1]).max()
    if ref.Dif is not None and n:
        assert got.Dif.shape[0] == got.it
        # compare the evolution record above the rounding floor of the type (it is a squared relative change)
        sel = ref.Dif[:n] > (1e-9 if dtype is np.float32 else 1e-24)
        assert np.allclose(got.Dif[:n][sel], ref.Dif[:n][sel], rtol=5e-2 if dtype is np.float32 else 1e-4)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name", [n for n in problems.QUAD_CASES if n.startswith(("grid", "dense"))])
def test_stencil_layout_matches_table(gpu_api, monkeypatch, name, dtype):
    """grid problems on the one-kernel-per-iteration path use the constant-offset layout; the half-edge table
    (PFDR_B200_DIA=0) must give the same bits: same arithmetic, same summation order"""
    from cp_pfdr_graph_d1_b200.api import _b
    monkeypatch.setenv("PFDR_B200_SMALL", "0")
    got = problems.run_quad(_b(), name, dtype)
    assert gpu_api.last_stats()["layout"] == 1
    monkeypatch.setenv("PFDR_B200_DIA", "0")
    tab = problems.run_quad(_b(), name, dtype)
    assert gpu_api.last_stats()["layout"] == 0
    assert got.it == tab.it and np.array_equal(got.X, tab.X)
    if got.Obj is not None:
        assert np.array_equal(got.Obj, tab.Obj)
    if got.Dif is not None:      # the evolution sums are associated differently (per-lane double accumulators)
        assert np.allclose(got.Dif, tab.Dif, rtol=1e-5 if dtype is np.float32 else 1e-12, atol=0)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("variant", ["plain", "l1_pos", "box", "identity", "chain", "image"])
def test_stencil_layout_matches_table_larger_grids(gpu_api, monkeypatch, variant, dtype):
    """16^3 grids, 8192-vertex chains and 64 x 96 images (several trips per persistent warp) with every vertex prox:
    the stencil kernels give the same bits as the half-edge table"""
    from cp_pfdr_graph_d1_b200 import synth
    from cp_pfdr_graph_d1_b200.api import _b
    monkeypatch.setenv("PFDR_B200_SMALL", "0")
    if variant == "chain":
        V = 8192
        Eu = np.arange(V - 1, dtype=np.int32); Ev = Eu + 1
    elif variant == "image":
        nx, ny = 64, 96
        V = nx * ny
        v = np.arange(V).reshape(ny, nx)
        eu = np.stack([v, v], -1); ev = np.stack([v + 1, v + nx], -1)
        ok = np.zeros(eu.shape, bool); ok[:, :-1, 0] = True; ok[:-1, :, 1] = True
        Eu, Ev = eu[ok].astype(np.int32), ev[ok].astype(np.int32)
    else:
        V, Eu, Ev = synth.grid_graph(16)
    rng = np.random.default_rng(5)
    A = None if variant == "identity" else rng.uniform(0.5, 1.5, V).astype(dtype)
    Y = (np.round(4 * np.arange(V) / V) + rng.normal(0, 0.3, V)).astype(dtype)
    if A is not None:
        Y = (A * Y).astype(dtype)
    La_d1 = rng.uniform(0.1, 0.5, Eu.shape[0]).astype(dtype)
    kw = dict(N=0, rho=1.5, condMin=1e-3, difRcd=0.0, difTol=1e-6, itMax=150, want_obj=True, want_dif=True)
    Lt, L = (1, A) if A is not None else (0, None)
    if variant == "box":
        call = lambda: _b().PFDR_graph_quadratic_d1_bounds(np.zeros(V, dtype), Y, A, Eu, Ev, La_d1, 0.5, 3.0, Lt, L, **kw)
    else:
        l1 = np.full(V, 0.05, dtype) if variant == "l1_pos" else None
        call = lambda: _b().PFDR_graph_quadratic_d1_l1(np.zeros(V, dtype), Y, A, Eu, Ev, La_d1, l1, int(variant == "l1_pos"), Lt, L, **kw)
    got = call()
    assert gpu_api.last_stats()["layout"] == 1
    monkeypatch.setenv("PFDR_B200_DIA", "0")
    tab = call()
    assert gpu_api.last_stats()["layout"] == 0
    assert got.it == tab.it and np.array_equal(got.X, tab.X) and np.array_equal(got.Obj, tab.Obj)
    assert np.allclose(got.Dif, tab.Dif, rtol=1e-5 if dtype is np.float32 else 1e-12, atol=0)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("kind", ["chain", "image", "image8"])
@pytest.mark.parametrize("recond", [0.0, 1e-3])
def test_stencil_layout_chain_and_image(gpu_api, oracle, monkeypatch, kind, dtype, recond):
    """1-D chains (one offset), 2-D images (two offsets) and 8-connected images (four offsets: more than the stencil
    layout holds, so the table is used) against the oracle, l1 + positivity"""
    from cp_pfdr_graph_d1_b200.api import _b
    monkeypatch.setenv("PFDR_B200_SMALL", "0")
    rng = np.random.default_rng(len(kind))
    if kind == "chain":
        V = 3001
        Eu = np.arange(V - 1, dtype=np.int32); Ev = Eu + 1
    else:
        nx, ny = 61, 47
        V = nx * ny
        v = np.arange(V).reshape(ny, nx)
        offs = [(1, (slice(None), slice(None, -1))), (nx, (slice(None, -1), slice(None)))]
        if kind == "image8":
            offs = [(1, (slice(None), slice(None, -1))), (nx - 1, (slice(None, -1), slice(1, None))),
                    (nx, (slice(None, -1), slice(None))), (nx + 1, (slice(None, -1), slice(None, -1)))]
        eu = np.stack([v] * len(offs), -1); ev = np.stack([v + d for d, _ in offs], -1)
        ok = np.zeros(eu.shape, bool)
        for i, (_, sl) in enumerate(offs):
            ok[sl + (i,)] = True
        Eu, Ev = eu[ok].astype(np.int32), ev[ok].astype(np.int32)
    A = rng.uniform(0.5, 1.5, V).astype(dtype)
    Y = (A * (np.round(3 * np.arange(V) / V) + rng.normal(0, 0.3, V))).astype(dtype)
    La_d1 = rng.uniform(0.05, 0.5, Eu.shape[0]).astype(dtype)
    args = (Y, A, Eu, Ev, La_d1, np.full(V, 0.02, dtype), 1, 1, A)
    kw = dict(N=0, rho=1.5, condMin=1e-3, difRcd=recond, difTol=1e-6, itMax=400, want_obj=True, want_dif=True)
    got = _b().PFDR_graph_quadratic_d1_l1(np.zeros(V, dtype), *args, **kw)
    assert gpu_api.last_stats()["layout"] == (0 if kind == "image8" else 1)
    ref = oracle.PFDR_graph_quadratic_d1_l1(np.zeros(V, dtype), *args, **kw)
    tol = TOL[dtype]
    slack = 0 if dtype is np.float64 else max(2, ref.it // 10)
    assert abs(got.it - ref.it) <= slack, (got.it, ref.it)
    assert rel(got.X, ref.X) < tol, rel(got.X, ref.X)
    n = min(got.it, ref.it)
    assert np.allclose(got.Obj[:n + 1], ref.Obj[:n + 1], rtol=tol, atol=tol * abs(ref.Obj[0]) * 1e-2)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_exact_zeros_and_bounds(gpu_api, dtype):
    """cut-pursuit branches on exact zeros (reference src/CP_PFDR_graph_quadratic_d1_l1.cpp:357): the l1 prox
    must emit exact 0 and the box prox exact bounds"""
    from cp_pfdr_graph_d1_b200.api import _b
    r = problems.run_quad(_b(), "dense_direct", dtype)
    assert (r.X >= 0).all() and (r.X == 0).sum() > 100
    r = problems.run_quad(_b(), "grid_diag_box2", dtype)
    assert r.X.min() == dtype(0.5) and r.X.max() == dtype(2.5)


@pytest.mark.parametrize("csize", ["auto", "1", "2", "16"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name", [n for n in problems.QUAD_CASES if "recond" not in n])
def test_quadratic_parity_push(gpu_api, oracle, monkeypatch, name, dtype, csize):
    """the way cut-pursuit calls PFDR (no objective record, no reconditioning) small problems run in the push
    kernel: one launch of a thread-block cluster, CTAs exchanging through distributed shared memory"""
    from cp_pfdr_graph_d1_b200.api import _b
    monkeypatch.setenv("PFDR_B200_SMALL", "1")
    if csize != "auto":
        monkeypatch.setenv("PFDR_B200_SMALL_C", csize)
    got = problems.run_quad(_b(), name, dtype, want_obj=False)
    st = gpu_api.last_stats()
    ref = problems.run_quad(oracle, name, dtype, want_obj=False)
    tol = TOL[dtype]
    slack = 0 if dtype is np.float64 else max(2, ref.it // 50)
    assert st["loop_launches"] == 1 or ref.it == 0, st
    assert abs(got.it - ref.it) <= slack, (got.it, ref.it)
    assert rel(got.X, ref.X) < tol, rel(got.X, ref.X)
    n = min(got.it, ref.it)
    if ref.Dif is not None and n:
        assert got.Dif.shape[0] == got.it
        sel = ref.Dif[:n] > (1e-9 if dtype is np.float32 else 1e-24)
        assert np.allclose(got.Dif[:n][sel], ref.Dif[:n][sel], rtol=5e-2 if dtype is np.float32 else 1e-4)


@pytest.mark.parametrize("name", ["grid_diag", "dense_direct", "dense_ata_box"])
def test_small_mode_is_one_launch(gpu_api, monkeypatch, name):
    """small reduced problems (cut-pursuit) must not pay one launch per iteration"""
    from cp_pfdr_graph_d1_b200.api import _b
    if name not in problems.QUAD_CASES:
        pytest.skip("no such case")
    monkeypatch.setenv("PFDR_B200_SMALL", "1")
    r = problems.run_quad(_b(), name, np.float64)
    st = gpu_api.last_stats()
    assert r.it > 1 and st["loop_launches"] == 1, st
    monkeypatch.setenv("PFDR_B200_SMALL", "0")
    r2 = problems.run_quad(_b(), name, np.float64)
    assert gpu_api.last_stats()["loop_launches"] >= r2.it
    assert abs(r.it - r2.it) <= 1 and rel(r.X, r2.X) < 1e-9


def test_deterministic(gpu_api):
    from cp_pfdr_graph_d1_b200.api import _b
    a = problems.run_quad(_b(), "rand_big", np.float32)
    b = problems.run_quad(_b(), "rand_big", np.float32)
    assert a.it == b.it and np.array_equal(a.X, b.X) and np.array_equal(a.Dif, b.Dif)


def test_device_pointers(gpu_api):
    """inputs already resident in HBM (torch CUDA tensors) are used in place, X written in place"""
    import torch
    from cp_pfdr_graph_d1_b200.api import _b
    fam, args, kw = problems.quad_case("grid_diag", np.float32)
    ref = problems.run_quad(_b(), "grid_diag", np.float32)
    dev = [torch.from_numpy(np.ascontiguousarray(a)).cuda() if isinstance(a, np.ndarray) else a for a in args]
    torch.cuda.synchronize()
    r = _b().PFDR_graph_quadratic_d1_l1(*dev, **kw)
    assert r.it == ref.it
    assert np.array_equal(dev[0].cpu().numpy(), ref.X)
    st = gpu_api.last_stats()
    assert st["h2d_bytes"] == 0 and st["launches"] > 0 and st["iterations"] == ref.it


def test_bad_edges_rejected(gpu_api):
    from cp_pfdr_graph_d1_b200.api import _b, PfdrError
    fam, args, kw = problems.quad_case("grid_diag", np.float32)
    args = list(args)
    args[3] = args[3].copy()
    args[3][5] = 10 ** 7
    with pytest.raises(PfdrError):
        _b().PFDR_graph_quadratic_d1_l1(*args, **kw)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_resident_handle_matches_one_shot(gpu_api, dtype):
    """create + (reset, run) x 2 + fetch == the one-shot entry point, bit for bit"""
    fam, args, kw = problems.quad_case("grid_diag_l1_pos", dtype)
    ref = problems.run_quad(gpu_api._b(), "grid_diag_l1_pos", dtype)
    X0, Y, A, Eu, Ev, La_d1, La_l1, pos, Ltype, L = args
    s = gpu_api.QuadraticSolver(X0, Y, A, Eu, Ev, La_d1, La_l1, pos, Ltype=Ltype, L=L, rho=kw["rho"],
                                condMin=kw["condMin"], difRcd=kw["difRcd"], difTol=kw["difTol"], itMax=kw["itMax"],
                                N=0, want_obj=True, want_dif=True)
    for _ in range(2):
        s.reset()
        st = s.run()
        r = s.fetch(want_obj=True, want_dif=True)
        assert r.it == ref.it == st["iterations"]
        assert np.array_equal(r.X, ref.X) and np.array_equal(r.Obj, ref.Obj) and np.array_equal(r.Dif, ref.Dif)
    s.close()


@pytest.mark.parametrize("small", ["0", "1"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name", ["grid_diag_recond", "rand_big_recond", "dense_ata_recond"])
def test_resident_handle_with_reconditioning(gpu_api, monkeypatch, name, dtype, small):
    """(reset, run) x 3 on a handle whose solve reconditions (difRcd > 0): reset() drops the per-slot thresholds and
    the next reconditioning must bring them back (both the one-kernel-per-iteration and the cluster path)"""
    monkeypatch.setenv("PFDR_B200_SMALL", small)
    fam, args, kw = problems.quad_case(name, dtype)
    ref = problems.run_quad(gpu_api._b(), name, dtype)
    X0, Y, A, Eu, Ev, La_d1, La_l1, pos, Ltype, L = args
    s = gpu_api.QuadraticSolver(X0, Y, A, Eu, Ev, La_d1, La_l1, pos, Ltype=Ltype, L=L, rho=kw["rho"],
                                condMin=kw["condMin"], difRcd=kw["difRcd"], difTol=kw["difTol"], itMax=kw["itMax"],
                                N=kw["N"], want_obj=True, want_dif=True)
    for _ in range(3):
        s.reset()
        s.run()
        r = s.fetch(want_obj=True, want_dif=True)
        assert r.it == ref.it
        assert np.array_equal(r.X, ref.X) and np.array_equal(r.Dif, ref.Dif)
    s.close()


@pytest.mark.parametrize("V", [2, 5, 31, 32, 33, 64, 100, 257, 1000, 3000])
@pytest.mark.parametrize("kind", ["diag", "dense", "ata"])
def test_paths_agree_on_irregular_graphs(gpu_api, oracle, monkeypatch, V, kind):
    """the three schedules of the quadratic solver (one kernel per iteration; cluster kernel exchanging through L2;
    cluster kernel exchanging through shared-memory pushes) on multigraphs with self-loops, duplicate edges and a hub,
    from two vertices up: same iteration count, same minimiser to rounding (float64), and the oracle's minimiser
    (pinned against the reference on the same inputs by tests/test_oracle_vs_ref.py)"""
    from cp_pfdr_graph_d1_b200.api import _b
    if kind != "diag" and V > 1000:
        pytest.skip("dense operator: small sizes only")
    fam, args, kw = problems.irregular_case(V, kind)

    def fresh():
        return tuple(x.copy() if isinstance(x, np.ndarray) else x for x in args)
    res = {}
    for mode, env in (("large", {"PFDR_B200_SMALL": "0"}), ("l2", {"PFDR_B200_SMALL": "1", "PFDR_B200_PUSH": "0"}),
                      ("push", {"PFDR_B200_SMALL": "1", "PFDR_B200_PUSH": "1"})):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        res[mode] = _b().PFDR_graph_quadratic_d1_l1(*fresh(), **kw)
    ref = res["large"]
    assert np.isfinite(ref.X).all()
    for mode in ("l2", "push"):
        r = res[mode]
        assert abs(r.it - ref.it) <= 1, (mode, r.it, ref.it)
        assert rel(r.X, ref.X) < 1e-8, (mode, rel(r.X, ref.X))
    orc = oracle.PFDR_graph_quadratic_d1_l1(*fresh(), **kw)
    assert abs(orc.it - ref.it) <= 1 and rel(ref.X, orc.X) < 1e-5, (orc.it, ref.it, rel(ref.X, orc.X))


def test_verbose_progress_lines(gpu_api, monkeypatch, capfd):
    """`verbose` of the reference (src/PFDR_graph_quadratic_d1_l1.cpp:43-55, :425-428): a message every `verbose`
    iterations with the iteration count and the evolution measure against its thresholds, and the final information"""
    import re
    from cp_pfdr_graph_d1_b200.api import _b
    monkeypatch.setenv("PFDR_B200_SMALL", "0")
    r = problems.run_quad(_b(), "grid_diag", np.float64, verbose=25, want_obj=False)
    out = capfd.readouterr().out
    its = [int(m) for m in re.findall(r"^iteration (\d+) \(max\. 300\)$", out, flags=re.M)]
    assert its[:-1] == list(range(25, r.it + 1, 25))[:len(its) - 1] and its[-1] == r.it, (its, r.it)
    evo = re.findall(r"^iterate evolution (\S+) \(recond\. (\S+); tol\. (\S+)\)$", out, flags=re.M)
    assert len(evo) == len(its)
    assert abs(float(evo[0][0]) - float(r.Dif[24])) <= 1e-5 * float(r.Dif[24])          # the value of THAT iteration
    assert abs(float(evo[0][2]) - 1e-10) < 1e-16                                        # difTol^2, as the reference prints it
