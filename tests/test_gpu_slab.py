"""Slab-sharded solves on real GPUs (csrc/slab.cuh): the ranks exchange their halos by storing straight into each
other's memory (CUDA IPC between processes, peer access inside one process), one flag per rank pair and iteration.
Every case must give the single-GPU iterate - bit for bit when the evolution is not monitored, to rounding of the
stopping sums otherwise - the same iteration count and the same Dif record, and agree with the oracle.
Needs >= 2 GPUs (`gpurun --gpus 2`); skipped otherwise."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def case_inputs(case):
    from cp_pfdr_graph_d1_b200 import synth
    fam, _, tag = case.partition("_")
    dt = np.float32 if "f32" in tag else np.float64
    if fam == "simplex":
        K = 16 if "k16" in tag else 5
        n, nz = (12, 16) if "tall" in tag else (14, 14)
        w = synth.c4_simplex(n, K, dt, nz=nz)
        kw = dict(rho=1.5, condMin=1e-2, difTol=0.0 if "notrack" in tag else 1e-4, itMax=40 if "notrack" in tag else 200)
        if "recond" in tag:            # reconditions on the way (difRcd > 0) and records the objective
            kw.update(difRcd=1e-2, difTol=1e-5, itMax=300, want_obj=True)
        return fam, w, kw, dt
    w = synth.c2_denoise(20, dt, diag=True)
    w["Ys"] = (w["La_l2"] * w["Y"]).astype(dt)
    kw = dict(rho=1.5, condMin=1e-3, difTol=0.0 if "notrack" in tag else 1e-5, itMax=40 if "notrack" in tag else 300)
    if "recond" in tag:
        kw.update(difRcd=1e-3, difTol=1e-6, itMax=500, want_obj=True)
    return fam, w, kw, dt


def make_solver(pd, plan, fam, w, kw, dt, **extra):
    kw = dict(kw)
    if fam == "simplex":
        return pd.SlabSimplexSolver(plan, w["Q"], w["Q"], 0.05, w["La_d1"], want_dif=kw["difTol"] > 0, **kw, **extra)
    V = w["V"]
    return pd.SlabQuadraticSolver(plan, np.zeros(V, dt), w["Ys"], w["La_l2"], w["La_d1"], np.full(V, 0.05, dt), 1, Ltype=1,
                                  L=w["La_l2"], want_dif=kw["difTol"] > 0, **kw, **extra)


def _worker(rank, world, port, case, out):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cp_pfdr_graph_d1_b200 import api, dist as pd, shard
        api._lib.load().pfdr_b200_set_device(rank)
        fam, w, kw, dt = case_inputs(case)
        plan = shard.make_slab_plan(w["V"], w["Eu"], w["Ev"], world, rank)
        s = make_solver(pd, plan, fam, w, kw, dt)
        res = []
        for _ in range(2):                       # a second (reset, run) on the same handle: flags carry the run number
            st = s.run()
            r = s.fetch(want_dif=kw["difTol"] > 0, want_obj=kw.get("want_obj", False))
            res.append((pd.gather_owned(r.X, plan), r.it, r.Dif, r.Obj))
            s.reset()
        if rank == 0:
            assert np.array_equal(res[0][0], res[1][0]) and res[0][1] == res[1][1]
            np.savez(out, X=res[0][0], it=res[0][1], Dif=res[0][2] if res[0][2] is not None else np.zeros(0),
                     Obj=res[0][3] if res[0][3] is not None else np.zeros(0), gpus=st["gpus"], layout=st["layout"])
        s.close()
    finally:
        dist.destroy_process_group()


def single_gpu(gpu_api, oracle, case):
    fam, w, kw, dt = case_inputs(case)
    b = gpu_api._b()
    if fam == "simplex":
        args = (w["Q"], 0.05, w["Eu"], w["Ev"], w["La_d1"])
        k2 = dict(dict(difRcd=0.0), **kw, want_dif=kw["difTol"] > 0)
        return b.PFDR_graph_loss_d1_simplex(w["Q"].copy(), *args, **k2), oracle.PFDR_graph_loss_d1_simplex(w["Q"].copy(), *args, **k2)
    V = w["V"]
    args = (w["Ys"], w["La_l2"], w["Eu"], w["Ev"], w["La_d1"], np.full(V, 0.05, dt), 1, 1, w["La_l2"])
    k2 = dict(dict(difRcd=0.0), **kw, N=0, want_dif=kw["difTol"] > 0)
    return b.PFDR_graph_quadratic_d1_l1(np.zeros(V, dt), *args, **k2), oracle.PFDR_graph_quadratic_d1_l1(np.zeros(V, dt), *args, **k2)


def check(got_X, got_it, got_Dif, one, ref, dt, exact=True):
    rel = lambda a, c: float(np.linalg.norm(a.astype(np.float64) - c) / np.linalg.norm(c.astype(np.float64)))
    assert got_it == one.it, (got_it, one.it)
    tracked = one.Dif is not None and one.Dif.shape[0] > 0
    if not tracked and exact:
        assert np.array_equal(got_X, one.X)                    # no global sum involved: the very same arithmetic
    else:
        assert rel(got_X, one.X) < (1e-12 if dt is np.float64 else 1e-6)
    if tracked:
        assert np.allclose(got_Dif, one.Dif, rtol=1e-9 if dt is np.float64 else 1e-4), np.abs(got_Dif / one.Dif - 1).max()
    assert abs(one.it - ref.it) <= (0 if dt is np.float64 else max(2, ref.it // 20))
    assert rel(got_X, ref.X) < (1e-5 if dt is np.float64 else 1e-3)


CASES = ["simplex_k5", "simplex_k16", "simplex_k16notrack", "simplex_k16f32", "simplex_k5tall", "simplex_k16recond", "simplex_k5tallrecond",
         "quad_f64", "quad_f32", "quad_notrack", "quad_recond"]


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("case", CASES)
def test_slab_ranks_match_single_gpu(gpu_api, oracle, tmp_path, case, world):
    """one process per GPU (the way bench.py --gpus N runs under torchrun): arenas mapped through CUDA IPC"""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    if world == 4 and "tall" not in case and "quad" not in case:
        pytest.skip("14 planes over 4 ranks: slabs thinner than they need to be for a useful test")
    out = str(tmp_path / "r.npz")
    mp.spawn(_worker, args=(world, _free_port(), case, out), nprocs=world, join=True)
    got = np.load(out)
    assert int(got["gpus"]) == world and int(got["layout"]) == 1
    one, ref = single_gpu(gpu_api, oracle, case)
    # (quadratic family: the preconditioner's mean amplitude is a sum over all vertices, added rank by rank)
    check(got["X"], int(got["it"]), got["Dif"], one, ref, case_inputs(case)[3], exact=case.startswith("simplex"))
    if one.Obj is not None:                                      # objective record: every entry, the final iterate's included
        assert got["Obj"].shape == one.Obj.shape and np.allclose(got["Obj"], one.Obj, rtol=1e-10)


@pytest.mark.parametrize("case", ["simplex_k16", "simplex_k16notrack", "quad_f64", "quad_notrack"])
def test_slab_ranks_in_one_process(gpu_api, oracle, case):
    """all ranks inside ONE process, one host thread per GPU (the way the one-shot C entry points use several GPUs):
    arenas reached by plain peer access instead of CUDA IPC"""
    import threading
    import torch
    world = 2
    if torch.cuda.device_count() < world:
        pytest.skip("needs 2 GPUs")
    from cp_pfdr_graph_d1_b200 import api, dist as pd, shard
    lib = api._lib.load()
    fam, w, kw, dt = case_inputs(case)
    plans = [shard.make_slab_plan(w["V"], w["Eu"], w["Ev"], world, r) for r in range(world)]
    gate, blobs, out, errs = threading.Barrier(world), [None] * world, [None] * world, []

    def rank_thread(r):
        try:
            lib.pfdr_b200_set_device(r)
            s = make_solver(pd, plans[r], fam, w, kw, dt, connect=False)
            import ctypes as C
            mine = (C.c_ubyte * pd.SLAB_BLOB)()
            lib.pfdr_b200_slab_export.argtypes = [C.c_void_p, C.c_void_p]
            assert lib.pfdr_b200_slab_export(s._h, mine) == 0
            blobs[r] = bytes(mine)
            gate.wait()
            s.connect(b"".join(blobs), same_process=True)
            s.run()
            res = s.fetch(want_dif=kw["difTol"] > 0)
            out[r] = (plans[r].v0, res)
            gate.wait()
            s.close()
        except Exception as e:            # noqa: BLE001 - reported by the main thread
            errs.append(e)
            gate.abort()

    ts = [threading.Thread(target=rank_thread, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    lib.pfdr_b200_set_device(0)
    assert not errs, errs
    X = np.concatenate([o[1].X for o in sorted(out, key=lambda t: t[0])])
    one, ref = single_gpu(gpu_api, oracle, case)
    check(X, out[0][1].it, out[0][1].Dif, one, ref, dt, exact=case.startswith("simplex"))


@pytest.mark.parametrize("family", ["simplex", "quad"])
def test_one_shot_entry_points_on_two_gpus(gpu_api, monkeypatch, family):
    """PFDR_B200_GPUS=2: the reference-shaped one-shot call cuts a large grid problem into slabs itself (one host thread
    per GPU, sub-ranges of the caller's host arrays) - same iterate as on one GPU, bit for bit (difTol = 0)"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from cp_pfdr_graph_d1_b200 import synth
    b = gpu_api._b()
    if family == "simplex":
        w = synth.c4_simplex(64, 16, np.float32)
        call = lambda: b.PFDR_graph_loss_d1_simplex(w["Q"].copy(), w["Q"], 0.05, w["Eu"], w["Ev"], w["La_d1"], rho=1.5, condMin=1e-2,
                                                    difRcd=0.0, difTol=0.0, itMax=30)
    else:
        w = synth.c2_denoise(130, np.float32, diag=True)
        Ys = (w["La_l2"] * w["Y"]).astype(np.float32)
        call = lambda: b.PFDR_graph_quadratic_d1_l1(np.zeros(w["V"], np.float32), Ys, w["La_l2"], w["Eu"], w["Ev"], w["La_d1"], None, 0, 1,
                                                    w["La_l2"], rho=1.5, condMin=1e-3, difRcd=0.0, difTol=0.0, itMax=30, N=0)
    one = call()
    assert gpu_api.last_stats()["gpus"] == 1
    monkeypatch.setenv("PFDR_B200_GPUS", "2")
    two = call()
    st = gpu_api.last_stats()
    assert st["gpus"] == 2 and st["layout"] == 1, st
    assert two.it == one.it == 30
    if family == "simplex":
        assert np.array_equal(two.X, one.X)
    else:       # the preconditioner's mean amplitude is a sum over all vertices, added rank by rank
        assert np.linalg.norm(two.X.astype(np.float64) - one.X) / np.linalg.norm(one.X.astype(np.float64)) < 1e-6


@pytest.mark.parametrize("family", ["simplex", "quad"])
def test_one_shot_entry_points_on_two_gpus_full_semantics(gpu_api, monkeypatch, family):
    """the same with everything the reference's signature offers: stopping rule, Obj and Dif records, a reconditioning"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from cp_pfdr_graph_d1_b200 import synth
    b = gpu_api._b()
    if family == "simplex":
        w = synth.c4_simplex(64, 16, np.float64)
        call = lambda: b.PFDR_graph_loss_d1_simplex(w["Q"].copy(), w["Q"], 0.05, w["Eu"], w["Ev"], w["La_d1"], rho=1.5, condMin=1e-2,
                                                    difRcd=3e-3, difTol=1e-4, itMax=60, want_obj=True, want_dif=True)
    else:
        w = synth.c2_denoise(130, np.float64, diag=True)
        Ys = w["La_l2"] * w["Y"]
        call = lambda: b.PFDR_graph_quadratic_d1_l1(np.zeros(w["V"]), Ys, w["La_l2"], w["Eu"], w["Ev"], w["La_d1"], np.full(w["V"], 0.05), 1, 1,
                                                    w["La_l2"], rho=1.5, condMin=1e-3, difRcd=3e-2, difTol=1e-4, itMax=60, N=0,
                                                    want_obj=True, want_dif=True)
    one = call()
    monkeypatch.setenv("PFDR_B200_GPUS", "2")
    two = call()
    assert gpu_api.last_stats()["gpus"] == 2
    assert two.it == one.it and one.it > 5, (two.it, one.it)
    assert np.linalg.norm(two.X - one.X) / np.linalg.norm(one.X) < 1e-12
    assert two.Obj.shape == one.Obj.shape and np.allclose(two.Obj, one.Obj, rtol=1e-10)
    assert np.allclose(two.Dif, one.Dif, rtol=1e-9)


def test_one_shot_entry_points_fall_back_to_one_gpu(gpu_api, monkeypatch):
    """PFDR_B200_GPUS set, but the problem is not one the slab path takes (an irregular graph; a small grid): solved on one
    GPU, same result as without the variable"""
    import problems
    b = gpu_api._b()
    monkeypatch.setenv("PFDR_B200_GPUS", "2")
    r = problems.run_quad(b, "rand_big", np.float64)
    assert gpu_api.last_stats()["gpus"] == 1
    monkeypatch.delenv("PFDR_B200_GPUS")
    r1 = problems.run_quad(b, "rand_big", np.float64)
    assert r.it == r1.it and np.array_equal(r.X, r1.X)
    V = 1 << 21                                                  # large enough, but the edge list is not vertex-major: table layout
    rng = np.random.default_rng(0)
    Eu = np.arange(V - 1, dtype=np.int32); Ev = Eu + 1
    perm = rng.permutation(V - 1)
    Y = rng.normal(0, 1, V).astype(np.float32)
    monkeypatch.setenv("PFDR_B200_GPUS", "2")
    r = b.PFDR_graph_quadratic_d1_l1(np.zeros(V, np.float32), Y, None, Eu[perm], Ev[perm], np.full(V - 1, 0.3, np.float32), None, 0, 0, None,
                                     difTol=0.0, itMax=5, N=0)
    assert gpu_api.last_stats()["gpus"] == 1 and r.it == 5 and np.isfinite(r.X).all()


def test_template_symbol_of_the_shim_on_two_gpus(gpu_api, monkeypatch):
    """the drop-in boundary itself: the reference's C++ template instantiation PFDR_graph_loss_d1_simplex<float>
    (mangled symbol of lib/libpfdr_b200_shim.so, what the MEX gateways and CP_*.cpp link against) on two GPUs"""
    import ctypes as C
    import os
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from cp_pfdr_graph_d1_b200 import _lib, synth
    shim = C.CDLL(os.path.join(os.path.dirname(_lib.LIB_PATH), "libpfdr_b200_shim.so"))
    f = getattr(shim, "_Z26PFDR_graph_loss_d1_simplexIfEviiiT_PKS0_PS0_S2_PKiS5_S2_S0_S0_S0_S0_iPiS3_S3_i")
    PF, PI = C.POINTER(C.c_float), C.POINTER(C.c_int)
    f.restype = None
    f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, PF, PF, PF, PI, PI, PF, C.c_float, C.c_float, C.c_float, C.c_float,
                  C.c_int, PI, PF, PF, C.c_int]
    w = synth.c4_simplex(64, 16, np.float32)
    Eu, Ev = np.ascontiguousarray(w["Eu"], np.int32), np.ascontiguousarray(w["Ev"], np.int32)
    res = []
    for gpus in ("1", "2"):
        monkeypatch.setenv("PFDR_B200_GPUS", gpus)
        P = w["Q"].copy()
        it = C.c_int(0)
        f(16, w["V"], w["E"], 0.05, None, P.ctypes.data_as(PF), w["Q"].ctypes.data_as(PF), Eu.ctypes.data_as(PI), Ev.ctypes.data_as(PI),
          w["La_d1"].ctypes.data_as(PF), 1.5, 1e-2, 0.0, 0.0, 25, C.byref(it), None, None, 0)
        res.append((P, it.value, gpu_api.last_stats()["gpus"]))
    assert res[0][2] == 1 and res[1][2] == 2
    assert res[0][1] == res[1][1] == 25 and np.array_equal(res[0][0], res[1][0])
