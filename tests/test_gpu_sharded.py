"""Sharded (multi-GPU) path on real GPUs: 2 ranks over NCCL must reproduce the single-GPU solve and the oracle.
Needs >= 2 GPUs (`gpurun --gpus 2`); skipped otherwise."""
import os
import socket

import numpy as np
import pytest

import problems

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, out):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cp_pfdr_graph_d1_b200 import dist as pd, shard, synth
        pd.init_comm(rank)
        if case == "quad":
            w = synth.c2_denoise(20, np.float64, diag=True)
            A = w["La_l2"]
            Y = (A * w["Y"])
            plan = shard.make_plan(w["V"], w["Eu"], w["Ev"], world, rank)
            s = pd.ShardedQuadraticSolver(plan, np.zeros(w["V"]), Y, A, w["La_d1"], np.full(w["V"], 0.05), 1, Ltype=1, L=A,
                                          rho=1.5, condMin=1e-3, difTol=1e-5, itMax=300, want_dif=True)
        elif case == "quad_rand":
            V = 3000
            Eu, Ev = problems.random_graph(V, 3 * V, 5)
            rng = np.random.default_rng(3)
            A = rng.uniform(0.5, 1.5, V)
            Y = A * (np.round(rng.uniform(0, 3, V)) + rng.normal(0, 0.2, V))
            plan = shard.make_plan(V, Eu, Ev, world, rank)
            s = pd.ShardedQuadraticSolver(plan, np.zeros(V), Y, A, rng.uniform(0.05, 0.5, Eu.shape[0]), None, 0, 0.2, 2.5,
                                          "bounds", Ltype=1, L=A, difTol=1e-5, itMax=300, want_dif=True)
        elif case == "dense":
            w = synth.dense_problem(12, 40, np.float64)
            lam = np.array([float(np.linalg.norm(w["A"], 2) ** 2)])
            plan = shard.make_plan(w["V"], w["Eu"], w["Ev"], world, rank)
            s = pd.ShardedQuadraticSolver(plan, np.zeros(w["V"]), w["Y"], w["A"], w["La_d1"], np.full(w["V"], 0.01), 1,
                                          Ltype=0, L=lam, rho=1.5, condMin=1e-3, difTol=1e-5, itMax=200, want_dif=True)
        else:
            w = synth.c4_simplex(14, 16 if case == "simplex16" else 5, np.float64)
            plan = shard.make_plan(w["V"], w["Eu"], w["Ev"], world, rank)
            s = pd.ShardedSimplexSolver(plan, w["Q"], w["Q"], 0.05, w["La_d1"], rho=1.5, condMin=1e-2, difTol=1e-4,
                                        itMax=200, want_dif=True)
        st = s.run()
        r = s.fetch(want_dif=True)
        full = pd.gather_owned(r.X, plan)
        if rank == 0:
            np.savez(out, X=full, it=r.it, Dif=r.Dif, launches=st["launches"])
        s.close()
        pd.destroy_comm()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["quad", "quad_rand", "dense", "simplex5", "simplex16"])
def test_two_ranks_match_single_gpu(gpu_api, oracle, tmp_path, case):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from cp_pfdr_graph_d1_b200 import synth
    out = str(tmp_path / "r.npz")
    mp.spawn(_worker, args=(2, _free_port(), case, out), nprocs=2, join=True)
    got = np.load(out)
    b = gpu_api._b()
    if case == "quad":
        w = synth.c2_denoise(20, np.float64, diag=True)
        A = w["La_l2"]
        args = (np.zeros(w["V"]), A * w["Y"], A, w["Eu"], w["Ev"], w["La_d1"], np.full(w["V"], 0.05), 1, 1, A)
        kw = dict(rho=1.5, condMin=1e-3, difRcd=0.0, difTol=1e-5, itMax=300, N=0, want_dif=True)
        one = b.PFDR_graph_quadratic_d1_l1(args[0].copy(), *args[1:], **kw)
        ref = oracle.PFDR_graph_quadratic_d1_l1(args[0].copy(), *args[1:], **kw)
    elif case == "quad_rand":
        V = 3000
        Eu, Ev = problems.random_graph(V, 3 * V, 5)
        rng = np.random.default_rng(3)
        A = rng.uniform(0.5, 1.5, V)
        Y = A * (np.round(rng.uniform(0, 3, V)) + rng.normal(0, 0.2, V))
        args = (np.zeros(V), Y, A, Eu, Ev, rng.uniform(0.05, 0.5, Eu.shape[0]), 0.2, 2.5, 1, A)
        kw = dict(rho=1.5, condMin=1e-3, difRcd=0.0, difTol=1e-5, itMax=300, N=0, want_dif=True)
        one = b.PFDR_graph_quadratic_d1_bounds(args[0].copy(), *args[1:], **kw)
        ref = oracle.PFDR_graph_quadratic_d1_bounds(args[0].copy(), *args[1:], **kw)
    elif case == "dense":
        w = synth.dense_problem(12, 40, np.float64)
        lam = np.array([float(np.linalg.norm(w["A"], 2) ** 2)])
        args = (np.zeros(w["V"]), w["Y"], w["A"], w["Eu"], w["Ev"], w["La_d1"], np.full(w["V"], 0.01), 1, 0, lam)
        kw = dict(rho=1.5, condMin=1e-3, difRcd=0.0, difTol=1e-5, itMax=200, N=w["N"], want_dif=True)
        one = b.PFDR_graph_quadratic_d1_l1(args[0].copy(), *args[1:], **kw)
        ref = oracle.PFDR_graph_quadratic_d1_l1(args[0].copy(), *args[1:], **kw)
    else:
        w = synth.c4_simplex(14, 16 if case == "simplex16" else 5, np.float64)
        args = (w["Q"].copy(), w["Q"], 0.05, w["Eu"], w["Ev"], w["La_d1"])
        kw = dict(rho=1.5, condMin=1e-2, difRcd=0.0, difTol=1e-4, itMax=200, want_dif=True)
        one = b.PFDR_graph_loss_d1_simplex(*args, **kw)
        ref = oracle.PFDR_graph_loss_d1_simplex(w["Q"].copy(), *args[1:], **kw)
    rel = lambda a, c: float(np.linalg.norm(a - c) / np.linalg.norm(c))
    assert int(got["it"]) == one.it == ref.it
    # same arithmetic; only reduction trees differ (the stopping scalars; for dense A also the N-vector A x)
    assert rel(got["X"], one.X) < (1e-12 if case != "dense" else 1e-9)
    assert rel(got["X"], ref.X) < 1e-5
    assert np.allclose(got["Dif"], one.Dif, rtol=1e-9 if case != "dense" else 1e-5)
