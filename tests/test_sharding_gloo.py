"""The N>1 path on CPU: world_size-2 (and 3) `gloo` process groups run the sharded schedule of
cp_pfdr_graph_d1_b200/shard.py (numpy per-rank compute, halo exchange + two-scalar all-reduce over gloo) and
must reproduce the unsharded oracle.  Also checks the plan's invariants without any process group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cp_pfdr_graph_d1_b200 import shard, synth

import dia_model
import problems
import sell_model


class GlooComm:
    def allreduce(self, a):
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
        dist.all_reduce(t)
        return t.numpy()

    def exchange(self, sends, recv_counts):
        ops, recv = [], {}
        for peer, buf in sends.items():
            s = torch.from_numpy(np.ascontiguousarray(buf))
            r = torch.empty(recv_counts[peer], dtype=s.dtype)
            recv[peer] = r
            ops += [dist.P2POp(dist.isend, s, peer), dist.P2POp(dist.irecv, r, peer)]
        for w in dist.batch_isend_irecv(ops) if ops else []:
            w.wait()
        return {p: r.numpy() for p, r in recv.items()}


def _worker(rank, world, port, case, dtype_name, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dtype = np.dtype(dtype_name).type
        V, Eu, Ev, X0, Y, A, La_d1, box = _problem(case, dtype)
        plan = shard.make_plan(V, Eu, Ev, world, rank)
        x, it = sell_model.quad_n0_sharded(plan, GlooComm(), X0, Y, A, La_d1, 1.5, 1e-5, 120, box)
        parts = [None] * world
        dist.all_gather_object(parts, (plan.v0, x, it))
        if rank == 0:
            full = np.concatenate([p[1] for p in sorted(parts, key=lambda t: t[0])])
            np.save(out, full)
            np.save(out + ".it.npy", np.array([p[2] for p in parts]))
    finally:
        dist.destroy_process_group()


def _problem(case, dtype):
    if case == "grid":
        w = synth.c2_denoise(9, dtype, diag=True)
        A = w["La_l2"]
        return w["V"], w["Eu"], w["Ev"], np.zeros(w["V"], dtype), (A * w["Y"]).astype(dtype), A, w["La_d1"], None
    V = 240
    Eu, Ev = problems.random_graph(V, 3 * V, 5)
    rng = np.random.default_rng(3)
    A = rng.uniform(0.5, 1.5, V).astype(dtype)
    Y = (A * (np.round(rng.uniform(0, 3, V)) + rng.normal(0, 0.2, V))).astype(dtype)
    return V, Eu, Ev, np.zeros(V, dtype), Y, A, rng.uniform(0.05, 0.5, Eu.shape[0]).astype(dtype), (dtype(0.2), dtype(2.5))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", ["grid", "random"])
def test_sharded_schedule_matches_oracle(oracle, tmp_path, world, case):
    dtype = np.float64
    out = str(tmp_path / "x.npy")
    mp.spawn(_worker, args=(world, _free_port(), case, np.dtype(dtype).name, out), nprocs=world, join=True)
    got, its = np.load(out), np.load(out + ".it.npy")
    V, Eu, Ev, X0, Y, A, La_d1, box = _problem(case, dtype)
    kw = dict(rho=1.5, condMin=1e-3, difRcd=0.0, difTol=1e-5, itMax=120, N=0)
    if box is None:
        ref = oracle.PFDR_graph_quadratic_d1_l1(X0.copy(), Y, A, Eu, Ev, La_d1, None, 0, 1, A, **kw)
    else:
        ref = oracle.PFDR_graph_quadratic_d1_bounds(X0.copy(), Y, A, Eu, Ev, La_d1, float(box[0]), float(box[1]), 1, A, **kw)
    assert (its == ref.it).all(), (its, ref.it)                 # every rank stops at the same iteration
    assert np.linalg.norm(got - ref.X) / np.linalg.norm(ref.X) < 1e-10


@pytest.mark.parametrize("world", [2, 4, 7])
def test_plan_invariants(world):
    V = 500
    Eu, Ev = problems.random_graph(V, 4 * V, 9)
    plans = [shard.make_plan(V, Eu, Ev, world, r) for r in range(world)]
    owners = np.zeros(V, int)
    covered = np.zeros(Eu.shape[0], int)
    for p in plans:
        owners[p.v0:p.v1] += 1
        covered[p.edge_ids] += 1
        assert (np.diff(p.edge_ids) > 0).all() and (np.diff(p.ghosts) > 0).all()
        # local numbering round-trips
        gl = np.concatenate([np.arange(p.v0, p.v1), p.ghosts])
        assert np.array_equal(gl[p.Eu], Eu[p.edge_ids]) and np.array_equal(gl[p.Ev], Ev[p.edge_ids])
        for nb in p.neighbors:
            q = next(m for m in plans[nb.rank].neighbors if m.rank == p.rank)
            gq = np.concatenate([np.arange(plans[nb.rank].v0, plans[nb.rank].v1), plans[nb.rank].ghosts])
            # what we send is what they expect, element for element
            assert np.array_equal(gl[nb.send_vertices], gq[q.recv_vertices])
            assert np.array_equal(p.edge_ids[nb.send_halfedges >> 1], plans[nb.rank].edge_ids[q.recv_halfedges >> 1])
            assert np.array_equal(nb.send_halfedges & 1, q.recv_halfedges & 1)
            assert (nb.send_vertices < p.Vown).all() and (nb.recv_vertices >= p.Vown).all()
    assert (owners == 1).all()
    assert ((covered == 1) | (covered == 2)).all()              # cut edges live on exactly two ranks


# ---------------------------------------------------------------------------------------------------------------
# slab sharding of grid graphs (shard.SlabPlan, csrc/slab.cuh): the ranks WRITE into each other's ghost zones
# ---------------------------------------------------------------------------------------------------------------
class GlooSlabComm(GlooComm):
    """the peer stores of the CUDA path as messages: every rank sends its neighbours (array, plane, position, values)"""

    def __init__(self, plan):
        r = [None] * plan.world
        dist.all_gather_object(r, (plan.lo, plan.hi, plan.v0, plan.v1))
        self.ranges, self.plan = r, plan

    def push(self, msgs):
        got = []
        box = [None] * self.plan.world
        dist.all_gather_object(box, msgs)                      # tiny problems: an all-gather of the per-destination lists
        for sender in box:
            got += sender.get(self.plan.rank, [])
        return got


def _slab_worker(rank, world, port, n, dtype_name, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dtype = np.dtype(dtype_name).type
        w = synth.c2_denoise(n, dtype, diag=True)
        A = w["La_l2"]
        plan = shard.make_slab_plan(w["V"], w["Eu"], w["Ev"], world, rank)
        x, it = dia_model.quad_n0_slab(plan, GlooSlabComm(plan), np.zeros(w["V"], dtype), (A * w["Y"]).astype(dtype), A, w["La_d1"],
                                       1.5, 1e-5, 120, (dtype(0.2), dtype(2.5)))
        parts = [None] * world
        dist.all_gather_object(parts, (plan.v0, x, it))
        if rank == 0:
            np.save(out, np.concatenate([p[1] for p in sorted(parts, key=lambda t: t[0])]))
            np.save(out + ".it.npy", np.array([p[2] for p in parts]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 8), (3, 9), (4, 8)])
def test_slab_schedule_matches_oracle(oracle, tmp_path, world, n):
    """world-size 2 / 3 / 4 gloo groups run the slab schedule (numpy compute, ghost-zone writes as messages) and must
    reproduce the unsharded oracle: checks which vertices / planes travel, where they land, and the shared stop rule"""
    dtype = np.float64
    out = str(tmp_path / "x.npy")
    mp.spawn(_slab_worker, args=(world, _free_port(), n, np.dtype(dtype).name, out), nprocs=world, join=True)
    got, its = np.load(out), np.load(out + ".it.npy")
    w = synth.c2_denoise(n, dtype, diag=True)
    A = w["La_l2"]
    ref = oracle.PFDR_graph_quadratic_d1_bounds(np.zeros(w["V"], dtype), (A * w["Y"]).astype(dtype), A, w["Eu"], w["Ev"], w["La_d1"],
                                                0.2, 2.5, 1, A, rho=1.5, condMin=1e-3, difRcd=0.0, difTol=1e-5, itMax=120, N=0)
    assert (its == ref.it).all(), (its, ref.it)
    assert np.linalg.norm(got - ref.X) / np.linalg.norm(ref.X) < 1e-10


@pytest.mark.parametrize("world", [2, 3, 5])
def test_slab_plan_invariants(world):
    n, nz = 6, 11
    V, Eu, Ev = synth.grid_graph(n, nz)
    plans = [shard.make_slab_plan(V, Eu, Ev, world, r) for r in range(world)]
    owners = np.zeros(V, int)
    covered = np.zeros(Eu.shape[0], int)
    for p in plans:
        g = shard.grid_slab_plan(n, world, p.rank, nz)             # built from the rank's planes alone: same plan
        assert (g.v0, g.v1, g.lo, g.hi) == (p.v0, p.v1, p.lo, p.hi) and np.array_equal(g.Eu, p.Eu) and np.array_equal(g.Ev, p.Ev)
        assert p.band == n * n and p.v0 % p.band == 0 and p.lo == max(p.v0 - p.band, 0) and p.hi == min(p.v1 + p.band, V)
        owners[p.v0:p.v1] += 1
        covered[p.edge_ids] += 1
        assert np.array_equal(p.Eu + p.lo, Eu[p.edge_ids]) and np.array_equal(p.Ev + p.lo, Ev[p.edge_ids])
        assert (np.diff(p.edge_ids) > 0).all() and p.Eu.min() >= 0 and p.Ev.max() < p.Vloc
        # the local graph is again a vertex-major constant-offset graph: what the stencil layout needs
        d = p.Ev - p.Eu
        assert set(d.tolist()) <= {1, n, n * n} and ((np.diff(p.Eu) > 0) | ((np.diff(p.Eu) == 0) & (np.diff(d) > 0))).all()
    assert (owners == 1).all() and ((covered == 1) | (covered == 2)).all()
    with pytest.raises(ValueError):
        shard.make_slab_plan(V, Eu, Ev, 12, 0)                      # slabs thinner than the band
