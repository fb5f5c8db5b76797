"""Vertex sharding of a PFDR problem across ranks - host logic, numpy only.

Rank r owns a contiguous range of vertices (a z-slab of a grid), all their vertex arrays and every edge
with at least one owned end.  Vertices at the far end of a cut edge are *ghosts*: numbered after the owned
ones, ordered by global id (= grouped by owner, since ranges are contiguous).  Every half-edge copy Z is
computed by exactly one rank - the owner of the vertex it belongs to.  After each iteration a rank sends to
each neighbour (i) the forward point of its boundary vertices and (ii) its half-edge copies on the cut edges,
and receives the same; that is all the data path needs (the reductions are two scalars).

Both ends of a pair derive the message layout independently from the global edge list:
  vertices    ascending global vertex id
  half-edges  ascending global edge id (an edge is cut between at most one pair of ranks)
"""
from dataclasses import dataclass, field
from typing import List

import numpy as np


def slab_bounds(V, world):
    """balanced contiguous ranges; for a z-major grid these are z-slabs"""
    return np.array([(V * r) // world for r in range(world + 1)], dtype=np.int64)


@dataclass
class Neighbor:
    rank: int
    send_vertices: np.ndarray     # local ids (owned) whose forward point the neighbour needs
    send_halfedges: np.ndarray    # local half-edge keys 2*e_local+side of OUR end of each cut edge
    recv_vertices: np.ndarray     # local ids (ghost) filled by the neighbour's message
    recv_halfedges: np.ndarray    # local half-edge keys of THEIR end of each cut edge


@dataclass
class ShardPlan:
    rank: int
    world: int
    V_global: int
    v0: int
    v1: int
    ghosts: np.ndarray            # global ids of the ghost vertices (ascending)
    edge_ids: np.ndarray          # global ids of the local edges (ascending: keeps the reference's summation order)
    Eu: np.ndarray                # local edge list in local vertex numbering (int32)
    Ev: np.ndarray
    neighbors: List[Neighbor] = field(default_factory=list)

    @property
    def Vown(self):
        return int(self.v1 - self.v0)

    @property
    def Vloc(self):
        return self.Vown + int(self.ghosts.shape[0])

    def to_local(self, g):
        g = np.asarray(g, dtype=np.int64)
        own = (g >= self.v0) & (g < self.v1)
        return np.where(own, g - self.v0, self.Vown + np.searchsorted(self.ghosts, g)).astype(np.int32)

    def vertex_array(self, a):
        """slice of a per-vertex array (leading axis = vertex): owned rows followed by the ghosts' rows"""
        if a is None:
            return None
        a = np.asarray(a)
        return np.ascontiguousarray(np.concatenate([a[self.v0:self.v1], a[self.ghosts]], axis=0))

    def edge_array(self, a):
        return None if a is None else np.ascontiguousarray(np.asarray(a)[self.edge_ids])

    def owned(self, a_local):
        return a_local[:self.Vown]


def make_plan(V, Eu, Ev, world, rank, bounds=None):
    bounds = slab_bounds(V, world) if bounds is None else np.asarray(bounds, dtype=np.int64)
    Eu = np.asarray(Eu, dtype=np.int64)
    Ev = np.asarray(Ev, dtype=np.int64)
    v0, v1 = int(bounds[rank]), int(bounds[rank + 1])
    ou = np.searchsorted(bounds, Eu, side="right") - 1
    ov = np.searchsorted(bounds, Ev, side="right") - 1
    mine = (ou == rank) | (ov == rank)
    eids = np.nonzero(mine)[0]
    gu, gv, lu, lv = Eu[eids], Ev[eids], ou[eids], ov[eids]
    ghosts = np.unique(np.concatenate([gu[lu != rank], gv[lv != rank]]))
    plan = ShardPlan(rank, world, int(V), v0, v1, ghosts, eids, None, None)
    plan.Eu, plan.Ev = plan.to_local(gu), plan.to_local(gv)
    gowner = np.searchsorted(bounds, ghosts, side="right") - 1
    e_loc = np.arange(eids.shape[0], dtype=np.int64)
    for s in np.unique(gowner):
        s = int(s)
        cut_u = (lu == rank) & (lv == s)          # our end is side 0, theirs side 1
        cut_v = (lv == rank) & (lu == s)          # our end is side 1, theirs side 0
        cut = cut_u | cut_v                       # ascending local edge id == ascending global edge id
        ours = 2 * e_loc[cut] + np.where(cut_v[cut], 1, 0)
        send_v = np.unique(np.concatenate([gu[cut_u], gv[cut_v]]))
        plan.neighbors.append(Neighbor(
            rank=s,
            send_vertices=plan.to_local(send_v),
            send_halfedges=ours.astype(np.int32),
            recv_vertices=(plan.Vown + np.nonzero(gowner == s)[0]).astype(np.int32),
            recv_halfedges=(ours ^ 1).astype(np.int32)))
    return plan


def make_plan_local(V, Eu, Ev, world, rank, bounds=None):
    """as make_plan, but Eu/Ev hold only the edges touching this rank (global vertex ids, in global edge order):
    for problems too large to enumerate every edge on every rank.  edge_ids are then positions in that list."""
    return make_plan(V, Eu, Ev, world, rank, bounds)


# ---------------------------------------------------------------------------------------------------------------
# slab sharding of constant-offset (grid) graphs: csrc/slab.cuh.  Rank r owns the contiguous range [v0, v1) and
# holds every vertex array over [lo, hi) = [v0 - band, v1 + band) clipped to the graph, numbered from lo - the local
# graph is again a constant-offset graph with the same offsets, so the index-free kernels run on it unchanged and a
# neighbour's ghost zone is a contiguous range of each array.
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class SlabPlan:
    rank: int
    world: int
    V_global: int
    band: int
    v0: int
    v1: int
    lo: int
    hi: int
    Eu: np.ndarray                # local edge list (vertex numbers relative to lo, int32), global order
    Ev: np.ndarray
    edge_ids: np.ndarray          # positions of the local edges in the edge list the plan was made from

    @property
    def Vown(self):
        return self.v1 - self.v0

    @property
    def Vloc(self):
        return self.hi - self.lo

    def vertex_array(self, a):
        """rows [lo, hi) of a global per-vertex array (the ghost rows are carried but never read)"""
        return None if a is None else np.ascontiguousarray(np.asarray(a)[self.lo:self.hi])

    def edge_array(self, a):
        return None if a is None else np.ascontiguousarray(np.asarray(a)[self.edge_ids])

    def owned(self, a_local):
        return a_local[self.v0 - self.lo:self.v1 - self.lo]


def slab_bounds_aligned(V, band, world):
    """contiguous ranges whose boundaries are multiples of `band` (whole z-planes of a grid) when there are enough planes"""
    planes = V // band if band > 0 and V % band == 0 else 0
    if planes >= 2 * world:
        return np.array([band * ((planes * r) // world) for r in range(world + 1)], dtype=np.int64)
    return slab_bounds(V, world)


def make_slab_plan(V, Eu, Ev, world, rank, bounds=None):
    """Eu < Ev in vertex-major order (the numbering the stencil layout needs); Eu/Ev may hold only the edges that
    touch this rank's range (global vertex numbers, global order)"""
    Eu = np.asarray(Eu, dtype=np.int64)
    Ev = np.asarray(Ev, dtype=np.int64)
    band = int((Ev - Eu).max()) if Eu.shape[0] else 1
    bounds = slab_bounds_aligned(V, band, world) if bounds is None else np.asarray(bounds, dtype=np.int64)
    v0, v1 = int(bounds[rank]), int(bounds[rank + 1])
    lo, hi = max(v0 - band, 0), min(v1 + band, int(V))
    if world > 1 and (bounds[1:] - bounds[:-1]).min() < band:
        raise ValueError("slab sharding: a rank's range is thinner than the band of the graph (too many ranks)")
    mine = ((Eu >= v0) & (Eu < v1)) | ((Ev >= v0) & (Ev < v1))
    eids = np.nonzero(mine)[0]
    return SlabPlan(rank, world, int(V), band, v0, v1, lo, hi, (Eu[eids] - lo).astype(np.int32), (Ev[eids] - lo).astype(np.int32), eids)


def grid_slab_plan(n, world, rank, nz=None):
    """slab plan of the n x n x nz grid of synth.grid_graph built from this rank's planes only (no global edge list)"""
    nz = n if nz is None else nz
    V, band = n * n * nz, n * n
    b = slab_bounds_aligned(V, band, world)
    v0, v1 = int(b[rank]), int(b[rank + 1])
    lo, hi = max(v0 - band, 0), min(v1 + band, V)
    g = np.arange(lo, hi, dtype=np.int64)
    z, y, x = g // band, (g // n) % n, g % n
    eu, ev = [], []
    for ok, d in ((x < n - 1, 1), (y < n - 1, n), (z < nz - 1, band)):
        eu.append(np.where(ok, g, -1)); ev.append(np.where(ok, g + d, -1))
    Eu, Ev = np.stack(eu, 1).reshape(-1), np.stack(ev, 1).reshape(-1)
    keep = (Eu >= 0) & (Ev < hi) & (((Eu >= v0) & (Eu < v1)) | ((Ev >= v0) & (Ev < v1)))
    Eu, Ev = Eu[keep], Ev[keep]
    return SlabPlan(rank, world, V, band, v0, v1, lo, hi, (Eu - lo).astype(np.int32), (Ev - lo).astype(np.int32), None)
